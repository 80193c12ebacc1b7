"""Turn an `ncu --set full` report that a gpurun call left in gpurun_out/ into a tracked summary under profiles/ (and, for
cnn_k9_kernel, the JSON bench.py reads `roofline.traffic` from).  Needs the `ncu` CLI (no GPU).

    python tools/ncu_extract.py gpurun_out/r02_k9.ncu-rep profiles/r02_k9_ncu_summary.txt "what was run" [--k9-json SEQS]
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

rep, out_path, note = sys.argv[1], sys.argv[2], sys.argv[3]
k9_seqs = int(sys.argv[sys.argv.index("--k9-json") + 1]) if "--k9-json" in sys.argv else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
out = [f"ncu --set full --clock-control none --import-source on  ({Path(rep).name})", note]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    out += ["", "== " + d["Kernel Name"]]
    for h in want:
        if h in d and d[h] != "":
            out.append(f"  {h:84s} {d[h]:>18s} {units[hdr.index(h)]}")
    out.append("  warp stall reasons (warps per issue-active cycle):")
    st = {h: float(d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and d[h]}
    for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
        out.append(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:6.2f}")
    if k9_seqs and ("cnn_k9_kernel" in d["Kernel Name"] or "cnn_k9_pair_kernel" in d["Kernel Name"]):
        def num(key):
            v = float(d[key].replace(",", ""))
            u = units[hdr.index(key)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        total = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        js = {"kernel": d["Kernel Name"].split("(")[0].split("::")[-1].strip(), "capture": Path(rep).name, "sequences_per_launch": k9_seqs, "dram_bytes_per_launch": total,
              "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
              "table_bytes": 425984 * 128, "algorithmic_bytes_per_launch": 104 * k9_seqs,
              "tensor_pipe_active_pct": float(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "nan") or "nan"),
              "gpu_time_us": float(d["gpu__time_duration.sum"].replace(",", "")) *
                             {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6}.get(
                                 units[hdr.index("gpu__time_duration.sum")], 1.0)}
        Path(out_path).with_name(Path(out_path).name.replace("_ncu_summary.txt", "_ncu.json")).write_text(json.dumps(js, indent=1) + "\n")
Path(out_path).write_text("\n".join(out) + "\n")
print("\n".join(out[:70]))
