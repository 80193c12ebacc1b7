"""Turns the raw ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py <tag>      # e.g. r01_k9 -> profiles/r01_k9_launch_list.txt, profiles/r01_k9_ncu_summary.txt

Inputs: gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum over bench.py) and gpurun_out/<tag>.ncu-rep
(ncu --set full of the forward kernels).  Needs the `ncu` CLI to read the report (no GPU)."""
import collections
import csv
import io
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_k9"
note = sys.argv[2] if len(sys.argv) > 2 else ""

# ---- launch list ---------------------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(REPO / "gpurun_out" / "launches.csv")) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, data = r, rows[i + 1:]
        break
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[ik])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
out = ["ncu launch list of `python bench.py --steps 2 --warmup 1 --skip-extras` (1x B200, north-star workload: 2^22 100-mers/step)",
       "command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py ...",
       "Per-launch times are cold-cache and serialised: the SHARES are what bench.py's live CUDA-event numbers must agree with.",
       note, "", f"{'ms total':>10s} {'share':>6s} {'launches':>8s} {'avg ms':>9s}  kernel"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{t:10.3f} {100 * t / tot:5.1f}% {c:8d} {t / c:9.4f}  {k}")
fwd = {k: v[1] for k, v in agg.items() if "cnn_k9" in k}
if fwd:
    ft = sum(fwd.values())
    out += ["", "share of the forward pass: " + ", ".join(f"{re.sub('.*::', '', k)} {100 * v / ft:.1f} %" for k, v in fwd.items())]
out += ["per step: ceil(2^22 / 1 060 864) = 4 x (cnn_k9_kernel + cnn_k9_dense_kernel), 1 gated cnn_tiled_kernel (fp16 range guard, returns",
        "immediately), 11 top-k launches; k9_build_kernel once per weight set; the torch random kernel is bench.py creating the synthetic batch."]
(REPO / "profiles" / f"{tag}_launch_list.txt").write_text("\n".join(out) + "\n")

# ---- full capture ---------------------------------------------------------------------------------------------------
rep = REPO / "gpurun_out" / f"{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_requests_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum", "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
out = [f"ncu --set full --clock-control none --import-source on, one launch each of the forward kernels (tools/k9_perf.py 100 20: 2^20 100-mers)", note]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    out += ["", "== " + d["Kernel Name"]]
    for h in want:
        if h in d:
            out.append(f"  {h:84s} {d[h]:>18s} {units[hdr.index(h)]}")
    out.append("  warp stall reasons (warps per issue-active cycle):")
    st = {h: float(d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and d[h]}
    for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
        out.append(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:6.2f}")
(REPO / "profiles" / f"{tag}_ncu_summary.txt").write_text("\n".join(out) + "\n")
print("\n".join(out[:60]))
