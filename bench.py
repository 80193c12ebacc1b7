#!/usr/bin/env python
"""bench.py — virtual-screen throughput of the B200-native FLEXS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path

Metric (BASELINE.json): sequences scored per second in a virtual screen.  One "step" = one pass of
the hot path over one synthetic candidate batch per GPU: fused CNN forward (uint8 residue indices in
HBM -> fp32 scores), per-shard top-k over distinct sequences (one launch), and — for N > 1 — ONE NCCL
all-gather of the per-shard top-k messages followed by the one-launch merge on every rank, all through
the product class flexs_b200.screen.VirtualScreen.  Weak scaling: the per-GPU batch is fixed.

Headline workload: the configuration the metric's target is quoted on in BASELINE.json's north_star
("100-mer x 4-alphabet CNN surrogate"), canonical hyper-parameters F=32, H=100, k=5
(paper_code/cloud/figure2a_data.py:19-30).  configs[1] (TF-binding 8-mer CNN, 1M candidates) and
configs[2] (RNA 14-mer, Ensemble(3xCNN)) are measured in the same run and reported under
"other_workloads".  Data: synthetic (iid uniform residues, fixed seeds), random-init weights.

Output: ONE JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "sequences_scored_per_sec_virtual_screen"
UNIT = "sequences/s"
L_NS, A_NS, F_NS, H_NS, K_NS = 100, 4, 32, 100, 5
PER_GPU_BATCH = 1 << 22          # 4,194,304 candidates per GPU per step (419 MB of uint8 > 126 MB L2)
TOPK = 99                        # sequences_batch_size=100 -> the [: -B : -1] slice keeps B-1
EXCHANGE = "peer"                # N > 1: how the per-shard top-k messages travel (--exchange); falls back to nccl by itself
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}   # B200_PROFILING.md fallback


def flop_alg(L, A, F, H, K):
    """conv1 as gather-add, everything else 2 flop per MAC."""
    T, K3 = L - K + 1, A - 1
    return T * F * K + 2 * (T * F * K * F + T * F * K3 * F + F * H + H * H + H)


def plugin_api_timings(device):
    """The two other calls an explorer makes through the drop-in classes, wall clock:
    AdaLead's ``model.get_fitness`` on eval_batch_size=20 strings (adalead.py:123) and ``model.train`` on a round's
    history (explorer.py:157; 1000 sequences, batch 256, 20 epochs = 80 Adam steps)."""
    import torch

    import flexs_b200 as flexs

    rng = np.random.default_rng(7)
    out = {}
    for tag, L, alphabet in (("cnn_100x4", 100, "ACGT"), ("cnn_237x20", 237, "ACDEFGHIKLMNPQRSTVWY")):
        model = flexs.baselines.models.CNN(L, num_filters=F_NS, hidden_size=H_NS, alphabet=alphabet,
                                           loss="MSE", device=device.index or 0, seed=0)
        letters = np.array(list(alphabet))
        seqs = ["".join(r) for r in letters[rng.integers(0, len(alphabet), size=(1000, L))]]
        labels = rng.random(1000)
        model.train(seqs[:64], labels[:64])          # warm-up: workspace allocation
        torch.cuda.synchronize()
        fits = []
        for _ in range(3):  # the first full-size fit also grows the training workspace: report the median
            t0 = time.perf_counter()
            model.train(seqs, labels)
            fits.append(time.perf_counter() - t0)
        fit_s = float(np.median(fits))
        small = seqs[:20]
        for _ in range(5):
            model.get_fitness(small)
        lat = []
        for _ in range(50):
            t0 = time.perf_counter()
            model.get_fitness(small)
            lat.append(time.perf_counter() - t0)
        out[tag] = {"train_1000seq_20epochs_s": fit_s, "train_ms_per_adam_step": fit_s / 80 * 1e3,
                    "get_fitness_20seq_latency_us_median": float(np.median(lat) * 1e6),
                    "get_fitness_20seq_latency_us_p90": float(np.quantile(lat, 0.9) * 1e6)}
        del model
    out["vae_generator"] = vae_timings(device)
    out["table_landscapes"] = landscape_timings(device)
    out["unique_ranking"] = dedup_timing(device)
    return out


def vae_timings(device):
    """K9: one refit of the CbAS generator as cbas_dbas.py:183 drives it — 1000 14-mers, batch 10,
    10 epochs = 800 optimiser steps (validation split 0.2) — and one 100-proposal cycle's decode + log-probability calls."""
    from flexs_b200.utils import sequence_utils as su
    from flexs_b200.utils.VAE_utils import VAE

    seqs = su.generate_random_sequences(14, 1000, su.RNAA)
    vae = VAE(seq_length=14, alphabet=su.RNAA, batch_size=10, latent_dim=2, intermediate_dim=250, epochs=10, verbose=False, seed=0)
    vae.train_model(seqs[:100], np.ones(100))                # warm-up: workspace allocation
    t0 = time.perf_counter(); vae.train_model(seqs, np.ones(1000)); fit_s = time.perf_counter() - t0
    steps = len(vae.last_fit_losses) * 80
    t0 = time.perf_counter()
    for _ in range(10):
        vae.calculate_log_probability(seqs[:100])
    lp_s = (time.perf_counter() - t0) / 10
    return {"refit_1000seq_s": fit_s, "epochs_run": int(len(vae.last_fit_losses)), "ms_per_optimizer_step": 1e3 * fit_s / max(steps, 1),
            "log_probability_100seq_us": lp_s * 1e6, "native": bool(vae.native)}


def dedup_timing(device):
    """K3 / K3b / K3c: ranking of a device-resident batch over DISTINCT sequences (what VirtualScreen adds after the
    forward pass), CUDA events: the single-launch selection (select.cu) beside the hash de-duplication + 11-launch radix
    select it replaces.  1M draws from the 65 536 8-mers (>= 93 % repeats) and 2^22 random 100-mers (no repeats)."""
    import torch

    from flexs_b200 import _native

    out = {}
    for tag, L, n in (("8mer_1M", 8, 1 << 20), ("100mer_4M", 100, 1 << 22)):
        idx = torch.randint(0, 4, (n, L), dtype=torch.uint8, device=device)
        # equal rows must carry equal scores (a deterministic surrogate): score = a function of the row
        w = torch.randn(L, 4, device=device)
        scores = w[torch.arange(L, device=device), idx.long()].sum(dim=1).contiguous()
        masked = torch.empty_like(scores)
        dwork = torch.empty(_native.dedup_workspace_bytes(n), dtype=torch.uint8, device=device)
        twork = torch.empty(_native.topk_workspace_bytes(n, TOPK), dtype=torch.uint8, device=device)
        swork = torch.empty(_native.topk_select_workspace_bytes(), dtype=torch.uint8, device=device)
        ts = torch.empty(TOPK, dtype=torch.float32, device=device)
        ti = torch.empty(TOPK, dtype=torch.int64, device=device)
        ts2, ti2 = torch.empty_like(ts), torch.empty_like(ti)
        status = torch.zeros(8, dtype=torch.int32, device=device)
        s = torch.cuda.current_stream().cuda_stream

        def two_call():
            _native.dedup_scores_dev(idx.data_ptr(), n, L, scores.data_ptr(), masked.data_ptr(), dwork.data_ptr(), s)
            _native.topk_dev(masked.data_ptr(), n, TOPK, 0, 0, ts.data_ptr(), ti.data_ptr(), twork.data_ptr(), s)

        def single():
            _native.topk_select_dev(scores.data_ptr(), n, TOPK, 0, idx.data_ptr(), L, True, ts2.data_ptr(), ti2.data_ptr(), 0,
                                    status.data_ptr(), swork.data_ptr(), s)

        res = {}
        for name, fn in (("dedup_plus_topk_ms", two_call), ("select_single_launch_ms", single)):
            for _ in range(3):
                fn()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(5):
                fn()
            ev[1].record()
            torch.cuda.synchronize()
            res[name] = ev[0].elapsed_time(ev[1]) / 5
        st = status.cpu().tolist()
        res["identical_winners"] = bool(torch.equal(ti, ti2) and torch.equal(ts, ts2)) and st[0] == 0
        res["select_diagnostics"] = {"radix_levels": st[1], "candidates": st[2], "select_us": st[3] / 1e3, "final_cta_us": st[4] / 1e3,
                                     "final_narrow_us": st[5] / 1e3, "final_sort_us": st[6] / 1e3, "final_dedup_us": st[7] / 1e3}
        res["sequences_per_s"] = n / (res["select_single_launch_ms"] / 1e3)
        res["distinct_in_top"] = int(len({bytes(r) for r in idx[ti2.clamp(min=0)].cpu().numpy()}))
        out[tag] = res
    return out


def landscape_timings(device):
    """K6: the additive AAV landscape on device-resident candidates, CUDA events.
    HBM-bound byte work: 90 B read + 8 B written per sequence."""
    import torch

    import flexs_b200 as flexs

    fixture = REPO / "tests" / "golden" / "aav_450_540_subs.json"
    if not fixture.exists():
        return None
    land = flexs.landscapes.AdditiveAAVPackaging(phenotype="heart", start=450, end=540, data_file=str(fixture),
                                                 device=device.index or 0)
    n = 1 << 22
    cols = torch.randint(0, len(land.residues), (n, land.seq_len), dtype=torch.uint8, device=device)
    for _ in range(3):
        land.get_fitness_device(cols, columns=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        land.get_fitness_device(cols, columns=True)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    return {"additive_aav_90mer": {"value": n / (ms / 1e3), "unit": UNIT, "kernel_ms": ms,
                                   "hbm_gbs_alg": n * (land.seq_len + 8) / (ms / 1e3) / 1e9, "dtype": "f64"}}


def gather_rank_stats(stats, world):
    """Every rank's small dict of timings -> list on every rank (N > 1: explains `ms_per_step`, which follows the slowest
    rank).  Never fails the bench: returns None if the collective does."""
    if world <= 1:
        return [stats]
    try:
        import torch.distributed as dist

        out = [None] * world
        dist.all_gather_object(out, stats)
        return out
    except Exception as exc:  # noqa: BLE001
        sys.stderr.write(f"per-rank statistics unavailable: {exc}\n")
        return None


def load_peaks():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            d = json.load(open(path))
            flat = {}

            def walk(x):
                if isinstance(x, dict):
                    for k, v in x.items():
                        if isinstance(v, (int, float)):
                            flat.setdefault(k, float(v))
                        else:
                            walk(v)
            walk(d)
            if "bf16_tflops" in flat and "hbm_gbs" in flat:
                return {"hbm_gbs": flat["hbm_gbs"], "bf16_tflops": flat["bf16_tflops"],
                        "bf16_tflops_sustained": flat.get("bf16_tflops_sustained")}, "measured"
        except Exception:  # noqa: BLE001
            pass
    return dict(FALLBACK_PEAKS), "fallback"


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons of one GPU through NVML during the timed region."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, torch_index: int):
        super().__init__(daemon=True)
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            handle = None
            try:
                uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:  # noqa: BLE001
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[torch_index]) if vis and vis.split(",")[torch_index].isdigit() else torch_index
                handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv, self._h = pynvml, handle
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM)))
                try:
                    bits = self._nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:  # noqa: BLE001
                    bits = self._nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.reason_bits |= int(bits)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        reasons = [name for bit, name in self.REASONS.items() if self.reason_bits & bit]
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def make_weights(shapes, seed):
    """Keras default init (glorot-uniform kernels, zero biases), numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    out = []
    for shp in shapes:
        if len(shp) == 1:
            out.append(np.zeros(shp, dtype=np.float32))
        else:
            rec = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            lim = np.sqrt(6.0 / (rec * shp[-2] + rec * shp[-1]))
            out.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
    return out


def cnn_shapes(L, A, F, H, K):
    return [(K, A, F), (F,), (K, F, F), (F,), (A - 1, F, F), (F,), (F, H), (H,), (H, H), (H,), (H, 1), (1,)]


ALPHABETS = {4: "TGCA", 20: "ILVAGMFYWEDQNHCRKSTP"}


def mlp_shapes(L, A, H):
    return [(L * A, H), (H,), (H, H), (H,), (H, H), (H,), (H, 1), (1,)]


class Screen:
    """One rank's share of the virtual screen, through the PRODUCT classes: a ``flexs_b200`` surrogate (``CNN`` / ``MLP`` /
    fused ``Ensemble``) under ``flexs_b200.screen.VirtualScreen`` (unique=True: the reference ranks the keys of a dict).
    A step = forward over the rank's shard + one selection launch (+ ONE all-gather + one merge launch for N > 1)."""

    def __init__(self, L, A, F, H, K, members, batch, rank, world, device, kind="cnn"):
        import torch

        import flexs_b200 as flexs
        from flexs_b200.screen import VirtualScreen

        self.torch = torch
        self.L, self.A, self.batch, self.rank, self.world, self.device = L, A, batch, rank, world, device
        alphabet = ALPHABETS[A]
        models = []
        for mem in range(members):
            if kind == "cnn":
                m = flexs.baselines.models.CNN(L, F, H, alphabet, kernel_size=K, device=device.index, seed=mem)
                m.set_weights(make_weights(cnn_shapes(L, A, F, H, K), mem))
            else:
                m = flexs.baselines.models.MLP(L, H, alphabet, device=device.index, seed=mem)
                m.set_weights(make_weights(mlp_shapes(L, A, H), mem))
            models.append(m)
        self.surrogate = models[0] if members == 1 else flexs.Ensemble(models)
        self.model = models[0].native if members == 1 else self.surrogate._fused_model().native   # the native object scored
        gen = torch.Generator(device=device)
        gen.manual_seed(1234 + rank)
        self.idx = torch.randint(0, A, (batch, L), dtype=torch.uint8, device=device, generator=gen)
        self.k = TOPK
        # N > 1: EXCHANGE = "peer": one-sided message stores over NVLink peer memory, the merge of step i issued with step
        # i + 1 (csrc/peer.cu); "nccl": one all_gather per step; "nccl-overlap": that all_gather on a side stream
        self.vs = VirtualScreen(self.surrogate, k=self.k, unique=True, overlap=(world > 1 and EXCHANGE == "nccl-overlap"),
                                exchange="peer" if (world > 1 and EXCHANGE == "peer") else "nccl")
        self.fwd_ms = []
        self.status_sum = torch.zeros(1, dtype=torch.int32, device=device)
        self.scores = None
        self.result = None

    @property
    def launches(self):
        return self._model_launches + self.vs.launches

    def reset_counters(self):
        self._l0 = self.model.launch_count
        self._model_launches = 0
        self.vs.launches = 0
        self.vs.forward_events = []
        self.status_sum.zero_()

    def step(self):
        # check=False: no host sync inside the timed region; the selection's status flags are summed on the device and
        # read after it (0 = the lazy de-duplication found k distinct sequences every time)
        _, _, self.scores = self.vs.local_topk(self.idx, self.rank * self.batch, check=False)
        self.status_sum += self.vs.last_status[:1]
        self.result = self.vs.merge(self.L, self.device)
        self._model_launches = self.model.launch_count - self._l0


def timed_steps(screen, steps, warmup, world, device):
    """W warm-up steps, then exactly K steps between barrier+synchronize, CUDA events, max over ranks."""
    import torch

    screen.reset_counters()
    for _ in range(warmup):
        screen.step()
    torch.cuda.synchronize(device)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize(device)
    sampler = ClockSampler(device.index)
    sampler.start()
    screen.reset_counters()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        screen.step()
    screen.vs.wait()   # overlap mode: the last steps' all-gather + merge launches are inside the timed region
    end.record()
    torch.cuda.synchronize(device)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize(device)
    clocks = sampler.stop()
    ms = start.elapsed_time(end)
    screen.own_ms = ms   # this rank's own device time (the reported one is the max over ranks)
    screen.fwd_ms = [a.elapsed_time(b) for a, b in screen.vs.forward_events]
    screen.vs.forward_events = None
    assert int(screen.status_sum.item()) == 0, "selection fell short of k distinct sequences: the step must use check=True"
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, clocks


def measure_e2e(screen, L, A, steps, world, device):
    """Same metric through the reference-facing C-ABI call with HOST buffers (what Model.get_fitness calls for a large
    list): candidates in pinned host memory in the packed wire format of include/flexs_b200.h (2 bits per DNA residue),
    host scores out; H2D + unpack + forward + D2H inside the timing."""
    import torch

    from flexs_b200 import _native

    rb = _native.packed_row_bytes(L, A)
    d_packed = torch.empty((screen.batch, rb), dtype=torch.uint8, device=device)
    _native.pack_dev(screen.idx.data_ptr(), screen.batch, L, A, d_packed.data_ptr(), torch.cuda.current_stream().cuda_stream)
    packed = d_packed.cpu().pin_memory()
    out = torch.empty(screen.batch, dtype=torch.float32).pin_memory()
    packed_np, out_np = packed.numpy(), out.numpy()
    screen.model.score_host_packed(packed_np[:4096])             # allocate staging, warm up
    screen.model.score_host_packed(packed_np, out_np)
    torch.cuda.synchronize(device)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    l0 = screen.model.launch_count
    t0 = time.perf_counter()
    for _ in range(steps):
        screen.model.score_host_packed(packed_np, out_np)
    dt = time.perf_counter() - t0
    launches = screen.model.launch_count - l0
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([dt], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    same = bool(np.array_equal(out_np, screen.scores.cpu().numpy()))   # identical to the device-resident path
    return {"value": world * screen.batch * steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(screen.batch * rb),
            "d2h_bytes_per_step": int(screen.batch * 4), "steps": steps, "matches_device_path": same,
            "api": "flexs_model_score_host_packed: pinned host rows in the packed wire format (2 bits per residue)",
            "gpu_launches": int(launches)}


def measure_e2e_strings(screen, L, A, device, n=1 << 20):
    """The reference's own signature, end to end: ``CNN.get_fitness(list[str])`` (flexs/landscape.py:29-45) on a Python
    list of n strings — the C pass over the str objects (alphabet lookup + bit packing, multi-threaded), pageable ->
    pinned staging, H2D, unpack, forward, D2H, all inside the timing.  Rank 0 only."""
    import torch

    from flexs_b200.utils import sequence_utils as su

    alphabet = ALPHABETS[A]
    idx = screen.idx[:n].cpu().numpy()
    seqs = su.decode_indices(idx, alphabet).tolist()
    model = screen.surrogate
    want = screen.scores[:n].cpu().numpy()
    model.get_fitness(seqs[:8192])
    t0 = time.perf_counter(); su.pack_sequences(seqs, alphabet); pack_s = time.perf_counter() - t0
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        got = model.get_fitness(seqs)
        times.append(time.perf_counter() - t0)
    torch.cuda.synchronize(device)
    dt = float(np.median(times))
    return {"value": n / dt, "unit": UNIT, "n": n, "seconds": dt, "host_pack_seconds": pack_s,
            "host_threads": su.host_threads(), "matches_device_path": bool(np.array_equal(got, want)),
            "api": "flexs_b200.baselines.models.CNN.get_fitness(list[str])",
            "note": "bounded by touching one Python str object per sequence on the host, not by PCIe or the GPU"}


def cpu_baseline(L, A, F, H, K, members=1, seconds_target=12.0):
    """The oracle's plain-C restatement (OpenMP, all host cores) on a bounded sample of the workload."""
    from oracle import c_oracle as co

    ws = [make_weights(cnn_shapes(L, A, F, H, K), m) for m in range(members)]
    rng = np.random.default_rng(99)
    probe = rng.integers(0, A, size=(2048, L), dtype=np.uint8)
    co.cnn_forward(probe[:64], ws, K)
    t0 = time.perf_counter(); co.cnn_forward(probe, ws, K); rate = len(probe) / (time.perf_counter() - t0)
    n = int(max(4096, min(rate * seconds_target, 2_000_000)))
    idx = rng.integers(0, A, size=(n, L), dtype=np.uint8)
    t0 = time.perf_counter(); co.cnn_forward(idx, ws, K); dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} synthetic {L}-mers (A={A}) through oracle/c/oracle.c (fp32, OpenMP static schedule), "
                      f"forward only, {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of get_fitness, restated (TensorFlow is
    absent): Python-loop float64 one-hot per sequence (sequence_utils.py:32-47), np.array -> fp32
    tensor (keras_model.py:70-75), predict in batches of 256 (keras_model.py:20,77) through
    oneDNN-backed torch-CPU conv1d/linear standing in for TF's CPU kernels, squeeze, nan_to_num."""
    if rank != 0:
        return
    import torch

    from oracle import ref_path

    # torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host core.  Batch-256 convolutions
    # do not scale to 128 threads (oversubscription makes them slower), so pick the thread count that maximises the
    # reference's own throughput on a probe, as TF's intra-op pool sizing would.
    ncpu = os.cpu_count() or 1
    probe_w = make_weights(cnn_shapes(L_NS, A_NS, F_NS, H_NS, K_NS), 0)
    probe_seqs = ["".join("TGCA"[i] for i in row) for row in np.random.default_rng(7).integers(0, 4, size=(1024, L_NS))]
    best_threads, best_rate = 1, 0.0
    for nthreads in sorted({1, 4, 8, 16, 32, 64, ncpu}):
        if nthreads > ncpu:
            continue
        torch.set_num_threads(nthreads)
        probe = ref_path.ReferenceCNN(L_NS, "TGCA", F_NS, H_NS, K_NS, probe_w)
        probe.get_fitness(probe_seqs[:256])
        t0 = time.perf_counter(); probe.get_fitness(probe_seqs); rate = len(probe_seqs) / (time.perf_counter() - t0)
        if rate > best_rate:
            best_threads, best_rate = nthreads, rate
    torch.set_num_threads(best_threads)
    model = ref_path.ReferenceCNN(L_NS, "TGCA", F_NS, H_NS, K_NS, make_weights(cnn_shapes(L_NS, A_NS, F_NS, H_NS, K_NS), 0))
    n = 4096
    rng = np.random.default_rng(1234)
    seqs = ["".join("TGCA"[i] for i in row) for row in rng.integers(0, 4, size=(n, L_NS))]
    for _ in range(args.warmup):
        model.get_fitness(seqs[:512])
    t0 = time.perf_counter()
    enc_s = 0.0
    for _ in range(args.steps):
        model.get_fitness(seqs)
        enc_s += model.last_encode_seconds
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "north_star_cnn_100x4", "seq_len": L_NS, "alphabet": 4, "num_filters": F_NS,
                   "hidden": H_NS, "kernel_size": K_NS, "per_step_sample": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": model.threads, "kind": "port",
                         "sample": f"{n} sequences/step through the restated reference get_fitness "
                                   f"(Python one-hot loop {100 * enc_s / dt:.0f}% of the time, then batch-256 predict on "
                                   f"{model.threads} torch-CPU threads); TensorFlow itself is not installed"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


_JSON_OUT = sys.stdout


def main():
    global EXCHANGE
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--exchange", choices=["peer", "nccl", "nccl-overlap"], default=EXCHANGE,
                    help="N > 1: how the per-shard top-k messages travel (see flexs_b200/screen.py)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="candidates per GPU per step")
    ap.add_argument("--variant", default="auto", choices=["auto", "simple", "tiled", "umma", "lut"])
    ap.add_argument("--skip-extras", action="store_true", help="skip cpu_baseline / e2e / other workloads")
    args = ap.parse_args()
    EXCHANGE = args.exchange
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch

    from flexs_b200 import _native

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: flexs_b200 has no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner out of stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=device)
    assert world == max(1, args.gpus) or world == 1, "launch N>1 through torch.distributed.run"

    peaks, peak_src = load_peaks()
    screen = Screen(L_NS, A_NS, F_NS, H_NS, K_NS, 1, args.batch, rank, world, device)
    variant = {"auto": 0, "simple": 1, "tiled": 2, "umma": 3, "lut": 4}[args.variant]
    screen.model.set_variant(variant)
    ms, clocks = timed_steps(screen, args.steps, args.warmup, world, device)
    total_seqs = world * args.batch * args.steps
    value = total_seqs / (ms / 1e3)
    fwd_ms = float(np.mean(screen.fwd_ms))
    fa = flop_alg(L_NS, A_NS, F_NS, H_NS, K_NS)
    achieved_tf = fa * args.batch / (fwd_ms / 1e3) / 1e12
    kernel_name = _native.VARIANT_NAMES[screen.model.active_variant(args.batch)]
    roofline = {
        "bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
        "frac": achieved_tf / peaks["bf16_tflops"], "traffic": None, "peak_source": peak_src,
        "kernel": kernel_name,
        "kernel_ms": fwd_ms, "kernel_share_of_step": fwd_ms * args.steps / ms,
        "flop_alg_per_seq": fa, "bytes_alg_per_seq": L_NS + 4,
        "hbm_gbs_alg": (L_NS + 4) * args.batch / (fwd_ms / 1e3) / 1e9,
        "hbm_frac": (L_NS + 4) * args.batch / (fwd_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
        "fp32_ffma_peak_tflops_nominal": 148 * 128 * 2 * 1.965e9 / 1e12,
        "note": "compute-bound path (1.55e4 flop/B): fraction is of the dense bf16 tensor peak; "
                "algorithmic flops count each MAC once even though the tcgen05 kernels run 3 fp16 hi/lo split products per MAC",
    }
    if kernel_name == "lut9_umma_tcgen05":
        # One forward = ceil(batch / 1 060 864) launches of cnn_k9_kernel (conv1+conv2 as one 128-byte gather per
        # position from an L2-resident table, conv3 on tcgen05) each followed by cnn_k9_dense_kernel; kernel_ms covers
        # both (ncu launch list, profiles/r01_k9_launch_list.txt: 91.5 % / 8.5 %).  Tensor work actually issued per
        # sequence: conv3 as 6 (M128 N64 K16) + 6 (M128 N32 K16) MMAs per 128-row tile of 8 sequences x 16 positions,
        # the dense head as 2 + 7 pairs of (N224, N112) MMAs per 128 sequences.
        T = L_NS - K_NS + 1
        tiles_per_seq = ((T + 15) // 16) / 8
        mac_exec = tiles_per_seq * 6 * (128 * 64 * 16 + 128 * 32 * 16) + 9 * (128 * 224 * 16 + 128 * 112 * 16) / 128
        exec_tf = 2 * mac_exec * args.batch / (fwd_ms / 1e3) / 1e12
        launches_per_fwd = -(-args.batch // (148 * 56 * 128))
        seq_per_launch = args.batch / launches_per_fwd
        roofline.update({
            "kernel": "cnn_k9_pair_kernel (cta_group::2 CTA pairs; + cnn_k9_dense_kernel)",
            "tensor_flop_executed_per_seq": 2 * mac_exec, "tensor_tflops_executed": exec_tf,
            "tensor_frac_executed": exec_tf / peaks["bf16_tflops"],
            "sequences_per_launch": seq_per_launch,
            "note": "compute-bound path: `achieved` counts the algorithmic flops of the whole layer stack once per MAC "
                    "(SURVEY.md 8d), although conv1+conv2 are served by a table lookup and conv3/dense run 3 fp16 hi/lo "
                    "split products per MAC; tensor_tflops_executed is what the tensor pipe really ran",
        })
        # traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE cnn_k9_kernel launch, read from the committed
        # `ncu --set full` capture of the shipped kernel (profiles/r02_k9_ncu.json, written by tools/ncu_extract.py from
        # the .ncu-rep): per-sequence bytes scale with the launch size, the first touch of the L2-resident table does not
        cap_path = REPO / "profiles" / "r02_k9_ncu.json"
        if cap_path.exists():
            cap = json.load(open(cap_path))
            per_seq = (cap["dram_bytes_per_launch"] - cap["table_bytes"]) / cap["sequences_per_launch"]
            roofline.update({"traffic": cap["table_bytes"] + per_seq * seq_per_launch,
                             "traffic_unit": "bytes per conv-kernel launch",
                             "traffic_source": f"{cap_path.name}: {cap['dram_bytes_per_launch']:.4g} B measured for "
                                               f"{cap['sequences_per_launch']} sequences, scaled to this launch size",
                             "traffic_over_algorithmic": (cap["table_bytes"] + per_seq * seq_per_launch) /
                                                         ((L_NS + 4) * seq_per_launch)})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "north_star_cnn_100x4", "seq_len": L_NS, "alphabet": A_NS, "num_filters": F_NS,
                   "hidden": H_NS, "kernel_size": K_NS, "per_gpu_batch": args.batch, "topk": TOPK,
                   "parallelism": f"candidate-shard x{world}, one exchange of the per-shard top-k messages per step" if world > 1 else "single GPU",
                   "screen": "flexs_b200.screen.VirtualScreen(unique=True): forward + one selection launch"
                             + ({"peer": " + one push launch (stores into every rank's mailbox over NVLink peer memory) + wait + merge launches, issued one step late",
                                 "nccl": " + one all-gather + one merge launch",
                                 "nccl-overlap": " + one all-gather + one merge launch on a side stream"}[EXCHANGE] if world > 1 else ""),
                   "l2_policy": f"inputs larger than L2 ({args.batch * L_NS / 1e6:.0f} MB of uint8 per GPU per step)"},
        "clocks": clocks, "roofline": roofline, "gpu_launches": int(screen.launches),
    }
    if world > 1:
        # ranks are coupled once per step (the exchange of the per-shard winners): the step follows the slowest GPU of the box
        per_rank = gather_rank_stats({"rank": rank, "forward_ms": round(float(np.mean(screen.fwd_ms)), 4),
                                      "forward_ms_max": round(float(np.max(screen.fwd_ms)), 4),
                                      "own_step_ms": round(screen.own_ms / args.steps, 4),
                                      "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")}, world)
        if per_rank is not None:
            line["per_rank"] = per_rank
            line["slowest_rank_forward_ms"] = max(r["forward_ms"] for r in per_rank)
        line["config"]["exchange"] = getattr(screen.vs, "exchange", EXCHANGE)
    if not args.skip_extras:
        line["e2e"] = measure_e2e(screen, L_NS, A_NS, max(2, min(args.steps, 4)), world, device)
        if rank == 0:
            line["e2e_strings"] = measure_e2e_strings(screen, L_NS, A_NS, device)
        if rank == 0 and world == 1:
            line["cpu_baseline"] = cpu_baseline(L_NS, A_NS, F_NS, H_NS, K_NS)
        elif rank == 0:
            line["cpu_baseline"] = None
        others = {}
        for tag, (L, A, members, batch, kind) in {"tfbind8_cnn_1M (configs[1])": (8, 4, 1, 1 << 20, "cnn"),
                                                  "tfbind8_mlp_1M (configs[0] surrogate)": (8, 4, 1, 1 << 20, "mlp"),
                                                  "mlp_100x4_H100": (100, 4, 1, 1 << 21, "mlp"),
                                                  "rna14_ens3cnn_1M (configs[2])": (14, 4, 3, 1 << 20, "cnn"),
                                                  "aav735_cnn (configs[3], per-GPU shard)": (735, 20, 1, 1 << 15, "cnn"),
                                                  "aav90_cnn (AAV registry window)": (90, 20, 1, 1 << 18, "cnn"),
                                                  "gfp237_cnn (configs[4], per-GPU shard)": (237, 20, 1, 1 << 17, "cnn")}.items():
            sc = Screen(L, A, F_NS, H_NS, K_NS, members, batch, rank, world, device, kind=kind)
            try:
                sc.model.set_variant(variant)
            except ValueError:  # a forced variant that this shape does not have: let the library choose
                sc.model.set_variant(0)
            oms, _ = timed_steps(sc, max(3, args.steps), args.warmup, world, device)
            st = max(3, args.steps)
            fa_o = flop_alg(L, A, F_NS, H_NS, K_NS) if kind == "cnn" else L * H_NS + 2 * (2 * H_NS * H_NS + H_NS)
            tf_o = members * fa_o * batch / (np.mean(sc.fwd_ms) / 1e3) / 1e12
            others[tag] = {"value": world * batch * st / (oms / 1e3), "unit": UNIT, "ms_per_step": oms / st,
                           "kernel_ms": float(np.mean(sc.fwd_ms)), "members": members,
                           "achieved_tflops": tf_o, "frac_of_bf16_peak": tf_o / peaks["bf16_tflops"],
                           "hbm_gbs_alg": (L + 4) * batch / (np.mean(sc.fwd_ms) / 1e3) / 1e9,
                           "kernel": ("mlp" if kind == "mlp" else "") + _native.VARIANT_NAMES[sc.model.active_variant(batch)],
                           "note": "input <= L2 size (re-read from L2 between steps); compute-bound path, so unaffected"}
            del sc
        line["other_workloads"] = others
        if rank == 0:
            line["plugin_api"] = plugin_api_timings(device)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), file=_JSON_OUT, flush=True)


if __name__ == "__main__":
    main()
