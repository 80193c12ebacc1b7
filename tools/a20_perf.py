"""Timing of the A = 20 (k3 = 19) forward kernels on one GPU (device-resident batch, CUDA events), and agreement between
the tcgen05 kernel and the FP32 FFMA kernel on the same batch.

    python tools/a20_perf.py [tag ...]        # tags: aav90 gfp237 gfp238 aav735 ; FLEXS_UMMA_V1=1 times the round-1 kernel
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

from flexs_b200 import _native
from _weights import cnn_shapes, trained_like

SHAPES = {"aav90": (90, 1 << 18), "gfp237": (237, 1 << 17), "gfp238": (238, 1 << 17), "aav735": (735, 1 << 15),
          "aa300": (300, 1 << 16)}
A = 20


def flop_alg(L, A=20, F=32, H=100, K=5):
    T, K3 = L - K + 1, A - 1
    return T * F * K + 2 * (T * F * K * F + T * F * K3 * F + F * H + H * H + H)


for tag in (sys.argv[1:] or ["aav90", "gfp237", "aav735"]):
    L, n = SHAPES[tag]
    ws = trained_like(cnn_shapes(L, A), 5)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(ws)
    g = torch.Generator(device="cuda").manual_seed(1234)
    idx = torch.randint(0, A, (n, L), dtype=torch.uint8, device="cuda", generator=g)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    res = {}
    for v, reps in ((_native.VARIANT_UMMA, 5), (_native.VARIANT_TILED, 1)):
        m.set_variant(v)
        nn = n if v == _native.VARIANT_UMMA else min(n, 1 << 14)
        for _ in range(2):
            m.forward_dev(idx.data_ptr(), nn, out.data_ptr(), s)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            m.forward_dev(idx.data_ptr(), nn, out.data_ptr(), s)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        res[v] = out[:nn].cpu().numpy().copy()
        print(f"{tag} L={L} n={nn} {_native.VARIANT_NAMES[v]:>14s}: {ms:9.3f} ms  {nn / ms * 1e3:.4g} seq/s  "
              f"{flop_alg(L) * nn / ms * 1e-9:.1f} TFLOP/s alg", flush=True)
    a, b = res[_native.VARIANT_UMMA][: 1 << 14], res[_native.VARIANT_TILED]
    print(f"{tag}: max |tcgen05 - ffma| / scale = {float(np.abs(a - b).max() / np.abs(b).max()):.3e}", flush=True)
    m.close()
