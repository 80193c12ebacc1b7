"""A/B timing of the A = 4 forward kernels on one GPU (device-resident batch, CUDA events).

    python tools/k9_perf.py [L] [log2 n]      # FLEXS_UMMA_PROF=1 adds the kernels' phase counters on stderr
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

from flexs_b200 import _native
from _weights import cnn_shapes, trained_like

L = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
A = 4
ws = trained_like(cnn_shapes(L, A), 5)
m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
m.set_weights(ws)
g = torch.Generator(device="cuda").manual_seed(1234)
idx = torch.randint(0, A, (n, L), dtype=torch.uint8, device="cuda", generator=g)
out = torch.empty(n, dtype=torch.float32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
res = {}
for v in (_native.VARIANT_UMMA, _native.VARIANT_UMMA_LUT):
    try:
        m.set_variant(v)
    except ValueError as e:
        print("skip", _native.VARIANT_NAMES[v], e)
        continue
    for _ in range(3):
        m.forward_dev(idx.data_ptr(), n, out.data_ptr(), s)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    t0.record()
    for _ in range(reps):
        m.forward_dev(idx.data_ptr(), n, out.data_ptr(), s)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    res[v] = out.cpu().numpy().copy()
    print(f"L={L} n={n} {_native.VARIANT_NAMES[v]:>20s}: {ms:8.3f} ms  {n / ms * 1e3:.4g} seq/s", flush=True)
if len(res) == 2:
    a, b = res[_native.VARIANT_UMMA], res[_native.VARIANT_UMMA_LUT]
    print("max |diff| / scale between the two kernels:", float(np.abs(a - b).max() / np.abs(a).max()))
# table rebuild cost: set_weights + first large forward
m.set_variant(_native.VARIANT_UMMA_LUT)
for _ in range(2):
    m.set_weights(ws)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    m.forward_dev(idx.data_ptr(), 1024, out.data_ptr(), s)
    t1.record()
    torch.cuda.synchronize()
    print(f"operand blob + table rebuild + 1024-sequence forward: {t0.elapsed_time(t1):.3f} ms")
