"""Run the CNN forward a few times (for ncu / phase-counter runs): python tools/fwd_once.py VARIANT N [L A reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexs_b200 import _native  # noqa: E402

variant = {"tiled": 2, "umma": 3, "auto": 0, "simple": 1}[sys.argv[1]]
n = int(sys.argv[2])
L = int(sys.argv[3]) if len(sys.argv) > 3 else 100
A = int(sys.argv[4]) if len(sys.argv) > 4 else 4
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rng = np.random.default_rng(0)
shapes = [(5, A, 32), (32,), (5, 32, 32), (32,), (A - 1, 32, 32), (32,), (32, 100), (100,), (100, 100), (100,), (100, 1), (1,)]
ws = [rng.uniform(-0.1, 0.1, size=s).astype(np.float32) for s in shapes]
m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
m.set_weights(ws)
m.set_variant(variant)
d = torch.randint(0, A, (n, L), dtype=torch.uint8, device="cuda")
out = torch.empty(n, dtype=torch.float32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    m.forward_dev(d.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print(f"{sys.argv[1]} L={L} A={A} n={n}: ms per call {['%.3f' % x for x in ms]} -> {n / (min(ms) / 1e3):.3e} seq/s", flush=True)
