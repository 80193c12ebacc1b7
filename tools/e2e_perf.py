"""End-to-end throughput of flexs_model_score_host (pinned host characters in, host scores out), wall clock."""
import sys, time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from flexs_b200 import _native
from _weights import cnn_shapes, trained_like

L, n = 100, 1 << 22
m = _native.NativeModel("cnn", seq_len=L, alphabet_size=4, num_filters=32, hidden_size=100, kernel_size=5)
m.set_weights(trained_like(cnn_shapes(L, 4), 5))
idx = torch.randint(0, 4, (n, L), dtype=torch.uint8)
chars = torch.tensor(list(b"TGCA"), dtype=torch.uint8)[idx.long()].contiguous().pin_memory()
out = torch.empty(n, dtype=torch.float32).pin_memory()
c, o = chars.numpy(), out.numpy()
m.score_host(c, "TGCA", o)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); m.score_host(c, "TGCA", o); ts.append(time.perf_counter() - t0)
print(f"e2e: best {n / min(ts):.4g} seq/s, median {n / float(np.median(ts)):.4g} seq/s ({min(ts) * 1e3:.2f} ms best)")
