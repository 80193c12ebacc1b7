"""Stress of the mbarrier protocol of cnn_k9: many launches of random sizes and lengths (ragged groups, odd tile
counts carried across groups, one- and multi-tile items), each checked against cnn_umma2 on the same inputs."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from flexs_b200 import _native
from _weights import cnn_shapes, trained_like

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
worst = 0.0
for L in (8, 9, 14, 20, 21, 36, 37, 100, 116, 120, 170):
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=4, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(trained_like(cnn_shapes(L, 4), L))
    nmax = 400_000 if L <= 40 else 150_000
    idx = torch.randint(0, 4, (nmax, L), dtype=torch.uint8, device="cuda")
    a = torch.empty(nmax, dtype=torch.float32, device="cuda")
    b = torch.empty(nmax, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for it in range(iters):
        n = int(rng.integers(1, nmax)) if it % 3 else int(rng.integers(1, 3000))
        m.set_variant(_native.VARIANT_UMMA_LUT)
        m.forward_dev(idx.data_ptr(), n, a.data_ptr(), s)
        m.set_variant(_native.VARIANT_UMMA)
        m.forward_dev(idx.data_ptr(), n, b.data_ptr(), s)
        torch.cuda.synchronize()
        err = float((a[:n] - b[:n]).abs().max() / b[:n].abs().max())
        worst = max(worst, err)
        assert err < 1e-4, (L, n, err)
    m.close()
    print(f"L={L}: {iters} launches ok", flush=True)
print("worst |k9 - umma2| / scale:", worst)
