"""Device-resident forward throughput of the A = 4 configurations, AUTO kernel choice vs the forced tcgen05 kernel."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

from flexs_b200 import _native
from _weights import cnn_shapes, trained_like

for L, M, n in ((8, 1, 1 << 20), (14, 3, 1 << 20), (100, 1, 1 << 22)):
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=4, num_filters=32, hidden_size=100, kernel_size=5, n_members=M)
    for i in range(M):
        m.set_weights(trained_like(cnn_shapes(L, 4), 5 + i), i)
    idx = torch.randint(0, 4, (n, L), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for v in (_native.VARIANT_UMMA, _native.VARIANT_AUTO):
        m.set_variant(v)
        for _ in range(3):
            m.forward_dev(idx.data_ptr(), n, out.data_ptr(), s)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            m.forward_dev(idx.data_ptr(), n, out.data_ptr(), s)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        print(f"L={L} M={M} n={n} {_native.VARIANT_NAMES[m.active_variant(n)]:>20s}: {ms:8.3f} ms  {n / ms * 1e3:.4g} seq/s", flush=True)
    m.close()
