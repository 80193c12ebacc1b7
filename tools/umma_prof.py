"""Device-resident forward throughput of the A = 20 shapes (cnn_umma.cu); FLEXS_UMMA_PROF=1 adds its phase counters."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from flexs_b200 import _native
from _weights import cnn_shapes, trained_like

for L, n in ((237, 1 << 17), (90, 1 << 18), (735, 1 << 15)):
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=20, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(trained_like(cnn_shapes(L, 20), 5))
    idx = torch.randint(0, 20, (n, L), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        m.forward_dev(idx.data_ptr(), n, out.data_ptr(), s)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        m.forward_dev(idx.data_ptr(), n, out.data_ptr(), s)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    print(f"L={L} A=20 n={n}: {ms:8.3f} ms  {n / ms * 1e3:.4g} seq/s", flush=True)
    m.close()
