"""Small forward passes of cnn_k9 for compute-sanitizer (memcheck / racecheck / synccheck), each compared with the
first-generation tcgen05 kernel on the same inputs."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from flexs_b200 import _native
from _weights import cnn_shapes, trained_like

for L, n in ((100, 300), (14, 700), (21, 259)):
    ws = trained_like(cnn_shapes(L, 4), L)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=4, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(ws)
    m.set_variant(_native.VARIANT_UMMA_LUT)
    idx = np.random.default_rng(L).integers(0, 4, size=(n, L), dtype=np.uint8)
    d = torch.from_numpy(idx).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    m.forward_dev(d.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    m.set_variant(_native.VARIANT_UMMA)
    m.forward_dev(d.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = out.cpu().numpy()
    print(L, n, "max |k9 - umma2| / scale", float(np.abs(got - ref).max() / np.abs(ref).max()), flush=True)
    m.close()
