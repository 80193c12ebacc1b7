"""Small forward passes of cnn_k9 for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from flexs_b200 import _native
from oracle import c_oracle as co
from oracle import flexs_oracle as fo

for L, n in ((100, 300), (14, 700), (21, 259)):
    ws = fo.trained_like_weights(fo.CNNShape(L, 4, 32, 100, 5).weight_shapes(), L)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=4, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(ws)
    m.set_variant(_native.VARIANT_UMMA_LUT)
    idx = np.random.default_rng(L).integers(0, 4, size=(n, L), dtype=np.uint8)
    d = torch.from_numpy(idx).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    m.forward_dev(d.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = co.cnn_forward(idx, [ws])
    print(L, n, "max err / scale", float(np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()), flush=True)
    m.close()
