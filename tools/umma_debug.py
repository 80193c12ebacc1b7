"""GPU debug helper: compare the UMMA (tcgen05) CNN kernel with the FFMA kernel and the oracle.
Prints errors instead of asserting so one gpurun call gives the whole picture."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexs_b200 import _native  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import flexs_oracle as fo  # noqa: E402


def fwd(m, idx):
    d = torch.from_numpy(idx).cuda()
    out = torch.empty(len(idx), dtype=torch.float32, device="cuda")
    m.forward_dev(d.data_ptr(), len(idx), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def main():
    shapes = [(100, 4, 1000), (8, 4, 3000), (14, 4, 999), (90, 20, 300), (237, 20, 50), (735, 20, 9), (5, 4, 100)]
    if len(sys.argv) > 1:
        shapes = shapes[: int(sys.argv[1])]
    for L, A, n in shapes:
        for wname, fn in (("glorot", fo.glorot_weights), ("trained", fo.trained_like_weights)):
            ws = fn(fo.CNNShape(L, A, 32, 100, 5).weight_shapes(), 3)
            idx = np.random.default_rng(0).integers(0, A, size=(n, L), dtype=np.uint8)
            ref = co.cnn_forward(idx, [ws])
            m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
            m.set_weights(ws)
            scale = np.abs(ref).max()
            res = {}
            for name, v in (("tiled", 2), ("umma", 3)):
                try:
                    m.set_variant(v)
                    y = fwd(m, idx)
                    res[name] = np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 0.1 * scale))
                    if name == "umma" and res[name] > 1e-4:
                        bad = np.argsort(-np.abs(y - ref))[:5]
                        print("   worst:", bad, y[bad], ref[bad], "nan:", np.isnan(y).sum())
                except Exception as e:  # noqa: BLE001
                    res[name] = repr(e)
            print(f"L={L} A={A} n={n} {wname}: scale={scale:.3g} " + " ".join(f"{k}={v}" for k, v in res.items()), flush=True)
            m.close()
    # throughput
    L, A, n = 100, 4, 1 << 20
    ws = fo.glorot_weights(fo.CNNShape(L, A, 32, 100, 5).weight_shapes(), 0)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(ws)
    d = torch.randint(0, A, (n, L), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    for name, v in (("tiled", 2), ("umma", 3)):
        m.set_variant(v)
        for _ in range(2):
            m.forward_dev(d.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            m.forward_dev(d.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"throughput {name}: {n / dt:.3e} seq/s ({dt * 1e3:.2f} ms per {n})", flush=True)


if __name__ == "__main__":
    main()
