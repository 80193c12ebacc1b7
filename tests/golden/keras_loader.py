"""Reader of the files tools/export_keras_golden.py writes (weights in Keras ``get_weights()`` order, residue indices and
``model.predict`` outputs).  Shared by the GPU parity test (tests/test_gpu_parity.py::test_real_keras_golden_if_present)
and the CPU format test (tests/test_oracle_golden.py::test_keras_golden_format_roundtrip)."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def committed_files():
    return sorted(glob.glob(os.path.join(HERE, "keras_*.npz")))


def load(path):
    g = np.load(path)
    nw = len([k for k in g.files if k[0] == "w" and k[1:].isdigit()])
    ws = [g[f"w{i}"] for i in range(nw)]
    kind = str(g["kind"])
    idx, y = g["idx"], g["y"]
    if kind == "cnn":
        assert nw == 12 and ws[0].ndim == 3
        cfg = dict(seq_len=int(idx.shape[1]), alphabet_size=int(ws[0].shape[1]), num_filters=int(ws[0].shape[2]),
                   hidden_size=int(ws[6].shape[1]), kernel_size=int(g["kernel_size"]))
        assert ws[0].shape[0] == cfg["kernel_size"] and ws[4].shape[0] == cfg["alphabet_size"] - 1
    else:
        assert kind == "mlp" and nw == 8
        cfg = dict(seq_len=int(idx.shape[1]), alphabet_size=int(ws[0].shape[0] // idx.shape[1]), hidden_size=int(ws[0].shape[1]))
    assert idx.dtype == np.uint8 and y.shape == (idx.shape[0],)
    return dict(kind=kind, cfg=cfg, weights=ws, idx=idx, y=np.asarray(y, dtype=np.float32), backend=str(g["backend"]) if "backend" in g.files else "tensorflow")
