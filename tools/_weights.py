"""Synthetic weight sets for the measurement scripts under tools/ (Keras get_weights() order and layout).

Deliberately independent of ``oracle/``: the oracle is test infrastructure, and only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may use it.  These scripts measure or stress the
kernels and, where they check results, compare one kernel with another.
"""
import numpy as np


def cnn_shapes(L, A, F=32, H=100, K=5):
    return [(K, A, F), (F,), (K, F, F), (F,), (A - 1, F, F), (F,), (F, H), (H,), (H, H), (H,), (H, 1), (1,)]


def trained_like(shapes, seed):
    """Glorot-uniform kernels scaled up by 1.7 and small non-zero biases of both signs: ReLU clipping and the max path
    are exercised the way a trained surrogate exercises them."""
    rng = np.random.default_rng(seed)
    out = []
    for shp in shapes:
        if len(shp) == 1:
            out.append(rng.normal(0.0, 0.15, size=shp).astype(np.float32))
        else:
            rec = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            lim = 1.7 * np.sqrt(6.0 / (rec * shp[-2] + rec * shp[-1]))
            out.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
    return out
