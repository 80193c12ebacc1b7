"""CPU oracle for the FLEXS virtual-screen hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU and in plain numpy, what the reference computes on the
path ``Explorer.propose_sequences -> Model.get_fitness -> Keras predict``.  It exists so
that the CUDA path in ``flexs_b200/`` can be *checked*; it is never the thing shipped or
measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package never imports ``oracle``.

PARITY STATUS: **parity unpinned** for the floating-point part.  The arithmetic of the
reference path lives in TensorFlow/Keras (``tensorflow>=2``, setup.py:29; docs pin
``tensorflow==2.3.1``, docs/requirements.txt:99), which is not installed here and is not
under /root/reference, and the reference's tests assert no values for it
(tests/test_models.py:55-77).  The layer semantics below are the published Keras ones
(channels-last cross-correlation, ``valid`` / ``same`` padding rules, Dense ``x@W+b``),
anchored on the reference's own call sites.  The *integer* parts (encode, decode, top-k
slices, cost accounting, mutation) ARE pinned: tests/golden/ holds vectors produced by
importing the reference's pure-Python modules in the authoring container
(tests/golden/make_golden.py).

Reference call sites restated here (paths relative to /root/reference):
  * flexs/utils/sequence_utils.py:32-47   string_to_one_hot   -> encode / one_hot
  * flexs/utils/sequence_utils.py:50-66   one_hot_to_string   -> decode_argmax
  * flexs/baselines/models/keras_model.py:69-79  _fitness_function (squeeze + nan_to_num)
  * flexs/baselines/models/cnn.py:23-54   CNN layer stack      -> cnn_forward
  * flexs/baselines/models/mlp.py:21-31   MLP layer stack      -> mlp_forward
  * flexs/ensemble.py:54-59, :24          Ensemble mean        -> ensemble_mean
  * flexs/baselines/models/keras_model.py:49-67 + cnn.py:56    -> train_step (Adam/MSE)
  * flexs/baselines/explorers/adalead.py:171-175 (and cbas/cmaes) -> top_slice_bm1
  * flexs/baselines/explorers/dyna_ppo.py:315-319               -> top_slice_b
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# Alphabets, flexs/utils/sequence_utils.py:7-17
AAS = "ILVAGMFYWEDQNHCRKSTP"
RNAA = "UGCA"
DNAA = "TGCA"
BA = "01"


# --------------------------------------------------------------------------------------
# integer pieces
# --------------------------------------------------------------------------------------
def encode(sequences: Sequence[str], alphabet: str) -> np.ndarray:
    """Residue indices ``uint8[N, L]``; index = ``alphabet.index(ch)``.

    Follows sequence_utils.py:44-47 (the position of the 1 in each one-hot row).  Like
    ``str.index`` it raises ``ValueError`` for a character that is not in the alphabet.
    """
    n = len(sequences)
    length = len(sequences[0]) if n else 0
    out = np.zeros((n, length), dtype=np.uint8)
    for i, seq in enumerate(sequences):
        if len(seq) != length:
            raise ValueError("ragged sequences")
        for j, ch in enumerate(seq):
            out[i, j] = alphabet.index(ch)
    return out


def one_hot(idx: np.ndarray, num_classes: int, dtype=np.float64) -> np.ndarray:
    """``(N, L) uint8 -> (N, L, A)``; sequence_utils.py:44-47 stacked (keras_model.py:70-75)."""
    out = np.zeros(idx.shape + (num_classes,), dtype=dtype)
    np.put_along_axis(out, idx[..., None].astype(np.int64), 1, axis=-1)
    return out


def decode_argmax(x: np.ndarray, alphabet: str) -> List[str]:
    """``(N, L, A) float -> N strings`` by per-position ``np.argmax`` (first max wins).

    sequence_utils.py:50-66 as used by cmaes.py:61-67 and environments/dyna_ppo.py:144-147.
    """
    idx = np.argmax(x, axis=-1)
    return ["".join(alphabet[i] for i in row) for row in idx]


def top_slice_bm1(preds: np.ndarray, batch: int) -> np.ndarray:
    """``np.argsort(preds)[: -B : -1]`` — the B-1 best, descending (adalead.py:173,
    cbas_dbas.py:199, cmaes.py:120)."""
    return np.argsort(preds)[: -batch: -1]


def top_slice_b(preds: np.ndarray, batch: int) -> np.ndarray:
    """``np.argsort(preds)[::-1][:B]`` — the B best, descending (dyna_ppo.py:317)."""
    return np.argsort(preds)[::-1][:batch]


# --------------------------------------------------------------------------------------
# weights in Keras ``get_weights()`` order and layout
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class CNNShape:
    """Hyper-parameters of cnn.py:10-21 (kernel_size3 = len(alphabet) - 1, cnn.py:43)."""

    seq_len: int
    alphabet_size: int
    num_filters: int
    hidden_size: int
    kernel_size: int = 5

    @property
    def kernel_size3(self) -> int:
        return self.alphabet_size - 1

    @property
    def conv_len(self) -> int:  # T after the ``valid`` conv
        return self.seq_len - self.kernel_size + 1

    def weight_shapes(self) -> List[Tuple[int, ...]]:
        k, a, f, h, k3 = (self.kernel_size, self.alphabet_size, self.num_filters,
                          self.hidden_size, self.kernel_size3)
        return [(k, a, f), (f,), (k, f, f), (f,), (k3, f, f), (f,),
                (f, h), (h,), (h, h), (h,), (h, 1), (1,)]

    def flop_alg(self) -> int:
        """conv1 as gather-add, everything else as 2 flop per MAC."""
        t, f, k, k3, h = (self.conv_len, self.num_filters, self.kernel_size,
                          self.kernel_size3, self.hidden_size)
        return t * f * k + 2 * (t * f * k * f + t * f * k3 * f + f * h + h * h + h)


@dataclasses.dataclass
class MLPShape:
    """Hyper-parameters of mlp.py:10-19."""

    seq_len: int
    alphabet_size: int
    hidden_size: int

    def weight_shapes(self) -> List[Tuple[int, ...]]:
        d, h = self.seq_len * self.alphabet_size, self.hidden_size
        return [(d, h), (h,), (h, h), (h,), (h, h), (h,), (h, 1), (1,)]

    def flop_alg(self) -> int:
        """Layer 1 as L row gather-adds of H, the rest 2 flop per MAC."""
        l, h = self.seq_len, self.hidden_size
        return l * h + 2 * (h * h + h * h + h)


def _glorot_limit(shape: Tuple[int, ...]) -> float:
    """Keras glorot_uniform: limit = sqrt(6 / (fan_in + fan_out)); for a conv kernel
    (k, in, out): fan_in = k*in, fan_out = k*out."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        receptive = int(np.prod(shape[:-2]))
        fan_in, fan_out = receptive * shape[-2], receptive * shape[-1]
    return float(np.sqrt(6.0 / (fan_in + fan_out)))


def glorot_weights(shapes: List[Tuple[int, ...]], seed: int) -> List[np.ndarray]:
    """Keras default init (glorot-uniform kernels, zero biases) from ``default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    out = []
    for shp in shapes:
        if len(shp) == 1:
            out.append(np.zeros(shp, dtype=np.float32))
        else:
            lim = _glorot_limit(shp)
            out.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
    return out


def trained_like_weights(shapes: List[Tuple[int, ...]], seed: int) -> List[np.ndarray]:
    """A weight set that exercises ReLU clipping and the max path: glorot kernels scaled
    up, non-zero biases of both signs."""
    rng = np.random.default_rng(seed)
    out = []
    for shp in shapes:
        if len(shp) == 1:
            out.append(rng.normal(0.0, 0.15, size=shp).astype(np.float32))
        else:
            lim = 1.7 * _glorot_limit(shp)
            out.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
    return out


# --------------------------------------------------------------------------------------
# Keras layer semantics (TF-Keras 2.x), restated
# --------------------------------------------------------------------------------------
def _windows(x: np.ndarray, k: int) -> np.ndarray:
    """``(N, T, C) -> (N, T-k+1, k, C)`` sliding windows along time."""
    w = np.lib.stride_tricks.sliding_window_view(x, k, axis=1)  # (N, T-k+1, C, k)
    return np.moveaxis(w, -1, 2)


def conv1d(x: np.ndarray, w: np.ndarray, b: np.ndarray, padding: str) -> np.ndarray:
    """Keras ``Conv1D(strides=1)``: channels-last cross-correlation, kernel ``(k, in, out)``.

    ``valid``: T_out = T - k + 1.  ``same``: zero pad k-1 in total, ``(k-1)//2`` on the
    left and the remainder on the right (TF's SAME rule for stride 1).
    """
    k = w.shape[0]
    if padding == "same":
        left = (k - 1) // 2
        right = (k - 1) - left
        x = np.pad(x, ((0, 0), (left, right), (0, 0)))
    elif padding != "valid":
        raise ValueError(padding)
    win = _windows(x, k)
    return np.einsum("ntkc,kcf->ntf", win, w, optimize=True) + b


def relu(x: np.ndarray) -> np.ndarray:
    return np.maximum(x, 0)


def cnn_features(onehot: np.ndarray, weights: Sequence[np.ndarray]) -> np.ndarray:
    """cnn.py:25-48: three convs (+ the identity MaxPooling1D(1)) and GlobalMaxPooling1D."""
    w1, b1, w2, b2, w3, b3 = weights[:6]
    h = relu(conv1d(onehot, w1, b1, "valid"))          # cnn.py:25-32
    h = relu(conv1d(h, w2, b2, "same"))                # cnn.py:33-39 ; :40 is identity
    h = relu(conv1d(h, w3, b3, "same"))                # cnn.py:41-47
    return h.max(axis=1)                               # cnn.py:48


def cnn_forward(idx: np.ndarray, weights: Sequence[np.ndarray], dtype=np.float64) -> np.ndarray:
    """Whole CNN forward at inference (Dropout is the identity), output ``(N,)``.

    ``dtype=np.float64`` is the definition; ``np.float32`` is the "as the reference would
    compute it" precision.  keras_model.py:77-79 squeezes axis 1 and applies nan_to_num.
    """
    ws = [np.asarray(w, dtype=dtype) for w in weights]
    a = ws[0].shape[1]
    p = cnn_features(one_hot(idx, a, dtype), ws)
    d = relu(p @ ws[6] + ws[7])                        # cnn.py:49
    d = relu(d @ ws[8] + ws[9])                        # cnn.py:50 ; :51 Dropout no-op
    y = d @ ws[10] + ws[11]                            # cnn.py:52
    return y[:, 0]


def mlp_forward(idx: np.ndarray, weights: Sequence[np.ndarray], dtype=np.float64) -> np.ndarray:
    """mlp.py:21-31: Flatten (row-major ``l*A + c``) -> Dense(H,relu) x3 -> Dense(1)."""
    ws = [np.asarray(w, dtype=dtype) for w in weights]
    n, length = idx.shape
    a = ws[0].shape[0] // length
    x = one_hot(idx, a, dtype).reshape(n, length * a)
    h = relu(x @ ws[0] + ws[1])
    h = relu(h @ ws[2] + ws[3])
    h = relu(h @ ws[4] + ws[5])
    return (h @ ws[6] + ws[7])[:, 0]


def nan_to_num_f32(y: np.ndarray) -> np.ndarray:
    """keras_model.py:77: ``np.nan_to_num`` on the float32 prediction."""
    return np.nan_to_num(np.asarray(y, dtype=np.float32))


def ensemble_mean(member_scores: Sequence[np.ndarray]) -> np.ndarray:
    """ensemble.py:54-59 with the default ``combine_with`` (:24): stack on axis 1, mean."""
    scores = np.stack([np.asarray(s) for s in member_scores], axis=1)
    return np.mean(scores, axis=1)


# --------------------------------------------------------------------------------------
# training step (keras_model.py:61-67, compile at cnn.py:56 / mlp.py:33)
# --------------------------------------------------------------------------------------
ADAM_LR, ADAM_B1, ADAM_B2, ADAM_EPS = 1e-3, 0.9, 0.999, 1e-7  # Keras "adam" defaults


def _conv1d_backward(x, w, padding, gout):
    """Gradients of ``conv1d`` w.r.t. x, w, b."""
    k = w.shape[0]
    if padding == "same":
        left = (k - 1) // 2
        right = (k - 1) - left
    else:
        left = right = 0
    xp = np.pad(x, ((0, 0), (left, right), (0, 0)))
    win = _windows(xp, k)                                   # (N, To, k, C)
    gw = np.einsum("ntkc,ntf->kcf", win, gout, optimize=True)
    gb = gout.sum(axis=(0, 1))
    gxp = np.zeros_like(xp)
    t_out = gout.shape[1]
    for j in range(k):
        gxp[:, j:j + t_out, :] += np.einsum("ntf,cf->ntc", gout, w[j], optimize=True)
    gx = gxp[:, left:xp.shape[1] - right, :]
    return gx, gw, gb


def cnn_loss_and_grads(idx, labels, weights, dropout_mask=None, dtype=np.float64):
    """MSE loss (mean over the batch) and its gradient for every CNN weight.

    ``dropout_mask``: ``(N, H)`` of {0,1} applied after the second Dense with the Keras
    inverted-dropout scale 1/(1-0.25) (cnn.py:51); ``None`` = no dropout.
    Returns ``(loss, grads, predictions)``.
    """
    ws = [np.asarray(w, dtype=dtype) for w in weights]
    y = np.asarray(labels, dtype=dtype)
    a = ws[0].shape[1]
    x0 = one_hot(idx, a, dtype)
    z1 = conv1d(x0, ws[0], ws[1], "valid"); h1 = relu(z1)
    z2 = conv1d(h1, ws[2], ws[3], "same"); h2 = relu(z2)
    z3 = conv1d(h2, ws[4], ws[5], "same"); h3 = relu(z3)
    am = h3.argmax(axis=1)                                  # (N, F) first max wins
    p = np.take_along_axis(h3, am[:, None, :], axis=1)[:, 0, :]
    u1 = p @ ws[6] + ws[7]; d1 = relu(u1)
    u2 = d1 @ ws[8] + ws[9]; d2 = relu(u2)
    if dropout_mask is not None:
        keep = np.asarray(dropout_mask, dtype=dtype) / 0.75
        d2d = d2 * keep
    else:
        keep = None
        d2d = d2
    out = (d2d @ ws[10] + ws[11])[:, 0]
    n = len(y)
    loss = float(np.mean((out - y) ** 2))
    gout = (2.0 / n) * (out - y)                            # (N,)
    g = [None] * 12
    g[10] = d2d.T @ gout[:, None]; g[11] = np.array([gout.sum()], dtype=dtype)
    gd2 = gout[:, None] * ws[10][:, 0][None, :]
    if keep is not None:
        gd2 = gd2 * keep
    gu2 = gd2 * (u2 > 0)
    g[8] = d1.T @ gu2; g[9] = gu2.sum(0)
    gu1 = (gu2 @ ws[8].T) * (u1 > 0)
    g[6] = p.T @ gu1; g[7] = gu1.sum(0)
    gp = gu1 @ ws[6].T                                      # (N, F)
    gh3 = np.zeros_like(h3)
    np.put_along_axis(gh3, am[:, None, :], gp[:, None, :], axis=1)
    gz3 = gh3 * (z3 > 0)
    gh2, g[4], g[5] = _conv1d_backward(h2, ws[4], "same", gz3)
    gz2 = gh2 * (z2 > 0)
    gh1, g[2], g[3] = _conv1d_backward(h1, ws[2], "same", gz2)
    gz1 = gh1 * (z1 > 0)
    _, g[0], g[1] = _conv1d_backward(x0, ws[0], "valid", gz1)
    return loss, g, out


def mlp_loss_and_grads(idx, labels, weights, dtype=np.float64):
    """MSE loss and gradients for the MLP (mlp.py:21-33; no dropout in that stack)."""
    ws = [np.asarray(w, dtype=dtype) for w in weights]
    y = np.asarray(labels, dtype=dtype)
    n, length = idx.shape
    a = ws[0].shape[0] // length
    x = one_hot(idx, a, dtype).reshape(n, length * a)
    u1 = x @ ws[0] + ws[1]; h1 = relu(u1)
    u2 = h1 @ ws[2] + ws[3]; h2 = relu(u2)
    u3 = h2 @ ws[4] + ws[5]; h3 = relu(u3)
    out = (h3 @ ws[6] + ws[7])[:, 0]
    loss = float(np.mean((out - y) ** 2))
    gout = (2.0 / n) * (out - y)
    g = [None] * 8
    g[6] = h3.T @ gout[:, None]; g[7] = np.array([gout.sum()], dtype=dtype)
    gu3 = (gout[:, None] * ws[6][:, 0][None, :]) * (u3 > 0)
    g[4] = h2.T @ gu3; g[5] = gu3.sum(0)
    gu2 = (gu3 @ ws[4].T) * (u2 > 0)
    g[2] = h1.T @ gu2; g[3] = gu2.sum(0)
    gu1 = (gu2 @ ws[2].T) * (u1 > 0)
    g[0] = x.T @ gu1; g[1] = gu1.sum(0)
    return loss, g, out


def adam_update(weights, grads, m, v, step, lr=ADAM_LR, b1=ADAM_B1, b2=ADAM_B2, eps=ADAM_EPS):
    """One Keras-Adam update (TF2 ``Adam._resource_apply_dense``):
    ``lr_t = lr*sqrt(1-b2^t)/(1-b1^t)``; ``w -= lr_t * m / (sqrt(v) + eps)``.  ``step`` is
    1-based.  Updates m, v in place and returns the new weights list."""
    lr_t = lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    new = []
    for i, (w, g) in enumerate(zip(weights, grads)):
        g = np.asarray(g, dtype=np.float64).reshape(np.shape(w))
        m[i] = b1 * m[i] + (1 - b1) * g
        v[i] = b2 * v[i] + (1 - b2) * g * g
        new.append(np.asarray(w, dtype=np.float64) - lr_t * m[i] / (np.sqrt(v[i]) + eps))
    return new


# --------------------------------------------------------------------------------------
# the reference call, end to end (what get_fitness does)
# --------------------------------------------------------------------------------------
def get_fitness_cnn(sequences: Sequence[str], alphabet: str, weights, dtype=np.float32):
    """keras_model.py:69-79 for a CNN: encode -> forward -> squeeze -> nan_to_num."""
    return nan_to_num_f32(cnn_forward(encode(sequences, alphabet), weights, dtype))


def get_fitness_mlp(sequences: Sequence[str], alphabet: str, weights, dtype=np.float32):
    return nan_to_num_f32(mlp_forward(encode(sequences, alphabet), weights, dtype))


def percentile_gamma(scores: np.ndarray, q: float, gamma: float) -> float:
    """cbas_dbas.py:163: ``gamma = max(np.percentile(scores, Q*100), gamma)``."""
    return max(float(np.percentile(scores, q * 100)), gamma)
