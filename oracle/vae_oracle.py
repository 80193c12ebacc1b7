"""CPU restatement of the CbAS / DbAS generator (flexs/utils/VAE_utils.py) in float64.  TEST INFRASTRUCTURE ONLY: only
tests/ may import it; it checks the CUDA kernels of flexs_b200/csrc/vae.cu and is never on a product path.

PARITY STATUS: **parity unpinned**.  The arithmetic of the reference lives in TensorFlow/Keras (absent here, and the
reference's tests assert no values for the VAE), so this follows the reference's layer definitions line by line and states
the two semantics the reference leaves to Keras explicitly:

  * VAEModel.__init__ (VAE_utils.py:40-63): encoder Dense(I, elu) -> Dropout(0.3) -> Dense(I, elu) ->
    BatchNormalization() -> Dense(I, elu) -> z_mean / z_log_var (Dense(Z)) -> Sampling (:12-25:
    z = mean + exp(0.5 * log_var) * eps); decoder Dense(I, elu) x2 -> Dropout(0.3) -> Dense(I, elu) -> Dense(D, sigmoid).
  * train_step (:75-92): reconstruction = D * mean(binary_crossentropy(x, out)) (Keras clips the probabilities to
    [1e-7, 1 - 1e-7]), kl = -0.5 * mean(1 + lv - m^2 - exp(lv)), total = reconstruction + kl.
  * fit(..., sample_weight=w) (:141-151): stated here as  loss = sum_b w_b * (reconstruction_b + kl_b) / B.
  * BatchNormalization (Keras defaults momentum 0.99, epsilon 1e-3): training mode normalises with the batch mean and the
    BIASED batch variance and moves the moving statistics towards them with weight 0.01 (the non-fused path Keras takes
    for a 2-D input).
  * compile (:127): Adam(lr 1e-4, clipvalue 0.5): every gradient element is clipped to [-0.5, 0.5] first; Keras Adam
    defaults beta 0.9 / 0.999, epsilon 1e-7, lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t).
  * calculate_log_probability (:189-217): sum_l log(1e-9 + out[l, x_l] / sum_a out[l, a]), nan_to_num.

Weights: the 22 arrays of include/flexs_b200.h (Keras get_weights() order).
"""
import numpy as np
import torch

NAMES = ["W1", "b1", "W2", "b2", "gamma", "beta", "mov_mean", "mov_var", "W3", "b3", "Wm", "bm", "Wv", "bv",
         "W4", "b4", "W5", "b5", "W6", "b6", "W7", "b7"]
TRAINABLE = [i for i, n in enumerate(NAMES) if n not in ("mov_mean", "mov_var")]


def shapes(seq_len, alphabet_size, intermediate, latent):
    d, i, z = seq_len * alphabet_size, intermediate, latent
    return [(d, i), (i,), (i, i), (i,), (i,), (i,), (i,), (i,), (i, i), (i,), (i, z), (z,), (i, z), (z,),
            (z, i), (i,), (i, i), (i,), (i, i), (i,), (i, d), (d,)]


def init_weights(seq_len, alphabet_size, intermediate, latent, seed):
    """Keras defaults: glorot-uniform kernels, zero biases, BatchNorm gamma 1 / beta 0 / moving mean 0 / variance 1;
    biases and BatchNorm parameters perturbed so that every path carries signal in the parity tests."""
    rng = np.random.default_rng(seed)
    out = []
    for name, shp in zip(NAMES, shapes(seq_len, alphabet_size, intermediate, latent)):
        if len(shp) == 2:
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            out.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
        elif name in ("gamma", "mov_var"):
            out.append((1.0 + 0.2 * rng.random(shp)).astype(np.float32))
        else:
            out.append(rng.normal(0, 0.1, size=shp).astype(np.float32))
    return out


def _one_hot(idx, alphabet_size):
    return torch.from_numpy(np.eye(alphabet_size, dtype=np.float64)[np.asarray(idx)].reshape(len(idx), -1))


def forward(weights, idx, alphabet_size, eps=None, mask1=None, mask2=None, train=False):
    """Returns (out, z_mean, z_log_var, batch_mean, batch_var); torch float64 tensors (weights may require grad)."""
    w = dict(zip(NAMES, weights))
    x = _one_hot(idx, alphabet_size)
    elu = torch.nn.functional.elu
    h1 = elu(x @ w["W1"] + w["b1"])
    if train:
        h1 = h1 * mask1
    h2 = elu(h1 @ w["W2"] + w["b2"])
    if train:
        mean, var = h2.mean(dim=0), h2.var(dim=0, unbiased=False)
    else:
        mean, var = w["mov_mean"], w["mov_var"]
    bn = (h2 - mean) / torch.sqrt(var + 1e-3) * w["gamma"] + w["beta"]
    h3 = elu(bn @ w["W3"] + w["b3"])
    zm, zlv = h3 @ w["Wm"] + w["bm"], h3 @ w["Wv"] + w["bv"]
    z = zm if eps is None else zm + torch.exp(0.5 * zlv) * eps
    g1 = elu(z @ w["W4"] + w["b4"])
    g2 = elu(g1 @ w["W5"] + w["b5"])
    if train:
        g2 = g2 * mask2
    g3 = elu(g2 @ w["W6"] + w["b6"])
    out = torch.sigmoid(g3 @ w["W7"] + w["b7"])
    return out, zm, zlv, mean, var


def loss_and_grads(weights, idx, alphabet_size, sample_weights, eps, mask1, mask2):
    """One training batch: (loss, gradients in the 22-array layout (zeros for the moving statistics), new moving mean,
    new moving variance), all float64 numpy."""
    tw = [torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=(i in TRAINABLE)) for i, a in enumerate(weights)]
    eps_t, m1, m2 = (torch.from_numpy(np.asarray(a, dtype=np.float64)) for a in (eps, mask1, mask2))
    out, zm, zlv, mean, var = forward(tw, idx, alphabet_size, eps_t, m1, m2, train=True)
    x = _one_hot(idx, alphabet_size)
    oc = out.clamp(1e-7, 1 - 1e-7)
    recon = -(x * torch.log(oc) + (1 - x) * torch.log(1 - oc)).sum(dim=1)          # = D * mean over features
    kl = -0.5 * (1 + zlv - zm ** 2 - torch.exp(zlv)).mean(dim=1)
    sw = torch.from_numpy(np.asarray(sample_weights, dtype=np.float64))
    loss = (sw * (recon + kl)).sum() / len(idx)
    loss.backward()
    grads = [t.grad.numpy().copy() if t.grad is not None else np.zeros(t.shape) for t in tw]
    new_mean = 0.99 * np.asarray(weights[6], dtype=np.float64) + 0.01 * mean.detach().numpy()
    new_var = 0.99 * np.asarray(weights[7], dtype=np.float64) + 0.01 * var.detach().numpy()
    return float(loss.detach()), grads, new_mean, new_var


def adam_clip_update(weights, grads, m, v, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-7, clip=0.5):
    lr_t = lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    new_w, new_m, new_v = [], [], []
    for i, (w, g, mi, vi) in enumerate(zip(weights, grads, m, v)):
        w = np.asarray(w, dtype=np.float64)
        if i not in TRAINABLE:
            new_w.append(w); new_m.append(mi); new_v.append(vi)
            continue
        g = np.clip(g, -clip, clip)
        mi = b1 * mi + (1 - b1) * g
        vi = b2 * vi + (1 - b2) * g * g
        new_w.append(w - lr_t * mi / (np.sqrt(vi) + eps)); new_m.append(mi); new_v.append(vi)
    return new_w, new_m, new_v


def decode(weights, z):
    w = dict(zip(NAMES, [torch.from_numpy(np.asarray(a, dtype=np.float64)) for a in weights]))
    elu = torch.nn.functional.elu
    z = torch.from_numpy(np.asarray(z, dtype=np.float64))
    g = elu(elu(elu(z @ w["W4"] + w["b4"]) @ w["W5"] + w["b5"]) @ w["W6"] + w["b6"])
    return torch.sigmoid(g @ w["W7"] + w["b7"]).numpy()


def log_probability(weights, idx, alphabet_size, eps=None):
    tw = [torch.from_numpy(np.asarray(a, dtype=np.float64)) for a in weights]
    e = None if eps is None else torch.from_numpy(np.asarray(eps, dtype=np.float64))
    out = forward(tw, idx, alphabet_size, e)[0].numpy().reshape(len(idx), -1, alphabet_size)
    sel = np.take_along_axis(out, np.asarray(idx)[..., None].astype(np.int64), axis=2)[..., 0]
    return np.nan_to_num(np.log(1e-9 + sel / out.sum(axis=2)).sum(axis=1))
