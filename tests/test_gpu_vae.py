"""GPU: K9, the CbAS / DbAS generator (flexs_b200/csrc/vae.cu) against the float64 restatement in oracle/vae_oracle.py:
one optimiser step from identical weights, batch, dropout masks and latent noise (loss, every gradient, the Adam update with
clipvalue, the BatchNorm moving statistics), the decoder pass, the reconstruction log-probability, and a fit that learns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from flexs_b200 import _native  # noqa: E402
from flexs_b200.utils import sequence_utils as su  # noqa: E402
from oracle import vae_oracle as vo  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device; the CUDA path has no CPU fallback")


def _cuda(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


@pytest.mark.parametrize("L,A,I,Z,B", [(14, 4, 50, 2, 10), (8, 4, 250, 2, 10), (30, 20, 64, 3, 7), (5, 4, 16, 2, 2)])
def test_vae_train_step_matches_oracle(L, A, I, Z, B):
    rng = np.random.default_rng(L * 100 + I)
    ws = vo.init_weights(L, A, I, Z, seed=3)
    idx = rng.integers(0, A, size=(B, L), dtype=np.uint8)
    sw = rng.uniform(0.0, 2.0, size=B)
    sw[0] = 0.0                                            # CbAS zeroes the weight of proposals below the threshold
    eps = rng.normal(size=(B, Z))
    m1 = (rng.random((B, I)) >= 0.3) / 0.7
    m2 = (rng.random((B, I)) >= 0.3) / 0.7
    loss, grads, new_mean, new_var = vo.loss_and_grads(ws, idx, A, sw, eps, m1, m2)
    vae = _native.NativeVAE(L, A, I, Z)
    vae.set_weights(ws)
    dev = [_cuda(idx, np.uint8), _cuda(sw, np.float32), _cuda(m1, np.float32), _cuda(m2, np.float32), _cuda(eps, np.float32)]
    got_loss = vae.train_step_dev(dev[0].data_ptr(), dev[1].data_ptr(), B, dev[2].data_ptr(), dev[3].data_ptr(), dev[4].data_ptr())
    assert abs(got_loss - loss) <= 2e-5 * abs(loss)
    got_grads = vae.get_gradients()
    for name, g, ref in zip(vo.NAMES, got_grads, grads):
        if name in ("mov_mean", "mov_var"):
            continue
        scale = max(float(np.abs(ref).max()), 1e-6)
        assert np.abs(g.astype(np.float64) - ref).max() <= 2e-4 * scale, name
    # the Adam step with clipvalue 0.5 (VAE_utils.py:127) and the moving statistics
    zeros = [np.zeros_like(np.asarray(w, dtype=np.float64)) for w in ws]
    want_w, _, _ = vo.adam_clip_update(ws, grads, zeros, [z.copy() for z in zeros], step=1)
    got_w = vae.get_weights()
    for i, name in enumerate(vo.NAMES):
        ref = want_w[i] if name not in ("mov_mean", "mov_var") else (new_mean if name == "mov_mean" else new_var)
        # the first Adam step moves every weight by ~lr * sign(g): compare the step itself where the gradient is not ~0
        if name in ("mov_mean", "mov_var"):
            np.testing.assert_allclose(got_w[i], ref, rtol=1e-5, atol=1e-6, err_msg=name)
        else:
            step_ref = ref - np.asarray(ws[i], dtype=np.float64)
            step_got = got_w[i].astype(np.float64) - np.asarray(ws[i], dtype=np.float64)
            big = np.abs(np.clip(grads[i], -0.5, 0.5)) > 1e-4
            np.testing.assert_allclose(step_got[big], step_ref[big], rtol=2e-2, atol=2e-7, err_msg=name)
    vae.close()


def test_vae_decode_and_log_probability_match_oracle():
    L, A, I, Z, n = 14, 4, 50, 2, 333
    rng = np.random.default_rng(0)
    ws = vo.init_weights(L, A, I, Z, seed=5)
    vae = _native.NativeVAE(L, A, I, Z)
    vae.set_weights(ws)
    z = rng.normal(size=(17, Z))
    out = torch.empty((17, L * A), dtype=torch.float32, device="cuda")
    d_z = _cuda(z, np.float32)
    vae.decode_dev(d_z.data_ptr(), 17, out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), vo.decode(ws, z), rtol=0, atol=2e-6)
    idx = rng.integers(0, A, size=(n, L), dtype=np.uint8)
    eps = rng.normal(size=(n, Z))
    d_idx, d_eps = _cuda(idx, np.uint8), _cuda(eps, np.float32)
    for e in (None, eps):
        lp = torch.empty(n, dtype=torch.float64, device="cuda")
        vae.log_prob_dev(d_idx.data_ptr(), n, d_eps.data_ptr() if e is not None else 0, lp.data_ptr())
        torch.cuda.synchronize()
        np.testing.assert_allclose(lp.cpu().numpy(), vo.log_probability(ws, idx, A, e), rtol=2e-5, atol=1e-4)
    vae.close()


def test_vae_fit_learns_a_motif_and_generator_interface():
    """flexs_b200.utils.VAE_utils.VAE on the native kernels: fitting on sequences that share a motif lowers the loss and
    raises their reconstruction log-probability over random sequences; generate() returns new, distinct sequences;
    get/set_weights round-trips (CbAS clones the prior that way, cbas_dbas.py:130-144)."""
    import random

    from flexs_b200.utils.VAE_utils import VAE

    random.seed(0); np.random.seed(0)
    L, alphabet = 14, su.RNAA
    rng = np.random.default_rng(1)
    motif = "GCUAGCUAGCUAGC"
    train = []
    for _ in range(300):
        s = list(motif)
        for p in rng.integers(0, L, size=2):
            s[p] = alphabet[rng.integers(0, 4)]
        train.append("".join(s))
    vae = VAE(seq_length=L, alphabet=alphabet, batch_size=10, latent_dim=2, intermediate_dim=50, epochs=30, verbose=False, seed=0)
    assert vae.native
    before = vae.calculate_log_probability(train[:100]).mean()
    vae.train_model(train, np.ones(len(train)))
    losses = vae.last_fit_losses
    assert len(losses) >= 4 and losses[-1] < 0.8 * losses[0] and np.isfinite(losses).all()
    after = vae.calculate_log_probability(train[:100]).mean()
    randoms = su.generate_random_sequences(L, 100, alphabet)
    assert after > before + 1.0 and after > vae.calculate_log_probability(randoms).mean() + 1.0
    proposals = vae.generate(50, train, np.ones(len(train)))
    assert len(proposals) == 50 == len(set(proposals)) and not set(proposals) & set(train)
    assert all(len(p) == L and set(p) <= set(alphabet) for p in proposals)
    twin = VAE(seq_length=L, alphabet=alphabet, batch_size=10, latent_dim=2, intermediate_dim=50, epochs=1, verbose=False)
    twin.vae.set_weights(vae.vae.get_weights())
    np.random.seed(5); a = vae.calculate_log_probability(train[:20])
    np.random.seed(5); b = twin.calculate_log_probability(train[:20])
    np.testing.assert_array_equal(a, b)
    # zero weights freeze the model: nothing to learn from
    w0 = vae.vae.get_weights()
    vae.train_model(train[:40], np.zeros(40))
    for x, y in zip(w0, vae.vae.get_weights()):
        if x.ndim == 2:
            assert np.abs(x - y).max() < 5e-3   # Adam's epsilon-sized drift only
