"""GPU: K4 training kernels (flexs_model_train_step_dev / flexs_model_fit_dev) against the oracle.

Keras' fit is RNG-dependent (shuffle order, dropout masks, glorot init) and unpinned, so parity is
defined on ONE optimiser step from identical weights / batch / dropout mask — loss, every gradient
(read back through the Adam first moment) and the updated weights — plus end-metric checks for the
whole fit loop.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from flexs_b200 import _native  # noqa: E402
from flexs_b200.utils import sequence_utils as su  # noqa: E402
from oracle import flexs_oracle as fo  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")


def _step(model, idx, y, mask=None, member=0):
    d_idx = torch.from_numpy(idx).cuda()
    d_y = torch.from_numpy(y.astype(np.float32)).cuda()
    d_m = torch.from_numpy(mask.astype(np.float32)).cuda() if mask is not None else None
    return model.train_step_dev(member, d_idx.data_ptr(), d_y.data_ptr(), len(y), d_m.data_ptr() if d_m is not None else 0,
                                torch.cuda.current_stream().cuda_stream)


def _check_grads(m_arrays, grads, shapes):
    for i, (mm, g, shp) in enumerate(zip(m_arrays, grads, shapes)):
        g = np.asarray(g, dtype=np.float64).reshape(-1)
        got = mm.astype(np.float64) / (1.0 - fo.ADAM_B1)   # m_1 = (1 - beta1) * g
        scale = max(np.abs(g).max(), 1e-12)
        assert np.abs(got - g).max() <= 2e-4 * scale, (i, shp, np.abs(got - g).max(), scale)


@pytest.mark.parametrize("L,A,F,H,K,n,use_mask", [(12, 4, 32, 100, 5, 37, True), (8, 4, 32, 100, 5, 64, False),
                                                 (3, 4, 1, 1, 2, 9, True), (30, 20, 32, 100, 5, 11, True),
                                                 (11, 7, 8, 10, 3, 20, True)])
def test_cnn_train_step_matches_oracle(L, A, F, H, K, n, use_mask):
    shp = fo.CNNShape(L, A, F, H, K)
    ws = fo.trained_like_weights(shp.weight_shapes(), 21)
    rng = np.random.default_rng(4)
    idx = rng.integers(0, A, size=(n, L), dtype=np.uint8)
    y = rng.normal(size=n)
    mask = (rng.random((n, H)) >= 0.25).astype(np.float64) if use_mask else None
    loss, grads, _ = fo.cnn_loss_and_grads(idx, y, ws, mask)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=F, hidden_size=H, kernel_size=K)
    m.set_weights(ws)
    got_loss = _step(m, idx, y, mask)
    assert abs(got_loss - loss) <= 1e-4 * max(loss, 1e-6)
    ms, vs, step = m.optimizer_state(0)
    assert step == 1
    _check_grads(ms, grads, shp.weight_shapes())
    # updated weights vs Keras-Adam on the oracle gradients
    mo = [np.zeros(s) for s in shp.weight_shapes()]
    vo = [np.zeros(s) for s in shp.weight_shapes()]
    want = fo.adam_update([w.astype(np.float64) for w in ws], grads, mo, vo, 1)
    # Adam's first update is lr * g / (|g| + 3.2e-6): a +-1e-3 step whose sign is decided by the sign of g, so
    # an element whose gradient is ~1e-8 may legitimately flip.  Bound the worst case by one full step and
    # require everything but a sliver of near-zero-gradient elements to agree to 2 % of a step.
    diffs = np.concatenate([np.abs(a.astype(np.float64) - b.reshape(-1)) for a, b in zip(m.get_weights(0), want)])
    assert diffs.max() < 2.1e-3 and np.mean(diffs > 2e-5) < 2e-3
    # three more steps: moments and step count persist, like the compiled Keras optimiser's
    wcur = want
    for t in (2, 3, 4):
        _, g, _ = fo.cnn_loss_and_grads(idx, y, wcur, mask)
        wcur = fo.adam_update(wcur, g, mo, vo, t)
        _step(m, idx, y, mask)
    assert m.optimizer_state(0)[2] == 4
    diffs = np.concatenate([np.abs(a.astype(np.float64) - b.reshape(-1)) for a, b in zip(m.get_weights(0), wcur)])
    assert diffs.max() < 8.1e-3 and np.mean(diffs > 1e-4) < 5e-3
    m.reset_optimizer()
    assert m.optimizer_state(0)[2] == 0 and all(np.all(x == 0) for x in m.optimizer_state(0)[0])
    m.close()


def test_mlp_train_step_matches_oracle():
    L, A, H, n = 8, 4, 100, 50
    ms_ = fo.MLPShape(L, A, H)
    ws = fo.trained_like_weights(ms_.weight_shapes(), 3)
    rng = np.random.default_rng(1)
    idx = rng.integers(0, A, size=(n, L), dtype=np.uint8)
    y = rng.normal(size=n)
    loss, grads, _ = fo.mlp_loss_and_grads(idx, y, ws)
    m = _native.NativeModel("mlp", seq_len=L, alphabet_size=A, hidden_size=H)
    m.set_weights(ws)
    got_loss = _step(m, idx, y)
    assert abs(got_loss - loss) <= 1e-4 * loss
    _check_grads(m.optimizer_state(0)[0], grads, ms_.weight_shapes())
    with pytest.raises(ValueError):
        _step(m, idx, y, mask=np.ones((n, H)))  # the MLP has no Dropout layer
    m.close()


def _additive_problem(L, A, n, seed):
    rng = np.random.default_rng(seed)
    table = rng.normal(size=(L, A))
    idx = rng.integers(0, A, size=(n, L), dtype=np.uint8)
    y = table[np.arange(L)[None, :], idx].sum(axis=1)
    return idx, (y - y.mean()) / y.std()


@pytest.mark.parametrize("kind", ["cnn", "mlp"])
def test_fit_learns_and_is_deterministic(kind):
    """End metric of the whole fit loop: training loss falls, held-out r^2 is high, the Adam step
    counter equals epochs * ceil(n / batch) (weights + moments persist across fit calls), and the same
    seed reproduces the same weights."""
    L, A, n = 14, 4, 900
    idx, y = _additive_problem(L, A, n + 300, 0)
    tr_idx, tr_y, te_idx, te_y = idx[:n], y[:n], idx[n:], y[n:]

    def run():
        if kind == "cnn":
            m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
            m.set_weights(fo.glorot_weights(fo.CNNShape(L, A, 32, 100, 5).weight_shapes(), 1))
        else:
            m = _native.NativeModel("mlp", seq_len=L, alphabet_size=A, hidden_size=100)
            m.set_weights(fo.glorot_weights(fo.MLPShape(L, A, 100).weight_shapes(), 1))
        d_idx = torch.from_numpy(tr_idx).cuda()
        d_y = torch.from_numpy(tr_y.astype(np.float32)).cuda()
        l1 = m.fit_dev(d_idx.data_ptr(), d_y.data_ptr(), n, 256, 20, 7)
        l2 = m.fit_dev(d_idx.data_ptr(), d_y.data_ptr(), n, 256, 20, 8)
        return m, l1[:20], l2[:20]

    m, l1, l2 = run()
    assert l1[-1] < 0.5 * l1[0] and l2[-1] <= l1[-1] * 1.05
    assert m.optimizer_state(0)[2] == 2 * 20 * 4   # ceil(900/256) = 4 steps per epoch
    d = torch.from_numpy(te_idx).cuda()
    out = torch.empty(len(te_idx), dtype=torch.float32, device="cuda")
    m.forward_dev(d.data_ptr(), len(te_idx), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    pred = out.cpu().numpy()
    r2 = np.corrcoef(pred, te_y)[0, 1] ** 2
    assert r2 > 0.8, r2
    m2, l1b, _ = run()
    np.testing.assert_array_equal(l1, l1b)
    for a, b in zip(m.get_weights(0), m2.get_weights(0)):
        np.testing.assert_array_equal(a, b)
    m.close(); m2.close()


def test_python_surrogate_train_and_ensemble():
    """Model.train through the drop-in classes (what Explorer.run calls each round)."""
    import flexs_b200 as flexs

    L = 14
    idx, y = _additive_problem(L, 4, 600, 3)
    seqs = su.decode_indices(idx, su.RNAA)
    cnn = flexs.baselines.models.CNN(L, 32, 100, su.RNAA, seed=0)
    before = np.corrcoef(cnn.get_fitness(seqs), y)[0, 1] ** 2
    cnn.train(seqs[:500], y[:500])
    assert cnn.last_fit_losses.shape == (20,) and cnn.last_fit_losses[-1] < cnn.last_fit_losses[0]
    after = np.corrcoef(cnn.get_fitness(seqs[500:]), y[500:])[0, 1] ** 2
    assert after > 0.4 and after > before + 0.25
    ens = flexs.Ensemble([flexs.baselines.models.CNN(L, 32, 100, su.RNAA, seed=i) for i in range(3)])
    ens.train(seqs[:500], y[:500])
    r2 = np.corrcoef(ens.get_fitness(seqs[500:]), y[500:])[0, 1] ** 2
    assert r2 > 0.4
    mlp = flexs.baselines.models.MLP(L, 100, su.RNAA, seed=0)
    mlp.train(seqs[:500], y[:500])
    assert np.corrcoef(mlp.get_fitness(seqs[500:]), y[500:])[0, 1] ** 2 > 0.3
