"""GPU: parity of the CUDA path (through the C ABI) with the CPU oracle.

Tolerance (north_star: 1e-4 relative): ``|gpu - ref| <= 1e-4 * max(|ref|, FLOOR)``
with FLOOR = the batch score scale (max |ref|), because the final Dense(1) output may be
arbitrarily close to zero.  Integer work (encode, decode, ranking) is compared bit-exact.
"""
import numpy as np
import pytest

from tests.conftest import elem_rel_err, parity_report, rel_err

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from flexs_b200 import _native  # noqa: E402
from flexs_b200.utils import sequence_utils as su  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import flexs_oracle as fo  # noqa: E402

TOL = 1e-4


def _floor(ref):
    """Absolute floor of the relative test: the batch's score scale (max |ref|), i.e. errors are
    measured relative to max(|ref_i|, scale).  A pure per-element relative bound is meaningless for
    a Dense(1) output: it is a small difference of O(1) hidden activations, so ANY fp32 evaluation
    (numpy, C, TF) carries an absolute error of ~1e-7 x the hidden scale, which is ~1e-5 of the
    score scale for glorot-initialised long proteins (measured: the fp32 numpy/C oracles differ from
    the float64 definition by up to 2e-5 of the score scale on the 238-mer)."""
    return max(float(np.abs(ref).max()), 1e-7)


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device; the CUDA path has no CPU fallback")


def _device_forward(model, idx):
    d_idx = torch.from_numpy(np.ascontiguousarray(idx)).cuda()
    out = torch.empty(len(idx), dtype=torch.float32, device="cuda")
    model.forward_dev(d_idx.data_ptr(), len(idx), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _variants(model):
    vs = [_native.VARIANT_SIMPLE]
    for v in (_native.VARIANT_TILED, _native.VARIANT_UMMA, _native.VARIANT_UMMA_LUT, _native.VARIANT_ENUM):
        try:
            model.set_variant(v)
            vs.append(v)
        except ValueError:
            pass
    model.set_variant(_native.VARIANT_AUTO)
    return vs


CNN_SHAPES = {
    # tag: (L, A, F, H, K, N)
    "test_shape": (3, 4, 1, 1, 2, 37),      # tests/test_models.py:56-63 of the reference (even k)
    "tf8": (8, 4, 32, 100, 5, 1000),        # config 2
    "rna14": (14, 4, 32, 100, 5, 777),      # config 3 member
    "ns100": (100, 4, 32, 100, 5, 403),     # north-star shape
    "aav90": (90, 20, 32, 100, 5, 150),     # AAV registry window
    "gfp237": (237, 20, 32, 100, 5, 41),    # config 5
    "gfp238": (238, 20, 32, 100, 5, 40),
    "aav735": (735, 20, 32, 100, 5, 13),    # config 4
    "minlen": (5, 4, 32, 100, 5, 65),       # L == k: a single conv position
    "odd": (11, 7, 8, 10, 3, 50),           # generic shape only the simple kernel covers
    "dna600": (600, 4, 32, 100, 5, 33),     # A=4 longer than one 508-row chunk: chunked path + partial last chunk
    "dna1100": (1100, 4, 32, 100, 5, 19),   # three chunks per sequence
    "aa300": (300, 20, 32, 100, 5, 37),     # k3=19: chunks of 236 rows with partial tails (barrier phases per tile)
}


@pytest.mark.parametrize("tag", list(CNN_SHAPES))
@pytest.mark.parametrize("wname", ["glorot", "trained"])
def test_cnn_forward_parity_all_variants(tag, wname):
    L, A, F, H, K, n = CNN_SHAPES[tag]
    shp = fo.CNNShape(L, A, F, H, K)
    ws = (fo.glorot_weights if wname == "glorot" else fo.trained_like_weights)(shp.weight_shapes(), 11)
    idx = np.random.default_rng(5).integers(0, A, size=(n, L), dtype=np.uint8)
    ref32 = co.cnn_forward(idx, [ws], K)
    ref = fo.cnn_forward(idx, ws, np.float64)        # the float64 definition is the reference for every shape
    assert rel_err(ref32, ref, _floor(ref)) < 2e-5
    cpu_elem, _ = elem_rel_err(ref32, ref)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=F, hidden_size=H, kernel_size=K)
    m.set_weights(ws)
    for v in _variants(m):
        m.set_variant(v)
        try:
            got = _device_forward(m, idx)
        except ValueError as e:
            # the shape-generic kernel keeps a whole sequence's activations in shared memory: L <~ 900 at F=32
            assert v == _native.VARIANT_SIMPLE and "too long" in str(e) and L > 800
            continue
        scale_err = rel_err(got, ref, _floor(ref))
        elem_err, n_elem = elem_rel_err(got, ref)
        parity_report(f"{tag}/{wname}/{_native.VARIANT_NAMES[v]}", scale_rel=scale_err, elem_rel=elem_err, n_elem=n_elem,
                      n=n, cpu_fp32_elem_rel=cpu_elem, min_abs_ref_over_scale=float(np.abs(ref).min() / _floor(ref)))
        assert scale_err < TOL, (tag, wname, _native.VARIANT_NAMES[v])
        # north_star's bound element by element: every score of at least 1 % of the batch's scale is within 1e-4
        # RELATIVE of the float64 definition
        assert elem_err < TOL, (tag, wname, _native.VARIANT_NAMES[v], elem_err)
    m.close()


def test_cnn_forward_golden_vectors(golden):
    g = golden("oracle_forward.npz")
    for tag in ["test_shape", "tf8", "rna14", "ns100", "aav90", "gfp237", "gfp238", "aav735"]:
        for wname, fn in (("glorot", fo.glorot_weights), ("trained", fo.trained_like_weights)):
            L, A, F, H, K, wseed = [int(v) for v in g[f"{tag}_{wname}_cfg"]]
            ws = fn(fo.CNNShape(L, A, F, H, K).weight_shapes(), wseed)
            m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=F, hidden_size=H, kernel_size=K)
            m.set_weights(ws)
            ref = g[f"{tag}_{wname}_y"]
            for v in _variants(m):
                m.set_variant(v)
                got = _device_forward(m, g[f"{tag}_{wname}_idx"])
                assert rel_err(got, ref, _floor(ref)) < TOL, (tag, wname, v)
            m.close()


@pytest.mark.parametrize("n", [1, 2, 4, 5, 6, 63, 64, 65, 147, 148 * 5 + 3, 4099])
def test_cnn_batch_sizes_and_tile_boundaries(n):
    """Ragged batch sizes around the item/tile sizes of the tiled kernels (S=5 for 100-mers)."""
    L, A = 100, 4
    shp = fo.CNNShape(L, A, 32, 100, 5)
    ws = fo.trained_like_weights(shp.weight_shapes(), 3)
    idx = np.random.default_rng(n).integers(0, A, size=(n, L), dtype=np.uint8)
    ref = co.cnn_forward(idx, [ws])
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(ws)
    for v in _variants(m):
        if v == _native.VARIANT_SIMPLE and n > 1000:
            continue
        m.set_variant(v)
        assert rel_err(_device_forward(m, idx), ref, _floor(ref)) < TOL, (n, v)
    m.close()


def test_cnn_unaligned_device_pointer_and_empty_batch():
    """The idx staging copy rounds to 16-byte boundaries itself; any device pointer must work."""
    L, A, n = 100, 4, 333
    ws = fo.trained_like_weights(fo.CNNShape(L, A, 32, 100, 5).weight_shapes(), 3)
    idx = np.random.default_rng(0).integers(0, A, size=(n, L), dtype=np.uint8)
    ref = co.cnn_forward(idx, [ws])
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(ws)
    for off in (1, 7, 13):
        buf = torch.zeros(n * L + 64, dtype=torch.uint8, device="cuda")
        buf[off: off + n * L] = torch.from_numpy(idx.reshape(-1)).cuda()
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        m.forward_dev(buf.data_ptr() + off, n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert rel_err(out.cpu().numpy(), ref, _floor(ref)) < TOL
    m.forward_dev(0, 0, 0, 0)  # n == 0 is a no-op
    m.close()


def test_cnn_full_size_properties():
    """BASELINE-size batch (1M 8-mers; 2^18 100-mers): size-independent properties instead of a CPU
    re-computation — permutation equivariance, duplicate rows score identically, a strided sample
    matches the oracle, and the two kernel variants agree with each other everywhere."""
    for L, n in ((8, 1 << 20), (100, 1 << 18)):
        A = 4
        ws = fo.trained_like_weights(fo.CNNShape(L, A, 32, 100, 5).weight_shapes(), 9)
        rng = np.random.default_rng(L)
        idx = rng.integers(0, A, size=(n, L), dtype=np.uint8)
        idx[1::2] = idx[0::2]  # every odd row duplicates the row before it
        m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
        m.set_weights(ws)
        base = _device_forward(m, idx)
        np.testing.assert_array_equal(base[1::2], base[0::2])
        perm = rng.permutation(n)
        np.testing.assert_array_equal(_device_forward(m, idx[perm]), base[perm])
        sample = np.arange(0, n, 997)
        ref = co.cnn_forward(idx[sample], [ws])
        assert rel_err(base[sample], ref, _floor(ref)) < TOL
        vs = _variants(m)
        if len(vs) > 2:
            m.set_variant(vs[1])
            other = _device_forward(m, idx)
            assert rel_err(other, base, _floor(base)) < TOL
        m.close()


@pytest.mark.parametrize("L,n", [(8, 1000), (9, 333), (14, 2000), (20, 1000), (21, 259), (37, 1031), (100, 128 * 148 + 77), (120, 515), (170, 300)])
def test_cnn_table_kernel_lengths_and_ragged_groups(L, n):
    """cnn_k9.cu (conv1+conv2 as an L2-resident table over 9 residues): lengths whose conv positions do and do not
    fill the last 16-row tile, batches that end inside a group of 128 / an item of 8 sequences, unaligned pointers."""
    A = 4
    ws = fo.trained_like_weights(fo.CNNShape(L, A, 32, 100, 5).weight_shapes(), L)
    idx = np.random.default_rng(L).integers(0, A, size=(n, L), dtype=np.uint8)
    ref = co.cnn_forward(idx, [ws])
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(ws)
    m.set_variant(_native.VARIANT_UMMA_LUT)
    got = _device_forward(m, idx)
    assert rel_err(got, ref, _floor(ref)) < TOL
    for cut in (1, 7, 8, 9, 127, 129):
        np.testing.assert_array_equal(_device_forward(m, idx[:cut]), got[:cut])
    buf = torch.zeros(n * L + 64, dtype=torch.uint8, device="cuda")
    buf[5: 5 + n * L] = torch.from_numpy(idx.reshape(-1)).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    m.forward_dev(buf.data_ptr() + 5, n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), got)
    m.close()


@pytest.mark.parametrize("L,n", [(19, 333), (24, 1), (90, 9), (90, 1001), (237, 17), (238, 130), (300, 47), (735, 25)])
def test_cnn_protein_kernel_ragged_items_and_pairs(L, n):
    """cnn_a20.cu runs as CTA pairs (cta_group::2) on items of 8 sequences: batches with an odd number of items (one CTA
    of the last pair has no item), items with fewer than 8 sequences, a single sequence, the shortest supported
    sequence; prefixes of a batch score exactly like the batch (no cross-item state), unaligned pointers."""
    A = 20
    ws = fo.trained_like_weights(fo.CNNShape(L, A, 32, 100, 5).weight_shapes(), L)
    idx = np.random.default_rng(L * 7 + n).integers(0, A, size=(n, L), dtype=np.uint8)
    ref = co.cnn_forward(idx, [ws])
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(ws)
    m.set_variant(_native.VARIANT_UMMA)
    got = _device_forward(m, idx)
    assert rel_err(got, ref, _floor(ref)) < TOL
    assert elem_rel_err(got, ref)[0] < TOL
    for cut in (1, 7, 8, 9, 16, 17, 129):
        if cut < n:
            np.testing.assert_array_equal(_device_forward(m, idx[:cut]), got[:cut])
    buf = torch.zeros(n * L + 64, dtype=torch.uint8, device="cuda")
    buf[3: 3 + n * L] = torch.from_numpy(idx.reshape(-1)).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    m.forward_dev(buf.data_ptr() + 3, n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), got)
    m.close()


def test_cnn_table_kernel_selection_rebuild_and_ensemble():
    """AUTO picks the table kernel for large batches only; new weights rebuild the table; an ensemble keeps one
    table per member; residues outside [0, A) cannot index outside the table."""
    L, A = 40, 4
    shp = fo.CNNShape(L, A, 32, 100, 5)
    wss = [fo.trained_like_weights(shp.weight_shapes(), 30 + i) for i in range(3)]
    idx = np.random.default_rng(3).integers(0, A, size=(70_000, L), dtype=np.uint8)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(wss[0])
    assert m.active_variant(100) == _native.VARIANT_UMMA and m.active_variant(70_000) == _native.VARIANT_UMMA_LUT
    small = _device_forward(m, idx[:100])                      # cnn_umma2
    big = _device_forward(m, idx)                              # builds the table, cnn_k9
    assert m.active_variant(10_000) == _native.VARIANT_UMMA_LUT  # the table of these weights exists now
    sample = np.arange(0, len(idx), 499)
    ref = co.cnn_forward(idx[sample], [wss[0]])
    assert rel_err(big[sample], ref, _floor(ref)) < TOL
    assert rel_err(big[:100], small, _floor(small)) < TOL
    m.set_weights(wss[1])                                      # stale table must not be used
    assert m.active_variant(10_000) == _native.VARIANT_UMMA
    ref = co.cnn_forward(idx[sample], [wss[1]])
    assert rel_err(_device_forward(m, idx)[sample], ref, _floor(ref)) < TOL
    bad = idx[:300].copy(); bad[:, ::7] |= 0xF0               # garbage high bits: the kernel masks residues to 2 bits
    m.set_variant(_native.VARIANT_UMMA_LUT)
    np.testing.assert_array_equal(_device_forward(m, bad), _device_forward(m, bad & 3))
    m.close()
    ens = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5,
                              n_members=3)
    for i, ws in enumerate(wss):
        ens.set_weights(ws, i)
    ens.set_variant(_native.VARIANT_UMMA_LUT)
    sub = idx[:1000]
    ref = fo.ensemble_mean([fo.nan_to_num_f32(fo.cnn_forward(sub, ws, np.float64)) for ws in wss])
    assert rel_err(_device_forward(ens, sub), ref, _floor(ref)) < TOL
    ens.close()


@pytest.mark.parametrize("L", [8, 14, 20, 21, 37, 100, 120])
def test_cnn_table_kernel_random_batch_sizes_agree_with_umma2(L):
    """Many launches of random sizes (ragged groups, odd tile counts carried from group to group inside a CTA, one- and
    multi-tile items): the mbarrier protocol of cnn_k9 must neither hang nor drift from cnn_umma2 on the same inputs.
    (tools/k9_stress.py is the longer form of this test.)"""
    rng = np.random.default_rng(L)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=4, num_filters=32, hidden_size=100, kernel_size=5)
    m.set_weights(fo.trained_like_weights(fo.CNNShape(L, 4, 32, 100, 5).weight_shapes(), L))
    nmax = 120_000
    idx = torch.randint(0, 4, (nmax, L), dtype=torch.uint8, device="cuda")
    a = torch.empty(nmax, dtype=torch.float32, device="cuda")
    b = torch.empty(nmax, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for it in range(10):
        n = int(rng.integers(1, nmax)) if it % 2 else int(rng.integers(1, 3000))
        m.set_variant(_native.VARIANT_UMMA_LUT)
        m.forward_dev(idx.data_ptr(), n, a.data_ptr(), s)
        m.set_variant(_native.VARIANT_UMMA)
        m.forward_dev(idx.data_ptr(), n, b.data_ptr(), s)
        torch.cuda.synchronize()
        assert float((a[:n] - b[:n]).abs().max() / b[:n].abs().max()) < TOL, (L, n)
    m.close()


def test_whole_model_table_for_tiny_sequence_spaces():
    """enum_table.cu: for A^L <= 2^20 the model is evaluated once on every sequence and a batch becomes a gather.  AUTO
    switches to it for batches at least as large as the space, for CNNs, MLPs and ensembles alike; new weights
    invalidate the table; a residue outside the alphabet cannot read outside it."""
    L, A = 8, 4
    rng = np.random.default_rng(12)
    idx = rng.integers(0, A, size=(70_000, L), dtype=np.uint8)
    sample = np.arange(0, len(idx), 131)
    shp = fo.CNNShape(L, A, 32, 100, 5)
    wss = [fo.trained_like_weights(shp.weight_shapes(), 40 + i) for i in range(2)]
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5, n_members=2)
    for i, ws in enumerate(wss):
        m.set_weights(ws, i)
    assert m.active_variant(1000) != _native.VARIANT_ENUM and m.active_variant(4 ** L) == _native.VARIANT_ENUM
    direct = _device_forward(m, idx[:1000])
    big = _device_forward(m, idx)                              # scores all 65 536 8-mers once, then gathers
    assert m.active_variant(20) == _native.VARIANT_ENUM        # the table of these weights exists now
    ref = fo.ensemble_mean([fo.nan_to_num_f32(fo.cnn_forward(idx[sample], ws, np.float64)) for ws in wss])
    assert rel_err(big[sample], ref, _floor(ref)) < TOL
    assert rel_err(big[:1000], direct, _floor(direct)) < TOL
    np.testing.assert_array_equal(_device_forward(m, idx[:20]), big[:20])
    bad = idx[:64].copy(); bad[:, 3] = 200
    clamp = bad.copy(); clamp[:, 3] = A - 1
    np.testing.assert_array_equal(_device_forward(m, bad), _device_forward(m, clamp))
    m.set_weights(wss[0], 1)                                   # both members equal now; the old table must go
    assert m.active_variant(20) != _native.VARIANT_ENUM
    ref = fo.nan_to_num_f32(fo.cnn_forward(idx[sample], wss[0], np.float64))
    assert rel_err(_device_forward(m, idx)[sample], ref, _floor(ref)) < TOL
    m.close()
    ms = fo.MLPShape(L, A, 100)
    wm = fo.trained_like_weights(ms.weight_shapes(), 6)
    mlp = _native.NativeModel("mlp", seq_len=L, alphabet_size=A, hidden_size=100)
    mlp.set_weights(wm)
    direct = _device_forward(mlp, idx[:500])
    mlp.set_variant(_native.VARIANT_ENUM)
    np.testing.assert_array_equal(_device_forward(mlp, idx[:500]), direct)  # same kernel filled the table
    mlp.close()
    with pytest.raises(ValueError):
        big_space = _native.NativeModel("cnn", seq_len=14, alphabet_size=4, num_filters=32, hidden_size=100, kernel_size=5)
        big_space.set_variant(_native.VARIANT_ENUM)


@pytest.mark.parametrize("members", [2, 3])
def test_ensemble_mean_fused(members):
    """Ensemble of CNNs in one launch == np.mean over the members' oracle scores (ensemble.py:54-59)."""
    L, A, n = 14, 4, 500
    shp = fo.CNNShape(L, A, 32, 100, 5)
    wss = [fo.trained_like_weights(shp.weight_shapes(), 20 + i) for i in range(members)]
    idx = np.random.default_rng(1).integers(0, A, size=(n, L), dtype=np.uint8)
    per_member = [fo.nan_to_num_f32(fo.cnn_forward(idx, ws, np.float64)) for ws in wss]
    ref = fo.ensemble_mean(per_member)
    assert ref.dtype == np.float32
    np.testing.assert_allclose(co.cnn_forward(idx, wss), ref, rtol=0, atol=2e-6)
    m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5,
                            n_members=members)
    for i, ws in enumerate(wss):
        m.set_weights(ws, i)
    for v in _variants(m):
        m.set_variant(v)
        assert rel_err(_device_forward(m, idx), ref, _floor(ref)) < TOL, v
    # get_weights returns what was set, member by member, in Keras order
    for i, ws in enumerate(wss):
        for a, b in zip(m.get_weights(i), ws):
            np.testing.assert_array_equal(a, b.reshape(-1))
    m.close()


def _mlp_variants(model):
    vs = []
    for v in (_native.VARIANT_TILED, _native.VARIANT_UMMA):      # FP32 FFMA kernel / every layer on tcgen05 (H <= 112)
        try:
            model.set_variant(v)
            vs.append(v)
        except ValueError:
            pass
    model.set_variant(_native.VARIANT_AUTO)
    return vs


@pytest.mark.parametrize("L,A,H,n", [(8, 4, 100, 1000), (3, 4, 1, 5), (14, 4, 200, 130), (90, 20, 100, 77), (100, 4, 100, 2500),
                                     (237, 20, 100, 300), (8, 4, 112, 129), (5, 7, 33, 128)])
@pytest.mark.parametrize("wname", ["glorot", "trained"])
def test_mlp_forward_parity(L, A, H, n, wname):
    """Both MLP kernels (FP32 FFMA; every layer on tcgen05 with the one-hot built as an exact fp16 operand) against the
    float64 definition of mlp.py:21-31, scale-normalised and element by element; ensembles too."""
    ms = fo.MLPShape(L, A, H)
    make = fo.glorot_weights if wname == "glorot" else fo.trained_like_weights
    ws = make(ms.weight_shapes(), 4)
    idx = np.random.default_rng(2).integers(0, A, size=(n, L), dtype=np.uint8)
    ref = fo.mlp_forward(idx, ws, np.float64)
    m = _native.NativeModel("mlp", seq_len=L, alphabet_size=A, hidden_size=H)
    m.set_weights(ws)
    vs = _mlp_variants(m)
    assert _native.VARIANT_TILED in vs and ((_native.VARIANT_UMMA in vs) == (H <= 112))
    for v in vs:
        m.set_variant(v)
        got = _device_forward(m, idx)
        parity_report(f"mlp{L}x{A}_H{H}/{wname}/{_native.VARIANT_NAMES[v]}", scale_rel=rel_err(got, ref, _floor(ref)),
                      elem_rel=elem_rel_err(got, ref)[0], n=n)
        assert rel_err(got, ref, _floor(ref)) < TOL, v
        assert elem_rel_err(got, ref)[0] < TOL, v
    m.set_variant(_native.VARIANT_AUTO)
    assert m.active_variant(n) in ((_native.VARIANT_UMMA if H <= 112 else _native.VARIANT_TILED), _native.VARIANT_ENUM)
    m.close()
    # ensemble of MLPs
    ws2 = make(ms.weight_shapes(), 5)
    m2 = _native.NativeModel("mlp", seq_len=L, alphabet_size=A, hidden_size=H, n_members=2)
    m2.set_weights(ws, 0); m2.set_weights(ws2, 1)
    ref2 = fo.ensemble_mean([fo.nan_to_num_f32(ref), fo.nan_to_num_f32(fo.mlp_forward(idx, ws2, np.float64))])
    for v in _mlp_variants(m2):
        m2.set_variant(v)
        assert rel_err(_device_forward(m2, idx), ref2, _floor(ref2)) < TOL, v
    m2.close()
    m2.close()


def test_nan_to_num_epilogue():
    """keras_model.py:77: NaN -> 0, +-inf -> +-float32 max."""
    L, A = 8, 4
    shp = fo.CNNShape(L, A, 32, 100, 5)
    idx = np.zeros((4, L), dtype=np.uint8)
    for bias, want in ((np.inf, np.finfo(np.float32).max), (-np.inf, np.finfo(np.float32).min), (np.nan, 0.0)):
        ws = fo.glorot_weights(shp.weight_shapes(), 0)
        ws[11] = np.array([bias], dtype=np.float32)
        m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
        m.set_weights(ws)
        for v in _variants(m):
            m.set_variant(v)
            np.testing.assert_array_equal(_device_forward(m, idx), np.full(4, want, dtype=np.float32))
        m.close()


# ------------------------------------------------------------------------------------ encode
def test_encode_kernel_bit_exact(golden):
    g = golden("ref_encode_decode.json")
    for name, case in g["encode"].items():
        chars = su.sequences_to_char_array(case["seqs"])
        d_chars = torch.from_numpy(chars).cuda()
        d_idx = torch.empty_like(d_chars)
        status = torch.zeros(2, dtype=torch.int64, device="cuda")
        _native.encode_dev(d_chars.data_ptr(), chars.size, case["alphabet"], d_idx.data_ptr(), status.data_ptr())
        torch.cuda.synchronize()
        np.testing.assert_array_equal(d_idx.cpu().numpy(), np.array(case["idx"], dtype=np.uint8), err_msg=name)
        assert status[0].item() == 0


def test_encode_kernel_large_unaligned_and_bad_chars():
    rng = np.random.default_rng(0)
    alphabet = su.AAS
    n_bytes = 1_000_003
    idx = rng.integers(0, 20, size=n_bytes, dtype=np.uint8)
    chars = np.frombuffer(alphabet.encode(), dtype=np.uint8)[idx].copy()
    bad_positions = [17, 999_999, 123_456]
    for p in bad_positions:
        chars[p] = ord("X")
    for off in (0, 3):
        buf = torch.zeros(n_bytes + 32, dtype=torch.uint8, device="cuda")
        buf[off: off + n_bytes] = torch.from_numpy(chars).cuda()
        out = torch.zeros(n_bytes + 32, dtype=torch.uint8, device="cuda")
        status = torch.zeros(2, dtype=torch.int64, device="cuda")
        _native.encode_dev(buf.data_ptr() + off, n_bytes, alphabet, out.data_ptr() + off, status.data_ptr())
        torch.cuda.synchronize()
        got = out[off: off + n_bytes].cpu().numpy()
        mask = np.ones(n_bytes, dtype=bool); mask[bad_positions] = False
        np.testing.assert_array_equal(got[mask], idx[mask])
        assert status.tolist() == [3, 17]


def test_score_host_strings_and_value_error():
    """The reference-facing call: list[str] / ndarray[str] in, float32 ndarray out; a character
    outside the alphabet raises ValueError like str.index (sequence_utils.py:46)."""
    import flexs_b200 as flexs

    L = 14
    cnn = flexs.baselines.models.CNN(L, 32, 100, su.RNAA, seed=3)
    ws = fo.trained_like_weights(fo.CNNShape(L, 4, 32, 100, 5).weight_shapes(), 2)
    cnn.set_weights(ws)
    seqs = su.generate_random_sequences(L, 3000, su.RNAA)
    ref = fo.get_fitness_cnn(seqs, su.RNAA, ws, np.float64)
    for container in (seqs, np.array(seqs), su.encode_sequences(seqs, su.RNAA)):
        got = cnn.get_fitness(container)
        assert isinstance(got, np.ndarray) and got.dtype == np.float32 and got.shape == (3000,)
        assert rel_err(got, ref, _floor(ref)) < TOL
    assert cnn.cost == 9000
    with pytest.raises(ValueError):
        cnn.get_fitness(seqs[:5] + ["UGCAUGCAUGCAUX"])
    with pytest.raises(ValueError):
        cnn.get_fitness(["UGCA"])  # wrong length
    assert cnn.get_fitness([]).shape == (0,)
    for a, b in zip(cnn.get_weights(), ws):
        np.testing.assert_array_equal(a, b)
    # device-resident entry used by the explorers
    d = torch.from_numpy(su.encode_sequences(seqs, su.RNAA)).cuda()
    got = cnn.get_fitness_device(d).cpu().numpy()
    assert rel_err(got, ref, _floor(ref)) < TOL


def test_score_host_multi_chunk_pipeline():
    """More sequences than one staging slot holds: chunks alternate between the two streams."""
    import flexs_b200 as flexs

    L, n = 100, 700_000  # three equal chunks of 233 344 sequences (a staging slot holds 335 488 100-mers)
    cnn = flexs.baselines.models.CNN(L, 32, 100, su.DNAA, seed=1)
    idx = np.random.default_rng(0).integers(0, 4, size=(n, L), dtype=np.uint8)
    chars = np.frombuffer(su.DNAA.encode(), dtype=np.uint8)[idx]
    got = cnn.native.score_host(np.ascontiguousarray(chars), su.DNAA)
    d = torch.from_numpy(idx).cuda()
    want = cnn._score_device(d).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    chars = chars.copy(); chars[150_000, 7] = ord("N")
    with pytest.raises(ValueError, match="150000"):
        cnn.native.score_host(np.ascontiguousarray(chars), su.DNAA)


def test_fused_ensemble_drop_in_and_costs():
    import flexs_b200 as flexs

    L = 14
    members = [flexs.baselines.models.CNN(L, 32, 100, su.RNAA, seed=i) for i in range(3)]
    ens = flexs.Ensemble(members)
    seqs = su.generate_random_sequences(L, 257, su.RNAA)
    got = ens.get_fitness(seqs)
    ref = fo.ensemble_mean([fo.get_fitness_cnn(seqs, su.RNAA, m.get_weights(), np.float64) for m in members])
    assert got.dtype == np.float32 and rel_err(got, ref, _floor(ref)) < TOL
    assert ens.cost == 257 and [m.cost for m in members] == [257] * 3
    # the generic path (custom reducer) gives the same numbers through M separate launches
    ens2 = flexs.Ensemble(members, combine_with=lambda s: np.mean(s, axis=1))
    assert rel_err(ens2.get_fitness(seqs), ref, _floor(ref)) < TOL
    # weights changed on a member -> fused copy refreshes
    members[1].set_weights(fo.trained_like_weights(fo.CNNShape(L, 4, 32, 100, 5).weight_shapes(), 99))
    ref = fo.ensemble_mean([fo.get_fitness_cnn(seqs, su.RNAA, m.get_weights(), np.float64) for m in members])
    assert rel_err(ens.get_fitness(seqs), ref, _floor(ref)) < TOL


# ------------------------------------------------------------------------------------ top-k
def _topk(scores, k, offset=0, index_map=None):
    d = torch.from_numpy(scores).cuda()
    ws = torch.empty(_native.topk_workspace_bytes(len(scores), k), dtype=torch.uint8, device="cuda")
    ts = torch.empty(k, dtype=torch.float32, device="cuda")
    ti = torch.empty(k, dtype=torch.int64, device="cuda")
    dm = torch.from_numpy(index_map).cuda() if index_map is not None else None
    _native.topk_dev(d.data_ptr(), len(scores), k, offset, dm.data_ptr() if dm is not None else 0, ts.data_ptr(),
                     ti.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return ts.cpu().numpy(), ti.cpu().numpy()


def test_topk_matches_reference_slices(golden):
    g = golden("ref_topk_slices.json")
    preds = np.array(g["preds"], dtype=np.float32)
    for b in (2, 5, 100, 499, 500):
        ts, ti = _topk(preds, b - 1)      # adalead.py:173: B-1 items
        np.testing.assert_array_equal(ti, g[f"bm1_{b}"])
        np.testing.assert_array_equal(ts, preds[g[f"bm1_{b}"]])
        ts, ti = _topk(preds, b)          # dyna_ppo.py:317: B items
        np.testing.assert_array_equal(ti, g[f"b_{b}"])
    ts, ti = _topk(preds, 600)            # k > n: tail is (-inf, -1)
    np.testing.assert_array_equal(ti[:500], g["b_600"])
    assert (ti[500:] == -1).all() and np.isneginf(ts[500:]).all()


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 1), (4097, 100), (1 << 20, 99), (3_000_001, 1000), (5000, 4096)])
def test_topk_random_ties_and_scale(n, k):
    rng = np.random.default_rng(n)
    # coarse values -> many ties; negative values, zeros and -0.0 included
    scores = (rng.integers(-50, 50, size=n) / 8.0).astype(np.float32)
    scores[rng.integers(0, n, size=min(n, 10))] = -0.0
    ts, ti = _topk(scores, k, offset=1000)
    keff = min(n, k)
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))[:keff]  # score desc, index asc
    np.testing.assert_array_equal(ti[:keff] - 1000, order)
    np.testing.assert_array_equal(ts[:keff], scores[order] + 0.0)


def test_topk_index_map():
    scores = np.array([0.5, 2.0, 2.0, -1.0, 7.0], dtype=np.float32)
    imap = np.array([40, 10, 11, 99, 5], dtype=np.int64)
    ts, ti = _topk(scores, 3, index_map=imap)
    np.testing.assert_array_equal(ti, [5, 10, 11])
    np.testing.assert_array_equal(ts, [7.0, 2.0, 2.0])


# ------------------------------------------------------------------------------------ K5
def test_argmax_decode_bit_exact(golden):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(300, 14, 5)).round(1).astype(np.float32)  # ties; last column = mask channel
    d = torch.from_numpy(x).cuda()
    out = torch.empty((300, 14), dtype=torch.uint8, device="cuda")
    _native.argmax_decode_dev(d.data_ptr(), 300, 14, 5, 4, out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), np.argmax(x[:, :, :4], axis=2))
    for case in golden("ref_encode_decode.json")["decode"]:
        xx = np.array(case["x"], dtype=np.float32)
        d = torch.from_numpy(xx).cuda()
        out = torch.empty(xx.shape[:2], dtype=torch.uint8, device="cuda")
        _native.argmax_decode_dev(d.data_ptr(), xx.shape[0], xx.shape[1], xx.shape[2], xx.shape[2], out.data_ptr())
        torch.cuda.synchronize()
        assert list(su.decode_indices(out.cpu().numpy(), case["alphabet"])) == case["strings"]


def test_mutate_rates_and_determinism():
    """generate_random_mutant semantics (sequence_utils.py:87-108), distributional: each residue is
    redrawn with probability mu, uniformly over the WHOLE alphabet (so it changes w.p. mu*(A-1)/A)."""
    n, L, A = 20000, 100, 4
    parents = np.random.default_rng(0).integers(0, A, size=(n, L), dtype=np.uint8)
    d = torch.from_numpy(parents).cuda()
    out = torch.empty_like(d)
    for mu in (0.0, 1.0 / L, 0.2, 1.0):
        _native.mutate_dev(d.data_ptr(), n, L, A, mu, 1234, 0, out.data_ptr())
        torch.cuda.synchronize()
        child = out.cpu().numpy()
        assert child.max() < A
        changed = (child != parents).mean()
        expect = mu * (A - 1) / A
        assert abs(changed - expect) < 5 * np.sqrt(max(expect, 1e-9) / (n * L)) + 1e-9, (mu, changed, expect)
        if mu == 1.0:
            counts = np.bincount(child.reshape(-1), minlength=A) / child.size
            assert np.abs(counts - 1 / A).max() < 2e-3
    out2 = torch.empty_like(d)
    _native.mutate_dev(d.data_ptr(), n, L, A, 0.2, 1234, 0, out.data_ptr())
    _native.mutate_dev(d.data_ptr(), n, L, A, 0.2, 1234, 0, out2.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    _native.mutate_dev(d.data_ptr(), n, L, A, 0.2, 1234, 1, out2.data_ptr())
    torch.cuda.synchronize()
    assert not torch.equal(out, out2)


def test_real_keras_golden_if_present():
    """tools/export_keras_golden.py output (made where TensorFlow exists) pins the float parity."""
    from tests.golden import keras_loader

    files = keras_loader.committed_files()
    if not files:
        pytest.skip("no real-Keras vectors committed (TensorFlow is absent in the authoring container): parity unpinned")
    for path in files:
        g = keras_loader.load(path)
        m = _native.NativeModel(g["kind"], **g["cfg"])
        m.set_weights(g["weights"])
        got = _device_forward(m, g["idx"])
        assert rel_err(got, g["y"], _floor(g["y"])) < TOL, path
        assert elem_rel_err(got, g["y"])[0] < TOL, path
        m.close()


def test_fp16_range_guard_falls_back_to_fp32_kernel():
    """The tcgen05 kernels carry activations as scaled fp16 hi/lo pairs; values above 60000/8 raise a flag and a
    gated launch of the FP32 FFMA kernel recomputes the batch.  Huge conv1 weights force that path."""
    for L, A in ((100, 4), (30, 20)):
        shp = fo.CNNShape(L, A, 32, 100, 5)
        ws = fo.trained_like_weights(shp.weight_shapes(), 2)
        ws[0] = ws[0] * np.float32(4e4)     # conv1 activations ~1e4-1e5 -> beyond the fp16 window
        idx = np.random.default_rng(0).integers(0, A, size=(257, L), dtype=np.uint8)
        ref = co.cnn_forward(idx, [ws])
        m = _native.NativeModel("cnn", seq_len=L, alphabet_size=A, num_filters=32, hidden_size=100, kernel_size=5)
        m.set_weights(ws)
        for v in _variants(m):
            m.set_variant(v)
            got = _device_forward(m, idx)
            assert np.isfinite(got).all()
            assert rel_err(got, ref, _floor(ref)) < TOL, (L, A, v)
        m.close()


@pytest.mark.parametrize("L,A,n", [(8, 4, 300_000), (100, 4, 50_000), (3, 2, 1000), (237, 20, 4000), (1, 4, 77)])
def test_dedup_scores_exact(L, A, n):
    """flexs_dedup_scores_dev: the first occurrence of every distinct sequence keeps its score, every repeat gets
    -inf — bit-exact against numpy's unique (the reference's dict keys, adalead.py:157)."""
    rng = np.random.default_rng(L * 1000 + A)
    idx = rng.integers(0, A, size=(n, L), dtype=np.uint8)
    if L >= 100:                                   # long random sequences never repeat: plant repeats
        src = rng.integers(0, n // 2, size=n // 3)
        idx[n - len(src):] = idx[src]
    scores = rng.normal(size=n).astype(np.float32)
    _, first = np.unique(idx, axis=0, return_index=True)
    want = np.full(n, -np.inf, dtype=np.float32)
    want[first] = scores[first]
    d_idx = torch.from_numpy(idx).cuda()
    d_s = torch.from_numpy(scores).cuda()
    d_o = torch.empty_like(d_s)
    work = torch.empty(_native.dedup_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    for _ in range(2):                             # the workspace is reusable as is
        _native.dedup_scores_dev(d_idx.data_ptr(), n, L, d_s.data_ptr(), d_o.data_ptr(), work.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(d_o.cpu().numpy(), want)
    _native.dedup_scores_dev(d_idx.data_ptr(), n, L, d_s.data_ptr(), d_s.data_ptr(), work.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)   # in place
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_s.cpu().numpy(), want)


def test_virtual_screen_ranks_distinct_sequences():
    """A screen whose candidates repeat (1M draws from the 65 536 8-mers in the reference's setting) must return k
    DISTINCT winners, each through its first occurrence, like the reference's ranking of dict keys."""
    import flexs_b200 as flexs
    from flexs_b200.screen import VirtualScreen

    L, n, B = 8, 200_000, 100
    cnn = flexs.baselines.models.CNN(L, 32, 100, su.DNAA, seed=5)
    idx = np.random.default_rng(3).integers(0, 4, size=(n, L), dtype=np.uint8)
    scores = cnn.get_fitness(idx)
    _, first = np.unique(idx, axis=0, return_index=True)
    first = np.sort(first)
    order = first[np.lexsort((first, -scores[first].astype(np.float64)))][: B - 1]
    top_i, top_s = VirtualScreen(cnn, k=B - 1).screen(idx)
    np.testing.assert_array_equal(top_i, order)
    np.testing.assert_array_equal(top_s, scores[order])
    assert len({bytes(r) for r in idx[top_i]}) == B - 1
    # rows, not sequences: the best sequence's copies fill the list
    rows_i, _ = VirtualScreen(cnn, k=B - 1, unique=False).screen(idx)
    assert len({bytes(r) for r in idx[rows_i]}) < B - 1
    # fewer distinct candidates than k: the list is short, not padded with repeats
    few_i, few_s = VirtualScreen(cnn, k=B - 1).screen(idx[:40].repeat(5, axis=0))
    assert len(few_i) == len(np.unique(idx[:40], axis=0)) and np.isfinite(few_s).all()


def test_virtual_screen_single_rank_matches_numpy():
    """flexs_b200.screen.VirtualScreen (the sharded top-k of the path) on one rank: same winners, same order and
    scores as np.argsort over get_fitness; with >= 2 GPUs the NCCL form is covered by bench.py --gpus N."""
    import flexs_b200 as flexs
    from flexs_b200.screen import VirtualScreen

    L, n, B = 14, 5000, 100
    ens = flexs.Ensemble([flexs.baselines.models.CNN(L, 32, 100, su.RNAA, seed=i) for i in range(3)])
    seqs = np.array(su.generate_random_sequences(L, n, su.RNAA))
    scores = ens.get_fitness(seqs)
    screen = VirtualScreen(ens, k=B - 1)          # the [: -B : -1] slice keeps B-1
    top_seqs, top_scores = screen.screen(seqs)
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))[: B - 1]
    np.testing.assert_array_equal(top_seqs, seqs[order])
    np.testing.assert_array_equal(top_scores, scores[order])
    assert ens.cost == 2 * n and all(m.cost == 2 * n for m in ens.models)
    # tie-free batches agree with the reference's own slice
    if len(np.unique(scores)) == n:
        np.testing.assert_array_equal(top_seqs, seqs[np.argsort(scores)[: -B: -1]])


@pytest.mark.parametrize("L,alphabet", [(100, su.DNAA), (237, su.AAS), (8, su.DNAA), (11, "ABCDEFG")])
def test_packed_wire_format_device_round_trip_bit_exact(L, alphabet):
    """The packed wire format (include/flexs_b200.h): host packers (numpy and the C pass over str objects) and the
    device pack / unpack kernels agree bit for bit; a value outside the alphabet is reported with its position."""
    A = len(alphabet)
    n = 3001
    idx = np.random.default_rng(L).integers(0, A, size=(n, L), dtype=np.uint8)
    packed = su.pack_indices(idx, A)
    assert packed.shape == (n, _native.packed_row_bytes(L, A))
    assert int(_native.lib().flexs_packed_row_bytes(L, A)) == packed.shape[1]
    assert int(_native.lib().flexs_bits_per_residue(A)) == _native.bits_per_residue(A)
    np.testing.assert_array_equal(su.pack_sequences(list(su.decode_indices(idx, alphabet)), alphabet), packed)
    np.testing.assert_array_equal(su.unpack_indices(packed, L, A), idx)
    d_packed = torch.from_numpy(packed).cuda()
    d_idx = torch.empty((n, L), dtype=torch.uint8, device="cuda")
    status = torch.zeros(2, dtype=torch.int64, device="cuda")
    _native.unpack_dev(d_packed.data_ptr(), n, L, A, d_idx.data_ptr(), status.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_idx.cpu().numpy(), idx)
    assert status.tolist()[0] == 0
    d_back = torch.zeros_like(d_packed)
    _native.pack_dev(d_idx.data_ptr(), n, L, A, d_back.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_back.cpu().numpy(), packed)
    if (1 << _native.bits_per_residue(A)) > A:   # a bit pattern that is not a residue
        bad = idx.copy().astype(np.uint16)
        bad[7, 3] = A
        planes = ((bad[:, :, None] >> np.arange(_native.bits_per_residue(A))) & 1).astype(np.uint8).reshape(n, -1)
        d_bad = torch.from_numpy(np.packbits(planes, axis=1, bitorder="little")).cuda()
        _native.unpack_dev(d_bad.data_ptr(), n, L, A, d_idx.data_ptr(), status.data_ptr())
        torch.cuda.synchronize()
        assert status.tolist() == [1, 7 * L + 3]


def test_score_host_packed_matches_byte_route_and_strings():
    """flexs_model_score_host_packed == flexs_model_score_host == the device-resident forward, bit for bit (the same
    kernels behind a different first stage), across several chunks; and Model.get_fitness(list[str]) takes that route
    for large lists, raising the reference's ValueError for a foreign character."""
    import flexs_b200 as flexs

    for L, alphabet, n in ((100, su.DNAA, 700_000), (237, su.AAS, 20_000)):
        A = len(alphabet)
        cnn = flexs.baselines.models.CNN(L, 32, 100, alphabet, seed=1)
        idx = np.random.default_rng(1).integers(0, A, size=(n, L), dtype=np.uint8)
        want = cnn._score_device(torch.from_numpy(idx).cuda()).cpu().numpy()
        got = cnn.native.score_host_packed(su.pack_indices(idx, A))
        np.testing.assert_array_equal(got, want)
        seqs = list(su.decode_indices(idx[:9000], alphabet))
        cost0 = cnn.cost
        np.testing.assert_array_equal(cnn.get_fitness([str(s) for s in seqs]), want[:9000])
        assert cnn.cost == cost0 + 9000
        bad = [str(s) for s in seqs]
        bad[4321] = bad[4321][:5] + "!" + bad[4321][6:]
        with pytest.raises(ValueError, match="4321"):
            cnn.get_fitness(bad)
        with pytest.raises(ValueError):
            cnn.native.score_host_packed(np.zeros((4, 3), dtype=np.uint8))


def _select(scores, k, rows=None, unique=False, offset=0, want_rows=False):
    n = len(scores)
    d = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)).cuda()
    d_rows = torch.from_numpy(np.ascontiguousarray(rows)).cuda() if rows is not None else None
    L = rows.shape[1] if rows is not None else 0
    work = torch.empty(_native.topk_select_workspace_bytes(), dtype=torch.uint8, device="cuda")
    ts = torch.empty(k, dtype=torch.float32, device="cuda")
    ti = torch.empty(k, dtype=torch.int64, device="cuda")
    tr = torch.full((k, max(L, 1)), 255, dtype=torch.uint8, device="cuda") if want_rows else None
    status = torch.full((8,), -7, dtype=torch.int32, device="cuda")
    _native.topk_select_dev(d.data_ptr(), n, k, offset, d_rows.data_ptr() if d_rows is not None else 0, L, unique,
                            ts.data_ptr(), ti.data_ptr(), tr.data_ptr() if tr is not None else 0, status.data_ptr(),
                            work.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return ts.cpu().numpy(), ti.cpu().numpy(), (tr.cpu().numpy() if tr is not None else None), int(status[0].item())


@pytest.mark.parametrize("n,k", [(1, 1), (5, 8), (1000, 99), (70_000, 4096), (3_000_001, 99), (1 << 22, 1000)])
def test_topk_select_single_launch_matches_numpy(n, k):
    """flexs_topk_select_dev == np.lexsort((position, -score)) bit for bit, ties included; also == flexs_topk_dev."""
    rng = np.random.default_rng(n + k)
    for kind in ("normal", "ties", "wall"):
        if kind == "normal":
            s = rng.normal(size=n).astype(np.float32)
        elif kind == "ties":
            s = (rng.integers(-30, 30, size=n) / 4).astype(np.float32)
            s[rng.integers(0, n, size=max(1, n // 50))] = -0.0
        else:
            s = np.full(n, 1.25, dtype=np.float32)          # every score equal: the descent runs into the position bits
        order = np.lexsort((np.arange(n), -s.astype(np.float64)))[:k]
        ts, ti, _, st = _select(s, k, offset=1000)
        m = min(n, k)
        np.testing.assert_array_equal(ti[:m], order + 1000, err_msg=kind)
        np.testing.assert_array_equal(ts[:m].view(np.int32) & 0x7fffffff, s[order].view(np.int32) & 0x7fffffff)
        assert np.all(ti[m:] == -1) and np.all(np.isneginf(ts[m:])) and st == 0
        if n <= 70_000:
            rs, ri = _topk(s, k, offset=1000)
            np.testing.assert_array_equal(ri, ti)


@pytest.mark.parametrize("L,A,n,k", [(8, 4, 200_000, 99), (100, 4, 300_000, 99), (14, 4, 50_000, 1000), (237, 20, 20_000, 99)])
def test_topk_select_distinct_rows_lazy_dedup_exact(L, A, n, k):
    """`unique`: the k best DISTINCT rows through their first occurrence == the reference's ranking of dict keys,
    with the winners' rows written next to them; equal rows carry equal scores (a deterministic surrogate)."""
    rng = np.random.default_rng(L)
    copies = 7 if k < 500 else 3     # the select looks at the best max(8k, 4096) rows (<= 8192)
    pool = np.unique(rng.integers(0, A, size=(max(64, n // copies), L), dtype=np.uint8), axis=0)   # distinct rows
    pool = pool[rng.permutation(len(pool))]
    pick = rng.integers(0, len(pool), size=n)
    rows = pool[pick]
    pool_scores = (rng.integers(-2000, 2000, size=len(pool)) / 16).astype(np.float32)   # ties between different rows too
    s = pool_scores[pick]
    _, first = np.unique(rows, axis=0, return_index=True)
    first = np.sort(first)
    order = first[np.lexsort((first, -s[first].astype(np.float64)))][:k]
    ts, ti, tr, st = _select(s, k, rows=rows, unique=True, offset=5, want_rows=True)
    assert st == 0
    np.testing.assert_array_equal(ti[: len(order)], order + 5)
    np.testing.assert_array_equal(ts[: len(order)], s[order])
    np.testing.assert_array_equal(tr[: len(order)], rows[order])
    # rows only (no de-duplication) still returns the winners' rows
    ts2, ti2, tr2, _ = _select(s, k, rows=rows, unique=False, want_rows=True)
    o2 = np.lexsort((np.arange(n), -s.astype(np.float64)))[:k]
    np.testing.assert_array_equal(ti2, o2)
    np.testing.assert_array_equal(tr2, rows[o2])


def test_topk_select_reports_when_the_best_rows_hold_too_few_distinct_sequences():
    """A batch that is almost all repeats: the best 4096 rows hold fewer than k distinct sequences although the batch has
    more — status 1, and VirtualScreen falls back to hashing every row (dedup.cu): still the exact answer."""
    import flexs_b200 as flexs
    from flexs_b200.screen import VirtualScreen

    L, n, k = 8, 120_000, 99
    cnn = flexs.baselines.models.CNN(L, 32, 100, su.DNAA, seed=5)
    base = np.random.default_rng(1).integers(0, 4, size=(150, L), dtype=np.uint8)
    base_scores = cnn.get_fitness(base)
    best = base[np.argsort(-base_scores)[:40]]
    idx = np.concatenate([np.repeat(best, 2900, axis=0), base])     # 116 000 copies of the 40 best, then everything once
    idx = idx[np.random.default_rng(2).permutation(len(idx))][:n]
    scores = cnn.get_fitness(idx)
    _, _, _, st = _select(scores, k, rows=idx, unique=True)
    assert st == 1
    _, first = np.unique(idx, axis=0, return_index=True)
    first = np.sort(first)
    order = first[np.lexsort((first, -scores[first].astype(np.float64)))][:k]
    vs = VirtualScreen(cnn, k=k)
    top_i, top_s = vs.screen(idx)
    assert vs.fallbacks == 1
    np.testing.assert_array_equal(top_i, order)
    np.testing.assert_array_equal(top_s, scores[order])


@pytest.mark.parametrize("world,k,L", [(8, 99, 100), (2, 1000, 14), (4, 99, 0), (3, 7, 237)])
def test_screen_merge_kernel_matches_reference_merge(world, k, L):
    """flexs_screen_merge_dev on synthetic gathered messages (repeats across ranks, absent winners, score ties) ==
    screen.merge_reference, the numpy statement the gloo tests pin."""
    from flexs_b200 import screen

    rng = np.random.default_rng(world * 1000 + k)
    mb = screen.message_bytes(k, L)
    assert mb == _native.screen_message_bytes(k, L)
    gathered = torch.zeros(world * mb, dtype=torch.uint8)
    shared_rows = rng.integers(0, 4, size=(k, max(L, 1)), dtype=np.uint8)
    shared_scores = np.sort((rng.integers(0, 400, size=k) / 8).astype(np.float32))[::-1]
    for r in range(world):
        idx_v, sc_v, rows_v = screen.message_views(gathered[r * mb:(r + 1) * mb], k, L)
        own = rng.random(k) < 0.6
        rows = np.where(own[:, None], rng.integers(0, 4, size=(k, max(L, 1)), dtype=np.uint8), shared_rows)
        sc = np.where(own, (rng.integers(0, 400, size=k) / 8).astype(np.float32), shared_scores).astype(np.float32)
        o = np.argsort(-sc, kind="stable")
        sc, rows = sc[o], rows[o]
        gi = np.sort(rng.choice(1 << 20, size=k, replace=False)) + (r << 20)
        if r == world - 1:
            sc[-3:] = -np.inf; gi[-3:] = -1
        sc_v.copy_(torch.from_numpy(sc)); idx_v.copy_(torch.from_numpy(gi))
        if L:
            rows_v.copy_(torch.from_numpy(rows))
    want_s, want_i, want_rows = screen.merge_reference(gathered.numpy(), world, k, L)
    d = gathered.cuda()
    out = torch.zeros(mb, dtype=torch.uint8, device="cuda")
    fi, fs, fr = screen.message_views(out, k, L)
    _native.screen_merge_dev(d.data_ptr(), world, k, L, fs.data_ptr(), fi.data_ptr(), fr.data_ptr() if L else 0,
                             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    m = len(want_i)
    np.testing.assert_array_equal(fi.cpu().numpy()[:m], want_i)
    np.testing.assert_array_equal(fs.cpu().numpy()[:m], want_s)
    if L:
        np.testing.assert_array_equal(fr.cpu().numpy()[:m], want_rows)
    assert np.all(fi.cpu().numpy()[m:] == -1)


def test_peer_mailbox_push_wait_merge_on_one_gpu():
    """csrc/peer.cu without a second process: the `world` ranks' pushes all target this GPU's own mailbox (the peer table
    lists it `world` times), over 9 steps so the 4 slots are reused; after the stream-side wait the slot holds the
    rank-major block flexs_screen_merge_dev takes, and the merge equals screen.merge_reference."""
    from flexs_b200 import screen

    world, k, L, depth = 4, 50, 37, 4
    mb = screen.message_bytes(k, L)
    own, handle = _native.peer_alloc(_native.peer_mailbox_bytes(mb, world, depth))
    assert len(handle) == 64
    bases = torch.tensor([own] * world, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    side = torch.cuda.Stream()
    rng = np.random.default_rng(5)
    try:
        for step in range(1, 10):
            slot = step % depth
            msgs = torch.zeros(world, mb, dtype=torch.uint8)
            for r in range(world):
                idx_v, sc_v, rows_v = screen.message_views(msgs[r], k, L)
                sc = np.sort((rng.integers(0, 300, size=k) / 4).astype(np.float32))[::-1].copy()
                sc_v.copy_(torch.from_numpy(sc))
                idx_v.copy_(torch.from_numpy(np.sort(rng.choice(1 << 16, size=k, replace=False)) + (r << 16)))
                rows_v.copy_(torch.from_numpy(rng.integers(0, 3, size=(k, L), dtype=np.uint8)))   # few letters: repeats across ranks
            d_msgs = msgs.cuda()
            # step 5: one rank's message arrives late (its push sits behind a ~10 ms spin on another stream; everything is
            # enqueued before anything waits, so the host never blocks on the spinning wait): the merge below can only be
            # right if flexs_screen_wait_dev really held the stream until that message was in
            last = world - 1 if step == 5 else None
            for r in range(world):
                if r != last:
                    _native.screen_push_dev(d_msgs[r].data_ptr(), mb, r, world, slot, depth, step, bases.data_ptr(), stream)
            if last is not None:
                with torch.cuda.stream(side):
                    torch.cuda._sleep(20_000_000)
                    _native.screen_push_dev(d_msgs[last].data_ptr(), mb, last, world, slot, depth, step, bases.data_ptr(),
                                            side.cuda_stream)
            status = torch.zeros(1, dtype=torch.int32, device="cuda")
            _native.screen_wait_dev(own, mb, world, slot, depth, step, status.data_ptr(), stream)
            out = torch.zeros(mb, dtype=torch.uint8, device="cuda")
            fi, fs, fr = screen.message_views(out, k, L)
            _native.screen_merge_dev(own + slot * world * mb, world, k, L, fs.data_ptr(), fi.data_ptr(), fr.data_ptr(), stream)
            torch.cuda.synchronize()
            assert int(status.item()) == 0
            want_s, want_i, want_rows = screen.merge_reference(msgs.reshape(-1).numpy(), world, k, L)
            m = len(want_i)
            np.testing.assert_array_equal(fi.cpu().numpy()[:m], want_i)
            np.testing.assert_array_equal(fs.cpu().numpy()[:m], want_s)
            np.testing.assert_array_equal(fr.cpu().numpy()[:m], want_rows)
    finally:
        torch.cuda.synchronize()
        _native.peer_free(own)


def test_mlp_fp16_range_guard_falls_back_to_fp32_kernel():
    """The tcgen05 MLP carries hidden activations as scaled fp16 hi/lo pairs: values above 60000/8 raise the per-stream
    flag and the gated FP32 kernel recomputes the batch — same contract as the CNN kernels."""
    L, A, H, n = 20, 4, 100, 300
    ms = fo.MLPShape(L, A, H)
    ws = fo.trained_like_weights(ms.weight_shapes(), 2)
    ws[0] = ws[0] * np.float32(3e4)          # layer-1 activations far beyond the fp16 window
    idx = np.random.default_rng(0).integers(0, A, size=(n, L), dtype=np.uint8)
    ref = fo.mlp_forward(idx, ws, np.float64)
    m = _native.NativeModel("mlp", seq_len=L, alphabet_size=A, hidden_size=H)
    m.set_weights(ws)
    m.set_variant(_native.VARIANT_UMMA)
    got = _device_forward(m, idx)
    assert np.isfinite(got).all() and rel_err(got, ref, _floor(ref)) < TOL
    m.close()
