"""CPU: the oracle against the reference-generated golden vectors and against itself."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import flexs_oracle as fo


def test_oracle_encode_matches_reference_one_hot_argmax(golden):
    g = golden("ref_encode_decode.json")
    for name, case in g["encode"].items():
        idx = fo.encode(case["seqs"], case["alphabet"])
        assert idx.dtype == np.uint8
        np.testing.assert_array_equal(idx, np.array(case["idx"], dtype=np.uint8), err_msg=name)
        # plain-C restatement too
        n, length = idx.shape
        c_idx = co.encode("".join(case["seqs"]).encode(), n, length, case["alphabet"])
        np.testing.assert_array_equal(c_idx, idx)
        # one_hot round trip is exact
        oh = fo.one_hot(idx, len(case["alphabet"]))
        assert fo.decode_argmax(oh, case["alphabet"]) == case["seqs"]


def test_oracle_encode_error_behaviour(golden):
    assert golden("ref_encode_decode.json")["bad_char_exception"] == "ValueError"
    with pytest.raises(ValueError):
        fo.encode(["ATXG"], "ATCG")
    with pytest.raises(ValueError):
        co.encode(b"ATXG", 1, 4, "ATCG")


def test_oracle_decode_first_max_wins(golden):
    for case in golden("ref_encode_decode.json")["decode"]:
        x = np.array(case["x"])
        assert fo.decode_argmax(x, case["alphabet"]) == case["strings"]


def test_oracle_top_slices(golden):
    g = golden("ref_topk_slices.json")
    preds = np.array(g["preds"], dtype=np.float32)
    for b in (1, 2, 5, 100, 499, 500, 600):
        np.testing.assert_array_equal(fo.top_slice_bm1(preds, b), g[f"bm1_{b}"])
        np.testing.assert_array_equal(fo.top_slice_b(preds, b), g[f"b_{b}"])
        assert len(g[f"bm1_{b}"]) == min(b - 1, len(preds))  # the B-1 quirk
        assert len(g[f"b_{b}"]) == min(b, len(preds))


def test_oracle_ensemble_mean_matches_reference(golden):
    g = golden("ref_ensemble.json")
    out = fo.ensemble_mean([np.full(2, c, dtype=np.float32) for c in (0.1, 0.7, 0.25)])
    assert out.dtype == np.float32 == np.dtype(g["out_dtype"])
    np.testing.assert_array_equal(out, np.array(g["out"], dtype=np.float32))
    # the fused kernels compute ((s0+s1)+s2)/3 in fp32: bit-identical to numpy's mean for M < 8
    rng = np.random.default_rng(0)
    s = rng.normal(size=(1000, 3)).astype(np.float32)
    manual = ((s[:, 0] + s[:, 1]) + s[:, 2]) / np.float32(3)
    np.testing.assert_array_equal(np.mean(s, axis=1), manual)


FORWARD_TAGS = ["test_shape", "tf8", "rna14", "ns100", "aav90", "gfp237", "gfp238", "aav735"]


@pytest.mark.parametrize("tag", FORWARD_TAGS)
@pytest.mark.parametrize("wname", ["glorot", "trained"])
def test_oracle_forward_regression_and_cross_check(golden, tag, wname):
    """numpy float64 definition == committed vectors; numpy fp32 and plain-C fp32 restatements
    (independent code: einsum over padded windows vs explicit loops) agree with it to ~1e-6."""
    g = golden("oracle_forward.npz")
    idx, y = g[f"{tag}_{wname}_idx"], g[f"{tag}_{wname}_y"]
    L, A, F, H, K, wseed = [int(v) for v in g[f"{tag}_{wname}_cfg"]]
    shp = fo.CNNShape(L, A, F, H, K)
    ws = (fo.glorot_weights if wname == "glorot" else fo.trained_like_weights)(shp.weight_shapes(), wseed)
    y64 = fo.cnn_forward(idx, ws, np.float64)
    np.testing.assert_allclose(y64, y, rtol=1e-12, atol=1e-14)
    scale = max(np.abs(y).max(), 1e-3)
    if L <= 100:  # the numpy fp32 einsum is slow for the long proteins; the C one covers them
        assert np.abs(fo.cnn_forward(idx, ws, np.float32) - y).max() <= 2e-5 * scale
    assert np.abs(co.cnn_forward(idx, [ws], K) - y).max() <= 2e-5 * scale


def test_oracle_same_padding_even_kernel():
    """TF 'same' rule for even k: pad_left=(k-1)//2, the extra zero goes on the right."""
    x = np.arange(1, 5, dtype=np.float64).reshape(1, 4, 1)
    w = np.array([1.0, 10.0]).reshape(2, 1, 1)
    out = fo.conv1d(x, w, np.zeros(1), "same")[0, :, 0]
    np.testing.assert_array_equal(out, [1 + 20, 2 + 30, 3 + 40, 4 + 0])
    w3 = np.array([1.0, 10.0, 100.0]).reshape(3, 1, 1)
    out3 = fo.conv1d(x, w3, np.zeros(1), "same")[0, :, 0]
    np.testing.assert_array_equal(out3, [0 + 10 + 200, 1 + 20 + 300, 2 + 30 + 400, 3 + 40 + 0])


def test_oracle_mlp_and_c_agree(golden):
    g = golden("oracle_forward.npz")
    ms = fo.MLPShape(8, 4, 100)
    ws = fo.trained_like_weights(ms.weight_shapes(), 5)
    y = fo.mlp_forward(g["mlp8_idx"], ws, np.float64)
    np.testing.assert_allclose(y, g["mlp8_y"], rtol=1e-12)
    assert np.abs(co.mlp_forward(g["mlp8_idx"], [ws]) - y).max() < 1e-5


def test_oracle_gradients_numerically():
    """cnn_loss_and_grads / mlp_loss_and_grads against central differences (float64)."""
    rng = np.random.default_rng(3)
    shp = fo.CNNShape(9, 4, 3, 5, 2)
    ws = [w.astype(np.float64) for w in fo.trained_like_weights(shp.weight_shapes(), 1)]
    idx = rng.integers(0, 4, size=(6, 9), dtype=np.uint8)
    y = rng.normal(size=6)
    mask = (rng.random((6, 5)) < 0.75).astype(np.float64)
    loss, grads, _ = fo.cnn_loss_and_grads(idx, y, ws, mask)
    for wi in range(12):
        flat = ws[wi].reshape(-1)
        for e in rng.choice(flat.size, size=min(3, flat.size), replace=False):
            old = flat[e]
            flat[e] = old + 1e-6; lp = fo.cnn_loss_and_grads(idx, y, ws, mask)[0]
            flat[e] = old - 1e-6; lm = fo.cnn_loss_and_grads(idx, y, ws, mask)[0]
            flat[e] = old
            num = (lp - lm) / 2e-6
            assert abs(num - np.asarray(grads[wi]).reshape(-1)[e]) < 1e-5 * max(1.0, abs(num)), (wi, e)
    ms = fo.MLPShape(5, 4, 6)
    ws = [w.astype(np.float64) for w in fo.trained_like_weights(ms.weight_shapes(), 2)]
    idx = rng.integers(0, 4, size=(7, 5), dtype=np.uint8)
    y = rng.normal(size=7)
    loss, grads, _ = fo.mlp_loss_and_grads(idx, y, ws)
    for wi in range(8):
        flat = ws[wi].reshape(-1)
        e = int(rng.integers(flat.size))
        old = flat[e]
        flat[e] = old + 1e-6; lp = fo.mlp_loss_and_grads(idx, y, ws)[0]
        flat[e] = old - 1e-6; lm = fo.mlp_loss_and_grads(idx, y, ws)[0]
        flat[e] = old
        assert abs((lp - lm) / 2e-6 - np.asarray(grads[wi]).reshape(-1)[e]) < 1e-5


@pytest.mark.parametrize("tag,L,alphabet", [("ns100", 100, "TGCA"), ("tf8", 8, "ACGT"), ("aav90", 90, "ILVAGMFYWEDQNHCRKSTP"),
                                            ("test_shape", 3, "ATCG")])
def test_reference_arm_port_matches_oracle(tag, L, alphabet):
    """bench.py --impl reference times oracle/ref_path.ReferenceCNN (the reference's own one-hot loop, then oneDNN-backed
    torch-CPU conv1d / linear standing in for TF's CPU kernels).  oneDNN is an implementation of the same Keras layer
    semantics that shares no code with oracle/flexs_oracle.py (numpy) or oracle/c/oracle.c: the three must agree, which
    both checks the denominator of the headline ratio and gives the oracle an independent third opinion (padding rule,
    channels-last, (k, in, out) kernels, cross-correlation without flip)."""
    torch = pytest.importorskip("torch")
    from oracle import ref_path

    A = len(alphabet)
    F, H, K = (1, 1, 2) if tag == "test_shape" else (32, 100, 5)   # tests/test_models.py:56-63: even kernel, asymmetric pad
    shp = fo.CNNShape(L, A, F, H, K)
    rng = np.random.default_rng(3)
    for wname, fn in (("glorot", fo.glorot_weights), ("trained", fo.trained_like_weights)):
        ws = fn(shp.weight_shapes(), 7)
        idx = rng.integers(0, A, size=(300, L), dtype=np.uint8)
        seqs = ["".join(alphabet[i] for i in row) for row in idx]
        torch.set_num_threads(2)
        port = ref_path.ReferenceCNN(L, alphabet, F, H, K, ws, batch_size=128)   # 300 = 2 full batches + a ragged one
        got = port.get_fitness(seqs)
        assert got.dtype == np.float32 and got.shape == (300,) and port.cost == 300
        ref64 = fo.cnn_forward(idx, ws, np.float64)
        scale = max(float(np.abs(ref64).max()), 1e-7)
        assert np.max(np.abs(got - ref64)) <= 2e-5 * scale, (tag, wname)
        np.testing.assert_allclose(got, fo.get_fitness_cnn(seqs, alphabet, ws), rtol=0, atol=2e-5 * scale)
        np.testing.assert_allclose(got, co.cnn_forward(idx, [ws], K), rtol=0, atol=2e-5 * scale)


def test_keras_golden_format_roundtrip(tmp_path):
    """tools/export_keras_golden.py -> tests/golden/keras_loader.py -> oracle, end to end on the CPU.  TensorFlow is
    absent here, so the file comes from the tool's torch backend (same writer, same loader, same consumers as a real
    Keras export); what a real export adds is the pin, not the plumbing."""
    import subprocess
    import sys
    from pathlib import Path

    from tests.golden import keras_loader

    tool = Path(__file__).resolve().parent.parent / "tools" / "export_keras_golden.py"
    for kind, L, alphabet in (("cnn", 23, "TGCA"), ("cnn", 30, "ILVAGMFYWEDQNHCRKSTP"), ("mlp", 8, "TGCA")):
        out = tmp_path / f"fmt_{kind}_{L}.npz"
        subprocess.run([sys.executable, str(tool), "--out", str(out), "--seq-len", str(L), "--alphabet", alphabet,
                        "--kind", kind, "--n", "64", "--backend", "torch"], check=True, stdout=subprocess.PIPE)
        g = keras_loader.load(str(out))
        assert g["backend"] == "torch" and g["kind"] == kind and g["cfg"]["seq_len"] == L
        ref = (fo.cnn_forward if kind == "cnn" else fo.mlp_forward)(g["idx"], g["weights"], np.float64)
        assert np.max(np.abs(g["y"] - ref)) <= 2e-5 * np.abs(ref).max()


def test_keras_golden_files_match_oracle():
    """Every committed tests/golden/keras_*.npz (real TensorFlow output) pins the oracle itself; none can be made in
    the authoring container, so this skips there — the oracle header and DESIGN.md say "parity unpinned"."""
    from tests.golden import keras_loader

    files = keras_loader.committed_files()
    if not files:
        pytest.skip("no real-Keras vectors committed: floating-point parity unpinned")
    for path in files:
        g = keras_loader.load(path)
        assert g["backend"] == "tensorflow", f"{path} was not written by TensorFlow: it pins nothing"
        ref = (fo.cnn_forward if g["kind"] == "cnn" else fo.mlp_forward)(g["idx"], g["weights"], np.float64)
        assert np.max(np.abs(g["y"] - ref)) <= 1e-4 * np.abs(ref).max(), path
