"""GPU: the table-landscape kernels (K6) through the drop-in classes, bit-exact against outputs of the reference's own
``AdditiveAAVPackaging`` / ``TFBinding`` classes (tests/golden/ref_landscapes.json) and, at full size, against a
vectorised numpy evaluation of the same left-to-right float64 sums."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import flexs_b200 as flexs  # noqa: E402
from flexs_b200 import _native  # noqa: E402

GOLD = Path(__file__).parent / "golden"
AAV_FILE = str(GOLD / "aav_450_540_subs.json")
TF_FILE = str(GOLD / "tfbind_4mers.txt")


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device; the CUDA path has no CPU fallback")


@pytest.fixture(scope="module")
def ref():
    return json.load(open(GOLD / "ref_landscapes.json"))


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def test_additive_aav_matches_reference_bit_exact(ref):
    for case in ref["aav"]:
        land = flexs.landscapes.AdditiveAAVPackaging(phenotype=case["phenotype"], minimum_fitness_multiplier=case["mfm"],
                                                     start=450, end=540, noise=case["noise"], data_file=AAV_FILE)
        np.random.seed(case["seed"])
        out = land.get_fitness(case["sequences"])
        assert out.dtype == np.dtype(case["dtype"]) and land.cost == case["cost"]
        np.testing.assert_array_equal(out, unhex(case["fitness"]))
        assert float(np.random.random()).hex() == case["next_uniform"]     # numpy's global stream left where the reference leaves it
    land = flexs.landscapes.AdditiveAAVPackaging(phenotype="heart", start=450, end=540, data_file=AAV_FILE)
    out = land.get_fitness(["P" * 90, "W" * 90])
    assert out.dtype == np.dtype(ref["aav_all_clipped"]["dtype"]) and out.tolist() == ref["aav_all_clipped"]["out"]
    with pytest.raises(KeyError, match="540"):
        land.get_fitness([land.wild_type + "A"])
    assert land.get_fitness([]).shape == (0,)


def test_additive_device_resident_full_size_and_long_table():
    """1M candidates that never leave the GPU vs the same sums in numpy; then a 735-position table (L2 path)."""
    land = flexs.landscapes.AdditiveAAVPackaging(phenotype="liver", start=450, end=540, data_file=AAV_FILE)
    rng = np.random.default_rng(5)
    wt = np.frombuffer(land.wild_type.encode(), dtype=np.uint8)
    n = 1 << 20
    chars = np.tile(wt, (n, 1))
    letters = np.frombuffer(b"ILVAGMFYWEDQNHCRKSPT*X", dtype=np.uint8)
    for _ in range(3):                                                    # three substitutions per candidate
        pos = rng.integers(0, 90, size=n)
        chars[np.arange(n), pos] = letters[rng.integers(0, len(letters), size=n)]
    d = torch.from_numpy(chars).cuda()
    got = land.get_fitness_device(d)
    assert got.is_cuda and got.dtype == torch.float64 and land.cost == n
    cols = land.column_of_char[chars]
    tab = np.concatenate([land.table, np.zeros((90, 1))], axis=1)         # column for "no column" (0xFF -> last)
    cols = np.where(cols == 0xFF, tab.shape[1] - 1, cols)
    total = np.zeros(n)
    for i in range(90):
        total = total + tab[i, cols[:, i]]
    want = (total + land._offset) / land._denom
    want = np.where(want > 0, want, 0.0)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert (want > 0).mean() > 0.3
    # column-index input (candidates generated on the device in table-column order)
    got2 = land.get_fitness_device(torch.from_numpy(np.where(cols == tab.shape[1] - 1, 0xFF, cols).astype(np.uint8)).cuda(),
                                   columns=True, charge=False)
    np.testing.assert_array_equal(got2.cpu().numpy(), want)

    # a table too large for shared memory: 1400 positions x 21 columns x 8 B = 235 KB
    L, C, m = 1400, 21, 4096
    table = rng.normal(size=(L, C))
    seq = rng.integers(0, C, size=(m, L), dtype=np.uint8)
    d_t, d_s = torch.from_numpy(table).cuda(), torch.from_numpy(seq).cuda()
    out = torch.empty(m, dtype=torch.float64, device="cuda")
    _native.additive_score_dev(d_s.data_ptr(), m, L, None, C, d_t.data_ptr(), 3.0, 7.0, 0, out.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    total = np.zeros(m)
    for i in range(L):
        total = total + table[i, seq[:, i]]
    want = (total + 3.0) / 7.0
    np.testing.assert_array_equal(out.cpu().numpy(), np.where(want > 0, want, 0.0))


def test_tfbinding_matches_reference_bit_exact(ref):
    land = flexs.landscapes.TFBinding(TF_FILE)
    out = land.get_fitness(ref["tf"]["query"])
    assert out.dtype == np.dtype(ref["tf"]["dtype"])
    np.testing.assert_array_equal(out, unhex(ref["tf"]["fitness"]))
    with pytest.raises(KeyError, match="ACGN"):
        land.get_fitness(["ACGT", "ACGN"])
    assert land.cost == ref["tf"]["cost"]        # the failed call was charged too (landscape.py:44 runs first)
    keys = sorted(ref["tf"]["dict"])
    np.testing.assert_array_equal(land.get_fitness(np.array(keys)), unhex([ref["tf"]["dict"][k] for k in keys]))
    with pytest.raises(KeyError, match="ACGN"):
        land.get_fitness(["ACGT", "ACGN"])
    with pytest.raises(KeyError):
        land.get_fitness(["ACG"])


def test_tfbinding_8mer_device_resident_full_table(tmp_path):
    """All 65 536 8-mers x 16 through the device path against a host gather of the same table."""
    import itertools

    rng = np.random.default_rng(8)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    seen, rows = set(), []
    for tup in itertools.product("ACGT", repeat=8):
        s = "".join(tup)
        if s in seen:
            continue
        r = "".join(comp[c] for c in reversed(s))
        seen.update((s, r))
        rows.append((s, r))
    esc = np.round(rng.uniform(-0.5, 0.5, size=len(rows)), 5)
    path = tmp_path / "SYN_8mers.txt"
    with open(path, "w") as f:
        f.write("8-mer\t8-mer\tE-score\tMedian\tZ-score\n")
        for (s, r), e in zip(rows, esc):
            f.write(f"{s}\t{r}\t{e:.5f}\t0\t0\n")
    land = flexs.landscapes.TFBinding(str(path))
    assert land.table.size == 65536 and not np.isnan(land.table).any()
    idx = rng.integers(0, 4, size=(1 << 20, 8), dtype=np.uint8)
    chars = np.frombuffer(b"ACGT", dtype=np.uint8)[idx]
    got = land.get_fitness_device(torch.from_numpy(chars).cuda())
    key = idx.astype(np.int64) @ (4 ** np.arange(7, -1, -1))
    np.testing.assert_array_equal(got.cpu().numpy(), land.table[key])
    got = land.get_fitness_device(torch.from_numpy(idx).cuda(), columns=True)
    np.testing.assert_array_equal(got.cpu().numpy(), land.table[key])
    assert land.cost == 2 << 20
    # Adalead runs on it through the plugin API exactly as in the reference's smoke test (tests/test_explorers.py)
    model = flexs.LandscapeAsModel(land)
    ex = flexs.baselines.explorers.Adalead(model, rounds=2, sequences_batch_size=5, model_queries_per_batch=20,
                                          starting_sequence="GCTCGAGC", alphabet="ACGT", eval_batch_size=1)
    table, _ = ex.run(land, verbose=False)
    assert table["round"].max() == 2 and np.isfinite(table["true_score"]).all()
