"""CPU: host-side logic of the drop-in package against reference-generated goldens, and the C ABI's
symbol table.  No compute call is made here (there is no GPU in the authoring container)."""
import random
import re
import warnings

import numpy as np
import pandas as pd
import pytest

import flexs_b200 as flexs
from flexs_b200.utils import sequence_utils as su


# ---------------------------------------------------------------------------- sequence_utils
def test_alphabets_match_reference():
    assert (su.AAS, su.RNAA, su.DNAA, su.BA) == ("ILVAGMFYWEDQNHCRKSTP", "UGCA", "TGCA", "01")


def test_encode_sequences_matches_reference(golden):
    g = golden("ref_encode_decode.json")
    for name, case in g["encode"].items():
        want = np.array(case["idx"], dtype=np.uint8)
        for container in (case["seqs"], np.array(case["seqs"]), np.array(case["seqs"]).astype("S")):
            np.testing.assert_array_equal(su.encode_sequences(container, case["alphabet"]), want, err_msg=name)
        # string_to_one_hot: float64 (L, A), same integer content
        oh = su.string_to_one_hot(case["seqs"][0], case["alphabet"])
        assert oh.dtype == np.float64 and oh.shape == (len(case["seqs"][0]), len(case["alphabet"]))
        np.testing.assert_array_equal(oh.argmax(1), want[0])
        assert su.one_hot_to_string(oh, case["alphabet"]) == case["seqs"][0]
        np.testing.assert_array_equal(su.decode_indices(want, case["alphabet"]), np.array(case["seqs"]))


def test_encode_errors_like_str_index():
    with pytest.raises(ValueError):
        su.string_to_one_hot("ATXG", "ATCG")
    with pytest.raises(ValueError):
        su.encode_sequences(["ATCG", "ATC"], "ATCG")  # ragged
    with pytest.raises(ValueError):
        su.encode_sequences(["AT☃G"], "ATCG")
    assert su.encode_sequences([], "ATCG").shape[0] == 0


def test_one_hot_to_string_first_max(golden):
    for case in golden("ref_encode_decode.json")["decode"]:
        for x, want in zip(case["x"], case["strings"]):
            assert su.one_hot_to_string(np.array(x), case["alphabet"]) == want


def test_mutation_helpers_reproduce_reference_rng_stream(golden):
    g = golden("ref_mutation.json")
    for case in g["mutants"]:
        random.seed(case["seed"])
        got = [su.generate_random_mutant(case["seq"], case["mu"], case["alphabet"]) for _ in range(20)]
        assert got == case["mutants"]
    rs = g["random_sequences"]
    random.seed(rs["seed"])
    assert su.generate_random_sequences(rs["length"], rs["number"], rs["alphabet"]) == rs["out"]
    sm = g["single_mutants"]
    assert su.generate_single_mutants(sm["wt"], sm["alphabet"]) == sm["out"]


def test_construct_mutant_from_sample():
    base = su.string_to_one_hot("ATC", "ATCG")
    sample = np.zeros((3, 4)); sample[1, 3] = 1
    out = su.construct_mutant_from_sample(sample, base)
    assert su.one_hot_to_string(out, "ATCG") == "AGC"


# ---------------------------------------------------------------------------- plugin ABCs
class HashModel(flexs.Model):
    def __init__(self):
        super().__init__("hash")

    def _fitness_function(self, sequences):
        from tests.golden.make_golden import hash_model_score

        return np.array([hash_model_score(s) for s in sequences])

    def train(self, *a, **k):
        pass


class Const(flexs.Model):
    def __init__(self, c):
        super().__init__(f"c{c}")
        self.c = c

    def _fitness_function(self, sequences):
        return np.full(len(sequences), self.c, dtype=np.float32)

    def train(self, *a, **k):
        pass


def test_landscape_cost_and_landscape_as_model():
    m = Const(1.0)
    assert m.cost == 0
    m.get_fitness(["A", "B", "C"])
    assert m.cost == 3
    wrapped = flexs.LandscapeAsModel(m)
    assert wrapped.name == "LandscapeAsModel=c1.0"
    wrapped.get_fitness(["A"])
    assert wrapped.cost == 1 and m.cost == 3  # wrapper is charged, not the landscape


def test_ensemble_matches_reference(golden):
    g = golden("ref_ensemble.json")
    members = [Const(0.1), Const(0.7), Const(0.25)]
    ens = flexs.Ensemble(members)
    out = ens.get_fitness(["AAA", "CCC"])
    assert ens.name == g["name"]
    assert str(out.dtype) == g["out_dtype"]
    np.testing.assert_array_equal(out, np.array(g["out"], dtype=np.float32))
    assert ens.cost == g["ens_cost"] and [m.cost for m in members] == g["member_costs"]
    # custom reducer
    ens2 = flexs.Ensemble(members, combine_with=lambda s: s.max(axis=1))
    np.testing.assert_allclose(ens2.get_fitness(["AAA"]), [0.7])


def test_adalead_reproduces_reference_proposals_and_cost(golden):
    from flexs_b200.baselines.explorers import Adalead

    for run in golden("ref_adalead.json")["runs"]:
        random.seed(run["seed"])
        np.random.seed(run["seed"])
        model = HashModel()
        ex = Adalead(model, rounds=1, sequences_batch_size=run["batch"], model_queries_per_batch=run["queries"],
                     starting_sequence=run["start"], alphabet=run["alphabet"], eval_batch_size=run["eval_batch_size"],
                     rho=run["rho"], recomb_rate=run["recomb_rate"])
        from tests.golden.make_golden import hash_model_score

        df = pd.DataFrame({"sequence": run["measured"], "true_score": [hash_model_score(s) for s in run["measured"]],
                           "model_score": np.nan, "round": 0})
        seqs, preds = ex.propose_sequences(df)
        assert list(map(str, seqs)) == run["proposed"], run["seed"]
        np.testing.assert_array_equal(np.asarray(preds, dtype=np.float64), np.array(run["preds"]))
        assert model.cost == run["model_cost"]
        assert len(seqs) <= run["batch"] - 1  # the B-1 quirk (adalead.py:173)
        assert ex.name == "Adalead_mu=1_threshold=0.05"


def test_adalead_raises_when_nothing_generated():
    from flexs_b200.baselines.explorers import Adalead

    ex = Adalead(HashModel(), rounds=1, sequences_batch_size=5, model_queries_per_batch=5,
                 starting_sequence="ATCATCAT", alphabet="ATCG", eval_batch_size=20)
    df = pd.DataFrame({"sequence": ["ATCATCAT"], "true_score": [0.5], "model_score": np.nan, "round": 0})
    with pytest.raises(ValueError, match="No sequences generated"):
        ex.propose_sequences(df)


def test_explorer_run_bookkeeping_and_log(tmp_path):
    """Explorer.run with fakes: columns, costs and the JSON+CSV log (explorer.py:92-184)."""
    from flexs_b200.baselines.explorers import Adalead

    random.seed(0)
    log = tmp_path / "sub" / "run.csv"
    model = HashModel()
    ex = Adalead(model, rounds=3, sequences_batch_size=5, model_queries_per_batch=20, starting_sequence="ATCATCAT",
                 alphabet="ATCG", eval_batch_size=1, log_file=str(log))
    landscape = HashModel()
    table, meta = ex.run(landscape, verbose=False)
    assert list(table.columns) == ["sequence", "model_score", "true_score", "round", "model_cost", "measurement_cost"]
    assert table.iloc[0]["round"] == 0 and np.isnan(table.iloc[0]["model_score"]) and table.iloc[0]["measurement_cost"] == 1
    assert set(table["round"]) == {0, 1, 2, 3}
    for r in (1, 2, 3):
        rows = table[table["round"] == r]
        assert 1 <= len(rows) <= 4  # B-1
        assert (rows["measurement_cost"] == (table["round"] <= r).sum()).all()
    assert landscape.cost == len(table)
    assert meta["exp_name"] == ex.name and meta["model_name"] == "hash" and meta["rounds"] == 3
    lines = log.read_text().splitlines()
    import json

    assert json.loads(lines[0])["sequences_batch_size"] == 5
    assert lines[1] == "sequence,model_score,true_score,round,model_cost,measurement_cost"
    assert len(lines) == 2 + len(table)


def test_explorer_warns_on_small_query_budget():
    from flexs_b200.baselines.explorers import Adalead

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        Adalead(HashModel(), 1, 10, 5, "ATCATCAT", "ATCG")
    assert any("model_queries_per_batch" in str(x.message) for x in w)


def test_surrogate_constructors_and_names():
    cnn = flexs.baselines.models.CNN(seq_len=8, num_filters=32, hidden_size=100, alphabet=su.DNAA)
    assert cnn.name == "CNN_hidden_size_100_num_filters_32" and cnn.batch_size == 256 and cnn.epochs == 20
    assert [tuple(s) for s in cnn._weight_shapes()] == [(5, 4, 32), (32,), (5, 32, 32), (32,), (3, 32, 32), (32,),
                                                         (32, 100), (100,), (100, 100), (100,), (100, 1), (1,)]
    mlp = flexs.baselines.models.MLP(seq_len=8, hidden_size=100, alphabet=su.DNAA)
    assert mlp.name == "MLP_hidden_size_100"
    assert sum(int(np.prod(s)) for s in mlp._weight_shapes()) == 23601
    assert sum(int(np.prod(s)) for s in cnn._weight_shapes()) == 22429
    ens = flexs.Ensemble([flexs.baselines.models.CNN(14, 32, 100, su.RNAA) for _ in range(3)])
    assert ens._fusable()
    assert not flexs.Ensemble([cnn, mlp])._fusable()


def test_no_silent_cpu_fallback():
    """Without a CUDA device the surrogate must raise, never compute on the CPU."""
    from flexs_b200 import _native

    if _native.device_count() > 0:
        pytest.skip("CUDA device present")
    cnn = flexs.baselines.models.CNN(seq_len=8, num_filters=32, hidden_size=100, alphabet=su.DNAA)
    with pytest.raises(_native.NativeError):
        cnn.get_fitness(["TGCATGCA"])


def test_product_never_imports_oracle():
    import pathlib

    root = pathlib.Path(flexs.__file__).parent
    for path in root.rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), path
    for path in list(root.rglob("*.cu")) + list(root.rglob("*.cuh")):
        assert "oracle" not in path.read_text().lower(), path


# ---------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol(native_lib):
    import pathlib

    header = (pathlib.Path(flexs.__file__).parent.parent / "include" / "flexs_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(flexs_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    from flexs_b200 import _native

    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(native_lib, name), name
    assert native_lib.flexs_abi_version() == 1
    assert native_lib.flexs_topk_workspace_bytes(1000, 100) > 0
    assert native_lib.flexs_topk_workspace_bytes(1000, 5000) < 0


def test_packstr_extension_matches_pure_python_route():
    """csrc/packstr.c (CPython helper built next to the CUDA library): one C pass over a list of str gives the same
    uint8[N, L] character array and the same ValueErrors as the "".join / encode route it short-cuts."""
    from flexs_b200 import _build
    from flexs_b200.utils import sequence_utils as su

    _build.build_packstr()
    su._PACKSTR = False                                   # look the module up again
    assert su._packstr() is not None
    rng = np.random.default_rng(0)
    letters = np.array(list("ACDEFGHIKLMNPQRSTVWY\xe9"))  # includes a Latin-1 character above 127
    seqs = ["".join(r) for r in letters[rng.integers(0, len(letters), size=(5000, 37))]]
    fast = su.sequences_to_char_array(seqs)
    su._PACKSTR = None                                    # force the pure-Python route
    try:
        slow = su.sequences_to_char_array(seqs)
        for bad in (["ACG", "AC"], ["AC", "ACG"], ["AĀC", "ACG"]):
            with pytest.raises(ValueError) as e_slow:
                su.sequences_to_char_array(bad)
            su._PACKSTR = False
            with pytest.raises(ValueError) as e_fast:
                su.sequences_to_char_array(bad)
            su._PACKSTR = None
            assert str(e_fast.value).split(":")[0] == str(e_slow.value).split(":")[0]
    finally:
        su._PACKSTR = False
    assert fast.dtype == np.uint8 and fast.shape == (5000, 37) and np.array_equal(fast, slow)
    assert su.sequences_to_char_array(tuple(seqs[:3])).tolist() == slow[:3].tolist()
    assert su.sequences_to_char_array(seqs[:4], seq_len=37).shape == (4, 37)
    with pytest.raises(ValueError):
        su.sequences_to_char_array(seqs[:4], seq_len=36)


def test_packed_wire_format_host_packers_agree():
    """The two host packers of the wire format of include/flexs_b200.h (numpy on index arrays, the multi-threaded C pass
    over str objects) produce identical rows; unpack inverts them; error behaviour follows str.index."""
    import numpy as np
    import pytest

    from flexs_b200.utils import sequence_utils as su

    rng = np.random.default_rng(0)
    for alphabet, L, n in ((su.DNAA, 100, 3000), (su.AAS, 237, 1200), (su.DNAA, 8, 50), (su.AAS, 735, 400), ("AB", 9, 10), ("ABCDEFG", 11, 77)):
        A = len(alphabet)
        idx = rng.integers(0, A, size=(n, L), dtype=np.uint8)
        seqs = [str(s) for s in su.decode_indices(idx, alphabet)]
        packed = su.pack_indices(idx, A)
        bits = su.bits_per_residue(A)
        assert packed.dtype == np.uint8 and packed.shape == (n, (L * bits + 7) // 8)
        np.testing.assert_array_equal(su.pack_sequences(seqs, alphabet), packed)
        np.testing.assert_array_equal(su.pack_sequences(np.array(seqs), alphabet), packed)   # numpy route
        np.testing.assert_array_equal(su.unpack_indices(packed, L, A), idx)
        # definition: residue i occupies bits [i*b, (i+1)*b) of the row's little-endian bit stream
        i = L - 1
        stream = int.from_bytes(packed[0].tobytes(), "little")
        assert (stream >> (i * bits)) & ((1 << bits) - 1) == idx[0, i]
    big = [su.DNAA[i % 4] * 100 for i in range(5000)]      # large enough for the threaded path
    with pytest.raises(ValueError, match="sequence 4999"):
        su.pack_sequences(big[:-1] + ["A" * 99 + "N"], su.DNAA)
    with pytest.raises(ValueError, match="same length"):
        su.pack_sequences(big[:-1] + ["A" * 99], su.DNAA)
    with pytest.raises(TypeError):
        su.pack_sequences(big[:-1] + [7], su.DNAA)
    with pytest.raises(ValueError):
        su.pack_indices(np.full((2, 5), 4, dtype=np.uint8), 4)


def test_virtual_screen_argument_validation():
    """screen.py: the exchange mode is validated before any GPU work; a model without a device path is refused."""
    from flexs_b200.screen import VirtualScreen

    class _Dev:
        def get_fitness_device(self, idx):
            raise AssertionError("not reached")

    with pytest.raises(ValueError):
        VirtualScreen(_Dev(), k=9, exchange="carrier-pigeon")
    with pytest.raises(TypeError):
        VirtualScreen(object(), k=9)
    vs = VirtualScreen(_Dev(), k=9, exchange="peer", overlap=False)
    assert vs.exchange == "peer" and vs.PEER_DEPTH == 4 and vs.launches == 0
    vs.wait()   # nothing pending, no stream: a no-op without a GPU


def test_peer_mailbox_layout_arithmetic():
    """csrc/peer.cu: size of a rank's mailbox (pure host arithmetic, no GPU): data [depth][world][msg] rounded to 128 bytes,
    then depth * world flags; bad arguments are refused."""
    from flexs_b200 import _native
    from flexs_b200.screen import message_bytes

    mb = message_bytes(99, 100)
    assert mb % 16 == 0 and mb == _native.screen_message_bytes(99, 100)
    for world in (1, 2, 8):
        for depth in (1, 4):
            total = _native.peer_mailbox_bytes(mb, world, depth)
            data = -(-depth * world * mb // 128) * 128
            assert total >= data + depth * world * 4 and total - data - depth * world * 4 <= 128
    for bad in ((0, 2, 4), (24, 2, 4), (mb, 0, 4), (mb, 2, 0)):
        with pytest.raises(ValueError):
            _native.peer_mailbox_bytes(*bad)
