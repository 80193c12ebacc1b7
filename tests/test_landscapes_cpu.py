"""CPU: table landscapes — oracle restatement and host-side table construction against outputs of the reference's own
classes (tests/golden/ref_landscapes.json, made by tests/golden/make_golden_landscapes.py)."""
import json
import os
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

import flexs_b200 as flexs
from oracle import landscapes as ol

GOLD = Path(__file__).parent / "golden"
AAV_FILE = str(GOLD / "aav_450_540_subs.json")
TF_FILE = str(GOLD / "tfbind_4mers.txt")


@pytest.fixture(scope="module")
def ref():
    return json.load(open(GOLD / "ref_landscapes.json"))


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def emulate_additive(land, seqs, noise):
    """What the kernel computes, in numpy: left-to-right float64 sum of table columns (absent pair -> +0.0)."""
    out = []
    for s, eps in zip(seqs, noise):
        total = np.float64(0.0)
        for i, ch in enumerate(s):
            c = land.column_of_char[ord(ch)]
            if c != 0xFF:
                total = total + land.table[i, c]
        f = (total + land._offset) / land._denom + eps
        out.append(f if f > 0 else 0.0)
    return np.array(out)


def test_oracle_additive_matches_reference_outputs(ref):
    window = {int(p): v for p, v in json.load(open(AAV_FILE)).items()}
    for case in ref["aav"]:
        ph = f"log2_{case['phenotype']}_v_wt"
        top, mx = ol.compute_max_possible(window, ph)
        assert top == case["top_seq"] and float(mx).hex() == case["max_possible"]
        np.random.seed(case["seed"])
        draws = np.random.normal(scale=case["noise"], size=len(case["sequences"]))
        assert float(np.random.random()).hex() == case["next_uniform"]   # vector draw == the reference's scalar draws
        out = ol.additive_fitness(case["sequences"], window, ph, 450, case["mfm"], mx, draws)
        np.testing.assert_array_equal(out, unhex(case["fitness"]))
        assert (out > 0).sum() > 20 and (out == 0).sum() > 5               # both branches of max(0, .) are exercised


def test_additive_host_tables_match_reference(ref):
    for case in ref["aav"]:
        land = flexs.landscapes.AdditiveAAVPackaging(phenotype=case["phenotype"], minimum_fitness_multiplier=case["mfm"],
                                                     start=450, end=540, noise=case["noise"], data_file=AAV_FILE)
        assert land.name == case["name"] and land.wild_type == case["wild_type"] and land.cost == 0
        assert land.top_seq == case["top_seq"] and float(land.max_possible).hex() == case["max_possible"]
        assert land.seq_len == 90 and land.table.shape == (90, len(land.residues)) and "*" in land.residues
        np.random.seed(case["seed"])
        draws = np.random.normal(scale=case["noise"], size=len(case["sequences"]))
        seqs = [s + "\0" * (90 - len(s)) for s in case["sequences"]]
        np.testing.assert_array_equal(emulate_additive(land, seqs, draws), unhex(case["fitness"]))
    assert flexs.landscapes.additive_aav_packaging.registry() == ref["aav_registry"]
    assert flexs.landscapes.additive_aav_packaging.AAV2_WT[450:540] == ref["aav"][0]["wild_type"]


def test_oracle_and_host_tfbinding_match_reference(ref):
    data = pd.read_csv(TF_FILE, sep="\t")
    d = ol.tfbinding_dict(data["8-mer"], data["8-mer.1"], data["E-score"])
    want = {k: float.fromhex(v) for k, v in ref["tf"]["dict"].items()}
    assert d == want
    np.testing.assert_array_equal(ol.tfbinding_fitness(ref["tf"]["query"], d), unhex(ref["tf"]["fitness"]))
    with pytest.raises(KeyError):
        ol.tfbinding_fitness(["ACGN"], d)

    land = flexs.landscapes.TFBinding(TF_FILE)
    assert land.name == ref["tf"]["name"] and land.seq_len == 4 and land.table.shape == (256,)
    assert land.sequences == want
    keys = land._keys(np.array(list(want)))
    np.testing.assert_array_equal(land.table[keys], np.array(list(want.values())))
    assert not np.isnan(land.table).any()


def test_tfbinding_registry_and_real_file(ref, tmp_path, monkeypatch):
    (tmp_path / "tf_binding").mkdir()
    for name in ("AAA_R1_8mers.txt", "BBB_REF_R2_8mers.txt"):
        (tmp_path / "tf_binding" / name).write_text(open(TF_FILE).read())
    monkeypatch.setenv("FLEXS_DATA_DIR", str(tmp_path))
    reg = flexs.landscapes.tf_binding.registry()
    assert set(reg) == {"AAA_R1", "BBB_REF_R2"} and len(reg["AAA_R1"]["starts"]) == 14
    assert reg["AAA_R1"]["params"]["landscape_file"].endswith("AAA_R1_8mers.txt")
    monkeypatch.setenv("FLEXS_DATA_DIR", str(tmp_path / "nope"))
    with pytest.raises(FileNotFoundError):
        flexs.landscapes.table_landscape.data_dir("tf_binding")
    real = "/root/reference/flexs/landscapes/data/tf_binding/" + ref["tf_real"]["file"]
    if not os.path.exists(real):
        pytest.skip("reference data not present (GPU box)")
    land = flexs.landscapes.TFBinding(real)
    assert land.seq_len == 8 and land.table.size == 65536 == ref["tf_real"]["n_keys"] and not np.isnan(land.table).any()
    for seq, v in ref["tf_real"]["probe"].items():
        assert land.table[land._keys(np.array([seq]))[0]] == float.fromhex(v)
    srt = sorted(land.sequences)
    assert float(np.sum([land.sequences[k] for k in srt])).hex() == ref["tf_real"]["sum"]
