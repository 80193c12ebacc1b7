"""Generates tests/golden/*.npz / *.json.  Run in the AUTHORING container only:

    python tests/golden/make_golden.py

Two kinds of fixtures:

* ``ref_*``  — produced by importing the reference's own pure-Python modules from /root/reference
  (sequence_utils, Landscape/Model/Explorer, Adalead) with a hand-built ``flexs`` namespace so that
  TensorFlow / tf-agents / cma are never imported.  These PIN the integer/string semantics:
  encode, decode, mutation RNG call order, the ``[: -B : -1]`` slice, Adalead's proposal set and
  cost accounting under a seeded RNG and a deterministic fake model.
* ``oracle_*`` — produced by oracle/flexs_oracle.py (float64 definition).  The reference cannot
  produce these here (TensorFlow absent) so they are regression vectors for the restatement, not
  reference outputs: floating-point parity stays "unpinned" (see oracle header).

/root/reference does not exist on the GPU box; tests read only the committed files.
"""
import importlib.util
import json
import random
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_namespace():
    """A ``flexs`` module object holding only the pure-Python reference pieces."""
    flexs = types.ModuleType("flexs")
    flexs.__path__ = []  # mark as package
    sys.modules["flexs"] = flexs
    flexs.types = _load("flexs.types", REF / "flexs/types.py")
    landscape = _load("flexs.landscape", REF / "flexs/landscape.py")
    flexs.Landscape = landscape.Landscape
    model = _load("flexs.model", REF / "flexs/model.py")
    flexs.Model, flexs.LandscapeAsModel = model.Model, model.LandscapeAsModel
    utils = types.ModuleType("flexs.utils")
    utils.__path__ = []
    sys.modules["flexs.utils"] = utils
    flexs.utils = utils
    utils.sequence_utils = _load("flexs.utils.sequence_utils", REF / "flexs/utils/sequence_utils.py")
    ensemble = _load("flexs.ensemble", REF / "flexs/ensemble.py")
    flexs.Ensemble = ensemble.Ensemble
    explorer = _load("flexs.explorer", REF / "flexs/explorer.py")
    flexs.Explorer = explorer.Explorer
    adalead = _load("flexs.baselines.explorers.adalead", REF / "flexs/baselines/explorers/adalead.py")
    return flexs, adalead


def hash_model_score(seq: str) -> float:
    """Deterministic, tie-free fake fitness: a fixed polynomial hash of the string in [0, 1)."""
    h = 1469598103934665603
    for ch in seq:
        h = ((h ^ ord(ch)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h >> 11) / float(1 << 53)


def main():
    flexs, adalead_mod = reference_namespace()
    su = flexs.utils.sequence_utils

    # ---- encode / decode ------------------------------------------------------------------
    rng = np.random.default_rng(20240925)
    enc = {}
    for name, alphabet, length, n in [("dna8", su.DNAA, 8, 64), ("rna14", su.RNAA, 14, 48), ("aa90", su.AAS, 90, 24),
                                      ("aa237", su.AAS, 237, 6), ("bin5", su.BA, 5, 16), ("dna100", su.DNAA, 100, 32)]:
        seqs = ["".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=length)) for _ in range(n)]
        onehots = np.array([su.string_to_one_hot(s, alphabet) for s in seqs])
        assert onehots.dtype == np.float64
        idx = onehots.argmax(axis=2).astype(np.uint8)
        back = [su.one_hot_to_string(oh, alphabet) for oh in onehots]
        assert back == seqs
        enc[name] = {"alphabet": alphabet, "seqs": seqs, "idx": idx.tolist()}
    # decode of non-one-hot float matrices (CMA-ES solutions): first maximum wins
    dec = []
    for a, length in [(4, 8), (20, 11)]:
        x = rng.normal(size=(12, length, a)).round(1)  # rounding creates ties
        alphabet = su.DNAA if a == 4 else su.AAS
        dec.append({"alphabet": alphabet, "x": x.tolist(), "strings": [su.one_hot_to_string(m, alphabet) for m in x]})
    # error behaviour
    try:
        su.string_to_one_hot("ATXG", "ATCG")
        raised = None
    except Exception as e:  # noqa: BLE001
        raised = type(e).__name__
    json.dump({"encode": enc, "decode": dec, "bad_char_exception": raised}, open(HERE / "ref_encode_decode.json", "w"))

    # ---- mutation helpers under a seeded `random` ---------------------------------------------
    mut = []
    for seed, seq, mu, alphabet in [(1, "ATCATCAT", 1 / 8, "ATCG"), (2, "UGCAUGCAUGCAUG", 0.2, su.RNAA),
                                    (3, su.AAS * 4, 0.05, su.AAS), (4, "TTTTTTTT", 1.0, su.DNAA), (5, "ACGT", 0.0, "ACGT")]:
        random.seed(seed)
        outs = [su.generate_random_mutant(seq, mu, alphabet) for _ in range(20)]
        mut.append({"seed": seed, "seq": seq, "mu": mu, "alphabet": alphabet, "mutants": outs})
    random.seed(11)
    rand_seqs = su.generate_random_sequences(9, 7, su.AAS)
    singles = su.generate_single_mutants("ATC", "ATCG")
    json.dump({"mutants": mut, "random_sequences": {"seed": 11, "length": 9, "number": 7, "alphabet": su.AAS,
                                                      "out": rand_seqs},
               "single_mutants": {"wt": "ATC", "alphabet": "ATCG", "out": singles}},
              open(HERE / "ref_mutation.json", "w"))

    # ---- ranking slices (tie-free inputs) ------------------------------------------------------
    preds = rng.permutation(500).astype(np.float32) / 7.0 - 20.0
    slices = {"preds": preds.tolist()}
    for b in (1, 2, 5, 100, 499, 500, 600):
        slices[f"bm1_{b}"] = np.argsort(preds)[: -b: -1].tolist()   # adalead.py:173 et al.
        slices[f"b_{b}"] = np.argsort(preds)[::-1][:b].tolist()      # dyna_ppo.py:317
    json.dump(slices, open(HERE / "ref_topk_slices.json", "w"))

    # ---- Adalead.propose_sequences: proposal set + cost under a seeded RNG --------------------
    import pandas as pd

    class HashModel(flexs.Model):
        def __init__(self):
            super().__init__("hash")

        def _fitness_function(self, sequences):
            return np.array([hash_model_score(s) for s in sequences])

        def train(self, *a, **k):
            pass

    runs = []
    for seed, start, alphabet, batch, queries, ebs, rho, recomb in [
        (0, "ATCATCAT", "ATCG", 5, 20, 1, 0, 0.0),
        (1, "ATCATCAT", "ATCG", 10, 200, 20, 0, 0.0),
        (2, "UGCAUGCAUGCAUG", "UGCA", 100, 2000, 20, 0, 0.0),
        (3, "ATCATCATGG", "ATCG", 8, 120, 4, 1, 0.2),
    ]:
        random.seed(seed)
        np.random.seed(seed)
        model = HashModel()
        ex = adalead_mod.Adalead(model, rounds=1, sequences_batch_size=batch, model_queries_per_batch=queries,
                                 starting_sequence=start, alphabet=alphabet, eval_batch_size=ebs, rho=rho,
                                 recomb_rate=recomb)
        # measured data: the start plus a few of its single mutants, true scores from the same hash
        measured = [start] + su.generate_single_mutants(start, alphabet)[1:7]
        measured = list(dict.fromkeys(measured))
        df = pd.DataFrame({"sequence": measured, "true_score": [hash_model_score(s) for s in measured],
                           "model_score": np.nan, "round": 0})
        seqs, preds_ = ex.propose_sequences(df)
        runs.append({"seed": seed, "start": start, "alphabet": alphabet, "batch": batch, "queries": queries,
                     "eval_batch_size": ebs, "rho": rho, "recomb_rate": recomb, "measured": measured,
                     "proposed": list(map(str, seqs)), "preds": [float(p) for p in preds_], "model_cost": int(model.cost)})
    json.dump({"runs": runs}, open(HERE / "ref_adalead.json", "w"))

    # ---- Ensemble + cost accounting -----------------------------------------------------------
    class Const(flexs.Model):
        def __init__(self, c):
            super().__init__(f"c{c}")
            self.c = c

        def _fitness_function(self, sequences):
            return np.full(len(sequences), self.c, dtype=np.float32)

        def train(self, *a, **k):
            pass

    members = [Const(0.1), Const(0.7), Const(0.25)]
    ens = flexs.Ensemble(members)
    out = ens.get_fitness(["AAA", "CCC"])
    json.dump({"name": ens.name, "out": [float(v) for v in out], "out_dtype": str(out.dtype), "ens_cost": ens.cost,
               "member_costs": [m.cost for m in members]}, open(HERE / "ref_ensemble.json", "w"))

    # ---- oracle regression vectors (float64 definition) --------------------------------------
    from oracle import flexs_oracle as fo

    vec = {}
    for tag, (L, A, F, H, K, n) in {"test_shape": (3, 4, 1, 1, 2, 16), "tf8": (8, 4, 32, 100, 5, 32),
                                    "rna14": (14, 4, 32, 100, 5, 32), "ns100": (100, 4, 32, 100, 5, 16),
                                    "aav90": (90, 20, 32, 100, 5, 8), "gfp237": (237, 20, 32, 100, 5, 4),
                                    "gfp238": (238, 20, 32, 100, 5, 4), "aav735": (735, 20, 32, 100, 5, 2)}.items():
        shp = fo.CNNShape(L, A, F, H, K)
        idx = np.random.default_rng(1000 + L).integers(0, A, size=(n, L), dtype=np.uint8)
        for wname, wseed, fn in (("glorot", 0, fo.glorot_weights), ("trained", 7, fo.trained_like_weights)):
            ws = fn(shp.weight_shapes(), wseed)
            vec[f"{tag}_{wname}_idx"] = idx
            vec[f"{tag}_{wname}_y"] = fo.cnn_forward(idx, ws, np.float64)
            vec[f"{tag}_{wname}_cfg"] = np.array([L, A, F, H, K, wseed], dtype=np.int64)
    ms = fo.MLPShape(8, 4, 100)
    idx = np.random.default_rng(77).integers(0, 4, size=(32, 8), dtype=np.uint8)
    ws = fo.trained_like_weights(ms.weight_shapes(), 5)
    vec["mlp8_idx"], vec["mlp8_y"] = idx, fo.mlp_forward(idx, ws, np.float64)
    np.savez_compressed(HERE / "oracle_forward.npz", **vec)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
