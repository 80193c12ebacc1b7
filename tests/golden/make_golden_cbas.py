"""Generates tests/golden/ref_cbas.json by running the REFERENCE's CbAS.propose_sequences (flexs/baselines/explorers/
cbas_dbas.py, imported from /root/reference with TensorFlow-free stubs) with the deterministic fake generator of
tests/golden/fake_vae.py and the hash model of make_golden.py.  Run in the authoring container only:

    python tests/golden/make_golden_cbas.py

Pins SURVEY.md §8 row a9: the percentile threshold gamma (:163), the importance weights exp(log p0 - log pt) with
nan_to_num (:170-174), the masking of proposals below gamma (:181), the growing sample pool the generator is re-fit on, the
[: -B : -1] ranking (:199) and model.cost, for both algo="cbas" and algo="dbas"."""
import json
import random
import sys
import types
from pathlib import Path

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from tests.golden import make_golden as mg  # noqa: E402
from tests.golden.fake_vae import FakeVAE  # noqa: E402


def measured_frame(su, alphabet, length):
    rng = np.random.default_rng(7)
    seqs = ["".join(alphabet[i] for i in row) for row in rng.integers(0, len(alphabet), size=(60, length))]
    return pd.DataFrame({"sequence": seqs, "true_score": [mg.hash_model_score(s[::-1]) for s in seqs],
                         "model_score": np.nan, "round": [1] * 60, "model_cost": 0, "measurement_cost": 60})


def main():
    flexs, _ = mg.reference_namespace()
    su = flexs.utils.sequence_utils
    vae_mod = types.ModuleType("flexs.utils.VAE_utils")
    vae_mod.VAE = FakeVAE
    sys.modules["flexs.utils.VAE_utils"] = vae_mod
    flexs.utils.VAE_utils = vae_mod
    cbas_mod = mg._load("flexs.baselines.explorers.cbas_dbas", mg.REF / "flexs/baselines/explorers/cbas_dbas.py")

    class HashModel(flexs.Model):
        def __init__(self):
            super().__init__("hash")

        def train(self, *a, **k):
            pass

        def _fitness_function(self, sequences):
            return np.array([mg.hash_model_score(s) for s in sequences])

    out = {}
    for algo in ("cbas", "dbas"):
        random.seed(11)
        model = HashModel()
        gen = FakeVAE(seq_length=12, alphabet=su.RNAA)
        ex = cbas_mod.CbAS(model, gen, rounds=1, starting_sequence="AUGCAUGCAUGC", sequences_batch_size=20,
                           model_queries_per_batch=350, alphabet=su.RNAA, algo=algo, Q=0.7, cycle_batch_size=100)
        seqs, preds = ex.propose_sequences(measured_frame(su, su.RNAA, 12))
        out[algo] = {"sequences": [str(s) for s in seqs], "preds": [float(p) for p in preds], "model_cost": int(model.cost),
                     "train_log": gen.train_log}
    (HERE / "ref_cbas.json").write_text(json.dumps(out, indent=1))
    print({k: (len(v["sequences"]), v["model_cost"], len(v["train_log"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
