"""CPU: explorer logic with fake models (the pattern of the reference's tests/test_explorers.py:7-47,
100-112 — rounds=3, sequences_batch_size=5, model_queries_per_batch=20, start "ATCATCAT")."""
import numpy as np
import pandas as pd
import pytest

import flexs_b200 as flexs
from flexs_b200.baselines.explorers import CMAES, Adalead
from flexs_b200.utils import cma

rng = np.random.default_rng(0)


class FakeModel(flexs.Model):
    def __init__(self):
        super().__init__("FakeModel")
        self.calls = []

    def _fitness_function(self, sequences):
        self.calls.append(len(sequences))
        return rng.random(size=len(sequences))

    def train(self, *a, **k):
        pass


class FakeLandscape(flexs.Landscape):
    def _fitness_function(self, sequences):
        return rng.random(size=len(sequences))


START, ALPHABET = "ATCATCAT", "ATCG"


def test_adalead_runs_like_reference_smoke_test():
    ex = Adalead(model=FakeModel(), rounds=3, sequences_batch_size=5, model_queries_per_batch=20,
                 starting_sequence=START, alphabet=ALPHABET, eval_batch_size=1)
    table, meta = ex.run(FakeLandscape("l"), verbose=False)
    assert table["round"].max() == 3


def test_cmaes_runs_and_accounts_cost():
    model = FakeModel()
    ex = CMAES(model, population_size=15, max_iter=200, rounds=3, sequences_batch_size=5,
               model_queries_per_batch=20, starting_sequence=START, alphabet=ALPHABET, seed=1)
    assert ex.name == "CMAES_popsize15"
    table, meta = ex.run(FakeLandscape("l"), verbose=False)
    assert set(table["round"]) == {0, 1, 2, 3}
    # one batched call per CMA iteration; never more than the per-round query budget
    for r in (1, 2, 3):
        prev = table[table["round"] == r - 1]["model_cost"].iloc[0]
        cur = table[table["round"] == r]["model_cost"].iloc[0]
        assert 0 < cur - prev <= 20
        assert len(table[table["round"] == r]) <= 4  # B-1 (cmaes.py:120)
    assert all(c <= 15 for c in model.calls)


def test_cmaes_decode_first_max_and_cache(golden):
    ex = CMAES(FakeModel(), 1, 5, 20, START, ALPHABET)
    for case in golden("ref_encode_decode.json")["decode"]:
        if len(case["alphabet"]) != 4:
            continue
        ex.alphabet, ex.starting_sequence = case["alphabet"], "A" * len(case["x"][0])
        for x, want in zip(case["x"], case["strings"]):
            assert ex._soln_to_string(np.array(x).reshape(-1)) == want
        assert ex._decode_population(np.array(case["x"]).reshape(len(case["x"]), -1)) == case["strings"]


def test_cmaes_uses_measured_cache_without_charging():
    class Const(FakeModel):
        def _fitness_function(self, sequences):
            self.calls.append(list(sequences))
            return np.full(len(sequences), 0.5)

    model = Const()
    ex = CMAES(model, 1, 5, 40, "AAAA", "AT", population_size=8, initial_variance=1e-12, seed=0)
    df = pd.DataFrame({"sequence": ["AAAA"], "true_score": [1.0], "model_score": np.nan, "round": 0})
    seqs, preds = ex.propose_sequences(df)
    # with a vanishing variance every sample decodes to the start sequence, which is answered from the cache
    assert model.cost == 0 and model.calls == []
    assert list(seqs) == ["AAAA"] and preds[0] == 1.0  # the start sequence itself is among the "seen" (cmaes.py:78-80)


def test_cma_sampler_minimises_sphere_full_and_separable():
    for dim, full in ((12, 512), (40, 8)):
        es = cma.CMAEvolutionStrategy(np.full(dim, 3.0), 1.0, {"popsize": 16, "seed": 3}, full_cov_max_dim=full)
        assert es.separable == (dim > full)
        f0 = None
        for _ in range(250):
            xs = es.ask()
            fs = [float(np.sum(x * x)) for x in xs]
            f0 = f0 or min(fs)
            es.tell(xs, fs)
        assert min(fs) < 1e-3 * f0
    xs, fs = es.ask_and_eval(lambda x: float(np.sum(x * x)))
    assert len(xs) == len(fs) == 16
