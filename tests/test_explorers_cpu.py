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


def test_cbas_and_dbas_run_like_reference_smoke_test():
    """tests/test_explorers.py:115-128 of the reference: 2-epoch VAE, rounds=3, B=5, Q=20."""
    from flexs_b200.baselines.explorers import VAE, CbAS

    for algo in ("cbas", "dbas"):
        vae = VAE(len(START), alphabet=ALPHABET, epochs=2, verbose=False)
        model = FakeModel()
        ex = CbAS(model, vae, rounds=3, starting_sequence=START, sequences_batch_size=5, model_queries_per_batch=20,
                  alphabet=ALPHABET, algo=algo, cycle_batch_size=10)
        assert ex.name == f"{algo}_Q=0.7_generator=VAE_latent_dim=2_intermediate_dim=250"
        table, _ = ex.run(FakeLandscape("l"), verbose=False)
        assert set(table["round"]) == {0, 1, 2, 3}
        assert len(table[table["round"] == 1]) == 5          # round 1: random neighbourhood, B sequences
        for r in (2, 3):
            rows = table[table["round"] == r]
            assert 1 <= len(rows) <= 4                        # B-1 (cbas_dbas.py:199)
            assert rows["model_cost"].iloc[0] - table[table["round"] == r - 1]["model_cost"].iloc[0] == 20
    with pytest.raises(ValueError):
        CbAS(FakeModel(), vae, 1, START, 5, 20, ALPHABET, algo="nope")


def test_vae_pieces():
    from flexs_b200.utils import VAE_utils

    w = VAE_utils.pwm_to_boltzmann_weights(np.array([[0.1, 0.9], [0.9, 0.1], [0.5, 0.5]]), 0.5)
    np.testing.assert_allclose(w.sum(axis=0), 1.0)
    vae = VAE_utils.VAE(6, "ATCG", epochs=1, verbose=False)
    seqs = ["ATCGAT", "TTTTTT", "ACACAC", "GGGGGG", "ATATAT", "CGCGCG", "AAAAAA", "TGTGTG", "CACACA", "GTGTGT"]
    vae.train_model(seqs, np.ones(len(seqs)))
    lp = vae.calculate_log_probability(seqs)
    assert lp.shape == (10,) and np.all(lp <= 0) and np.all(np.isfinite(lp))
    out = vae.generate(7, seqs, np.ones(len(seqs)))
    assert len(out) == len(set(out)) == 7 and not set(out) & set(seqs) and all(len(s) == 6 for s in out)
    clone = VAE_utils.VAE(6, "ATCG", epochs=1, verbose=False)
    clone.vae.set_weights(vae.vae.get_weights())
    for a, b in zip(clone.vae.get_weights(), vae.vae.get_weights()):
        np.testing.assert_array_equal(a, b)


def test_dynappo_runs_and_budgets():
    """tests/test_explorers.py:86-97 of the reference (DynaPPO with a fake model)."""
    from flexs_b200.baselines.explorers import DynaPPO
    from flexs_b200.baselines.explorers.dyna_ppo import bounded_edit_distance

    assert bounded_edit_distance("ATCG", "ATCG", 2) == 0
    assert bounded_edit_distance("ATCG", "ATGG", 2) == 1
    assert bounded_edit_distance("ATCG", "TCGA", 2) == 2
    assert bounded_edit_distance("ATCG", "GCTA", 2) == 3   # capped at radius + 1
    assert bounded_edit_distance("ATCGAAA", "ATCG", 2) == 3
    model, landscape = FakeModel(), FakeLandscape("l")
    ex = DynaPPO(landscape=landscape, model=model, rounds=3, sequences_batch_size=5, model_queries_per_batch=20,
                 starting_sequence=START, alphabet=ALPHABET)
    assert ex.name == "DynaPPO_Agent_10_1"
    table, _ = ex.run(landscape, verbose=False)
    assert set(table["round"]) <= {0, 1, 2, 3}
    for r in (1, 2, 3):
        rows = table[table["round"] == r]
        assert len(rows) <= 5
        assert all(s.endswith(ALPHABET[0]) for s in rows["sequence"])   # last position is never sampled (quirk 5)
    assert all(c == 4 for c in model.calls)                                # one env_batch_size call per episode end
    assert model.cost == 3 * 20                                            # 5 episodes x 4 per round


def test_noisy_abstract_model_and_evaluate_drivers():
    """tests/test_models.py:80-99 of the reference + the three flexs.evaluate sweeps with fakes."""
    from flexs_b200.baselines.models import NoisyAbstractModel
    from flexs_b200.baselines.models.noisy_abstract_model import edit_distance

    assert edit_distance("kitten", "sitting") == 3 and edit_distance("", "abc") == 3 and edit_distance("ab", "ab") == 0

    class Additive(flexs.Landscape):
        def _fitness_function(self, sequences):
            return np.array([sum(ch == "A" for ch in s) / len(s) for s in sequences])

    land = Additive("additive")
    nam = NoisyAbstractModel(land, signal_strength=1.0)
    np.testing.assert_allclose(nam.get_fitness(["AATT", "TTTT"]), [0.5, 0.0])   # ss = 1: exact ground truth
    nam = NoisyAbstractModel(land, signal_strength=0.5)
    first = nam.get_fitness(["AATT", "ATAT"])
    np.testing.assert_array_equal(nam.get_fitness(["AATT", "ATAT"]), first)       # cached -> deterministic
    nam.train(["CCCC"], [7.0])
    assert nam.get_fitness(["CCCC"])[0] == 7.0

    def mk(model, ss):
        return Adalead(model, rounds=2, sequences_batch_size=4, model_queries_per_batch=12, starting_sequence="ATCATCAT",
                       alphabet="ATCG", eval_batch_size=1)

    res = flexs.evaluate.robustness(land, mk, signal_strengths=[0, 1], verbose=False)
    assert [r[0] for r in res] == [0, 1] and all(len(r[1][0]) > 1 for r in res)
    res = flexs.evaluate.efficiency(land, lambda b, q: Adalead(FakeModel(), 1, b, q, "ATCATCAT", "ATCG", eval_batch_size=1),
                                    budgets=[(3, 9), (4, 12)])
    assert [r[0] for r in res] == [(3, 9), (4, 12)]
    res = flexs.evaluate.adaptivity(land, lambda r, b, q: Adalead(FakeModel(), r, b, q, "ATCATCAT", "ATCG", eval_batch_size=1),
                                    num_rounds=[1, 2], total_ground_truth_measurements=8, total_model_queries=24)
    assert [r[0] for r in res] == [1, 2]


def test_device_sep_cma_matches_host_sampler_update():
    """flexs_b200.utils.cma_device.SepCMA (torch tensors; here on the CPU) applies the same separable CMA update as the
    host sampler flexs_b200.utils.cma given the same samples and values."""
    import numpy as np
    import torch

    from flexs_b200.utils import cma, cma_device

    n, pop = 600, 24
    x0 = np.zeros(n); x0[::3] = 1
    host = cma.CMAEvolutionStrategy(x0, 0.4, {"popsize": pop, "seed": 1}, full_cov_max_dim=16)
    dev = cma_device.SepCMA(torch.from_numpy(x0), 0.4, pop, seed=1)
    assert host.separable and dev.mu == host.mu
    for _ in range(6):
        X = np.array(host.ask())
        f = ((X - 0.3) ** 2).sum(axis=1)
        host.tell(list(X), list(f))
        dev.tell(torch.from_numpy(X).float(), torch.from_numpy(f))
        assert np.abs(host.mean - dev.mean.numpy()).max() < 1e-5
        assert abs(host.sigma - dev.sigma) / host.sigma < 1e-6
        assert np.abs(host.diagC - dev.diagC.numpy()).max() < 1e-5
    samples = dev.ask()
    assert samples.shape == (pop, n) and samples.dtype == torch.float32 and torch.isfinite(samples).all()


def test_cbas_dbas_decisions_match_reference_code(golden, monkeypatch):
    """SURVEY.md §8 row a9, value parity: this repo's CbAS.propose_sequences and the REFERENCE's own (run by
    tests/golden/make_golden_cbas.py) drive the same deterministic fake generator and hash model from the same seed and
    must make identical decisions — the threshold gamma, the importance weights, which proposals are masked, the pool the
    generator is re-fit on (every train_model call is logged), the final ranking and model.cost."""
    import random

    import numpy as np
    import pandas as pd

    import flexs_b200 as flexs
    from flexs_b200.baselines.explorers import cbas_dbas
    from flexs_b200.utils import sequence_utils as su
    from tests.golden.fake_vae import FakeVAE
    from tests.golden.make_golden import hash_model_score

    class HashModel(flexs.Model):
        def __init__(self):
            super().__init__("hash")

        def train(self, *a, **k):
            pass

        def _fitness_function(self, sequences):
            return np.array([hash_model_score(s) for s in sequences])

    monkeypatch.setattr(cbas_dbas, "VAE", FakeVAE)     # the prior is cloned by constructing a VAE (cbas_dbas.py:125-144)
    ref = golden("ref_cbas.json")
    rng = np.random.default_rng(7)
    seqs = ["".join(su.RNAA[i] for i in row) for row in rng.integers(0, 4, size=(60, 12))]
    frame = pd.DataFrame({"sequence": seqs, "true_score": [hash_model_score(s[::-1]) for s in seqs], "model_score": np.nan,
                          "round": [1] * 60, "model_cost": 0, "measurement_cost": 60})
    for algo in ("cbas", "dbas"):
        random.seed(11)
        model = HashModel()
        gen = FakeVAE(seq_length=12, alphabet=su.RNAA)
        ex = flexs.baselines.explorers.CbAS(model, gen, rounds=1, starting_sequence="AUGCAUGCAUGC", sequences_batch_size=20,
                                           model_queries_per_batch=350, alphabet=su.RNAA, algo=algo, Q=0.7, cycle_batch_size=100)
        got_seqs, got_preds = ex.propose_sequences(frame)
        want = ref[algo]
        assert [str(s) for s in got_seqs] == want["sequences"]
        np.testing.assert_array_equal(np.asarray(got_preds, dtype=np.float64), np.asarray(want["preds"]))
        assert model.cost == want["model_cost"] == 400
        assert gen.train_log == want["train_log"]
        assert len(got_seqs) == 19                                  # the [: -B : -1] slice keeps B-1
