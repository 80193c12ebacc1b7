"""GPU: the BASELINE.json configurations as functional runs through the drop-in plugin API (reduced rounds / budgets).

Each test composes the whole hot path the way `flexs.evaluate` drives it — `Explorer.run` -> `model.train` (K4) ->
`propose_sequences` -> `model.get_fitness` (K0 + K1/K2) -> ranking — and then screens a large candidate batch with the
trained surrogate (K1e / K7 + K3b + K3), checking the invariants the reference's own smoke tests check (shapes, cost
accounting, finite scores, B-1 proposals) plus agreement of the screen with `get_fitness` on the same candidates.
Ground-truth landscapes are synthetic stand-ins where the reference needs ViennaRNA / data files.
"""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

AAV_FILE = str(Path(__file__).resolve().parent / "golden" / "aav_450_540_subs.json")

torch = pytest.importorskip("torch")

import flexs_b200 as flexs  # noqa: E402
from flexs_b200 import _native  # noqa: E402
from flexs_b200.screen import VirtualScreen  # noqa: E402
from flexs_b200.utils import sequence_utils as su  # noqa: E402


class MotifLandscape(flexs.Landscape):
    """A cheap ground truth: fraction of positions matching a hidden target plus a pair interaction."""

    def __init__(self, target: str, alphabet: str):
        super().__init__("motif")
        self.target = np.frombuffer(target.encode(), dtype=np.uint8)
        self.alphabet = alphabet

    def _fitness_function(self, sequences):
        chars = su.sequences_to_char_array(list(sequences))
        match = (chars == self.target[None, :]).astype(np.float64)
        return match.mean(axis=1) + 0.25 * match[:, 0] * match[:, -1]


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device; the CUDA path has no CPU fallback")


def _check_run(table, meta, rounds, batch):
    assert table["round"].max() == rounds
    assert np.isfinite(table["true_score"]).all()
    assert np.isfinite(table["model_score"][table["round"] > 0]).all()   # the starting sequence has no model score (NaN)
    per_round = table[table["round"] > 0].groupby("round").size()
    assert (per_round <= batch).all() and (per_round >= 1).all()
    assert table["measurement_cost"].iloc[-1] == len(table)


def test_config1_tfbinding8_cnn_adalead_then_1M_screen():
    """configs[1]: 8-mers over DNA, CNN surrogate, AdaLead; then a 1M-candidate virtual screen (>= 93 % repeats)."""
    rng = np.random.default_rng(0)
    land = MotifLandscape("GCTCGAGC", su.DNAA)
    cnn = flexs.baselines.models.CNN(8, num_filters=32, hidden_size=100, alphabet=su.DNAA, loss="MSE", seed=0)
    ex = flexs.baselines.explorers.Adalead(cnn, rounds=2, sequences_batch_size=50, model_queries_per_batch=400,
                                          starting_sequence="TTTTTTTT", alphabet=su.DNAA, eval_batch_size=20)
    table, meta = ex.run(land, verbose=False)
    _check_run(table, meta, 2, 50)
    assert cnn.cost > 0
    idx = rng.integers(0, 4, size=(1 << 20, 8), dtype=np.uint8)
    assert cnn.native.active_variant(len(idx)) == _native.VARIANT_ENUM      # whole-model table for the 65 536 8-mers
    top_i, top_s = VirtualScreen(cnn, k=99).screen(idx)
    assert len(top_i) == 99 and len({bytes(r) for r in idx[top_i]}) == 99      # distinct winners
    direct = cnn.get_fitness(idx[top_i])
    np.testing.assert_allclose(top_s, direct, rtol=0, atol=1e-4 * max(1e-6, float(np.abs(direct).max())))
    assert (np.diff(top_s) <= 0).all()
    # nothing outside the list beats its last member (sample check through the direct kernels)
    sample = idx[rng.integers(0, len(idx), size=20000)]
    s_scores = cnn.get_fitness(sample)
    winners = {bytes(r) for r in idx[top_i]}
    rest = np.array([sc for r, sc in zip(sample, s_scores) if bytes(r) not in winners])
    assert rest.max() <= top_s[-1] + 1e-4 * float(np.abs(top_s).max())


def test_config2_rna14_ensemble_cbas_then_screen():
    """configs[2]: 14-mers over RNA, Ensemble(3 x CNN), CbAS with the VAE generator; then a 300k screen (cnn_k9)."""
    rng = np.random.default_rng(1)
    land = MotifLandscape("GCUAGCUAGCUAGC", su.RNAA)
    members = [flexs.baselines.models.CNN(14, 32, 100, su.RNAA, loss="MSE", seed=i) for i in range(3)]
    ens = flexs.Ensemble(members)
    start = "AUAUAUAUAUAUAU"
    vae = flexs.baselines.explorers.VAE(seq_length=14, alphabet=su.RNAA, batch_size=10, latent_dim=2,
                                        intermediate_dim=50, epochs=2, verbose=False)
    ex = flexs.baselines.explorers.CbAS(ens, vae, rounds=1, starting_sequence=start, sequences_batch_size=20,
                                       model_queries_per_batch=200, alphabet=su.RNAA, cycle_batch_size=100)
    table, meta = ex.run(land, verbose=False)
    _check_run(table, meta, 1, 20)
    idx = rng.integers(0, 4, size=(300_000, 14), dtype=np.uint8)
    top_i, top_s = VirtualScreen(ens, k=19).screen(idx)
    assert len(top_i) == 19 and (np.diff(top_s) <= 0).all()
    direct = ens.get_fitness(idx[top_i])
    np.testing.assert_allclose(top_s, direct, rtol=0, atol=1e-4 * max(1e-6, float(np.abs(direct).max())))


def test_config3_aav_additive_cnn_cmaes():
    """configs[3] (registry size): the additive AAV landscape on the 90-mer window, CNN surrogate, CMA-ES over the
    relaxed one-hot (dimension 1800 -> separable sampler); the ground truth itself runs on the GPU (K6)."""
    land = flexs.landscapes.AdditiveAAVPackaging(phenotype="heart", start=450, end=540, data_file=AAV_FILE)
    start = land.wild_type
    assert len(start) == 90
    cnn = flexs.baselines.models.CNN(90, num_filters=32, hidden_size=100, alphabet=su.AAS, loss="MSE", seed=3)
    ex = flexs.baselines.explorers.CMAES(cnn, rounds=1, sequences_batch_size=10, model_queries_per_batch=60,
                                        starting_sequence=start, alphabet=su.AAS, population_size=15, max_iter=10, seed=0)
    table, meta = ex.run(land, verbose=False)
    _check_run(table, meta, 1, 10)
    assert cnn.cost <= 60 + 15


def test_config1_adalead_device_rollouts_produce_a_million_model_queries():
    """configs[1] as BASELINE words it — AdaLead ITSELF is the 1M-candidate virtual screen: 65 536 parallel rollouts on
    the GPU (mutate -> "not measured, not found" -> fused forward -> continue while the child is at least as good as its
    root), >= 1e6 model queries in ONE round, cost accounting as the reference's (every scored root and child is
    charged), B-1 distinct unmeasured proposals ranked by the model."""
    import random

    random.seed(0)
    land = MotifLandscape("GCTCGAGC", su.DNAA)
    cnn = flexs.baselines.models.CNN(8, num_filters=32, hidden_size=100, alphabet=su.DNAA, loss="MSE", seed=0)
    rng = np.random.default_rng(0)
    start_seqs = ["TTTTTTTT"] + ["".join(r) for r in np.array(list(su.DNAA))[rng.integers(0, 4, size=(63, 8))]]
    import pandas as pd

    measured = pd.DataFrame({"sequence": start_seqs, "true_score": land.get_fitness(start_seqs), "model_score": np.nan,
                             "round": 0, "model_cost": 0, "measurement_cost": len(start_seqs)})
    cnn.train(measured["sequence"].to_numpy(), measured["true_score"].to_numpy())
    B, Q, W = 100, 1_100_000, 65_536
    ex = flexs.baselines.explorers.Adalead(cnn, rounds=1, sequences_batch_size=B, model_queries_per_batch=Q,
                                          starting_sequence="TTTTTTTT", alphabet=su.DNAA, eval_batch_size=B,
                                          rollout_width=W)
    assert ex._use_device()
    cost0 = cnn.cost
    seqs, preds = ex.propose_sequences(measured)
    spent = cnn.cost - cost0
    stats = ex.last_device_stats
    assert 1_000_000 <= spent < Q + W                      # the reference overshoots by less than one step too
    assert stats["model_queries"] == spent and stats["found"] > 10_000
    assert len(seqs) == B - 1 == len(set(seqs)) and not set(seqs) & set(start_seqs)
    assert (np.diff(preds) <= 0).all()
    direct = cnn.get_fitness(list(seqs))                    # the scores are the surrogate's scores of those strings
    np.testing.assert_allclose(preds, direct, rtol=0, atol=1e-4 * max(1e-6, float(np.abs(direct).max())))
    # the space has 65 536 sequences: a million queries found essentially all of it, and nothing beats the proposals
    everything = np.array(np.meshgrid(*[np.arange(4)] * 8)).reshape(8, -1).T.astype(np.uint8)
    all_scores = cnn.get_fitness(everything)
    unmeasured = np.array([s not in set(start_seqs) for s in su.decode_indices(everything, su.DNAA)])
    if stats["found"] >= 65_536 - 64:
        best = np.sort(all_scores[unmeasured])[::-1][: B - 1]
        np.testing.assert_allclose(preds, best, rtol=0, atol=1e-4 * float(np.abs(best).max()))


def test_adalead_device_rollouts_small_budget_and_recombination():
    """The device path at the reference's own scale (B = 100, Q = 2000, rho = 1): same bookkeeping as the host path."""
    import random

    random.seed(1)
    land = MotifLandscape("GCUAGCUAGCUAGC", su.RNAA)
    cnn = flexs.baselines.models.CNN(14, 32, 100, su.RNAA, loss="MSE", seed=1)
    ex = flexs.baselines.explorers.Adalead(cnn, rounds=2, sequences_batch_size=100, model_queries_per_batch=2000,
                                          starting_sequence="AUAUAUAUAUAUAU", alphabet=su.RNAA, eval_batch_size=100,
                                          rho=1, recomb_rate=0.2)
    table, meta = ex.run(land, verbose=False)
    _check_run(table, meta, 2, 100)
    assert ex.last_device_stats["model_queries"] <= 2000 + 100
    per_round = table[table["round"] > 0].groupby("round").size()
    assert (per_round == 99).all()                          # B-1 proposals per round, all distinct and new
    assert table["sequence"].is_unique


def test_config3_aav735_cnn_cmaes_device_population():
    """configs[3] at full length: AAV 735-mers x 20 letters, CNN surrogate, CMA-ES with the population on the GPU
    (separable sampler at dimension 14 700, argmax decode, cache lookup, fused forward, tell — no strings)."""
    L = 735
    rng = np.random.default_rng(0)
    wt = "".join(np.array(list(su.AAS))[rng.integers(0, 20, size=L)])
    land = MotifLandscape(wt, su.AAS)
    cnn = flexs.baselines.models.CNN(L, num_filters=32, hidden_size=100, alphabet=su.AAS, loss="MSE", seed=3)
    start = su.generate_random_mutant(wt, 0.05, su.AAS) if hasattr(su, "generate_random_mutant") else wt
    pop, Q, B = 2048, 3 * 2048 + 100, 50
    ex = flexs.baselines.explorers.CMAES(cnn, rounds=1, sequences_batch_size=B, model_queries_per_batch=Q,
                                        starting_sequence=start, alphabet=su.AAS, population_size=pop, max_iter=10, seed=0)
    assert ex._use_device()
    table, meta = ex.run(land, verbose=False)
    _check_run(table, meta, 1, B)
    assert 2 * pop <= cnn.cost <= Q                          # three iterations fit the budget; cached members are free
    proposed = table[table["round"] == 1]
    assert len(proposed) >= 1 and proposed["sequence"].str.len().eq(L).all()
    direct = cnn.get_fitness(list(proposed["sequence"]))
    # proposals that the model scored carry the model's score (a re-proposed measured sequence carries its true score)
    new = ~proposed["sequence"].isin([start]).to_numpy()
    np.testing.assert_allclose(proposed["model_score"].to_numpy()[new], direct[new], rtol=0,
                               atol=1e-4 * float(np.abs(direct).max()))


def test_cmaes_device_matches_host_path_semantics_on_small_problem():
    """Same seed-free invariants on a problem both paths can run: budget, cache hits are free, B-1 proposals."""
    land = flexs.landscapes.AdditiveAAVPackaging(phenotype="heart", start=450, end=540, data_file=AAV_FILE)
    start = land.wild_type
    cnn = flexs.baselines.models.CNN(90, num_filters=32, hidden_size=100, alphabet=su.AAS, loss="MSE", seed=3)
    ex = flexs.baselines.explorers.CMAES(cnn, rounds=1, sequences_batch_size=10, model_queries_per_batch=600,
                                        starting_sequence=start, alphabet=su.AAS, population_size=64, max_iter=20,
                                        seed=0, device_population=True)
    table, meta = ex.run(land, verbose=False)
    _check_run(table, meta, 1, 10)
    assert cnn.cost <= 600 and cnn.cost > 0


def test_edit_density_kernel_matches_python_environment():
    """K8 (flexs_edit_density_dev) == the host environment's sequence_density (banded Levenshtein, radius 2, the
    reference's environments/dyna_ppo.py:106-114) on sequences with planted substitutions, insertions and deletions."""
    from flexs_b200.baselines.explorers.dyna_ppo import bounded_edit_distance

    rng = np.random.default_rng(0)
    L, A = 31, 20
    base = rng.integers(0, A, size=(40, L), dtype=np.uint8)
    seen = [base]
    for _ in range(6):                                   # neighbours at distance 1-3 of the base rows
        v = base.copy()
        for r in range(len(v)):
            kind = rng.integers(0, 4)
            p = int(rng.integers(1, L - 2))
            if kind == 0:
                v[r, p] = (v[r, p] + 1) % A
            elif kind == 1:
                v[r, p] = (v[r, p] + 1) % A; v[r, (p + 7) % L] = (v[r, (p + 7) % L] + 3) % A
            elif kind == 2:                              # delete one residue, append one: a shift (distance <= 2)
                v[r] = np.concatenate([v[r, :p], v[r, p + 1:], [rng.integers(0, A)]])
            else:
                v[r, p:p + 3] = (v[r, p:p + 3] + 5) % A  # three substitutions: outside the radius
        seen.append(v)
    seen = np.concatenate(seen)
    fit = rng.normal(size=len(seen))
    fresh = np.concatenate([base[:10], seen[45:75], rng.integers(0, A, size=(8, L), dtype=np.uint8)])
    strs = ["".join(chr(65 + c) for c in row) for row in seen]
    want = []
    for row in fresh:
        s = "".join(chr(65 + c) for c in row)
        dens = 0.0
        for o, f in zip(strs, fit):
            d = bounded_edit_distance(o, s, 2)
            if d != 0 and d <= 2:
                dens += f / d
        want.append(dens)
    d_new, d_seen = torch.from_numpy(fresh).cuda(), torch.from_numpy(seen).cuda()
    d_fit = torch.from_numpy(fit).cuda()
    out = torch.empty(len(fresh), dtype=torch.float64, device="cuda")
    _native.edit_density_dev(d_new.data_ptr(), len(fresh), d_seen.data_ptr(), d_fit.data_ptr(), len(seen), L, 2, out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), np.array(want), rtol=1e-12, atol=1e-12)
    assert np.count_nonzero(want) > 20


def test_config4_gfp237_cnn_dynappo_device_rollouts():
    """configs[4]: GFP-length proteins (237 x 20), CNN surrogate, DyNA-PPO with 1024 parallel constructive episodes on
    the GPU: every episode end is ONE fused forward over uint8[1024, 237] (no strings, no one-hot), rewards carry the
    density penalty (K8), proposals are the B best unmeasured sequences of the model-based round."""
    L, E, B, Q = 237, 1024, 100, 2048
    rng = np.random.default_rng(0)
    wt = "".join(np.array(list(su.AAS))[rng.integers(0, 20, size=L)])
    land = MotifLandscape(wt, su.AAS)
    cnn = flexs.baselines.models.CNN(L, num_filters=32, hidden_size=100, alphabet=su.AAS, loss="MSE", seed=4)
    ex = flexs.baselines.explorers.DynaPPO(land, rounds=1, sequences_batch_size=B, model_queries_per_batch=Q,
                                          starting_sequence=wt, alphabet=su.AAS, model=cnn, num_model_rounds=1,
                                          env_batch_size=E, seed=0)
    assert ex._dev is not None
    table, meta = ex.run(land, verbose=False)
    assert table["round"].max() == 1 and np.isfinite(table["true_score"]).all()
    proposed = table[table["round"] == 1]
    assert len(proposed) == B and proposed["sequence"].is_unique                 # B items (not B-1): dyna_ppo.py:317
    assert proposed["sequence"].str.len().eq(L).all()
    assert proposed["sequence"].str[-1].eq(su.AAS[0]).all()                      # the last residue is never sampled (quirk 5)
    assert cnn.cost == Q                                                         # two episodes of 1024 model queries
    direct = cnn.get_fitness(list(proposed["sequence"]))
    np.testing.assert_allclose(proposed["model_score"].to_numpy(), direct, rtol=0, atol=1e-4 * float(np.abs(direct).max()))
    assert (np.diff(proposed["model_score"].to_numpy()) <= 0).all()
    # the agent's first-layer shortcut is exact: incremental pre-activations == the dense product on the one-hot state
    d = ex._dev
    act = torch.from_numpy(rng.integers(0, 20, size=(3, L - 1))).to(d.dev)
    h = d._hidden(d.params["aW1"], d.params["ab1"], act)
    t = 100
    state = torch.zeros((L, 21), device=d.dev); state[:, 20] = 1
    state[torch.arange(t), 20] = 0; state[torch.arange(t), act[1, :t]] = 1
    dense = d.params["ab1"] + torch.einsum("lc,lch->h", state, d.params["aW1"])
    assert torch.allclose(h[1, t], dense, atol=1e-4)
