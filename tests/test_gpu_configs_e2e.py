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
