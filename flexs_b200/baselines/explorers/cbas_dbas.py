"""CbAS / DbAS explorer (reference: flexs/baselines/explorers/cbas_dbas.py:12-201).

Conditioning by adaptive sampling (Brookes et al. 2019): a generative model (VAE) is repeatedly re-fit
on its own samples, weighted by whether the surrogate scores them above a rising threshold gamma (and,
for CbAS, by the importance ratio p_0 / p_t).  Hot-path part per cycle: ``model.get_fitness`` on
``cycle_batch_size`` proposals, ``gamma = max(np.percentile(scores, 100 Q), gamma)``, masking, and the
final ``[: -B : -1]`` selection (cbas_dbas.py:159-163, :181, :197-201) — all values and ``model.cost``
identical to the reference; the surrogate call is one fused kernel launch.
"""
import random
from typing import Optional, Tuple

import numpy as np
import pandas as pd

from flexs_b200.explorer import Explorer
from flexs_b200.model import Model
from flexs_b200.utils import sequence_utils as s_utils
from flexs_b200.utils.VAE_utils import VAE


class CbAS(Explorer):
    """CbAS (``algo="cbas"``) and DbAS (``algo="dbas"``)."""

    def __init__(
        self,
        model: Model,
        generator: VAE,
        rounds: int,
        starting_sequence: str,
        sequences_batch_size: int,
        model_queries_per_batch: int,
        alphabet: str,
        algo: str = "cbas",
        Q: float = 0.7,
        cycle_batch_size: int = 100,
        mutation_rate: float = 0.2,
        log_file: Optional[str] = None,
    ):
        """
        Args:
            generator: the VAE.
            algo: "cbas" (importance-weighted) or "dbas".
            Q: percentile used as the fitness threshold.
            cycle_batch_size: proposals per adaptation cycle.
            mutation_rate: per-residue mutation probability when padding small sample sets.
        """
        super().__init__(model, f"{algo}_Q={Q}_generator={generator.name}", rounds, sequences_batch_size,
                         model_queries_per_batch, starting_sequence, log_file)
        if algo not in ("cbas", "dbas"):
            raise ValueError("`algo` must be one of 'cbas' or 'dbas'")
        self.algo = algo
        self.generator = generator
        self.alphabet = alphabet
        self.Q = Q
        self.cycle_batch_size = cycle_batch_size
        self.mutation_rate = mutation_rate

    def _extend_samples(self, samples, weights):
        """Pad a small sample set with random mutants (weight 1) until it has 100 members (:67-83)."""
        samples, weights = list(samples), list(weights)
        present = set(samples)
        while len(present) < 100:
            mutant = s_utils.generate_random_mutant(random.choice(samples), self.mutation_rate, alphabet=self.alphabet)
            if mutant not in present:
                samples.append(mutant)
                weights.append(1)
                present.add(mutant)
        return np.array(samples), np.array(weights)

    def _clone_generator(self) -> VAE:
        g = self.generator
        twin = VAE(seq_length=g.seq_length, alphabet=g.alphabet, batch_size=g.batch_size, latent_dim=g.latent_dim,
                   intermediate_dim=g.intermediate_dim, epochs=g.epochs, epsilon_std=g.epsilon_std, beta=g.beta,
                   validation_split=g.validation_split, verbose=g.verbose, device=getattr(g, "device", None))
        twin.vae.set_weights(g.vae.get_weights())
        return twin

    def propose_sequences(self, measured_sequences_data: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """Return the ``sequences_batch_size - 1`` best proposals of this round."""
        last_round = measured_sequences_data["round"].max()
        if last_round == 0:
            # no data for the model yet: random neighbourhood of the start (:91-104)
            pool = set()
            while len(pool) < self.sequences_batch_size:
                pool.add(s_utils.generate_random_mutant(self.starting_sequence, 2 / len(self.starting_sequence),
                                                        self.alphabet))
            pool = np.array(list(pool))
            return pool, self.model.get_fitness(pool)

        recent = measured_sequences_data[measured_sequences_data["round"] == last_round]
        gamma = np.percentile(recent["true_score"], 100 * self.Q)
        seed_batch = recent["sequence"][recent["true_score"] >= gamma].to_numpy()
        samples, weights = self._extend_samples(seed_batch, np.ones(len(seed_batch)))

        self.generator.train_model(samples, weights)
        vae_0 = self._clone_generator().vae   # frozen copy of the prior (:125-144)

        found = {}
        cost_at_start = self.model.cost
        while self.model.cost - cost_at_start < self.model_queries_per_batch:
            proposals = self.generator.generate(self.cycle_batch_size, samples, weights)
            scores = self.model.get_fitness(proposals)                      # HOT CALL (:159)
            gamma = max(np.percentile(scores, self.Q * 100), gamma)        # (:163)
            if self.algo == "cbas":
                log_p0 = self.generator.calculate_log_probability(proposals, vae=vae_0)
                log_pt = self.generator.calculate_log_probability(proposals)
                w = np.nan_to_num(np.exp(log_p0 - log_pt))
            else:
                w = np.ones(len(proposals))
            w[scores < gamma] = 0
            samples = np.append(samples, proposals)
            weights = np.append(weights, w)
            self.generator.train_model(samples, weights)
            found.update(zip(proposals, scores))

        new_seqs = np.array(list(found.keys()))
        preds = np.array(list(found.values()))
        order = np.argsort(preds)[: -self.sequences_batch_size: -1]
        return new_seqs[order], preds[order]
