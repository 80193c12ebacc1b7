"""AdaLead explorer (reference: flexs/baselines/explorers/adalead.py:11-175).

Greedy rollouts from the best measured sequences: parents within ``threshold`` of the best
measured fitness are (optionally recombined and) mutated; a child whose predicted fitness is at
least its root's keeps being mutated, otherwise that rollout stops.  Every model query goes
through ``self.model.get_fitness`` so ``model.cost`` evolves exactly as in the reference, and the
stdlib ``random`` calls happen in the reference's order (same seed -> same proposals; pinned by
tests/golden/ref_adalead.json).

With a B200 surrogate the per-call cost is one kernel launch whatever the batch, so raising
``eval_batch_size`` towards ``sequences_batch_size`` widens each rollout step to the whole frontier
(the reference default of 20 is kept for drop-in behaviour).
"""
import random
from typing import Optional, Tuple

import numpy as np
import pandas as pd

from flexs_b200.explorer import Explorer
from flexs_b200.model import Model
from flexs_b200.utils import sequence_utils as s_utils


class Adalead(Explorer):
    """Adaptive greedy evolutionary search with model-guided rollouts."""

    def __init__(
        self,
        model: Model,
        rounds: int,
        sequences_batch_size: int,
        model_queries_per_batch: int,
        starting_sequence: str,
        alphabet: str,
        mu: int = 1,
        recomb_rate: float = 0,
        threshold: float = 0.05,
        rho: int = 0,
        eval_batch_size: int = 20,
        log_file: Optional[str] = None,
    ):
        """
        Args:
            mu: expected number of mutations per sequence (``mu / L`` per residue).
            recomb_rate: per-position crossover probability during recombination.
            threshold: parents are the measured sequences with fitness >= (1 - threshold) * best.
            rho: number of recombination passes over the parent pool per outer iteration.
            eval_batch_size: how many rollouts advance together per model call.
        """
        super().__init__(model, f"Adalead_mu={mu}_threshold={threshold}", rounds, sequences_batch_size,
                         model_queries_per_batch, starting_sequence, log_file)
        self.threshold = threshold
        self.recomb_rate = recomb_rate
        self.alphabet = alphabet
        self.mu = mu
        self.rho = rho
        self.eval_batch_size = eval_batch_size

    def _recombine_population(self, gen):
        """Shuffle, then cross neighbouring pairs position by position (adalead.py:69-94)."""
        if len(gen) == 1:
            return gen
        random.shuffle(gen)
        offspring = []
        for first in range(0, len(gen) - 1, 2):
            mother, father = gen[first], gen[first + 1]
            a_chars, b_chars = [], []
            crossed = False
            for pos in range(len(mother)):
                if random.random() < self.recomb_rate:
                    crossed = not crossed
                if crossed:
                    a_chars.append(mother[pos]); b_chars.append(father[pos])
                else:
                    b_chars.append(mother[pos]); a_chars.append(father[pos])
            offspring.append("".join(a_chars))
            offspring.append("".join(b_chars))
        return offspring

    def propose_sequences(self, measured_sequences: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """Return the ``sequences_batch_size - 1`` best new sequences found by the rollouts
        (the reference's ``[: -B : -1]`` slice yields B-1 items, adalead.py:173)."""
        already_measured = set(measured_sequences["sequence"])
        best = measured_sequences["true_score"].max()
        cutoff = best * (1 - np.sign(best) * self.threshold)
        elite = measured_sequences["sequence"][measured_sequences["true_score"] >= cutoff].to_numpy()
        parents = np.resize(elite, self.sequences_batch_size)

        budget, step = self.model_queries_per_batch, self.eval_batch_size
        found = {}  # new sequence -> predicted fitness, in discovery order
        cost_at_start = self.model.cost

        def spent():
            return self.model.cost - cost_at_start

        while spent() < budget:
            for _ in range(self.rho):
                parents = self._recombine_population(parents)
            for lo in range(0, len(parents), step):
                roots = parents[lo: lo + step]
                root_scores = self.model.get_fitness(roots)
                frontier = list(enumerate(roots))
                while len(frontier) > 0 and spent() + step < budget:
                    owners, children = [], []
                    while len(children) < len(frontier):
                        # index -1 first: the rollout order of the reference (adalead.py:135)
                        owner, node = frontier[len(children) - 1]
                        child = s_utils.generate_random_mutant(node, self.mu * 1 / len(node), self.alphabet)
                        if child not in already_measured and child not in found:
                            owners.append(owner)
                            children.append(child)
                    child_scores = self.model.get_fitness(children)
                    found.update(zip(children, child_scores))
                    frontier = [(owner, child) for owner, child, score in zip(owners, children, child_scores)
                                if score >= root_scores[owner]]

        if len(found) == 0:
            raise ValueError(
                "No sequences generated. If `model_queries_per_batch` is small, try "
                "making `eval_batch_size` smaller"
            )
        new_seqs = np.array(list(found.keys()))
        preds = np.array(list(found.values()))
        order = np.argsort(preds)[: -self.sequences_batch_size: -1]
        return new_seqs[order], preds[order]
