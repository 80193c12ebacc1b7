"""CMA-ES explorer (reference: flexs/baselines/explorers/cmaes.py:12-122).

A Gaussian over the relaxed one-hot matrix ``(L, A)`` is adapted by CMA-ES; each sample is decoded to a
sequence by per-position argmax (first maximum wins) and scored by the model.  Reference quirks kept:
the model score is handed to a MINIMISING strategy un-negated (cmaes.py:108-110, comment :118), the
already-measured / already-seen sequences are answered from caches without charging the model, and
duplicates inside one population are each charged (the cache is only updated after the iteration).

What changed underneath: the reference scores one sequence per ``get_fitness`` call inside
``ask_and_eval`` (cmaes.py:83-91); here the whole population is decoded first and all uncached members
are scored by ONE ``get_fitness`` call — same values, same ``model.cost``, one kernel launch per
iteration instead of ``population_size``.  The sampler is flexs_b200.utils.cma (``cma`` is absent).
"""
from typing import Optional, Tuple

import numpy as np
import pandas as pd

from flexs_b200.explorer import Explorer
from flexs_b200.model import Model
from flexs_b200.utils import cma
from flexs_b200.utils import sequence_utils as s_utils


class CMAES(Explorer):
    """Covariance-matrix-adaptation evolution strategy over a continuous relaxation of the sequence."""

    def __init__(
        self,
        model: Model,
        rounds: int,
        sequences_batch_size: int,
        model_queries_per_batch: int,
        starting_sequence: str,
        alphabet: str,
        population_size: int = 15,
        max_iter: int = 400,
        initial_variance: float = 0.2,
        log_file: Optional[str] = None,
        seed: Optional[int] = None,
    ):
        """
        Args:
            population_size: solutions sampled per iteration.
            max_iter: iteration cap per round.
            initial_variance: initial variance of the search distribution.
            seed: optional seed of the sampler (the reference's is unseeded).
        """
        super().__init__(model, f"CMAES_popsize{population_size}", rounds, sequences_batch_size,
                         model_queries_per_batch, starting_sequence, log_file)
        self.alphabet = alphabet
        self.population_size = population_size
        self.max_iter = max_iter
        self.initial_variance = initial_variance
        self.round = 0
        self.seed = seed

    def _soln_to_string(self, soln) -> str:
        """Relaxed solution -> sequence: reshape ``(L, A)``, argmax per position (cmaes.py:61-67)."""
        x = np.asarray(soln).reshape((len(self.starting_sequence), len(self.alphabet)))
        return "".join(self.alphabet[i] for i in np.argmax(x, axis=1))

    def _decode_population(self, solutions) -> list:
        x = np.asarray(solutions).reshape((len(solutions), len(self.starting_sequence), len(self.alphabet)))
        return list(s_utils.decode_indices(np.argmax(x, axis=2).astype(np.uint8), self.alphabet))

    def propose_sequences(self, measured_sequences: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """Return the ``sequences_batch_size - 1`` best sequences seen this round (cmaes.py:117-122)."""
        measured = dict(zip(measured_sequences["sequence"], measured_sequences["true_score"]))
        best_row = measured_sequences["true_score"].argmax()
        top_seq = measured_sequences["sequence"].to_numpy()[best_row]
        top_val = measured_sequences["true_score"].to_numpy()[best_row]
        seen = {top_seq: top_val}

        x0 = s_utils.string_to_one_hot(top_seq, self.alphabet).flatten()
        opts = {"popsize": self.population_size, "verbose": -9, "verb_log": 0}
        if self.seed is not None:
            opts["seed"] = self.seed + self.round
        self.round += 1
        es = cma.CMAEvolutionStrategy(x0, np.sqrt(self.initial_variance), opts)

        cost_at_start = self.model.cost
        for _ in range(self.max_iter):
            if self.model.cost - cost_at_start + self.population_size > self.model_queries_per_batch:
                break
            solutions = es.ask()
            strings = self._decode_population(solutions)
            fitnesses = [None] * len(strings)
            to_score = []
            for i, seq in enumerate(strings):
                if seq in seen:
                    fitnesses[i] = seen[seq]
                elif seq in measured:
                    fitnesses[i] = measured[seq]
                else:
                    to_score.append(i)
            if to_score:
                scores = self.model.get_fitness([strings[i] for i in to_score])
                for i, sc in zip(to_score, scores):
                    fitnesses[i] = sc.item() if hasattr(sc, "item") else float(sc)
            es.tell(solutions, fitnesses)  # un-negated, as the reference does
            for seq, f in zip(strings, fitnesses):
                seen[seq] = f

        new_seqs = np.array(list(seen.keys()))
        preds = np.array(list(seen.values()))
        order = np.argsort(preds)[: -self.sequences_batch_size: -1]
        return new_seqs[order], preds[order]
