"""``NoisyAbstractModel`` (reference: flexs/baselines/models/noisy_abstract_model.py:9-101).

Not on the roofline path but ``evaluate.robustness`` builds one per signal strength
(evaluate.py:31), so the drop-in needs it.  Same behaviour: a ground-truth landscape corrupted by noise
that grows with the edit distance d to the nearest already-seen sequence,
``f_hat(x) = ss^d f(x) + (1 - ss^d) eps`` with ``eps ~ Exp(mean = f(nearest))``; answers are cached so the
model is deterministic per sequence.  The reference's ``editdistance`` C extension is absent: a plain
two-row Levenshtein is used.
"""
import numpy as np

from flexs_b200.landscape import Landscape
from flexs_b200.model import Model
from flexs_b200.types import SEQUENCES_TYPE


def edit_distance(a: str, b: str) -> int:
    """Levenshtein distance (what ``editdistance.eval`` returns)."""
    if a == b:
        return 0
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


class NoisyAbstractModel(Model):
    """Ground truth with distance-modulated noise."""

    def __init__(self, landscape: Landscape, signal_strength: float = 0.9):
        super().__init__(f"NAMb_ss{signal_strength}")
        self.landscape = landscape
        self.ss = signal_strength
        self.cache = {}

    def _get_min_distance(self, sequence):
        if len(self.cache) == 0:
            return 0, sequence
        best, closest = np.inf, None
        for seq in self.cache:
            dist = edit_distance(sequence, seq)
            if dist == 1:
                return dist, seq     # the reference stops at the first distance-1 neighbour (:51-52)
            if dist < best:
                best, closest = dist, seq
        return best, closest

    def train(self, sequences: SEQUENCES_TYPE, labels: np.ndarray):
        """Remember the measured sequences and their labels (:60-65)."""
        self.cache.update(zip(sequences, labels))

    def _fitness_function(self, sequences):
        sequences = np.array(sequences)
        fitnesses = np.empty(len(sequences))
        cached = np.array([seq in self.cache for seq in sequences], dtype=bool)
        fitnesses[cached] = np.array([self.cache[seq] for seq in sequences[cached]])
        fresh = []
        for seq in sequences[~cached]:
            distance, neighbour = self._get_min_distance(seq)
            signal = self.landscape.get_fitness([seq]).item()
            neighbour_fitness = self.landscape.get_fitness([neighbour]).item()
            if neighbour_fitness >= 0:
                noise = np.random.exponential(scale=neighbour_fitness)
            else:
                noise = np.random.choice(list(self.cache.values()))
            alpha = self.ss ** distance
            fresh.append(alpha * signal + (1 - alpha) * noise)
        fitnesses[~cached] = fresh
        self.cache.update(zip(sequences[~cached], fitnesses[~cached]))
        return np.array(fitnesses)
