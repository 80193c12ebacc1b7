"""``Model`` plugin base class: a :class:`Landscape` that can also be trained.

API parity with flexs/model.py:11-54 of the reference.
"""
import abc
from typing import Any, List

import numpy as np

from flexs_b200.landscape import Landscape
from flexs_b200.types import SEQUENCES_TYPE


class Model(Landscape, abc.ABC):
    """Learned surrogate of a landscape; adds ``train`` to the landscape interface."""

    @abc.abstractmethod
    def train(self, sequences: SEQUENCES_TYPE, labels: List[Any]):
        """Update the model from every (sequence, measured fitness) pair seen so far.

        ``Explorer.run`` calls this once per round with the whole history (explorer.py:157-160).
        """


class LandscapeAsModel(Model):
    """Adapter that lets a ground-truth landscape stand in as a perfect model (model.py:30-54).

    Queries are forwarded to the landscape's ``_fitness_function`` so that they are charged to
    this wrapper's ``cost`` and not to the landscape's.
    """

    def __init__(self, landscape: Landscape):
        super().__init__(f"LandscapeAsModel={landscape.name}")
        self.landscape = landscape

    def _fitness_function(self, sequences: SEQUENCES_TYPE) -> np.ndarray:
        return self.landscape._fitness_function(sequences)

    def train(self, sequences: SEQUENCES_TYPE, labels: List[Any]):
        """Nothing to learn."""
