"""``Landscape`` plugin base class.

API parity with the reference's flexs/landscape.py:9-45: a landscape has a ``name``, a running
``cost`` counter and a public ``get_fitness`` that first charges ``len(sequences)`` to ``cost`` and
then defers to the subclass hook ``_fitness_function``.  Subclasses override the hook only.
"""
import abc

import numpy as np

from flexs_b200.types import SEQUENCES_TYPE


class Landscape(abc.ABC):
    """Ground-truth oracle plugin (and base of :class:`flexs_b200.Model`).

    Attributes:
        cost: how many sequences have been scored through ``get_fitness`` so far.
        name: label used in explorer run logs.
    """

    def __init__(self, name: str):
        self.name = name
        self.cost = 0

    @abc.abstractmethod
    def _fitness_function(self, sequences: SEQUENCES_TYPE) -> np.ndarray:
        """Score ``sequences``; implemented by subclasses."""

    def get_fitness(self, sequences: SEQUENCES_TYPE) -> np.ndarray:
        """Charge ``len(sequences)`` queries to ``cost`` and return their scores.

        Not meant to be overridden (reference: landscape.py:29-45).
        """
        self.cost += len(sequences)
        return self._fitness_function(sequences)
