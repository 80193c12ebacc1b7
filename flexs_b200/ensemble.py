"""``Ensemble`` combinator (reference: flexs/ensemble.py:10-59).

Generic behaviour is the reference's: score with every member, stack to ``(N, M)``, reduce with
``combine_with`` (default: mean over members).  When every member is a B200 CNN/MLP of one
architecture and the default reducer is in use, the M forwards and the mean run as ONE fused
kernel launch over a native model holding all M weight sets — the candidate batch is staged once
instead of being re-encoded per member (ensemble.py:55-57 re-encodes M times).  Cost accounting is
unchanged: the ensemble AND each member are charged ``len(sequences)`` (landscape.py:44).
"""
from typing import Callable, List, Optional

import numpy as np

from flexs_b200.landscape import Landscape
from flexs_b200.model import Model
from flexs_b200.types import SEQUENCES_TYPE


def _mean_over_members(scores: np.ndarray) -> np.ndarray:
    return np.mean(scores, axis=1)


class Ensemble(Model):
    """Combine several landscapes/models into one model.

    Attributes:
        models: the members.
        combine_with: ``(N, M) -> (N,)`` reducer.
    """

    def __init__(self, models: List[Landscape], combine_with: Callable[[np.ndarray], np.ndarray] = _mean_over_members):
        super().__init__(f"Ens({'|'.join(m.name for m in models)})")
        self.models = models
        self.combine_with = combine_with
        self._fused = None
        self._fused_versions: Optional[list] = None

    def train(self, sequences: SEQUENCES_TYPE, labels: np.ndarray):
        """Train every member on the same data (ensemble.py:42-52)."""
        for member in self.models:
            member.train(sequences, labels)

    # ------------------------------------------------------------------ fused path
    def _fusable(self) -> bool:
        from flexs_b200.baselines.models.surrogate import B200Surrogate

        if self.combine_with is not _mean_over_members or len(self.models) < 2:
            return False
        first = self.models[0]
        if not isinstance(first, B200Surrogate):
            return False
        return all(
            isinstance(m, B200Surrogate) and type(m) is type(first) and m._native_kwargs == first._native_kwargs
            and m.alphabet == first.alphabet and m.device == first.device
            for m in self.models
        )

    def _fused_model(self):
        """A surrogate whose native object holds all members' weights (kept in sync lazily)."""
        from flexs_b200 import _native
        from flexs_b200.baselines.models.surrogate import B200Surrogate

        first = self.models[0]
        versions = [m.weights_version for m in self.models]
        if self._fused is None:
            shell = B200Surrogate.__new__(type(first))
            shell.__dict__.update({k: v for k, v in first.__dict__.items() if k not in ("_native", "cost")})
            shell.cost = 0
            shell._native = _native.NativeModel(first.kind, device=first.device, n_members=len(self.models),
                                                **first._native_kwargs)
            self._fused, self._fused_versions = shell, None
        if versions != self._fused_versions:
            for i, member in enumerate(self.models):
                self._fused._native.set_weights(member.native.get_weights(0), i)
            # reading .native may have initialised a member: re-read the versions
            self._fused_versions = [m.weights_version for m in self.models]
        return self._fused

    def _fitness_function(self, sequences):
        if self._fusable():
            for member in self.models:  # members are charged too (quirk kept: ensemble.py:55-57)
                member.cost += len(sequences)
            return self._fused_model()._fitness_function(sequences)
        scores = np.stack([member.get_fitness(sequences) for member in self.models], axis=1)
        return self.combine_with(scores)

    def get_fitness_device(self, idx):
        """Device-resident scoring (CUDA ``uint8[N, L]`` -> CUDA ``float32[N]``) for fused ensembles."""
        if not self._fusable():
            raise TypeError("get_fitness_device needs an ensemble of identical B200 surrogates with the default mean")
        n = int(idx.shape[0])
        self.cost += n
        for member in self.models:
            member.cost += n
        return self._fused_model()._score_device(idx)
