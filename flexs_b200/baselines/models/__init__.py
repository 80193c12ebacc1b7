"""Surrogate models on the virtual-screen hot path (``KerasModel`` -> ``B200Surrogate``)."""
from flexs_b200.baselines.models.cnn import CNN  # noqa: F401
from flexs_b200.baselines.models.mlp import MLP  # noqa: F401
from flexs_b200.baselines.models.surrogate import B200Surrogate  # noqa: F401

#: drop-in alias: code that subclasses / type-checks against ``baselines.models.KerasModel``
KerasModel = B200Surrogate
from flexs_b200.baselines.models.noisy_abstract_model import NoisyAbstractModel  # noqa: F401,E402
