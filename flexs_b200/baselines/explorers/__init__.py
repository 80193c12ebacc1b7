"""Explorers on the virtual-screen hot path (reference: flexs/baselines/explorers/)."""
from flexs_b200.baselines.explorers.adalead import Adalead  # noqa: F401
from flexs_b200.baselines.explorers.cmaes import CMAES  # noqa: F401
