"""Explorers on the virtual-screen hot path (reference: flexs/baselines/explorers/)."""
from flexs_b200.baselines.explorers.adalead import Adalead  # noqa: F401
from flexs_b200.baselines.explorers.cbas_dbas import VAE, CbAS  # noqa: F401
from flexs_b200.baselines.explorers.cmaes import CMAES  # noqa: F401
from flexs_b200.baselines.explorers.dyna_ppo import DynaPPO, DynaPPOEnsemble  # noqa: F401
