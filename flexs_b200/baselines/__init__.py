"""Baseline models and explorers (the subset on the virtual-screen hot path)."""
from flexs_b200.baselines import explorers, models  # noqa: F401
