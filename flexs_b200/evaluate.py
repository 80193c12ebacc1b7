"""Benchmark drivers (reference: flexs/evaluate.py:8-112): robustness, efficiency, adaptivity.

Each builds explorers through a user factory and calls ``explorer.run(landscape)``; the B200
surrogates plug in through the factory unchanged.
"""
from typing import Callable, List, Tuple

from flexs_b200.explorer import Explorer
from flexs_b200.landscape import Landscape
from flexs_b200.model import Model


def robustness(
    landscape: Landscape,
    make_explorer: Callable[[Model, float], Explorer],
    signal_strengths: List[float] = [0, 0.5, 0.75, 0.9, 1],
    verbose: bool = True,
):
    """Run the explorer against ``NoisyAbstractModel``s of each signal strength (evaluate.py:8-37).

    Returns ``[(signal_strength, (table, metadata)), ...]``.
    """
    from flexs_b200.baselines.models.noisy_abstract_model import NoisyAbstractModel

    results = []
    for ss in signal_strengths:
        print(f"Evaluating for robustness with model accuracy; signal_strength: {ss}")
        model = NoisyAbstractModel(landscape, signal_strength=ss)
        results.append((ss, make_explorer(model, ss).run(landscape, verbose=verbose)))
    return results


def efficiency(
    landscape: Landscape,
    make_explorer: Callable[[int, int], Explorer],
    budgets: List[Tuple[int, int]] = [(100, 500), (100, 5000), (1000, 5000), (1000, 10000)],
):
    """Sweep (``sequences_batch_size``, ``model_queries_per_batch``) budgets (evaluate.py:40-74)."""
    results = []
    for batch, queries in budgets:
        print(f"Evaluating for sequences_batch_size: {batch}, model_queries_per_batch: {queries}")
        results.append(((batch, queries), make_explorer(batch, queries).run(landscape)))
    return results


def adaptivity(
    landscape: Landscape,
    make_explorer: Callable[[int, int, int], Explorer],
    num_rounds: List[int] = [1, 10, 100],
    total_ground_truth_measurements: int = 1000,
    total_model_queries: int = 10000,
):
    """Fixed total budgets split over different numbers of rounds (evaluate.py:77-112)."""
    results = []
    for rounds in num_rounds:
        print(f"Evaluating for num_rounds: {rounds}")
        explorer = make_explorer(rounds, int(total_ground_truth_measurements / rounds),
                                 int(total_model_queries / rounds))
        results.append((rounds, explorer.run(landscape)))
    return results
