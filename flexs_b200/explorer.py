"""``Explorer`` plugin base class and its round loop.

API and log format parity with the reference's flexs/explorer.py:17-184: subclasses implement
``propose_sequences``; ``run`` alternates model training, proposal and ground-truth measurement
and keeps the same per-round bookkeeping columns.  (The reference grows its table with
``DataFrame.append``, explorer.py:170, which pandas >= 2 removed; ``pd.concat`` is used here.)
"""
import abc
import json
import os
import time
import warnings
from datetime import datetime
from typing import Dict, Optional, Tuple

import numpy as np
import pandas as pd

from flexs_b200.landscape import Landscape
from flexs_b200.model import Model


class Explorer(abc.ABC):
    """Search algorithm plugin.  Drive it with :meth:`run`; customise :meth:`propose_sequences`."""

    def __init__(
        self,
        model: Model,
        name: str,
        rounds: int,
        sequences_batch_size: int,
        model_queries_per_batch: int,
        starting_sequence: str,
        log_file: Optional[str] = None,
    ):
        """
        Args:
            model: surrogate the explorer may query ``model_queries_per_batch`` times per round.
            name: label written to the run log.
            rounds: number of (train, propose, measure) rounds.
            sequences_batch_size: ground-truth measurements allowed per round.
            model_queries_per_batch: surrogate queries allowed per round.
            starting_sequence: seed of the exploration.
            log_file: optional ``.csv`` path; rewritten after every round.
        """
        self.model = model
        self.name = name
        self.rounds = rounds
        self.sequences_batch_size = sequences_batch_size
        self.model_queries_per_batch = model_queries_per_batch
        self.starting_sequence = starting_sequence
        self.log_file = log_file
        if log_file is not None:
            folder = os.path.split(log_file)[0]
            if folder:
                os.makedirs(folder, exist_ok=True)
        if model_queries_per_batch < sequences_batch_size:
            warnings.warn("`model_queries_per_batch` should be >= `sequences_batch_size`")

    @abc.abstractmethod
    def propose_sequences(self, measured_sequences_data: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """Return ``(sequences, model_scores)`` to measure next.

        ``measured_sequences_data`` holds everything measured so far with columns
        ``sequence``, ``model_score``, ``true_score``, ``round`` (plus the cost columns).
        """

    def _log(self, sequences_data: pd.DataFrame, metadata: Dict, current_round: int, verbose: bool,
             round_start_time: float) -> None:
        # one JSON metadata line, then the table as CSV (explorer.py:100-107)
        if self.log_file is not None:
            with open(self.log_file, "w") as handle:
                json.dump(metadata, handle)
                handle.write("\n")
                sequences_data.to_csv(handle, index=False)
        if verbose:
            print(f"round: {current_round}, top: {sequences_data['true_score'].max()}, "
                  f"time: {time.time() - round_start_time:02f}s")

    def run(self, landscape: Landscape, verbose: bool = True) -> Tuple[pd.DataFrame, Dict]:
        """Run ``self.rounds`` rounds against ``landscape``; returns the measurement table and metadata."""
        self.model.cost = 0
        metadata = {
            "run_id": datetime.now().strftime("%H:%M:%S-%m/%d/%Y"),
            "exp_name": self.name,
            "model_name": self.model.name,
            "landscape_name": landscape.name,
            "rounds": self.rounds,
            "sequences_batch_size": self.sequences_batch_size,
            "model_queries_per_batch": self.model_queries_per_batch,
        }
        start_score = landscape.get_fitness([self.starting_sequence])
        table = pd.DataFrame(
            {
                "sequence": [self.starting_sequence],
                "model_score": [np.nan],
                "true_score": np.asarray(start_score).reshape(-1),
                "round": [0],
                "model_cost": [self.model.cost],
                "measurement_cost": [1],
            }
        )
        self._log(table, metadata, 0, verbose, time.time())

        if verbose:
            round_iter = range(1, self.rounds + 1)
        else:
            import tqdm

            round_iter = tqdm.trange(1, self.rounds + 1)
        for r in round_iter:
            t0 = time.time()
            # retrain on the full history; weights/optimiser state carry over between rounds
            self.model.train(table["sequence"].to_numpy(), table["true_score"].to_numpy())
            seqs, preds = self.propose_sequences(table)
            true_score = landscape.get_fitness(seqs)
            if len(seqs) > self.sequences_batch_size:
                warnings.warn("Must propose <= `self.sequences_batch_size` sequences per round")
            new_rows = pd.DataFrame(
                {
                    "sequence": seqs,
                    "model_score": preds,
                    "true_score": true_score,
                    "round": r,
                    "model_cost": self.model.cost,
                    "measurement_cost": len(table) + len(seqs),
                }
            )
            table = pd.concat([table, new_rows])
            self._log(table, metadata, r, verbose, t0)
        return table, metadata
