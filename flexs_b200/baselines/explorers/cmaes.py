"""CMA-ES explorer (reference: flexs/baselines/explorers/cmaes.py:12-122).

A Gaussian over the relaxed one-hot matrix ``(L, A)`` is adapted by CMA-ES; each sample is decoded to a
sequence by per-position argmax (first maximum wins) and scored by the model.  Reference quirks kept:
the model score is handed to a MINIMISING strategy un-negated (cmaes.py:108-110, comment :118), the
already-measured / already-seen sequences are answered from caches without charging the model, and
duplicates inside one population are each charged (the cache is only updated after the iteration).

What changed underneath: the reference scores one sequence per ``get_fitness`` call inside
``ask_and_eval`` (cmaes.py:83-91); here the whole population is decoded first and all uncached members
are scored by ONE ``get_fitness`` call — same values, same ``model.cost``, one kernel launch per
iteration instead of ``population_size``.  The sampler is flexs_b200.utils.cma (``cma`` is absent).

Device path (``device_population=True``, or automatically for ``population_size >= 256`` with a B200 surrogate): the
population never becomes strings.  ``flexs_b200.utils.cma_device.SepCMA`` samples ``float32[popsize, L*A]`` on the GPU,
``flexs_argmax_decode_dev`` decodes it (first maximum wins, bit-identical to ``np.argmax``), the ``seen`` / ``measured``
caches are one ``flexs_dedup_representatives_dev`` lookup over [measured rows; seen rows; population], the uncached
members are scored by one fused forward launch, and ``tell`` and the final ranking stay on the device.  Same cache and
``model.cost`` semantics as above.  Under ``torch.distributed`` every rank samples the same population (same seed),
scores its contiguous share, and ONE ``all_gather`` of the ``float32`` scores completes the fitness vector (SURVEY.md
§8e: the all-scores variant of the sharded screen) — BASELINE configs[3], AAV 735-mers sharded over 4 GPUs.
"""
from typing import Optional, Tuple

import numpy as np
import pandas as pd

from flexs_b200.explorer import Explorer
from flexs_b200.model import Model
from flexs_b200.utils import cma
from flexs_b200.utils import sequence_utils as s_utils


class CMAES(Explorer):
    """Covariance-matrix-adaptation evolution strategy over a continuous relaxation of the sequence."""

    def __init__(
        self,
        model: Model,
        rounds: int,
        sequences_batch_size: int,
        model_queries_per_batch: int,
        starting_sequence: str,
        alphabet: str,
        population_size: int = 15,
        max_iter: int = 400,
        initial_variance: float = 0.2,
        log_file: Optional[str] = None,
        seed: Optional[int] = None,
        device_population: Optional[bool] = None,
    ):
        """
        Args:
            population_size: solutions sampled per iteration.
            max_iter: iteration cap per round.
            initial_variance: initial variance of the search distribution.
            seed: optional seed of the sampler (the reference's is unseeded).
            device_population: keep the population on the GPU (``None``: for ``population_size >= 256`` with a B200
                surrogate).
        """
        super().__init__(model, f"CMAES_popsize{population_size}", rounds, sequences_batch_size,
                         model_queries_per_batch, starting_sequence, log_file)
        self.alphabet = alphabet
        self.population_size = population_size
        self.max_iter = max_iter
        self.initial_variance = initial_variance
        self.round = 0
        self.seed = seed
        self.device_population = device_population

    # ------------------------------------------------------------------ device path
    def _use_device(self) -> bool:
        if self.device_population is not None:
            return bool(self.device_population)
        return self.population_size >= 256 and hasattr(self.model, "get_fitness_device")

    def _propose_device(self, measured_sequences: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        import torch
        import torch.distributed as dist

        from flexs_b200 import _native
        from flexs_b200.utils.cma_device import SepCMA

        index = getattr(self.model, "device", None)
        if index is None and hasattr(self.model, "models"):
            index = getattr(self.model.models[0], "device", 0)
        dev = torch.device("cuda", int(index or 0))
        L, A, pop = len(self.starting_sequence), len(self.alphabet), self.population_size
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0

        seqs = measured_sequences["sequence"].to_numpy()
        truth = measured_sequences["true_score"].to_numpy(dtype=np.float64)
        best_row = int(np.argmax(truth))
        measured_rows = torch.from_numpy(s_utils.encode_sequences(list(seqs), self.alphabet)).to(dev)
        # the caches as (rows, values): [measured ...; seen ...]; `seen` starts with the best measured sequence
        # (cmaes.py:76-80) and is consulted BEFORE `measured` (:85-90) — it sits in front
        cache_rows = torch.cat([measured_rows[best_row: best_row + 1], measured_rows])
        cache_vals = torch.cat([torch.tensor([truth[best_row]], dtype=torch.float64),
                                torch.from_numpy(truth.copy())]).to(dev)
        n_static = int(cache_rows.shape[0])          # entries [1, n_static) are `measured`: never proposed
        seen_flag = torch.zeros(n_static, dtype=torch.bool, device=dev)
        seen_flag[0] = True

        x0 = torch.zeros((L, A), dtype=torch.float32, device=dev)
        x0[torch.arange(L, device=dev), measured_rows[best_row].long()] = 1.0
        seed = None if self.seed is None else self.seed + self.round
        if seed is None and world > 1:               # every rank must sample the same population
            box = torch.randint(0, 2 ** 31 - 1, (1,), device=dev)
            dist.broadcast(box, 0)
            seed = int(box.item())
        self.round += 1
        es = SepCMA(x0.reshape(-1), float(np.sqrt(self.initial_variance)), pop, seed)
        stream = torch.cuda.current_stream(dev).cuda_stream
        cost_at_start = self.model.cost
        for _ in range(self.max_iter):
            if self.model.cost - cost_at_start + pop > self.model_queries_per_batch:
                break
            X = es.ask()
            rows = torch.empty((pop, L), dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _native.argmax_decode_dev(X.data_ptr(), pop, L, A, A, rows.data_ptr(), stream)
                allrows = torch.cat([cache_rows, rows])
                rep = torch.empty(len(allrows), dtype=torch.int64, device=dev)
                work = torch.empty(_native.dedup_workspace_bytes(len(allrows)), dtype=torch.uint8, device=dev)
                _native.dedup_representatives_dev(allrows.data_ptr(), len(allrows), L, rep.data_ptr(), work.data_ptr(), stream)
            n_cache = int(cache_rows.shape[0])
            rep_pop = rep[n_cache:]
            cached = rep_pop < n_cache                      # answered from `seen` / `measured`, not charged
            fitness = torch.empty(pop, dtype=torch.float64, device=dev)
            fitness[cached] = cache_vals[rep_pop[cached]]
            todo = torch.nonzero(~cached).reshape(-1)       # duplicates inside one population are each scored (and charged)
            if len(todo):
                if world > 1:
                    lo, hi = (len(todo) * rank) // world, (len(todo) * (rank + 1)) // world
                    per = -(-len(todo) // world)
                    mine = torch.full((per,), float("nan"), dtype=torch.float32, device=dev)
                    if hi > lo:
                        mine[: hi - lo] = self.model.get_fitness_device(rows[todo[lo:hi]].contiguous())
                    gathered = torch.empty(world * per, dtype=torch.float32, device=dev)
                    dist.all_gather_into_tensor(gathered, mine)      # the one collective of an iteration
                    parts = [gathered[r * per: r * per + ((len(todo) * (r + 1)) // world - (len(todo) * r) // world)]
                             for r in range(world)]
                    scores = torch.cat(parts)
                    self.model.cost += len(todo) - (hi - lo)         # cost counts every query of the iteration on every rank
                else:
                    scores = self.model.get_fitness_device(rows[todo].contiguous())
                fitness[todo] = scores.to(torch.float64)
            es.tell(X, fitness)                               # un-negated, as the reference does
            # seen[seq] = f for the whole population: new sequences join the cache (first occurrence of each)
            first = todo[rep_pop[todo] == (n_cache + todo)]
            cache_rows = torch.cat([cache_rows, rows[first]])
            cache_vals = torch.cat([cache_vals, fitness[first]])
            seen_flag = torch.cat([seen_flag, torch.ones(len(first), dtype=torch.bool, device=dev)])
            hit = rep_pop[cached]
            seen_flag[hit] = True                             # a measured sequence that was proposed enters `seen` too

        cand = torch.nonzero(seen_flag).reshape(-1)
        preds = cache_vals[cand].to(torch.float32).contiguous()
        k = min(self.sequences_batch_size - 1, len(cand))     # [: -B : -1] keeps B-1
        if k <= 0:
            return np.array([], dtype=str), np.array([], dtype=np.float32)
        top_s = torch.empty(k, dtype=torch.float32, device=dev)
        top_i = torch.empty(k, dtype=torch.int64, device=dev)
        work = torch.empty(_native.topk_select_workspace_bytes(), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _native.topk_select_dev(preds.data_ptr(), len(cand), k, 0, 0, 0, False, top_s.data_ptr(), top_i.data_ptr(), 0, 0,
                                    work.data_ptr(), stream)
        winners = cache_rows[cand[top_i]].cpu().numpy()
        return s_utils.decode_indices(winners, self.alphabet), top_s.cpu().numpy()

    def _soln_to_string(self, soln) -> str:
        """Relaxed solution -> sequence: reshape ``(L, A)``, argmax per position (cmaes.py:61-67)."""
        x = np.asarray(soln).reshape((len(self.starting_sequence), len(self.alphabet)))
        return "".join(self.alphabet[i] for i in np.argmax(x, axis=1))

    def _decode_population(self, solutions) -> list:
        x = np.asarray(solutions).reshape((len(solutions), len(self.starting_sequence), len(self.alphabet)))
        return list(s_utils.decode_indices(np.argmax(x, axis=2).astype(np.uint8), self.alphabet))

    def propose_sequences(self, measured_sequences: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """Return the ``sequences_batch_size - 1`` best sequences seen this round (cmaes.py:117-122)."""
        if self._use_device():
            return self._propose_device(measured_sequences)
        measured = dict(zip(measured_sequences["sequence"], measured_sequences["true_score"]))
        best_row = measured_sequences["true_score"].argmax()
        top_seq = measured_sequences["sequence"].to_numpy()[best_row]
        top_val = measured_sequences["true_score"].to_numpy()[best_row]
        seen = {top_seq: top_val}

        x0 = s_utils.string_to_one_hot(top_seq, self.alphabet).flatten()
        opts = {"popsize": self.population_size, "verbose": -9, "verb_log": 0}
        if self.seed is not None:
            opts["seed"] = self.seed + self.round
        self.round += 1
        es = cma.CMAEvolutionStrategy(x0, np.sqrt(self.initial_variance), opts)

        cost_at_start = self.model.cost
        for _ in range(self.max_iter):
            if self.model.cost - cost_at_start + self.population_size > self.model_queries_per_batch:
                break
            solutions = es.ask()
            strings = self._decode_population(solutions)
            fitnesses = [None] * len(strings)
            to_score = []
            for i, seq in enumerate(strings):
                if seq in seen:
                    fitnesses[i] = seen[seq]
                elif seq in measured:
                    fitnesses[i] = measured[seq]
                else:
                    to_score.append(i)
            if to_score:
                scores = self.model.get_fitness([strings[i] for i in to_score])
                for i, sc in zip(to_score, scores):
                    fitnesses[i] = sc.item() if hasattr(sc, "item") else float(sc)
            es.tell(solutions, fitnesses)  # un-negated, as the reference does
            for seq, f in zip(strings, fitnesses):
                seen[seq] = f

        new_seqs = np.array(list(seen.keys()))
        preds = np.array(list(seen.values()))
        order = np.argsort(preds)[: -self.sequences_batch_size: -1]
        return new_seqs[order], preds[order]
