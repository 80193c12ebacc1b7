"""AdaLead explorer (reference: flexs/baselines/explorers/adalead.py:11-175).

Greedy rollouts from the best measured sequences: parents within ``threshold`` of the best
measured fitness are (optionally recombined and) mutated; a child whose predicted fitness is at
least its root's keeps being mutated, otherwise that rollout stops.  Every model query goes
through ``self.model.get_fitness`` so ``model.cost`` evolves exactly as in the reference, and the
stdlib ``random`` calls happen in the reference's order (same seed -> same proposals; pinned by
tests/golden/ref_adalead.json).

With a B200 surrogate the per-call cost is one kernel launch whatever the batch.  Two paths:

* the reference-order host path (default): the reference's control flow and ``random`` call order, string by string —
  what the golden test pins;
* the device path (``eval_batch_size >= sequences_batch_size`` with a B200 surrogate, or ``device_rollouts=True``):
  the WHOLE frontier advances per step and never leaves the GPU — parents as ``uint8[width, L]``, children from
  ``flexs_mutate_dev`` (Philox), "not measured, not found yet" through ``flexs_dedup_scores_dev`` against the rows seen
  so far, scores from the fused forward kernel, the ``score >= root score`` continuation test and the final ranking on
  the device.  ``rollout_width`` (default ``sequences_batch_size``) is the number of parallel rollouts: tens of thousands
  turn a round into a genuine virtual screen (BASELINE configs[1]: 1M model queries in a round).  Same algorithm, same
  ``model.cost`` accounting; the RNG stream differs (the reference's is unseeded, so parity is distributional).
"""
import random
from typing import Optional, Tuple

import numpy as np
import pandas as pd

from flexs_b200.explorer import Explorer
from flexs_b200.model import Model
from flexs_b200.utils import sequence_utils as s_utils


class Adalead(Explorer):
    """Adaptive greedy evolutionary search with model-guided rollouts."""

    def __init__(
        self,
        model: Model,
        rounds: int,
        sequences_batch_size: int,
        model_queries_per_batch: int,
        starting_sequence: str,
        alphabet: str,
        mu: int = 1,
        recomb_rate: float = 0,
        threshold: float = 0.05,
        rho: int = 0,
        eval_batch_size: int = 20,
        log_file: Optional[str] = None,
        device_rollouts: Optional[bool] = None,
        rollout_width: Optional[int] = None,
    ):
        """
        Args:
            mu: expected number of mutations per sequence (``mu / L`` per residue).
            recomb_rate: per-position crossover probability during recombination.
            threshold: parents are the measured sequences with fitness >= (1 - threshold) * best.
            rho: number of recombination passes over the parent pool per outer iteration.
            eval_batch_size: how many rollouts advance together per model call.
            device_rollouts: run the rollouts on the GPU (``None``: when ``eval_batch_size >= sequences_batch_size``
                and the model is a B200 surrogate).
            rollout_width: parallel rollouts of the device path (``None``: ``sequences_batch_size``, the size of
                the reference's parent pool).
        """
        super().__init__(model, f"Adalead_mu={mu}_threshold={threshold}", rounds, sequences_batch_size,
                         model_queries_per_batch, starting_sequence, log_file)
        self.threshold = threshold
        self.recomb_rate = recomb_rate
        self.alphabet = alphabet
        self.mu = mu
        self.rho = rho
        self.eval_batch_size = eval_batch_size
        self.device_rollouts = device_rollouts
        self.rollout_width = rollout_width
        self.last_device_stats = None

    # ------------------------------------------------------------------ device path
    MAX_MUTATION_TRIES = 16   # re-draws of a child that is already known (P(no residue changes) ~ 0.37 per draw at mu = 1)

    def _use_device(self) -> bool:
        if self.device_rollouts is not None:
            return bool(self.device_rollouts)
        return self.eval_batch_size >= self.sequences_batch_size and hasattr(self.model, "get_fitness_device")

    def _model_device(self):
        import torch

        index = getattr(self.model, "device", None)
        if index is None and hasattr(self.model, "models"):
            index = getattr(self.model.models[0], "device", 0)
        return torch.device("cuda", int(index or 0))

    @staticmethod
    def _first_occurrences(rows, n_known: int):
        """``rows`` = [known rows; candidates]: True for every candidate that equals no earlier row (exact, dedup.cu)."""
        import torch

        from flexs_b200 import _native

        n, L = int(rows.shape[0]), int(rows.shape[1])
        zeros = torch.zeros(n, dtype=torch.float32, device=rows.device)
        work = torch.empty(_native.dedup_workspace_bytes(n), dtype=torch.uint8, device=rows.device)
        with torch.cuda.device(rows.device):
            _native.dedup_scores_dev(rows.data_ptr(), n, L, zeros.data_ptr(), zeros.data_ptr(), work.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream)
        return zeros[n_known:] == 0          # repeats were set to -inf

    def _mutate_unseen(self, nodes, known, seed: int, counter: list):
        """One child per node that is neither measured nor found nor proposed twice in this step (adalead.py:137-150:
        the reference re-draws until the child is new; here up to MAX_MUTATION_TRIES draws, whatever is still known
        then leaves the frontier).  Returns ``(children, accepted mask)``."""
        import torch

        from flexs_b200 import _native

        n, L = int(nodes.shape[0]), int(nodes.shape[1])
        children = torch.empty_like(nodes)
        accepted = torch.zeros(n, dtype=torch.bool, device=nodes.device)
        pending = torch.arange(n, device=nodes.device)
        mu = float(self.mu) / L
        for _ in range(self.MAX_MUTATION_TRIES):
            src = nodes[pending].contiguous()
            cand = torch.empty_like(src)
            counter[0] += 1
            with torch.cuda.device(nodes.device):
                _native.mutate_dev(src.data_ptr(), len(src), L, len(self.alphabet), mu, seed, counter[0], cand.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
            rows = torch.cat([known, children[accepted], cand])
            fresh = self._first_occurrences(rows, len(rows) - len(cand))
            take = pending[fresh]
            children[take] = cand[fresh]
            accepted[take] = True
            pending = pending[~fresh]
            if len(pending) == 0:
                break
        return children, accepted

    def _recombine_device(self, gen, rng):
        """adalead.py:69-94 on ``uint8[n, L]``: shuffle, cross neighbouring pairs position by position (an odd one out is
        dropped, as the reference's ``range(0, len - 1, 2)`` does)."""
        import torch

        n, L = int(gen.shape[0]), int(gen.shape[1])
        if n == 1:
            return gen
        gen = gen[torch.randperm(n, device=gen.device, generator=rng)]
        pairs = (n // 2) * 2
        mother, father = gen[0:pairs:2], gen[1:pairs:2]
        toggles = torch.rand((pairs // 2, L), device=gen.device, generator=rng) < self.recomb_rate
        crossed = (torch.cumsum(toggles.to(torch.int32), dim=1) % 2) == 1
        a = torch.where(crossed, mother, father)
        b = torch.where(crossed, father, mother)
        return torch.stack([a, b], dim=1).reshape(pairs, L).contiguous()

    def _propose_device(self, measured_sequences: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        import torch

        from flexs_b200 import _native

        dev = self._model_device()
        seqs = measured_sequences["sequence"].to_numpy()
        truth = measured_sequences["true_score"].to_numpy(dtype=np.float64)
        best = truth.max()
        cutoff = best * (1 - np.sign(best) * self.threshold)
        measured_idx = torch.from_numpy(s_utils.encode_sequences(list(seqs), self.alphabet)).to(dev)
        elite = measured_idx[torch.from_numpy(np.flatnonzero(truth >= cutoff)).to(dev)]
        width = int(self.rollout_width or self.sequences_batch_size)
        parents = elite[torch.arange(width, device=dev) % len(elite)].contiguous()     # np.resize(elite, width)
        budget, step = self.model_queries_per_batch, width
        seed = random.getrandbits(62)          # random.seed(...) makes a run reproducible
        rng = torch.Generator(device=dev)
        rng.manual_seed(seed & 0x7FFFFFFF)
        counter = [0]
        known = measured_idx                   # measured + found so far (what the reference keeps in a set and a dict)
        found_rows, found_scores = [], []
        cost_at_start = self.model.cost
        steps = 0

        def spent():
            return self.model.cost - cost_at_start

        while spent() < budget:
            for _ in range(self.rho):
                parents = self._recombine_device(parents, rng)
            roots = parents
            root_scores = self.model.get_fitness_device(roots)
            nodes, owner = roots, torch.arange(len(roots), device=dev)
            grew = False
            while len(nodes) > 0 and spent() + step < budget:
                children, ok = self._mutate_unseen(nodes, known, seed, counter)
                children, own = children[ok].contiguous(), owner[ok]
                if len(children) == 0:
                    break
                scores = self.model.get_fitness_device(children)
                found_rows.append(children); found_scores.append(scores)
                known = torch.cat([known, children])
                keep = scores >= root_scores[own]
                nodes, owner = children[keep].contiguous(), own[keep]
                grew = True
                steps += 1
            # every pass charges the roots, so the loop ends with the budget like the reference's; a space with nothing
            # left to propose ends it early instead of re-scoring the roots until then
            if not grew and self._space_exhausted(known):
                break
        if len(found_rows) == 0:
            raise ValueError(
                "No sequences generated. If `model_queries_per_batch` is small, try "
                "making `eval_batch_size` smaller"
            )
        rows, preds = torch.cat(found_rows).contiguous(), torch.cat(found_scores).contiguous()
        k = min(self.sequences_batch_size - 1, len(rows))    # np.argsort(preds)[: -B : -1] keeps B-1
        self.last_device_stats = {"rollout_steps": steps, "found": int(len(rows)), "model_queries": int(spent())}
        if k <= 0:
            return np.array([], dtype=str), np.array([], dtype=np.float32)
        top_s = torch.empty(k, dtype=torch.float32, device=dev)
        top_i = torch.empty(k, dtype=torch.int64, device=dev)
        work = torch.empty(_native.topk_select_workspace_bytes(), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _native.topk_select_dev(preds.data_ptr(), len(rows), k, 0, 0, 0, False, top_s.data_ptr(), top_i.data_ptr(), 0, 0,
                                    work.data_ptr(), torch.cuda.current_stream().cuda_stream)
        winners = rows[top_i].cpu().numpy()
        return s_utils.decode_indices(winners, self.alphabet), top_s.cpu().numpy()

    def _space_exhausted(self, known) -> bool:
        """Every sequence of the space is already measured or found (only reachable for tiny spaces such as the 65 536
        8-mers): the reference would re-draw forever (adalead.py:137-150); the device path stops proposing."""
        space = float(len(self.alphabet)) ** int(known.shape[1])
        return space <= float(len(known))

    def _recombine_population(self, gen):
        """Shuffle, then cross neighbouring pairs position by position (adalead.py:69-94)."""
        if len(gen) == 1:
            return gen
        random.shuffle(gen)
        offspring = []
        for first in range(0, len(gen) - 1, 2):
            mother, father = gen[first], gen[first + 1]
            a_chars, b_chars = [], []
            crossed = False
            for pos in range(len(mother)):
                if random.random() < self.recomb_rate:
                    crossed = not crossed
                if crossed:
                    a_chars.append(mother[pos]); b_chars.append(father[pos])
                else:
                    b_chars.append(mother[pos]); a_chars.append(father[pos])
            offspring.append("".join(a_chars))
            offspring.append("".join(b_chars))
        return offspring

    def propose_sequences(self, measured_sequences: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """Return the ``sequences_batch_size - 1`` best new sequences found by the rollouts
        (the reference's ``[: -B : -1]`` slice yields B-1 items, adalead.py:173)."""
        if self._use_device():
            return self._propose_device(measured_sequences)
        already_measured = set(measured_sequences["sequence"])
        best = measured_sequences["true_score"].max()
        cutoff = best * (1 - np.sign(best) * self.threshold)
        elite = measured_sequences["sequence"][measured_sequences["true_score"] >= cutoff].to_numpy()
        parents = np.resize(elite, self.sequences_batch_size)

        budget, step = self.model_queries_per_batch, self.eval_batch_size
        found = {}  # new sequence -> predicted fitness, in discovery order
        cost_at_start = self.model.cost

        def spent():
            return self.model.cost - cost_at_start

        while spent() < budget:
            for _ in range(self.rho):
                parents = self._recombine_population(parents)
            for lo in range(0, len(parents), step):
                roots = parents[lo: lo + step]
                root_scores = self.model.get_fitness(roots)
                frontier = list(enumerate(roots))
                while len(frontier) > 0 and spent() + step < budget:
                    owners, children = [], []
                    while len(children) < len(frontier):
                        # index -1 first: the rollout order of the reference (adalead.py:135)
                        owner, node = frontier[len(children) - 1]
                        child = s_utils.generate_random_mutant(node, self.mu * 1 / len(node), self.alphabet)
                        if child not in already_measured and child not in found:
                            owners.append(owner)
                            children.append(child)
                    child_scores = self.model.get_fitness(children)
                    found.update(zip(children, child_scores))
                    frontier = [(owner, child) for owner, child, score in zip(owners, children, child_scores)
                                if score >= root_scores[owner]]

        if len(found) == 0:
            raise ValueError(
                "No sequences generated. If `model_queries_per_batch` is small, try "
                "making `eval_batch_size` smaller"
            )
        new_seqs = np.array(list(found.keys()))
        preds = np.array(list(found.values()))
        order = np.argsort(preds)[: -self.sequences_batch_size: -1]
        return new_seqs[order], preds[order]
