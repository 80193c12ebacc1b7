"""Sharded virtual screen: score a candidate batch on every GPU, keep the global top-k.

This is the multi-GPU form of the selection every explorer ends with
(``np.argsort(preds)[: -B : -1]``, adalead.py:171-175 / cbas_dbas.py:197-201 / cmaes.py:117-122, and
``[::-1][:B]``, dyna_ppo.py:315-319).  Candidates are independent, so rank r of G scores the contiguous
block ``[r*N/G, (r+1)*N/G)`` with the fused surrogate kernel, selects its own top-k with
``flexs_topk_dev`` (indices offset to global positions) and the ranks exchange ONE
``all_gather`` of ``k`` (score, index) pairs — ``G*k*16`` bytes (packed as int64 pairs) over NVLink — before
every rank runs the same final merge.  There is no other collective on the path; weights are replicated.

One process per GPU (``torch.distributed``, backend ``nccl``); the helpers that only move tensors work on
any backend, which is how the CPU test-suite exercises them with ``gloo`` at world size 2.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block split of ``n`` candidates: rank r gets ``[start, stop)``; sizes differ by <= 1."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_topk(scores, idx):
    """``float32[k]`` scores + ``int64[k]`` indices -> one ``int64[2k]`` message (indices, then score bits)."""
    import torch

    k = scores.shape[0]
    out = torch.empty(2 * k, dtype=torch.int64, device=scores.device)
    out[:k] = idx
    out[k:] = scores.contiguous().view(torch.int32).to(torch.int64)
    return out


def unpack_topk(gathered, world: int, k: int):
    """Inverse of :func:`pack_topk` for the concatenation of ``world`` messages: ``(scores[world*k], idx[world*k])``."""
    import torch

    g = gathered.view(world, 2, k)
    idx = g[:, 0, :].reshape(-1).contiguous()
    scores = g[:, 1, :].reshape(-1).to(torch.int32).contiguous().view(torch.float32)
    return scores, idx


def all_gather_topk(message, group=None):
    """The single collective of the path: every rank receives every rank's packed top-k list."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty(world * message.numel(), dtype=message.dtype, device=message.device)
    dist.all_gather_into_tensor(out, message, group=group)
    return out


class VirtualScreen:
    """Top-k of a (sharded) candidate batch under a B200 surrogate (``CNN`` / ``MLP`` / fused ``Ensemble``).

    ``k`` is the number of winners to return: explorers that reproduce the reference's ``[: -B : -1]``
    slice pass ``sequences_batch_size - 1``.
    """

    def __init__(self, model, k: int, group=None):
        if not hasattr(model, "get_fitness_device"):
            raise TypeError("VirtualScreen needs a B200 surrogate (CNN, MLP or an Ensemble of identical ones)")
        self.model, self.k, self.group = model, int(k), group

    def _world(self) -> Tuple[int, int]:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def local_topk(self, idx, index_offset: int = 0):
        """Score ``uint8[n, L]`` residue indices resident on this GPU and select the local top-k.

        Returns ``(top_scores[k], top_idx[k], scores[n])`` (CUDA tensors, no host sync).  Charges
        ``model.cost`` like ``get_fitness`` does.
        """
        import torch

        from flexs_b200 import _native

        scores = self.model.get_fitness_device(idx)
        n = int(scores.shape[0])
        dev = scores.device
        work = torch.empty(_native.topk_workspace_bytes(n, self.k), dtype=torch.uint8, device=dev)
        top_s = torch.empty(self.k, dtype=torch.float32, device=dev)
        top_i = torch.empty(self.k, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _native.topk_dev(scores.data_ptr(), n, self.k, index_offset, 0, top_s.data_ptr(), top_i.data_ptr(),
                             work.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return top_s, top_i, scores

    def merge(self, top_s, top_i):
        """All-gather the per-shard lists and reduce them to the global top-k (identical on every rank).

        Shards own increasing index ranges and each list is already ordered (score desc, index asc), so
        breaking score ties by position in the gathered array equals breaking them by global index."""
        import torch

        from flexs_b200 import _native

        rank, world = self._world()
        if world == 1:
            return top_s, top_i
        gathered = all_gather_topk(pack_topk(top_s, top_i), self.group)
        g_scores, g_idx = unpack_topk(gathered, world, self.k)
        dev = g_scores.device
        work = torch.empty(_native.topk_workspace_bytes(world * self.k, self.k), dtype=torch.uint8, device=dev)
        fin_s = torch.empty(self.k, dtype=torch.float32, device=dev)
        fin_i = torch.empty(self.k, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _native.topk_dev(g_scores.data_ptr(), world * self.k, self.k, 0, g_idx.data_ptr(), fin_s.data_ptr(),
                             fin_i.data_ptr(), work.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return fin_s, fin_i

    def screen_indices(self, idx_local, index_offset: int = 0):
        """``idx_local``: this rank's shard (CUDA ``uint8[n_local, L]``) whose first row has global index
        ``index_offset``.  Returns the global ``(scores[k], indices[k])`` as CUDA tensors."""
        top_s, top_i, _ = self.local_topk(idx_local, index_offset)
        return self.merge(top_s, top_i)

    def screen(self, sequences, alphabet: Optional[str] = None):
        """Host entry: every rank passes the SAME full candidate list (strings or ``uint8[N, L]`` indices);
        each scores its own block.  Returns ``(np.ndarray[str] | index array, np.float32 scores)`` of the
        k winners, best first."""
        import torch

        from flexs_b200.utils import sequence_utils as s_utils

        if alphabet is None:
            alphabet = self.model.alphabet if hasattr(self.model, "alphabet") else self.model.models[0].alphabet
        idx = sequences if (isinstance(sequences, np.ndarray) and sequences.dtype == np.uint8) \
            else s_utils.encode_sequences(sequences, alphabet)
        rank, world = self._world()
        start, stop = shard_bounds(len(idx), rank, world)
        device = torch.device("cuda", torch.cuda.current_device())
        shard = torch.from_numpy(np.ascontiguousarray(idx[start:stop])).to(device)
        top_s, top_i = self.screen_indices(shard, start)
        top_s, top_i = top_s.cpu().numpy(), top_i.cpu().numpy()
        keep = top_i >= 0
        top_s, top_i = top_s[keep], top_i[keep]
        if isinstance(sequences, np.ndarray) and sequences.dtype == np.uint8:
            return top_i, top_s
        return np.asarray(sequences)[top_i], top_s
