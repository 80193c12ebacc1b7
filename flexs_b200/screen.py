"""Sharded virtual screen: score a candidate batch on every GPU, keep the global top-k.

This is the multi-GPU form of the selection every explorer ends with
(``np.argsort(preds)[: -B : -1]``, adalead.py:171-175 / cbas_dbas.py:197-201 / cmaes.py:117-122, and
``[::-1][:B]``, dyna_ppo.py:315-319).  Candidates are independent, so rank r of G scores the contiguous
block ``[r*N/G, (r+1)*N/G)`` with the fused surrogate kernel, drops repeated sequences (the reference ranks the
keys of a dict: ``flexs_dedup_scores_dev``), selects its own top-k with ``flexs_topk_dev`` (indices offset to global
positions) and the ranks exchange ONE ``all_gather`` of ``k`` (score, index, sequence) triples — ``G*k*(16+L)``
bytes over NVLink — before every rank runs the same final merge, which de-duplicates once more (a sequence may have
reached the top-k of two shards).  There is no other collective on the path; weights are replicated.

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


def _seq_words(k: int, seq_len: int) -> int:
    """int64 words that hold ``k`` sequences of ``seq_len`` bytes."""
    return (k * seq_len + 7) // 8


def pack_topk(scores, idx, seqs=None):
    """``float32[k]`` scores + ``int64[k]`` indices (+ ``uint8[k, L]`` winner sequences) -> one int64 message:
    indices, score bits, then the sequence bytes padded to a whole number of words."""
    import torch

    k = scores.shape[0]
    extra = 0 if seqs is None else _seq_words(k, seqs.shape[1])
    out = torch.zeros(2 * k + extra, dtype=torch.int64, device=scores.device)
    out[:k] = idx
    out[k:2 * k] = scores.contiguous().view(torch.int32).to(torch.int64)
    if seqs is not None:
        out[2 * k:].view(torch.uint8)[: k * seqs.shape[1]] = seqs.contiguous().view(-1)
    return out


def unpack_topk(gathered, world: int, k: int, seq_len: int = 0):
    """Inverse of :func:`pack_topk` for the concatenation of ``world`` messages:
    ``(scores[world*k], idx[world*k])`` and, with ``seq_len``, the ``uint8[world*k, seq_len]`` sequences."""
    import torch

    extra = _seq_words(k, seq_len) if seq_len else 0
    g = gathered.view(world, 2 * k + extra)
    idx = g[:, :k].reshape(-1).contiguous()
    scores = g[:, k:2 * k].reshape(-1).to(torch.int32).contiguous().view(torch.float32)
    if not seq_len:
        return scores, idx
    seqs = g[:, 2 * k:].contiguous().view(torch.uint8).view(world, extra * 8)[:, : k * seq_len].reshape(world * k, seq_len)
    return scores, idx, seqs.contiguous()


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

    def __init__(self, model, k: int, group=None, unique: bool = True):
        """``unique``: rank distinct sequences, a repeated candidate competing once through its first occurrence —
        what the reference's explorers do by keeping scores in a dict (adalead.py:157, cmaes.py:112-115,
        dyna_ppo.py:310-314).  ``unique=False`` ranks rows."""
        if not hasattr(model, "get_fitness_device"):
            raise TypeError("VirtualScreen needs a B200 surrogate (CNN, MLP or an Ensemble of identical ones)")
        self.model, self.k, self.group, self.unique = model, int(k), group, bool(unique)

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
            stream = torch.cuda.current_stream().cuda_stream
            ranked = scores
            if self.unique:
                idx = idx.contiguous()
                ranked = torch.empty_like(scores)
                dwork = torch.empty(_native.dedup_workspace_bytes(n), dtype=torch.uint8, device=dev)
                _native.dedup_scores_dev(idx.data_ptr(), n, int(idx.shape[1]), scores.data_ptr(), ranked.data_ptr(),
                                         dwork.data_ptr(), stream)
            _native.topk_dev(ranked.data_ptr(), n, self.k, index_offset, 0, top_s.data_ptr(), top_i.data_ptr(),
                             work.data_ptr(), stream)
            if self.unique:
                top_i = torch.where(torch.isinf(top_s), torch.full_like(top_i, -1), top_i)  # fewer than k distinct
        return top_s, top_i, scores

    def merge(self, top_s, top_i, top_seqs=None):
        """All-gather the per-shard lists and reduce them to the global top-k (identical on every rank).

        Shards own increasing index ranges and each list is already ordered (score desc, index asc), so
        breaking score ties by position in the gathered array equals breaking them by global index.  With
        ``top_seqs`` (the winners' residues, ``uint8[k, L]``) the gathered list is de-duplicated first: the copy
        from the lower rank, i.e. the lower global index, survives."""
        import torch

        from flexs_b200 import _native

        rank, world = self._world()
        if world == 1:
            return top_s, top_i
        if top_seqs is None:
            gathered = all_gather_topk(pack_topk(top_s, top_i), self.group)
            g_scores, g_idx = unpack_topk(gathered, world, self.k)
        else:
            L = int(top_seqs.shape[1])
            gathered = all_gather_topk(pack_topk(top_s, top_i, top_seqs), self.group)
            g_scores, g_idx, g_seqs = unpack_topk(gathered, world, self.k, L)
            m = world * self.k
            dwork = torch.empty(_native.dedup_workspace_bytes(m), dtype=torch.uint8, device=g_scores.device)
            with torch.cuda.device(g_scores.device):
                _native.dedup_scores_dev(g_seqs.data_ptr(), m, L, g_scores.data_ptr(), g_scores.data_ptr(),
                                         dwork.data_ptr(), torch.cuda.current_stream().cuda_stream)
        dev = g_scores.device
        work = torch.empty(_native.topk_workspace_bytes(world * self.k, self.k), dtype=torch.uint8, device=dev)
        fin_s = torch.empty(self.k, dtype=torch.float32, device=dev)
        fin_i = torch.empty(self.k, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _native.topk_dev(g_scores.data_ptr(), world * self.k, self.k, 0, g_idx.data_ptr(), fin_s.data_ptr(),
                             fin_i.data_ptr(), work.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if top_seqs is not None:
            fin_i = torch.where(torch.isinf(fin_s), torch.full_like(fin_i, -1), fin_i)
        return fin_s, fin_i

    def screen_indices(self, idx_local, index_offset: int = 0):
        """``idx_local``: this rank's shard (CUDA ``uint8[n_local, L]``) whose first row has global index
        ``index_offset``.  Returns the global ``(scores[k], indices[k])`` as CUDA tensors."""
        top_s, top_i, _ = self.local_topk(idx_local, index_offset)
        if not self.unique or self._world()[1] == 1:
            return self.merge(top_s, top_i)
        if idx_local.shape[0] == 0:                         # an empty shard (fewer candidates than ranks)
            import torch

            seqs = torch.zeros((self.k, idx_local.shape[1]), dtype=torch.uint8, device=idx_local.device)
        else:
            rows = (top_i - index_offset).clamp(min=0)      # absent winners (-1) borrow row 0; their score is -inf
            seqs = idx_local[rows].contiguous()
        return self.merge(top_s, top_i, seqs)

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
