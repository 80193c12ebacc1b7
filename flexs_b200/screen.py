"""Sharded virtual screen: score a candidate batch on every GPU, keep the global top-k.

This is the multi-GPU form of the selection every explorer ends with
(``np.argsort(preds)[: -B : -1]``, adalead.py:171-175 / cbas_dbas.py:197-201 / cmaes.py:117-122, and
``[::-1][:B]``, dyna_ppo.py:315-319).  Candidates are independent, so rank r of G scores the contiguous
block ``[r*N/G, (r+1)*N/G)`` with the fused surrogate kernel and selects its own top-k over DISTINCT sequences (the
reference ranks the keys of a dict) with one launch of ``flexs_topk_select_dev``, which writes (global index, score,
sequence) triples straight into the rank's message; the ranks exchange ONE ``all_gather`` of those messages —
``G*k*(12+L)`` bytes over NVLink — and every rank runs the same one-launch merge (``flexs_screen_merge_dev``), which
de-duplicates once more (a sequence may have reached the top-k of two shards).  There is no other collective on the
path; weights are replicated.

One process per GPU (``torch.distributed``, backend ``nccl``); the helpers that only move tensors work on
any backend, which is how the CPU test-suite exercises them with ``gloo`` at world size 2.
"""
from __future__ import annotations

from typing import Optional, Tuple

import contextlib

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block split of ``n`` candidates: rank r gets ``[start, stop)``; sizes differ by <= 1."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def message_bytes(k: int, seq_len: int) -> int:
    """Bytes of one rank's message: ``[k] int64 global index | [k] float32 score | [k][seq_len] uint8 rows``, padded to
    16 bytes (mirrors ``flexs_screen_message_bytes``; pure arithmetic so the CPU tests can use it)."""
    return (k * 12 + k * seq_len + 15) // 16 * 16


def message_views(msg, k: int, seq_len: int):
    """``(idx int64[k], scores float32[k], rows uint8[k, seq_len])`` views INTO a uint8 message buffer (any device):
    ``flexs_topk_select_dev`` writes its three outputs straight through them, nothing is packed afterwards."""
    import torch

    idx = msg[: 8 * k].view(torch.int64)
    scores = msg[8 * k: 12 * k].view(torch.float32)
    rows = msg[12 * k: 12 * k + k * seq_len].view(k, seq_len) if seq_len else None
    return idx, scores, rows


def all_gather_messages(msg, group=None):
    """The single collective of the path: every rank receives every rank's message (rank-major)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty(world * msg.numel(), dtype=msg.dtype, device=msg.device)
    dist.all_gather_into_tensor(out, msg, group=group)
    return out


def merge_reference(gathered: np.ndarray, world: int, k: int, seq_len: int):
    """numpy statement of what ``flexs_screen_merge_dev`` computes from the gathered messages: rank by (score desc,
    gathered position asc), drop absent entries (index < 0) and — with ``seq_len`` — later copies of a sequence, keep k.
    Used by the CPU tests (gloo) and as the checker of the GPU tests."""
    mb = message_bytes(k, seq_len)
    g = np.ascontiguousarray(gathered, dtype=np.uint8).reshape(world, mb)
    idx = np.concatenate([g[r, : 8 * k].view(np.int64) for r in range(world)])
    sc = np.concatenate([g[r, 8 * k: 12 * k].view(np.float32) for r in range(world)])
    rows = np.concatenate([g[r, 12 * k: 12 * k + k * seq_len].reshape(k, seq_len) for r in range(world)]) if seq_len else None
    pos = np.flatnonzero(idx >= 0)
    order = pos[np.lexsort((pos, -sc[pos].astype(np.float64)))]
    out, seen = [], set()
    for j in order:
        if rows is not None:
            key = rows[j].tobytes()
            if key in seen:
                continue
            seen.add(key)
        out.append(j)
        if len(out) == k:
            break
    out = np.array(out, dtype=np.int64)
    return sc[out], idx[out], (rows[out] if rows is not None else None)


class VirtualScreen:
    """Top-k of a (sharded) candidate batch under a B200 surrogate (``CNN`` / ``MLP`` / fused ``Ensemble``).

    ``k`` is the number of winners to return: explorers that reproduce the reference's ``[: -B : -1]``
    slice pass ``sequences_batch_size - 1``.

    Per call and rank: the surrogate's forward launches, ONE selection launch (``flexs_topk_select_dev``: top-k over
    distinct sequences, winners' rows included, written straight into the message), ONE ``all_gather`` of
    ``message_bytes(k, L)`` bytes per rank, ONE merge launch (``flexs_screen_merge_dev``).  Buffers are cached.

    ``overlap=True`` (world > 1): the all-gather and the merge launch of a screen run on a side stream over
    double-buffered messages, so they overlap the NEXT screen's forward pass and the ranks stop meeting in a collective
    between two forwards (at 8 GPUs the exposed gather + the wait for the slowest rank cost 1.6 ms of an 8.8 ms step).
    The tensors ``merge`` returns are then complete after :meth:`wait` (or a device synchronize) and are overwritten
    by the second screen after theirs.  (Measured: fine at 2 GPUs; at 8 the NCCL kernel, spinning on a side stream until the
    slowest rank arrives, keeps SMs from the persistent forward kernel: 7.2 -> 8.5 ms.)

    ``exchange="peer"`` (world > 1, one node): no collective at all.  After its selection launch a rank writes its message
    into a slot of every rank's mailbox over NVLink peer memory (``flexs_screen_push_dev``: plain stores + a release store of
    the step number; it never waits), and the merge of screen i is issued when screen i + 1 is (or at :meth:`wait`): a
    ``flexs_screen_wait_dev`` launch that holds the stream until the slot's flags show step i, then the same merge launch.
    Ranks drift by up to one screen instead of meeting after every forward.  Results as in overlap mode: complete after
    :meth:`wait`.  Every rank must issue the same sequence of screens.
    """

    PEER_DEPTH = 4   # mailbox slots (csrc/peer.cu: why 4 is enough for a lag of one screen)

    def __init__(self, model, k: int, group=None, unique: bool = True, overlap: bool = False, exchange: str = "nccl"):
        """``unique``: rank distinct sequences, a repeated candidate competing once through its first occurrence —
        what the reference's explorers do by keeping scores in a dict (adalead.py:157, cmaes.py:112-115,
        dyna_ppo.py:310-314).  ``unique=False`` ranks rows."""
        if not hasattr(model, "get_fitness_device"):
            raise TypeError("VirtualScreen needs a B200 surrogate (CNN, MLP or an Ensemble of identical ones)")
        self.model, self.k, self.group, self.unique = model, int(k), group, bool(unique)
        self.overlap = bool(overlap)
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' or 'peer'")
        self.exchange = exchange
        self._peer = {}      # peer mode: (device, seq_len, world) -> mailbox state
        self._pending = None  # peer mode: the screen whose merge has not been issued yet
        self._calls = 0      # screens started: message slot of the current one = (_calls - 1) & 1 in overlap mode
        self._side = None    # side stream of the overlap mode
        self._buf = {}
        self.fallbacks = 0   # selections that needed the full hash de-duplication (almost-all-repeats batches)
        self.launches = 0    # selection + merge kernels launched by this object (the surrogate counts its own)
        self.forward_events = None   # set to a list: local_topk appends a (start, end) CUDA event pair around the forward

    def _world(self) -> Tuple[int, int]:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def _buffers(self, device, seq_len: int, world: int):
        import torch

        from flexs_b200 import _native

        slot = (self._calls - 1) & 1 if (self.overlap and world > 1) else 0
        key = (device, seq_len, world, slot)
        if key not in self._buf:
            mb = message_bytes(self.k, seq_len)
            self._buf[key] = dict(
                msg=torch.zeros(mb, dtype=torch.uint8, device=device),
                work=torch.empty(_native.topk_select_workspace_bytes(), dtype=torch.uint8, device=device),
                status=torch.zeros(8, dtype=torch.int32, device=device),   # [0] = fell short; [1..4] diagnostics
                gathered=torch.empty(world * mb, dtype=torch.uint8, device=device) if world > 1 else None,
                fin=torch.zeros(mb, dtype=torch.uint8, device=device) if world > 1 else None,
                merged=None)   # overlap mode: event recorded on the side stream after this slot's merge launch
        return self._buf[key]

    def wait(self):
        """Overlap mode: make the current stream wait for every all-gather + merge issued so far."""
        import torch

        if self._pending is not None:
            self._issue_peer_merge()
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

    def _peer_state(self, device, seq_len: int, rank: int, world: int):
        """Mailbox of this rank + the mapped mailboxes of its peers (cudaIpc handles exchanged once per shape)."""
        import torch
        import torch.distributed as dist

        from flexs_b200 import _native

        key = (device, seq_len, world)
        if key not in self._peer:
            mb = message_bytes(self.k, seq_len)
            own, bases, err = None, [], None
            with torch.cuda.device(device):
                # Any rank may fail to allocate / map (no peer access, IPC not permitted in this container): all ranks
                # must then agree to use the collective instead, so failures are exchanged, not raised.
                try:
                    own, handle = _native.peer_alloc(_native.peer_mailbox_bytes(mb, world, self.PEER_DEPTH))
                except Exception as exc:  # noqa: BLE001
                    handle, err = None, exc
                handles = [None] * world
                dist.all_gather_object(handles, handle, group=self.group)
                if all(h is not None for h in handles):
                    try:
                        bases = [own if r == rank else _native.peer_open(handles[r]) for r in range(world)]
                    except Exception as exc:  # noqa: BLE001
                        err = exc
                else:
                    err = err or RuntimeError("a peer could not allocate its mailbox")
                oks = [None] * world
                dist.all_gather_object(oks, err is None, group=self.group)
            if not all(oks):
                for b in bases:
                    if b != own:
                        _native.peer_close(b)
                if own is not None:
                    _native.peer_free(own)
                import warnings

                warnings.warn(f"VirtualScreen: peer-memory exchange unavailable ({err}); using the NCCL all_gather")
                self.exchange = "nccl"
                return None
            self._peer[key] = dict(own=own, bases=bases, seq=0, mb=mb,
                                   d_bases=torch.tensor(bases, dtype=torch.int64, device=device),
                                   status=torch.zeros(1, dtype=torch.int32, device=device),
                                   fin=[torch.zeros(mb, dtype=torch.uint8, device=device) for _ in range(2)])
        return self._peer[key]

    def _issue_peer_merge(self):
        import torch

        from flexs_b200 import _native

        st, seq, seq_len, device, world = self._pending
        self._pending = None
        slot = seq % self.PEER_DEPTH
        fin_i, fin_s, fin_rows = message_views(st["fin"][seq & 1], self.k, seq_len)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            _native.screen_wait_dev(st["own"], st["mb"], world, slot, self.PEER_DEPTH, seq, st["status"].data_ptr(), stream)
            _native.screen_merge_dev(st["own"] + slot * world * st["mb"], world, self.k, seq_len if self.unique else 0,
                                     fin_s.data_ptr(), fin_i.data_ptr(), fin_rows.data_ptr() if self.unique else 0, stream)
        self.launches += 2

    def close(self):
        """Peer mode: unmap the peers' mailboxes and free this rank's (collective: every rank calls it)."""
        import torch
        import torch.distributed as dist

        from flexs_b200 import _native

        self.wait()
        for (device, _, world), st in self._peer.items():
            torch.cuda.synchronize(device)
            dist.barrier(group=self.group)
            for b in st["bases"]:
                if b != st["own"]:
                    _native.peer_close(b)
            dist.barrier(group=self.group)
            _native.peer_free(st["own"])
        self._peer = {}

    def local_topk(self, idx, index_offset: int = 0, check: bool = True):
        """Score ``uint8[n, L]`` residue indices resident on this GPU and select the local top-k.

        Returns ``(top_scores[k], top_idx[k], scores[n])`` (CUDA tensors; the first two are views into this rank's
        message, whose third part holds the winners' rows).  Charges ``model.cost`` like ``get_fitness`` does.
        ``check`` reads one int back (a stream sync) to learn whether the lazy de-duplication found k distinct
        sequences among the best rows and, if not, re-selects after hashing the whole batch; callers that must not
        sync pass ``check=False`` and inspect ``last_status`` themselves."""
        import torch

        from flexs_b200 import _native

        idx = idx.contiguous()
        self._calls += 1
        if self.forward_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        scores = self.model.get_fitness_device(idx)
        if self.forward_events is not None:
            ev[1].record()
            self.forward_events.append(ev)
        n, L = int(idx.shape[0]), int(idx.shape[1])
        dev = scores.device
        buf = self._buffers(dev, L, self._world()[1])
        top_i, top_s, top_rows = message_views(buf["msg"], self.k, L)
        with torch.cuda.device(dev):
            if buf["merged"] is not None:   # the all-gather that last read this message slot (two screens ago) is done
                torch.cuda.current_stream().wait_event(buf["merged"])
            stream = torch.cuda.current_stream().cuda_stream
            _native.topk_select_dev(scores.data_ptr(), n, self.k, index_offset, idx.data_ptr(), L, self.unique,
                                    top_s.data_ptr(), top_i.data_ptr(), top_rows.data_ptr(), buf["status"].data_ptr(),
                                    buf["work"].data_ptr(), stream)
            self.last_status = buf["status"]
            self.launches += 1
            if check and self.unique and int(buf["status"][0].item()) != 0:
                # almost every row of the batch is a repeat: hash all rows (dedup.cu), then select among first occurrences
                self.fallbacks += 1
                ranked = torch.empty_like(scores)
                dwork = torch.empty(_native.dedup_workspace_bytes(n), dtype=torch.uint8, device=dev)
                _native.dedup_scores_dev(idx.data_ptr(), n, L, scores.data_ptr(), ranked.data_ptr(), dwork.data_ptr(), stream)
                _native.topk_select_dev(ranked.data_ptr(), n, self.k, index_offset, idx.data_ptr(), L, False,
                                        top_s.data_ptr(), top_i.data_ptr(), top_rows.data_ptr(), 0, buf["work"].data_ptr(), stream)
                absent = torch.isinf(top_s) & (top_s < 0)     # repeats carry -inf: fewer than k distinct sequences
                top_i.masked_fill_(absent, -1)
        return top_s, top_i, scores

    def merge(self, seq_len: int, device):
        """Exchange this rank's message with the other ranks (one all_gather, or the peer-memory mailboxes of
        ``exchange="peer"``) and reduce the ``world`` lists to the global top-k (identical on every rank): returns
        ``(scores[k], indices[k])`` views of the merged message — complete on return of the launches in the default
        mode, after :meth:`wait` in the overlap and peer modes."""
        import torch

        from flexs_b200 import _native

        rank, world = self._world()
        buf = self._buffers(device, seq_len, world)
        if world == 1:
            top_i, top_s, _ = message_views(buf["msg"], self.k, seq_len)
            return top_s, top_i
        import torch.distributed as dist

        st = self._peer_state(device, seq_len, rank, world) if self.exchange == "peer" else None
        if st is not None:
            if self._pending is not None:
                self._issue_peer_merge()   # the previous screen's messages have had a whole forward to arrive
            st["seq"] += 1
            seq = st["seq"]
            with torch.cuda.device(device):
                _native.screen_push_dev(buf["msg"].data_ptr(), st["mb"], rank, world, seq % self.PEER_DEPTH, self.PEER_DEPTH, seq,
                                        st["d_bases"].data_ptr(), torch.cuda.current_stream().cuda_stream)
            self.launches += 1
            self._pending = (st, seq, seq_len, device, world)
            fin_i, fin_s, _ = message_views(st["fin"][seq & 1], self.k, seq_len)
            return fin_s, fin_i
        fin_i, fin_s, fin_rows = message_views(buf["fin"], self.k, seq_len)
        with torch.cuda.device(device):
            if self.overlap:
                if self._side is None:
                    self._side = torch.cuda.Stream(device=device)
                self._side.wait_stream(torch.cuda.current_stream())   # the selection launch that wrote the message
                ctx = torch.cuda.stream(self._side)
            else:
                ctx = contextlib.nullcontext()
            with ctx:
                dist.all_gather_into_tensor(buf["gathered"], buf["msg"], group=self.group)   # the single collective of the path
                _native.screen_merge_dev(buf["gathered"].data_ptr(), world, self.k, seq_len if self.unique else 0,
                                         fin_s.data_ptr(), fin_i.data_ptr(), fin_rows.data_ptr() if self.unique else 0,
                                         torch.cuda.current_stream().cuda_stream)
                if self.overlap:
                    buf["merged"] = torch.cuda.Event()
                    buf["merged"].record()
        self.launches += 1
        return fin_s, fin_i

    def screen_indices(self, idx_local, index_offset: int = 0, check: bool = True):
        """``idx_local``: this rank's shard (CUDA ``uint8[n_local, L]``) whose first row has global index
        ``index_offset``.  Returns the global ``(scores[k], indices[k])`` as CUDA tensors."""
        self.local_topk(idx_local, index_offset, check)
        out = self.merge(int(idx_local.shape[1]), idx_local.device)
        self.wait()   # a synchronous entry: the caller reads the result next
        return out

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
        dev_index = getattr(self.model, "device", None)
        if dev_index is None and hasattr(self.model, "models"):
            dev_index = getattr(self.model.models[0], "device", None)
        device = torch.device("cuda", torch.cuda.current_device() if dev_index is None else dev_index)
        shard = torch.from_numpy(np.ascontiguousarray(idx[start:stop])).to(device)
        top_s, top_i = self.screen_indices(shard, start)
        top_s, top_i = top_s.cpu().numpy(), top_i.cpu().numpy()
        keep = top_i >= 0
        top_s, top_i = top_s[keep], top_i[keep]
        if isinstance(sequences, np.ndarray) and sequences.dtype == np.uint8:
            return top_i, top_s
        return np.asarray(sequences)[top_i], top_s
