"""CPU, world_size 2 over gloo: the host-side plumbing of the sharded virtual screen — block partition,
message packing, the single all-gather and the unpacking.  (The kernels themselves need a GPU; the final
merge is re-done here in numpy with the documented tie rule to pin what every rank must agree on.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flexs_b200 import screen


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 100, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [screen.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip_bit_exact():
    k, world = 7, 3
    rng = np.random.default_rng(0)
    msgs, all_s, all_i = [], [], []
    for r in range(world):
        s = torch.from_numpy(rng.normal(size=k).astype(np.float32))
        s[0] = float("-inf"); s[1] = -0.0
        i = torch.from_numpy(rng.integers(-1, 1 << 40, size=k))
        msgs.append(screen.pack_topk(s, i)); all_s.append(s); all_i.append(i)
    s2, i2 = screen.unpack_topk(torch.cat(msgs), world, k)
    assert torch.equal(i2, torch.cat(all_i))
    assert torch.equal(s2.view(torch.int32), torch.cat(all_s).view(torch.int32))


def test_pack_unpack_with_sequences_roundtrip():
    """The unique screen ships the winners' residues in the same message (one collective): indices, score bits,
    then k*L bytes padded to int64 words."""
    k, world, L = 5, 3, 13   # 65 bytes: not a multiple of 8
    rng = np.random.default_rng(1)
    msgs, all_s, all_i, all_q = [], [], [], []
    for r in range(world):
        s = torch.from_numpy(rng.normal(size=k).astype(np.float32))
        i = torch.from_numpy(rng.integers(-1, 1 << 40, size=k))
        q = torch.from_numpy(rng.integers(0, 20, size=(k, L), dtype=np.uint8))
        msgs.append(screen.pack_topk(s, i, q)); all_s.append(s); all_i.append(i); all_q.append(q)
    assert msgs[0].numel() == 2 * k + 9
    s2, i2, q2 = screen.unpack_topk(torch.cat(msgs), world, k, L)
    assert torch.equal(i2, torch.cat(all_i)) and torch.equal(q2, torch.cat(all_q))
    assert torch.equal(s2.view(torch.int32), torch.cat(all_s).view(torch.int32))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scores_all = np.random.default_rng(123).integers(-40, 40, size=n).astype(np.float32) / 4  # many ties
        start, stop = screen.shard_bounds(n, rank, world)
        local = scores_all[start:stop]
        order = np.lexsort((np.arange(len(local)), -local.astype(np.float64)))[:k]  # what flexs_topk_dev returns
        top_s = torch.full((k,), float("-inf")); top_i = torch.full((k,), -1, dtype=torch.int64)
        top_s[: len(order)] = torch.from_numpy(local[order]); top_i[: len(order)] = torch.from_numpy(order + start)
        gathered = screen.all_gather_topk(screen.pack_topk(top_s, top_i))
        g_s, g_i = screen.unpack_topk(gathered, world, k)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), s=g_s.numpy(), i=g_i.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(1000, 16), (5, 8)])
def test_all_gather_of_shard_topk_world2(tmp_path, n, k):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    np.testing.assert_array_equal(r0["s"], r1["s"])   # every rank holds the same gathered lists
    np.testing.assert_array_equal(r0["i"], r1["i"])
    # the merge every rank then performs (score desc, position asc == global index asc) equals the
    # single-process selection over the whole batch
    scores_all = np.random.default_rng(123).integers(-40, 40, size=n).astype(np.float32) / 4
    want = np.lexsort((np.arange(n), -scores_all.astype(np.float64)))[: min(k, n)]
    valid = r0["i"] >= 0
    pos = np.arange(len(r0["s"]))[valid]
    merged = pos[np.lexsort((pos, -r0["s"][valid].astype(np.float64)))][: min(k, n)]
    np.testing.assert_array_equal(r0["i"][merged], want)
