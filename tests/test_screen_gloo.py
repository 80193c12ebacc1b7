"""CPU, world_size 2 over gloo: the host-side plumbing of the sharded virtual screen — block partition, message
layout, the single all-gather.  (The kernels themselves need a GPU; the final merge is done here by
screen.merge_reference, the numpy statement of flexs_screen_merge_dev, which the GPU tests check the kernel against.)"""
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


def test_message_layout_views_and_reference_merge():
    """One rank's message is [k] int64 index | [k] float32 score | [k][L] rows, padded to 16 bytes; the selection kernel
    writes through views into it, the merge reads the all-gathered concatenation.  Pins the layout (bit-exact through
    -inf, -0.0 and 40-bit indices) and the merge rule (score desc, gathered position asc, later copies of a sequence and
    absent entries dropped) in numpy — what flexs_screen_merge_dev must reproduce on every rank."""
    k, world, L = 5, 3, 13
    assert screen.message_bytes(k, L) == 128 and screen.message_bytes(k, 0) == 64 and screen.message_bytes(99, 100) % 16 == 0
    rng = np.random.default_rng(1)
    msgs, all_s, all_i, all_q = [], [], [], []
    for r in range(world):
        msg = torch.zeros(screen.message_bytes(k, L), dtype=torch.uint8)
        idx_v, sc_v, rows_v = screen.message_views(msg, k, L)
        s = torch.from_numpy(np.sort(rng.normal(size=k).astype(np.float32))[::-1].copy())
        i = torch.from_numpy(np.sort(rng.integers(r << 38, (r + 1) << 38, size=k)))
        q = torch.from_numpy(rng.integers(0, 20, size=(k, L), dtype=np.uint8))
        if r == 1:
            s[-1] = float("-inf"); i[-1] = -1            # an absent winner
            s[0] = all_s[0][2]; q[0] = all_q[0][2]       # a sequence that also reached rank 0's list, same score
        sc_v.copy_(s); idx_v.copy_(i); rows_v.copy_(q)
        msgs.append(msg); all_s.append(s); all_i.append(i); all_q.append(q)
    gathered = torch.cat(msgs).numpy()
    fs, fi, fq = screen.merge_reference(gathered, world, k, L)
    cat_s, cat_i = torch.cat(all_s).numpy(), torch.cat(all_i).numpy()
    assert len(fs) == k and np.all(np.diff(fs) <= 0) and np.all(fi >= 0)
    assert all_i[1][0].item() not in fi.tolist()         # the repeated sequence is reported through rank 0's copy only
    for sc, gi in zip(fs, fi):                            # every winner is one of the messages' (score, index) pairs, bit-exact
        j = int(np.flatnonzero(cat_i == gi)[0])
        assert np.float32(sc).view(np.int32) == cat_s[j].view(np.int32)
    assert len({bytes(r) for r in fq}) == k
    # without rows: plain ranking of the valid entries
    m2 = torch.zeros(screen.message_bytes(k, 0), dtype=torch.uint8)
    i2, s2, r2 = screen.message_views(m2, k, 0)
    assert r2 is None
    s2.copy_(torch.tensor([3.0, 2.0, -0.0, 0.0, float("-inf")])); i2.copy_(torch.tensor([7, 1, 9, 4, -1]))
    fs, fi, _ = screen.merge_reference(m2.numpy(), 1, k, 0)
    assert fi.tolist() == [7, 1, 9, 4] and np.signbit(fs[2])


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
        msg = torch.zeros(screen.message_bytes(k, 0), dtype=torch.uint8)
        top_i, top_s, _ = screen.message_views(msg, k, 0)
        top_s.fill_(float("-inf")); top_i.fill_(-1)
        top_s[: len(order)] = torch.from_numpy(local[order]); top_i[: len(order)] = torch.from_numpy(order + start)
        gathered = screen.all_gather_messages(msg)                       # the single collective of the path
        f_s, f_i, _ = screen.merge_reference(gathered.numpy(), world, k, 0)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), g=gathered.numpy(), s=f_s, i=f_i)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(1000, 16), (5, 8)])
def test_all_gather_of_shard_topk_world2(tmp_path, n, k):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    np.testing.assert_array_equal(r0["g"], r1["g"])   # every rank holds the same gathered messages
    np.testing.assert_array_equal(r0["i"], r1["i"])
    np.testing.assert_array_equal(r0["s"], r1["s"])
    # the merge every rank then performs (score desc, position asc == global index asc) equals the
    # single-process selection over the whole batch
    scores_all = np.random.default_rng(123).integers(-40, 40, size=n).astype(np.float32) / 4
    want = np.lexsort((np.arange(n), -scores_all.astype(np.float64)))[: min(k, n)]
    np.testing.assert_array_equal(r0["i"], want)
    np.testing.assert_array_equal(r0["s"], scores_all[want])


def _stats_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import json
        import sys
        from pathlib import Path

        sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
        import bench

        got = bench.gather_rank_stats({"rank": rank, "forward_ms": 7.0 + rank}, world)
        with open(os.path.join(out_dir, f"s{rank}.json"), "w") as f:
            json.dump(got, f)
    finally:
        dist.destroy_process_group()


def test_bench_per_rank_statistics_world2(tmp_path):
    """bench.py's N > 1 line carries every rank's forward time (the step follows the slowest GPU): the gather helper."""
    import json

    world = 2
    mp.spawn(_stats_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = json.load(open(tmp_path / f"s{r}.json"))
        assert got == [{"rank": 0, "forward_ms": 7.0}, {"rank": 1, "forward_ms": 8.0}]
