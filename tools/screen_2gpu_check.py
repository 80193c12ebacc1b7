"""torchrun --nproc-per-node 2 tools/screen_2gpu_check.py — the sharded unique virtual screen against numpy.
Candidates repeat inside and ACROSS the shards; every rank must return the same k distinct winners, each through its
lowest global index, with the scores of a single-process ranking of the de-duplicated list."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

import flexs_b200 as flexs
from flexs_b200.screen import VirtualScreen
from flexs_b200.utils import sequence_utils as su

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl")
ok = True
for L, n, B in ((8, 300_001, 100), (100, 40_000, 50)):
    cnn = flexs.baselines.models.CNN(L, 32, 100, su.DNAA, seed=5, device=torch.cuda.current_device())
    rng = np.random.default_rng(11)
    idx = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
    if L == 100:                      # plant copies of first-half rows in the second half (the other shard)
        src = rng.integers(0, n // 2, size=n // 4)
        idx[n - len(src):] = idx[src]
    top_i, top_s = VirtualScreen(cnn, k=B - 1).screen(idx)
    cnn2 = flexs.baselines.models.CNN(L, 32, 100, su.DNAA, seed=5, device=torch.cuda.current_device())
    scores = cnn2.get_fitness(idx)
    _, first = np.unique(idx, axis=0, return_index=True)
    first = np.sort(first)
    order = first[np.lexsort((first, -scores[first].astype(np.float64)))][: B - 1]
    same = np.array_equal(top_i, order) and np.array_equal(top_s, scores[order])
    print(f"rank {rank}: L={L} n={n} k={B - 1}: {'OK' if same else 'MISMATCH'}", flush=True)
    ok = ok and same
# overlap mode: the all-gather + merge of screen i run on a side stream under the forward of screen i + 1; results are
# read one screen late (double-buffered messages) and must equal the synchronous screen's
L, n, B = 100, 60_000, 100
cnn = flexs.baselines.models.CNN(L, 32, 100, su.DNAA, seed=7, device=torch.cuda.current_device())
rng = np.random.default_rng(3)
batches = [rng.integers(0, 4, size=(n, L), dtype=np.uint8) for _ in range(5)]
lo, hi = rank * n // world, (rank + 1) * n // world
shards = [torch.from_numpy(b[lo:hi]).cuda() for b in batches]
sync_vs, pipe_vs = VirtualScreen(cnn, k=B - 1), VirtualScreen(cnn, k=B - 1, overlap=True)
want = []
for sh in shards:
    s_, i_ = sync_vs.screen_indices(sh, lo)
    want.append((s_.cpu().numpy().copy(), i_.cpu().numpy().copy()))
pending, same = None, True
for j, sh in enumerate(shards + [None]):
    if sh is not None:
        pipe_vs.local_topk(sh, lo, check=False)
        res = pipe_vs.merge(L, sh.device)
    if pending is not None:
        pipe_vs.wait()
        pj, (ps, pi) = pending
        same = same and np.array_equal(ps.cpu().numpy(), want[pj][0]) and np.array_equal(pi.cpu().numpy(), want[pj][1])
    pending = (j, res) if sh is not None else None
print(f"rank {rank}: overlap mode, 5 pipelined screens: {'OK' if same else 'MISMATCH'}", flush=True)
ok = ok and same
# peer mode: no collective; every rank stores its message into every mailbox over NVLink, merges one screen late
peer_vs = VirtualScreen(cnn, k=B - 1, exchange="peer")
results = []
for sh in shards + shards:            # 10 screens: the 4 mailbox slots are reused
    peer_vs.local_topk(sh, lo, check=False)
    s_, i_ = peer_vs.merge(L, sh.device)
    if results:                       # the previous screen's merge was issued by this call: its buffers are final after a sync
        torch.cuda.synchronize()
        results[-1] = (results[-1][0].cpu().numpy().copy(), results[-1][1].cpu().numpy().copy())
    results.append((s_, i_))
peer_vs.wait()
torch.cuda.synchronize()
results[-1] = (results[-1][0].cpu().numpy().copy(), results[-1][1].cpu().numpy().copy())
same = all(np.array_equal(results[j][0], want[j % 5][0]) and np.array_equal(results[j][1], want[j % 5][1]) for j in range(10))
s_, i_ = peer_vs.screen_indices(shards[2], lo)          # the synchronous entry in peer mode
same = same and np.array_equal(s_.cpu().numpy(), want[2][0]) and np.array_equal(i_.cpu().numpy(), want[2][1])
peer_vs.close()
print(f"rank {rank}: peer-memory exchange, 10 pipelined screens + 1 synchronous: {'OK' if same else 'MISMATCH'}", flush=True)
ok = ok and same
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
