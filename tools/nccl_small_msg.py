"""torchrun --nproc-per-node N tools/nccl_small_msg.py — latency of the screen's one collective (an 11 KB message per rank)
as NCCL all_gather vs all_to_all (every rank sends its message straight to every peer: one hop instead of N-1 ring steps),
timed with CUDA events on the device; run with NCCL_DEBUG=INFO to see the transport NCCL picked."""
import os
import sys

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl")
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 11088
msg = torch.full((mb,), rank, dtype=torch.uint8, device="cuda")
gathered = torch.empty(world * mb, dtype=torch.uint8, device="cuda")
outs = list(gathered.view(world, mb).unbind(0))
ins = [msg] * world


def timed(fn, iters=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


t_ag = timed(lambda: dist.all_gather_into_tensor(gathered, msg))
ok_ag = bool((gathered.view(world, mb)[:, 0].cpu() == torch.arange(world, dtype=torch.uint8)).all())
gathered.zero_()
t_a2a = timed(lambda: dist.all_to_all(outs, ins))
ok_a2a = bool((gathered.view(world, mb)[:, 0].cpu() == torch.arange(world, dtype=torch.uint8)).all())
if rank == 0:
    print(f"world {world}, {mb} B per rank: all_gather {t_ag:.1f} us ({'ok' if ok_ag else 'WRONG'}), all_to_all {t_a2a:.1f} us "
          f"({'ok' if ok_a2a else 'WRONG'})  [back-to-back launches, device time per call, max over ranks]", flush=True)
dist.destroy_process_group()
