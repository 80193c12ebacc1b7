"""torchrun --nproc-per-node 2 tools/cmaes_2gpu_check.py — CMA-ES with the population on the GPUs (BASELINE configs[3]:
AAV-length proteins, candidate batch sharded over the ranks).  Every rank samples the same population, scores its share
and the ranks exchange ONE all_gather of float32 scores per iteration; proposals, scores and model.cost must be identical
on every rank and equal to what the ranks' own surrogate gives for the proposed strings."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

import flexs_b200 as flexs
from flexs_b200.utils import sequence_utils as su

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl")
L, pop = 735, 4096
rng = np.random.default_rng(0)
wt = "".join(np.array(list(su.AAS))[rng.integers(0, 20, size=L)])
cnn = flexs.baselines.models.CNN(L, 32, 100, su.AAS, seed=3, device=local)      # same seed: replicated weights
ex = flexs.baselines.explorers.CMAES(cnn, rounds=1, sequences_batch_size=50, model_queries_per_batch=3 * pop + 10,
                                    starting_sequence=wt, alphabet=su.AAS, population_size=pop, max_iter=5, seed=1)
frame = pd.DataFrame({"sequence": [wt], "true_score": [0.5], "model_score": [np.nan], "round": [0], "model_cost": [0],
                      "measurement_cost": [1]})
seqs, preds = ex.propose_sequences(frame)
digest = torch.tensor([hash("".join(seqs)) % (1 << 40), int(cnn.cost), len(seqs)], dtype=torch.int64, device="cuda")
both = [torch.zeros_like(digest) for _ in range(world)]
dist.all_gather(both, digest)
same = all(torch.equal(b, both[0]) for b in both)
# PYTHONHASHSEED differs between processes: compare through the scores instead of hash() when it does
score_t = torch.tensor(preds, dtype=torch.float32, device="cuda")
gathered = [torch.zeros_like(score_t) for _ in range(world)]
dist.all_gather(gathered, score_t)
same_scores = all(torch.equal(g, gathered[0]) for g in gathered)
direct = cnn.get_fitness(list(seqs))
new = np.array([s != wt for s in seqs])
ok = same_scores and int(both[0][1]) == int(both[1][1]) and np.allclose(preds[new], direct[new], rtol=0, atol=1e-4 * np.abs(direct).max())
print(f"rank {rank}: {len(seqs)} proposals, model.cost {cnn.cost} (budget {3 * pop + 10}), identical on all ranks: {same_scores}, "
      f"scores match the surrogate: {ok}", flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
