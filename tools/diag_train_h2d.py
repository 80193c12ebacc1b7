import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import flexs_b200 as flexs
rng = np.random.default_rng(7)
for tag, L, alphabet in (("cnn_100x4", 100, "ACGT"), ("cnn_237x20", 237, "ACDEFGHIKLMNPQRSTVWY"), ("cnn_100x4_again", 100, "ACGT")):
    model = flexs.baselines.models.CNN(L, num_filters=32, hidden_size=100, alphabet=alphabet, loss="MSE", device=0, seed=0)
    letters = np.array(list(alphabet))
    seqs = ["".join(r) for r in letters[rng.integers(0, len(alphabet), size=(1000, L))]]
    labels = rng.random(1000)
    model.train(seqs[:64], labels[:64])
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); model.train(seqs, labels); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); idx = flexs.utils.sequence_utils.encode_sequences(seqs, alphabet); enc = time.perf_counter() - t0
    print(tag, "fit seconds:", [f"{t:.4f}" for t in ts], "host encode of 1000 strings:", f"{enc:.4f}")
# H2D bandwidth, pinned
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
print("H2D pinned GB/s:", 5 * x.numel() / (time.perf_counter() - t0) / 1e9)
