"""One small fit (for an ncu launch list of the training kernels): python tools/train_once.py L A [n epochs]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flexs_b200 as flexs  # noqa: E402

L, A = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
epochs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
alphabet = "ACGT" if A == 4 else "ACDEFGHIKLMNPQRSTVWY"[:A]
rng = np.random.default_rng(0)
letters = np.array(list(alphabet))
seqs = ["".join(r) for r in letters[rng.integers(0, A, size=(n, L))]]
model = flexs.baselines.models.CNN(L, num_filters=32, hidden_size=100, alphabet=alphabet, epochs=epochs, seed=0)
model.train(seqs, rng.random(n))
print("losses", model.last_fit_losses)
