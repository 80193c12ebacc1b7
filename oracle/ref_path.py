"""The reference's ``get_fitness`` call chain, restated for timing on the CPU.  TEST/BENCH
INFRASTRUCTURE ONLY (bench.py --impl reference); never imported by the product.

TensorFlow is not installed, so ``keras.Model.predict`` is stood in for by torch-CPU (oneDNN)
``conv1d`` / ``linear`` with the Keras layer semantics of oracle/flexs_oracle.py; everything around
it follows the reference line by line:

  * flexs/utils/sequence_utils.py:44-47  one float64 ``(L, A)`` one-hot per sequence, Python loop,
    ``alphabet.index(ch)``
  * flexs/baselines/models/keras_model.py:70-75  ``np.array([...])`` -> float32 tensor
  * keras_model.py:20,77-79  predict with batch_size=256, ``squeeze(axis=1)``, ``np.nan_to_num``
  * flexs/landscape.py:44-45  ``cost += len(sequences)``
"""
import time

import numpy as np
import torch
import torch.nn.functional as F


def string_to_one_hot(sequence: str, alphabet: str) -> np.ndarray:
    out = np.zeros((len(sequence), len(alphabet)))
    for i in range(len(sequence)):
        out[i, alphabet.index(sequence[i])] = 1
    return out


class ReferenceCNN:
    def __init__(self, seq_len, alphabet, num_filters, hidden_size, kernel_size, weights, batch_size=256):
        self.alphabet, self.batch_size, self.cost = alphabet, batch_size, 0
        self.k, self.k3 = kernel_size, len(alphabet) - 1
        self.threads = torch.get_num_threads()
        w = [torch.from_numpy(np.asarray(a, dtype=np.float32)) for a in weights]
        # Keras Conv1D kernel (k, in, out) -> torch (out, in, k); Dense (in, out) -> torch (out, in)
        self.c = [(w[i].permute(2, 1, 0).contiguous(), w[i + 1]) for i in (0, 2, 4)]
        self.d = [(w[i].t().contiguous(), w[i + 1]) for i in (6, 8, 10)]
        self.last_encode_seconds = 0.0

    @staticmethod
    def _same(x, k):
        left = (k - 1) // 2
        return F.pad(x, (left, (k - 1) - left))

    def _predict_batch(self, x):  # x: (B, L, A) float32, channels-last like Keras
        h = x.permute(0, 2, 1)
        h = F.relu(F.conv1d(h, *self.c[0]))
        h = F.relu(F.conv1d(self._same(h, self.k), *self.c[1]))
        h = F.relu(F.conv1d(self._same(h, self.k3), *self.c[2]))
        p = h.amax(dim=2)
        d = F.relu(F.linear(p, *self.d[0]))
        d = F.relu(F.linear(d, *self.d[1]))
        return F.linear(d, *self.d[2])

    def get_fitness(self, sequences):
        self.cost += len(sequences)
        t0 = time.perf_counter()
        one_hots = torch.from_numpy(
            np.array([string_to_one_hot(seq, self.alphabet) for seq in sequences]).astype(np.float32))
        self.last_encode_seconds = time.perf_counter() - t0
        with torch.no_grad():
            outs = [self._predict_batch(one_hots[i: i + self.batch_size])
                    for i in range(0, len(one_hots), self.batch_size)]
        return np.nan_to_num(torch.cat(outs).numpy().squeeze(axis=1))
