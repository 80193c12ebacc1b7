"""Baseline MLP surrogate, B200-native.

Same constructor and default name as the reference's flexs/baselines/models/mlp.py:10-19,:35-36;
architecture (mlp.py:21-31): ``Flatten -> Dense(H,relu) x3 -> Dense(1)``, MSE + Adam (mlp.py:33).
"""
from typing import List, Optional

from flexs_b200.baselines.models.surrogate import B200Surrogate


class MLP(B200Surrogate):
    """Three ReLU dense layers; the first one is a row gather because the input is one-hot."""

    kind = "mlp"

    def __init__(self, seq_len, hidden_size, alphabet, loss="MSE", name=None, batch_size=256, epochs=20,
                 device: int = 0, seed: Optional[int] = None):
        if str(loss).upper() not in ("MSE", "MEAN_SQUARED_ERROR"):
            raise ValueError("the B200 training kernels implement the MSE loss the reference scripts use")
        if name is None:
            name = f"MLP_hidden_size_{hidden_size}"
        self.seq_len, self.hidden_size = seq_len, hidden_size
        super().__init__(
            dict(seq_len=seq_len, alphabet_size=len(alphabet), hidden_size=hidden_size),
            alphabet=alphabet, name=name, batch_size=batch_size, epochs=epochs, device=device, seed=seed,
        )

    def _weight_shapes(self) -> List[tuple]:
        d, h = self.seq_len * len(self.alphabet), self.hidden_size
        return [(d, h), (h,), (h, h), (h,), (h, h), (h,), (h, 1), (1,)]
