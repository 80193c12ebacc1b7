"""Baseline CNN surrogate, B200-native.

Same constructor and default name as the reference's flexs/baselines/models/cnn.py:10-21,:58-59.
The architecture (cnn.py:23-54) is fixed by the native kernels:
``Conv1D(F,k,valid,relu) -> Conv1D(F,k,same,relu) -> MaxPooling1D(1) [identity] ->
Conv1D(F,len(alphabet)-1,same,relu) -> GlobalMaxPooling1D -> Dense(H,relu) -> Dense(H,relu) ->
Dropout(0.25) -> Dense(1)``, trained with MSE + Adam (cnn.py:56).
"""
from typing import List, Optional

from flexs_b200.baselines.models.surrogate import B200Surrogate


class CNN(B200Surrogate):
    """3 conv layers + 2 dense layers, evaluated by one fused sm_100a kernel."""

    kind = "cnn"

    def __init__(
        self,
        seq_len: int,
        num_filters: int,
        hidden_size: int,
        alphabet: str,
        loss="MSE",
        kernel_size: int = 5,
        name: Optional[str] = None,
        batch_size: int = 256,
        epochs: int = 20,
        device: int = 0,
        seed: Optional[int] = None,
    ):
        if str(loss).upper() not in ("MSE", "MEAN_SQUARED_ERROR"):
            raise ValueError("the B200 training kernels implement the MSE loss the reference scripts use")
        if name is None:
            name = f"CNN_hidden_size_{hidden_size}_num_filters_{num_filters}"
        self.seq_len, self.num_filters, self.hidden_size, self.kernel_size = seq_len, num_filters, hidden_size, kernel_size
        super().__init__(
            dict(seq_len=seq_len, alphabet_size=len(alphabet), num_filters=num_filters, hidden_size=hidden_size,
                 kernel_size=kernel_size),
            alphabet=alphabet, name=name, batch_size=batch_size, epochs=epochs, device=device, seed=seed,
        )

    def _weight_shapes(self) -> List[tuple]:
        k, a, f, h = self.kernel_size, len(self.alphabet), self.num_filters, self.hidden_size
        return [(k, a, f), (f,), (k, f, f), (f,), (a - 1, f, f), (f,), (f, h), (h,), (h, h), (h,), (h, 1), (1,)]
