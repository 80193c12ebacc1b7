"""``B200Surrogate``: the B200-native stand-in for the reference's ``KerasModel`` wrapper.

Replaces flexs/baselines/models/keras_model.py:12-79.  Same constructor knobs (``alphabet``,
``name``, ``batch_size``, ``epochs``), same ``train`` / ``_fitness_function`` contract, but the
Python one-hot loop + ``tf.convert_to_tensor`` + ``keras.Model.predict`` chain is one call into
``libflexs_b200.so``: residue characters go to the GPU as bytes, are mapped to residue indices
there, and a fused kernel evaluates the whole layer stack (no float one-hot is ever built).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np

from flexs_b200 import _native
from flexs_b200.model import Model
from flexs_b200.types import SEQUENCES_TYPE
from flexs_b200.utils import sequence_utils as s_utils


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


class B200Surrogate(Model):
    """Base of :class:`CNN` and :class:`MLP`; owns the native model object(s).

    Args mirror ``KerasModel.__init__`` (keras_model.py:15-47): ``batch_size`` / ``epochs`` drive
    ``train``; prediction has no batch size (the kernel tiles the batch itself).
    """

    #: subclasses set: "cnn" | "mlp"
    kind: str = ""
    #: lists of at least this many strings travel bit-packed (below it the byte route has one kernel launch less and
    #: PCIe volume is irrelevant)
    PACK_MIN_N: int = 4096

    def __init__(self, native_kwargs: dict, alphabet: str, name: str, batch_size: int = 256, epochs: int = 20,
                 device: int = 0, seed: Optional[int] = None):
        super().__init__(name)
        self.alphabet = alphabet
        self.batch_size = batch_size
        self.epochs = epochs
        self.device = device
        self._native_kwargs = dict(native_kwargs)
        self._native: Optional[_native.NativeModel] = None
        self._init_seed = seed
        self._fit_calls = 0
        self.weights_version = 0  # bumped whenever weights change (fused ensembles watch this)
        self.last_fit_losses: Optional[np.ndarray] = None

    # ------------------------------------------------------------------ native object
    @property
    def native(self) -> _native.NativeModel:
        """The native model, created (and glorot-initialised, like Keras does) on first use."""
        if self._native is None:
            self._native = _native.NativeModel(self.kind, device=self.device, **self._native_kwargs)
            self._native.set_weights(self._initial_weights())
            self.weights_version += 1
        return self._native

    def _weight_shapes(self) -> List[tuple]:
        raise NotImplementedError

    def _initial_weights(self) -> List[np.ndarray]:
        """Keras defaults: glorot-uniform kernels, zero biases (conv fan = k*in / k*out)."""
        rng = np.random.default_rng(self._init_seed)
        out = []
        for shp in self._weight_shapes():
            if len(shp) == 1:
                out.append(np.zeros(shp, dtype=np.float32))
                continue
            receptive = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            limit = np.sqrt(6.0 / (receptive * shp[-2] + receptive * shp[-1]))
            out.append(rng.uniform(-limit, limit, size=shp).astype(np.float32))
        return out

    def get_weights(self) -> List[np.ndarray]:
        """Weights in Keras ``get_weights()`` order and layout (see include/flexs_b200.h)."""
        flat = self.native.get_weights(0)
        return [a.reshape(shp) for a, shp in zip(flat, self._weight_shapes())]

    def set_weights(self, weights: Sequence[np.ndarray]) -> None:
        shapes = self._weight_shapes()
        if len(weights) != len(shapes):
            raise ValueError(f"expected {len(shapes)} weight arrays")
        for w, shp in zip(weights, shapes):
            if tuple(np.shape(w)) != tuple(shp):
                raise ValueError(f"weight of shape {np.shape(w)} does not match {shp}")
        self.native.set_weights(weights, 0)
        self.weights_version += 1

    # ------------------------------------------------------------------ scoring
    def _fitness_function(self, sequences: SEQUENCES_TYPE) -> np.ndarray:
        """Scores as a fresh host ``float32 (N,)`` array (what keras_model.py:77-79 returns)."""
        if _is_cuda_tensor(sequences):
            return self._score_device(sequences).cpu().numpy()
        n = len(sequences)
        if n == 0:
            return np.zeros(0, dtype=np.float32)
        alphabet = self.alphabet
        if isinstance(sequences, np.ndarray) and sequences.dtype == np.uint8 and sequences.ndim == 2:
            # pre-encoded residue INDICES: the identity alphabet maps index -> index on the device
            chars = np.ascontiguousarray(sequences)
            alphabet = bytes(range(len(self.alphabet)))
        elif isinstance(sequences, (list, tuple)) and isinstance(sequences[0], str) and n >= self.PACK_MIN_N:
            # A list of str crosses PCIe bit-packed (2 bits per DNA residue, 5 per amino acid): one multi-threaded C
            # pass maps the characters through the alphabet and packs them (csrc/packstr.c), the GPU unpacks.
            if len(sequences[0]) != self.seq_len:
                raise ValueError(f"{self.name} was built for sequences of length {self.seq_len}, got {len(sequences[0])}")
            return self.native.score_host_packed(s_utils.pack_sequences(sequences, self.alphabet))
        else:
            chars = s_utils.sequences_to_char_array(sequences)
        if chars.shape[1] != self.seq_len:
            raise ValueError(f"{self.name} was built for sequences of length {self.seq_len}, got {chars.shape[1]}")
        return self.native.score_host(chars, alphabet)

    def _score_device(self, idx):
        """``uint8[N, L]`` CUDA tensor of residue indices -> ``float32[N]`` CUDA tensor (no sync)."""
        import torch

        if idx.dtype != torch.uint8 or idx.dim() != 2 or idx.shape[1] != self.seq_len:
            raise ValueError(f"expected a uint8 [N, {self.seq_len}] CUDA tensor of residue indices")
        if idx.device.index != self.device:
            raise ValueError(f"{self.name} lives on cuda:{self.device} but the candidates are on {idx.device} "
                             f"(build one model per rank with device=LOCAL_RANK)")
        idx = idx.contiguous()
        out = torch.empty(idx.shape[0], dtype=torch.float32, device=idx.device)
        with torch.cuda.device(idx.device):
            stream = torch.cuda.current_stream().cuda_stream
            self.native.forward_dev(idx.data_ptr(), idx.shape[0], out.data_ptr(), stream)
        return out

    def get_fitness_device(self, idx):
        """Device-resident variant of ``get_fitness``: charges ``cost`` and returns a CUDA tensor."""
        self.cost += int(idx.shape[0])
        return self._score_device(idx)

    # ------------------------------------------------------------------ training
    def train(self, sequences: SEQUENCES_TYPE, labels: Union[np.ndarray, Sequence[float]], verbose: bool = False):
        """``model.fit(one_hots, labels, batch_size, epochs)`` (keras_model.py:49-67) on the GPU.

        MSE loss, Adam with Keras defaults; weights and optimiser moments persist across calls
        exactly as the compiled Keras model's do (explorer.py:157-160 retrains every round).
        """
        import torch

        idx = s_utils.encode_sequences(sequences, self.alphabet)
        if idx.shape[1] != self.seq_len:
            raise ValueError(f"{self.name} was built for sequences of length {self.seq_len}")
        y = np.asarray(labels, dtype=np.float32).reshape(-1)
        if len(y) != len(idx):
            raise ValueError("sequences and labels differ in length")
        dev = torch.device("cuda", self.device)
        d_idx = torch.from_numpy(np.ascontiguousarray(idx)).to(dev)
        d_y = torch.from_numpy(y).to(dev)
        seed = (0 if self._init_seed is None else int(self._init_seed)) * 1000003 + self._fit_calls
        self._fit_calls += 1
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            losses = self.native.fit_dev(d_idx.data_ptr(), d_y.data_ptr(), len(y), self.batch_size, self.epochs,
                                         seed & 0xFFFFFFFFFFFFFFFF, stream)
            torch.cuda.current_stream().synchronize()
        self.last_fit_losses = losses[: self.epochs]
        self.weights_version += 1
        if verbose:
            for e, l in enumerate(self.last_fit_losses):
                print(f"Epoch {e + 1}/{self.epochs} - loss: {l:.6f}")
