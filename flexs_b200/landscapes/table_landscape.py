"""Shared device plumbing of the table landscapes: characters in, float64 fitness out, all on the GPU."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from flexs_b200 import _native
from flexs_b200.landscape import Landscape
from flexs_b200.types import SEQUENCES_TYPE
from flexs_b200.utils import sequence_utils as s_utils


def data_dir(subdir: str) -> str:
    """Directory holding the reference's measurement files (``flexs/landscapes/data/<subdir>``).

    Resolution order: ``$FLEXS_DATA_DIR/<subdir>``, then a ``data/<subdir>`` directory next to this package.
    """
    roots = [os.environ.get("FLEXS_DATA_DIR"), os.path.join(os.path.dirname(__file__), "data")]
    for root in roots:
        if root and os.path.isdir(os.path.join(root, subdir)):
            return os.path.join(root, subdir)
    raise FileNotFoundError(
        f"landscape data '{subdir}' not found: set FLEXS_DATA_DIR to the reference checkout's flexs/landscapes/data")


class DeviceTableLandscape(Landscape):
    """A landscape whose ``_fitness_function`` is one kernel over a device-resident table.

    Subclasses build ``column_of_char`` (uint8[256], 0xFF = the character has no column) and a float64 table and
    implement ``_launch``.  ``get_fitness`` keeps the reference contract (host strings in, ``np.ndarray`` of float64
    out); ``get_fitness_device`` scores candidates that are already on the GPU (uint8 characters, or column indices
    with ``columns=True``) and returns a CUDA tensor, so an explorer round never leaves the device.
    """

    def __init__(self, name: str, device: int = 0):
        super().__init__(name)
        self.device = device
        self._d_table = None

    # -- subclass hooks ---------------------------------------------------------------------
    seq_len: int = 0
    column_of_char: np.ndarray
    table: np.ndarray

    def _launch(self, d_seq: int, n: int, lut: Optional[np.ndarray], d_out: int, stream: int) -> None:
        raise NotImplementedError

    # -- device path ------------------------------------------------------------------------
    def _table_on_device(self):
        import torch

        if self._d_table is None:
            if not torch.cuda.is_available():
                raise _native.NativeError("flexs_b200 landscapes run on a CUDA device; there is no CPU fallback")
            self._d_table = torch.from_numpy(np.ascontiguousarray(self.table, dtype=np.float64)).to(
                torch.device("cuda", self.device))
        return self._d_table

    def get_fitness_device(self, seq, columns: bool = False, charge: bool = True):
        """``seq``: CUDA uint8 tensor ``[n, seq_len]`` -> CUDA float64 tensor ``[n]`` (no host round trip)."""
        import torch

        if seq.dtype != torch.uint8 or seq.dim() != 2 or seq.shape[1] != self.seq_len or not seq.is_contiguous():
            raise ValueError(f"expected a contiguous uint8 [n, {self.seq_len}] CUDA tensor")
        self._table_on_device()
        if charge:
            self.cost += int(seq.shape[0])
        out = torch.empty(seq.shape[0], dtype=torch.float64, device=seq.device)
        with torch.cuda.device(seq.device):
            self._launch(seq.data_ptr(), int(seq.shape[0]), None if columns else self.column_of_char, out.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
        return out

    def _score_host(self, sequences: SEQUENCES_TYPE) -> np.ndarray:
        import torch

        if len(sequences) == 0:
            return np.array([])
        chars = s_utils.sequences_to_char_array(sequences, self.seq_len)
        self._table_on_device()
        d_seq = torch.from_numpy(chars if chars.flags.writeable else chars.copy()).to(torch.device("cuda", self.device))
        return self.get_fitness_device(d_seq, charge=False).cpu().numpy()
