"""``TFBinding`` landscape and its problem registry (reference: flexs/landscapes/tf_binding.py:12-93).

The reference holds the measurements in a ``dict`` keyed by the 8-mer; here they are a dense float64 table of
``4**L`` entries on the GPU indexed by the 2-bit-packed sequence (K6 ``flexs_lookup_score_dev``).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import pandas as pd

from flexs_b200 import _native
from flexs_b200.landscapes.table_landscape import DeviceTableLandscape, data_dir
from flexs_b200.types import SEQUENCES_TYPE

_BASES = "ACGT"


class TFBinding(DeviceTableLandscape):
    """Binding affinity of DNA 8-mers to a transcription factor (Barrera et al. 2016 measurement files).

    Same construction as the reference (tf_binding.py:22-41): min-max normalised E-score, both strands of a row map
    to the same value, the reverse-strand column written last.  ``self.sequences`` (the reference's dict attribute) is
    kept as a lazily built view for code that inspects it.
    """

    def __init__(self, landscape_file: str, device: int = 0):
        super().__init__(name="TF_Binding", device=device)
        data = pd.read_csv(landscape_file, sep="\t")
        score = data["E-score"]
        norm_score = ((score - score.min()) / (score.max() - score.min())).to_numpy(dtype=np.float64)
        fwd = data["8-mer"].to_numpy(dtype=str)
        rev = data["8-mer.1"].to_numpy(dtype=str)
        self.seq_len = len(fwd[0])
        self.column_of_char = np.full(256, 0xFF, dtype=np.uint8)
        for i, ch in enumerate(_BASES):
            self.column_of_char[ord(ch)] = i
        self.table = np.full(len(_BASES) ** self.seq_len, np.nan, dtype=np.float64)
        # dict(zip(fwd, v)) then .update(zip(rev, v)): later rows win, the reverse strand wins over the forward one
        self.table[self._keys(fwd)] = norm_score
        self.table[self._keys(rev)] = norm_score
        self._fwd, self._rev, self._norm = fwd, rev, norm_score
        self._dict: Optional[dict] = None

    def _keys(self, seqs: np.ndarray) -> np.ndarray:
        chars = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).reshape(len(seqs), -1)
        if chars.shape[1] != self.seq_len:
            raise ValueError("all sequences of a TFBinding file must have the same length")
        cols = self.column_of_char[chars]
        if (cols == 0xFF).any():
            raise ValueError(f"TFBinding file holds characters outside {_BASES}")
        weights = len(_BASES) ** np.arange(self.seq_len - 1, -1, -1, dtype=np.int64)
        return cols.astype(np.int64) @ weights

    @property
    def sequences(self) -> dict:
        """The reference's ``{sequence: normalised score}`` attribute (tf_binding.py:40-41)."""
        if self._dict is None:
            self._dict = dict(zip(self._fwd, self._norm))
            self._dict.update(zip(self._rev, self._norm))
        return self._dict

    def _launch(self, d_seq, n, lut, d_out, stream):
        _native.lookup_score_dev(d_seq, n, self.seq_len, lut, len(_BASES), self._d_table.data_ptr(), self.table.size,
                                 d_out, stream)

    def _fitness_function(self, sequences: SEQUENCES_TYPE) -> np.ndarray:
        for s in sequences:
            if len(s) != self.seq_len:
                raise KeyError(s)  # not a key of the reference's dict
        out = self._score_host(sequences)
        missing = np.isnan(out)
        if missing.any():
            raise KeyError(sequences[int(np.argmax(missing))])
        return out


def registry(data_directory: Optional[str] = None) -> Dict[str, Dict]:
    """Problems ``{name: {"params": {"landscape_file": ...}, "starts": [...]}}`` (tf_binding.py:47-93)."""
    tf_binding_data_dir = data_directory or data_dir("tf_binding")
    problems = {}
    for fname in os.listdir(tf_binding_data_dir):
        problem_name = fname.replace("_8mers.txt", "")
        problems[problem_name] = {
            "params": {"landscape_file": os.path.join(tf_binding_data_dir, fname)},
            "starts": [
                "GCTCGAGC", "GCGCGCGC", "TGCGCGCC", "ATATAGCC", "GTTTGGTA", "ATTATGTT", "CAGTTTTT",
                "AAAAATTT", "AAAAACGC", "GTTGTTTT", "TGCTTTTT", "AAAGATAG", "CCTTCTTT", "AAAGAGAG",
            ],
        }
    return problems
