"""``AdditiveAAVPackaging`` landscape and its registry (reference: flexs/landscapes/additive_aav_packaging.py:23-147).

Fitness = sum over positions of a per-(position, residue) table entry, normalised and clipped at 0.  The table
is built on the host exactly as the reference builds its nested dict; the per-sequence gather-sum runs on the GPU
(K6 ``flexs_additive_score_dev``) in float64, left to right, so results are bit-identical to the reference's loop.
"""
from __future__ import annotations

import json
import os
from typing import Optional

import numpy as np

from flexs_b200 import _native
from flexs_b200.landscapes.table_landscape import DeviceTableLandscape, data_dir

# AAV2 VP1 capsid protein, 735 residues (UniProt P03135) — `wild_type` is a window of it (:8-20, :62)
AAV2_WT = (
    "MAADGYLPDWLEDTLSEGIRQWWKLKPGPPPPKPAERHKDDSRGLVLPGYKYLGPFNGLDKGEPVNEADAAALEHDKAYDRQLDSGDNPYLKYNHADAEF"
    "QERLKEDTSFGGNLGRAVFQAKKRVLEPLGLVEEPVKTAPGKKRPVEHSPVEPDSSSGTGKAGQQPARKRLNFGQTGDADSVPDPQPLGQPPAAPSGLGT"
    "NTMATGSGAPMADNNEGADGVGNSSGNWHCDSTWMGDRVITTSTRTWALPTYNNHLYKQISSQSGASNDNHYFGYSTPWGYFDFNRFHCHFSPRDWQRLI"
    "NNNWGFRPKRLNFKLFNIQVKEVTQNDGTTTIANNLTSTVQVFTDSEYQLPYVLGSAHQGCLPPFPADVFMVPQYGYLTLNNGSQAVGRSSFYCLEYFPS"
    "QMLRTGNNFTFSYTFEDVPFHSSYAHSQSLDRLMNPLIDQYLYYLSRTNTPSGTTTQSRLQFSQAGASDIRDQSRNWLPGPCYRQQRVSKTSADNNNSEY"
    "SWTGATKYHLNGRDSLVNPGPAMASHKDDEEKFFPQSGVLIFGKQGSEKTNVDIEKVMITDEEEIRTTNPVATEQYGSVSTNLQRGNRQAATADVNTQGV"
    "LPGMVWQDRDVYLQGPIWAKIPHTDGHFHPSPLMGGFGLKHPPPQILIKNTPVPANPSTTFSAAKFASFITQYSTGQVSVEIEWELQKENSKRWNPEIQY"
    "TSNYNKSVNVDFTVDTNGVYSEPRPIGTRYLTRNL"
)


class AdditiveAAVPackaging(DeviceTableLandscape):
    """Additive landscape from AAV2 single-substitution measurements.

    Constructor arguments, attributes (``wild_type``, ``data``, ``top_seq``, ``max_possible``, ``mfm``, ``noise``) and
    results follow the reference (:38-85).  ``data_file`` (extension) names the ``AAV2_single_subs.json`` to read when
    it is not found through ``FLEXS_DATA_DIR``.
    """

    def __init__(self, phenotype: str = "heart", minimum_fitness_multiplier: float = 1, start: int = 0, end: int = 735,
                 noise: int = 0, data_file: Optional[str] = None, device: int = 0):
        super().__init__(f"AdditiveAAVPackaging_phenotype={phenotype}", device=device)
        self.sequences = {}
        self.phenotype = f"log2_{phenotype}_v_wt"
        self.mfm = minimum_fitness_multiplier
        self.start = start
        self.end = end
        self.noise = noise
        self.wild_type = AAV2_WT[start:end]
        if data_file is None:
            data_file = os.path.join(data_dir("additive_aav_packaging"), "AAV2_single_subs.json")
        with open(data_file) as f:
            self.data = {int(pos): val for pos, val in json.load(f).items() if self.start <= int(pos) < self.end}
        self.top_seq, self.max_possible = self.compute_max_possible()
        self._build_table()

    def compute_max_possible(self):
        """Max possible fitness of any sequence, used for normalisation (same loop as the reference, :79-96)."""
        best_seq = ""
        max_fitness = 0
        for pos in self.data:
            current_max = -10
            current_best = "M"
            for aa in self.data[pos]:
                current_fit = self.data[pos][aa][self.phenotype]
                if current_fit > current_max and self.data[pos][aa]["log2_packaging_v_wt"] > -6:
                    current_best = aa
                    current_max = current_fit
            best_seq += current_best
            max_fitness += current_max
        return best_seq, max_fitness

    def _build_table(self):
        # The kernel walks positions start, start+1, ... of the window; a position absent from the data is a KeyError
        # in the reference for any sequence that reaches it (:104), so the table covers the contiguous prefix only.
        self.seq_len = 0
        while self.start + self.seq_len in self.data:
            self.seq_len += 1
        residues = sorted({aa for i in range(self.seq_len) for aa in self.data[self.start + i]})
        if any(len(aa) != 1 or ord(aa) > 255 for aa in residues) or len(residues) > 254:
            raise ValueError("residue keys of the data file must be single latin-1 characters")
        self.residues = "".join(residues)
        self.column_of_char = np.full(256, 0xFF, dtype=np.uint8)
        for c, aa in enumerate(residues):
            self.column_of_char[ord(aa)] = c
        # absent (position, residue) pairs contribute +0.0, which equals the reference's "skip" for every running sum
        self.table = np.zeros((max(self.seq_len, 1), max(len(residues), 1)), dtype=np.float64)
        for i in range(self.seq_len):
            for aa, val in self.data[self.start + i].items():
                self.table[i, self.column_of_char[ord(aa)]] = float(val[self.phenotype])
        self._offset = float(self.mfm * self.max_possible)
        self._denom = float(self.max_possible * (self.mfm + 1))

    def _launch(self, d_seq, n, lut, d_out, stream, d_noise: int = 0):
        _native.additive_score_dev(d_seq, n, self.seq_len, lut, self.table.shape[1], self._d_table.data_ptr(),
                                   self._offset, self._denom, d_noise, d_out, stream)

    def get_fitness_device(self, seq, columns: bool = False, charge: bool = True):
        """Device-resident scoring (an extension: the reference has no such call).  With ``noise != 0`` the
        per-sequence normal draws come from numpy's global stream as in ``get_fitness`` and are uploaded (8 bytes per
        sequence); with ``noise == 0`` the stream is left untouched (``get_fitness`` advances it, like the reference)."""
        import torch

        if self.noise == 0:
            return super().get_fitness_device(seq, columns, charge)
        if seq.dtype != torch.uint8 or seq.dim() != 2 or seq.shape[1] != self.seq_len or not seq.is_contiguous():
            raise ValueError(f"expected a contiguous uint8 [n, {self.seq_len}] CUDA tensor")
        self._table_on_device()
        n = int(seq.shape[0])
        if charge:
            self.cost += n
        d_noise = torch.from_numpy(np.random.normal(scale=self.noise, size=n)).to(seq.device)
        out = torch.empty(n, dtype=torch.float64, device=seq.device)
        with torch.cuda.device(seq.device):
            self._launch(seq.data_ptr(), n, None if columns else self.column_of_char, out.data_ptr(),
                         torch.cuda.current_stream().cuda_stream, d_noise.data_ptr())
        return out

    def _fitness_function(self, sequences):
        import torch

        n = len(sequences)
        if n == 0:
            return np.array([])
        seqs = [str(s) for s in sequences]
        for s in seqs:
            if len(s) > self.seq_len:
                raise KeyError(self.start + self.seq_len)  # self.data[self.start + i] of the reference (:104)
        # shorter sequences sum over their own length (:103): pad with a character that has no column
        if any(len(s) != self.seq_len for s in seqs):
            seqs = [s + "\0" * (self.seq_len - len(s)) for s in seqs]
        # one np.random.normal draw per sequence, in order, also when noise == 0 (:113) — same global stream
        noise = np.random.normal(scale=self.noise, size=n)
        self._table_on_device()
        dev = torch.device("cuda", self.device)
        from flexs_b200.utils import sequence_utils as s_utils

        d_seq = torch.from_numpy(s_utils.sequences_to_char_array(seqs, self.seq_len).copy()).to(dev)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        d_noise = torch.from_numpy(noise).to(dev) if self.noise != 0 else None
        with torch.cuda.device(dev):
            self._launch(d_seq.data_ptr(), n, self.column_of_char, out.data_ptr(),
                         torch.cuda.current_stream().cuda_stream, d_noise.data_ptr() if d_noise is not None else 0)
        res = out.cpu().numpy()
        # max(0, x) yields the int 0 for x <= 0; a list of only those becomes an int64 array in the reference (:114-116)
        return res if (res > 0).any() else res.astype(np.int64)


def registry():
    """``{name: {"params": ...}}`` such that ``AdditiveAAVPackaging(**params)`` builds the problem (:121-147)."""
    return {
        "heart": {"params": {"phenotype": "heart", "start": 450, "end": 540}},
        "lung": {"params": {"phenotype": "lung", "start": 450, "end": 540}},
        "kidney": {"params": {"phenotype": "kidney", "start": 450, "end": 540}},
        "liver": {"params": {"phenotype": "liver", "start": 450, "end": 540}},
        "blood": {"params": {"phenotype": "blood", "start": 450, "end": 540}},
        "spleen": {"params": {"phenotype": "spleen", "start": 450, "end": 540}},
    }
