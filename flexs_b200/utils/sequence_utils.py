"""Sequence helpers with the names and semantics of the reference's flexs/utils/sequence_utils.py.

The surrogate kernels never consume a float one-hot; they work on residue INDICES
(``encode_sequences``).  The one-hot functions are kept because explorers and user code call them.
"""
import random
from typing import List, Sequence, Union

import numpy as np

AAS = "ILVAGMFYWEDQNHCRKSTP"
"""Protein alphabet, 20 amino acids, no stop codon (sequence_utils.py:7)."""

RNAA = "UGCA"
"""RNA alphabet (sequence_utils.py:10)."""

DNAA = "TGCA"
"""DNA alphabet (sequence_utils.py:13)."""

BA = "01"
"""Binary alphabet (sequence_utils.py:16)."""


def _lut(alphabet: str) -> np.ndarray:
    table = np.full(256, 255, dtype=np.uint8)
    for i in range(len(alphabet) - 1, -1, -1):  # first occurrence wins, like str.index
        table[ord(alphabet[i])] = i
    return table


_PACKSTR = False  # not looked up yet


def _packstr(name: str = "pack"):
    """``flexs_b200/_packstr`` (csrc/packstr.c, built next to the CUDA library): packs a list of str in one C pass.
    ``None`` when it has not been built — the pure-Python route below is equivalent, only ~6x slower."""
    global _PACKSTR
    if _PACKSTR is False:
        try:
            from flexs_b200 import _packstr as mod

            _PACKSTR = mod
        except ImportError:
            _PACKSTR = None
    return getattr(_PACKSTR, name, None) if _PACKSTR is not None else None


def bits_per_residue(alphabet_size: int) -> int:
    bits = 1
    while (1 << bits) < alphabet_size:
        bits += 1
    return bits


def pack_indices(idx: np.ndarray, alphabet_size: int) -> np.ndarray:
    """``uint8[N, L]`` residue indices -> the packed wire format of include/flexs_b200.h: ``ceil(log2 A)`` bits per
    residue, residue i in bits ``[i*b, (i+1)*b)`` of the row's little-endian bit stream, rows padded to whole bytes
    (``uint8[N, ceil(L*b/8)]``).  Vectorised numpy; the reference never builds anything like it (it ships a float32
    one-hot of 4*L*A bytes per sequence to TensorFlow, keras_model.py:70-75)."""
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    n, length = idx.shape
    bits = bits_per_residue(alphabet_size)
    if idx.size and int(idx.max()) >= alphabet_size:
        raise ValueError("residue index outside the alphabet")
    planes = ((idx[:, :, None] >> np.arange(bits, dtype=np.uint8)) & 1).reshape(n, length * bits)
    return np.packbits(planes, axis=1, bitorder="little")


def unpack_indices(packed: np.ndarray, seq_len: int, alphabet_size: int) -> np.ndarray:
    """Inverse of :func:`pack_indices`."""
    bits = bits_per_residue(alphabet_size)
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    planes = np.unpackbits(packed, axis=1, bitorder="little")[:, : seq_len * bits].reshape(len(packed), seq_len, bits)
    return (planes.astype(np.uint16) << np.arange(bits, dtype=np.uint16)).sum(axis=2).astype(np.uint8)


def host_threads() -> int:
    """Threads the C packer may use: the cores this process may run on (torchrun pins OMP_NUM_THREADS=1 per rank, which
    says nothing about the cores available to a short host-side pass), capped at 16."""
    import os

    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(16, n))


def pack_sequences(sequences: Union[Sequence[str], np.ndarray], alphabet: str) -> np.ndarray:
    """Strings (or a ``uint8[N, L]`` index array) -> packed wire format.  A list/tuple of ``str`` goes through ONE
    multi-threaded C pass over the string objects (csrc/packstr.c ``pack_bits``: alphabet lookup + bit packing, GIL
    released); everything else through numpy.  ``ValueError`` for a character outside the alphabet, like ``str.index``
    in the reference (sequence_utils.py:46)."""
    if isinstance(sequences, np.ndarray) and sequences.dtype == np.uint8:
        return pack_indices(sequences, len(alphabet))
    n = len(sequences)
    bits = bits_per_residue(len(alphabet))
    pack_bits = _packstr("pack_bits")
    if pack_bits is not None and n and len(alphabet) <= 127 and isinstance(sequences, (list, tuple)) and isinstance(sequences[0], str):
        width = len(sequences[0])
        out = np.empty((n, (width * bits + 7) // 8), dtype=np.uint8)
        pack_bits(sequences, alphabet.encode("latin-1"), out, host_threads())
        return out
    return pack_indices(encode_sequences(sequences, alphabet), len(alphabet))


def sequences_to_char_array(sequences: Union[Sequence[str], np.ndarray], seq_len: int = None) -> np.ndarray:
    """Pack sequences into a contiguous ``uint8[N, L]`` array of residue CHARACTER codes.

    Accepts a list/tuple of ``str``, a numpy array of dtype ``<U`` or ``S``, or (pass-through) a
    ``uint8[N, L]`` array.  Raises ``ValueError`` for ragged input or non-latin-1 characters.
    """
    if isinstance(sequences, np.ndarray) and sequences.dtype == np.uint8:
        arr = np.ascontiguousarray(sequences)
        return arr.reshape(len(arr), -1)
    n = len(sequences)
    if n == 0:
        return np.zeros((0, seq_len or 0), dtype=np.uint8)
    if isinstance(sequences, np.ndarray) and sequences.dtype.kind == "S":
        width = sequences.dtype.itemsize
        chars = np.ascontiguousarray(sequences).view(np.uint8).reshape(n, width)
    elif isinstance(sequences, np.ndarray) and sequences.dtype.kind == "U":
        width = sequences.dtype.itemsize // 4
        wide = np.ascontiguousarray(sequences).view(np.uint32).reshape(n, width)
        if wide.size and int(wide.max()) > 255:
            raise ValueError("substring not found: non latin-1 character in sequence")
        chars = wide.astype(np.uint8)
    else:
        seqs = [str(s) for s in sequences] if not isinstance(sequences[0], str) else sequences
        width = len(seqs[0])
        pack = _packstr("pack")
        if pack is not None and isinstance(seqs, (list, tuple)):
            chars = np.empty((n, width), dtype=np.uint8)
            pack(seqs, chars)  # one C pass over the str objects; raises the same ValueErrors as the route below
            if seq_len is not None and width != seq_len:
                raise ValueError(f"expected sequences of length {seq_len}, got {width}")
            return chars
        joined = "".join(seqs)
        if len(joined) != n * width or max(map(len, seqs)) != width:
            raise ValueError("all sequences must have the same length")
        try:
            raw = joined.encode("latin-1")
        except UnicodeEncodeError as e:
            raise ValueError("substring not found: non latin-1 character in sequence") from e
        chars = np.frombuffer(raw, dtype=np.uint8).reshape(n, width)
    if seq_len is not None and chars.shape[1] != seq_len:
        raise ValueError(f"expected sequences of length {seq_len}, got {chars.shape[1]}")
    return chars


def encode_sequences(sequences: Union[Sequence[str], np.ndarray], alphabet: str) -> np.ndarray:
    """Residue indices ``uint8[N, L]`` with ``idx = alphabet.index(ch)`` (host-side, vectorised).

    Same integer content as ``string_to_one_hot(s, alphabet).argmax(1)`` per sequence; raises
    ``ValueError`` for a character that is not in ``alphabet`` like ``str.index`` does
    (sequence_utils.py:46).  The GPU path does this inside ``flexs_encode_dev`` instead.
    """
    chars = sequences_to_char_array(sequences)
    idx = _lut(alphabet)[chars]
    if idx.size and int(idx.max()) == 255:
        pos = int(np.argmax(idx.reshape(-1) == 255))
        raise ValueError(f"substring not found: {chr(int(chars.reshape(-1)[pos]))!r} is not in alphabet {alphabet!r}")
    return idx


def decode_indices(idx: np.ndarray, alphabet: str) -> np.ndarray:
    """``uint8[N, L]`` residue indices -> numpy array of N strings (dtype ``<U{L}``)."""
    idx = np.asarray(idx)
    n, length = idx.shape
    table = np.frombuffer(alphabet.encode("latin-1"), dtype=np.uint8)
    chars = np.ascontiguousarray(table[idx])
    if length == 0:
        return np.array([""] * n)
    return chars.view(f"S{length}").reshape(n).astype(f"U{length}")


def construct_mutant_from_sample(pwm_sample: np.ndarray, one_hot_base: np.ndarray) -> np.ndarray:
    """One-hot mutant of ``one_hot_base`` with the rows named in ``pwm_sample`` replaced
    (sequence_utils.py:20-29; used by the BO explorer)."""
    mutant = np.array(one_hot_base, dtype=np.float64, copy=True)
    rows, cols = np.nonzero(pwm_sample)
    mutant[rows, :] = 0
    mutant[rows, cols] = 1
    return mutant


def string_to_one_hot(sequence: str, alphabet: str) -> np.ndarray:
    """``(len(sequence), len(alphabet))`` float64 one-hot (sequence_utils.py:32-47).
    Raises ``ValueError`` for characters outside the alphabet."""
    idx = encode_sequences([sequence], alphabet)[0] if len(sequence) else np.zeros(0, dtype=np.uint8)
    out = np.zeros((len(sequence), len(alphabet)))
    out[np.arange(len(sequence)), idx] = 1
    return out


def one_hot_to_string(one_hot: Union[List[List[int]], np.ndarray], alphabet: str) -> str:
    """Per-position argmax (first maximum wins) mapped back to characters (sequence_utils.py:50-66)."""
    return "".join(alphabet[i] for i in np.argmax(one_hot, axis=1))


def generate_single_mutants(wt: str, alphabet: str) -> List[str]:
    """``wt`` followed by every single-site substitution, including the identity ones
    (sequence_utils.py:69-77)."""
    out = [wt]
    for pos in range(len(wt)):
        for ch in alphabet:
            out.append(wt[:pos] + ch + wt[pos + 1:])
    return out


def generate_random_sequences(length: int, number: int, alphabet: str) -> List[str]:
    """``number`` uniform random sequences (sequence_utils.py:80-84)."""
    return ["".join(random.choice(alphabet) for _ in range(length)) for _ in range(number)]


def generate_random_mutant(sequence: str, mu: float, alphabet: str) -> str:
    """Mutate each residue with probability ``mu`` to a uniform draw from ``alphabet`` (which may
    be the same residue) — sequence_utils.py:87-108, same order of ``random`` calls."""
    out = []
    for ch in sequence:
        out.append(random.choice(alphabet) if random.random() < mu else ch)
    return "".join(out)
