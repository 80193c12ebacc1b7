"""CPU: the table decomposition behind ``flexs_b200/csrc/cnn_k9.cu`` restated in numpy and checked against the oracle.

For a 4-letter alphabet conv1 -> ReLU -> conv2("same") -> ReLU of cnn.py:23-40 depends on 9 residues (fewer at the two
positions next to either end, where the "same" padding truncates the window), so it is a table lookup.  This file
mirrors the kernel's index arithmetic — the segment layout (interior | o=0 | o=1 | o=T-2 | o=T-1), the big-endian base-4
code, the tap offset of the truncated windows, the row -> (tile, core matrix, stream) mapping of the producer and the
row/tap relation of conv3 — and requires the result to equal the oracle's activations.  The CUDA kernel itself is
compared with the oracle in tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from oracle import flexs_oracle as fo

K, A, F = 5, 4, 32
N_MAIN, N_E7, N_E8 = 1 << 18, 1 << 14, 1 << 16
ENT_EL0 = N_MAIN
ENT_EL1 = ENT_EL0 + N_E7
ENT_ER1 = ENT_EL1 + N_E8
ENT_ER0 = ENT_ER1 + N_E8
N_ENT = ENT_ER0 + N_E7


def table_rows(entries, ws):
    """k9_build_kernel: value (before the fp16 split and the activation scale) of the given table entries."""
    w1, b1, w2, b2 = [np.asarray(w, dtype=np.float64) for w in ws[:4]]
    out = np.zeros((len(entries), F))
    for n, e in enumerate(entries):
        length, joff, code = 9, 0, e
        if e >= ENT_ER0:
            length, code = 7, e - ENT_ER0
        elif e >= ENT_ER1:
            length, code = 8, e - ENT_ER1
        elif e >= ENT_EL1:
            length, joff, code = 8, 1, e - ENT_EL1
        elif e >= ENT_EL0:
            length, joff, code = 7, 2, e - ENT_EL0
        res = [(code >> (2 * (length - 1 - m))) & 3 for m in range(length)]
        acc = b2.copy()
        for i in range(length - (K - 1)):
            h1 = np.maximum(b1 + sum(w1[m, res[i + m]] for m in range(K)), 0.0)
            acc += h1 @ w2[i + joff]
        out[n] = np.maximum(acc, 0.0)
    return out


def pack_residues(seq):
    """Producer warps of cnn_k9_kernel: residues as 2 bits each, 16 per 32-bit word, first residue in the top bits,
    one zero word in front and zero padding behind."""
    L = len(seq)
    nw = (L + 15) // 16
    words = [0]
    for w in range(nw):
        word = 0
        for r in range(16):
            i = 16 * w + r
            word = (word << 2) | ((int(seq[i]) & 3) if i < L else 0)
        words.append(word)
    words.append(0)
    return words


def row_entry(seq, o, T):
    """Producer warps of cnn_k9_kernel: table entry of conv2 position ``o`` of one sequence (or -1: zero row), with the
    kernel's arithmetic: input row c of tile q is position o = 16 q + c - 1, its 9-residue window starts 2 (c + 13) bits
    into the three packed words q-1, q, q+1; truncated windows at the left end come out of the zero word in front,
    those at the right end drop the bits past the end."""
    if o < 0 or o >= T:
        return -1
    q, c = divmod(o + 1, 16)
    if c >= 16:  # rows 16, 17 of a tile are rows 0, 1 of the next one
        q, c = q + 1, c - 16
    pw = pack_residues(seq)
    w0, w1 = pw[q], pw[q + 1]
    w2 = pw[q + 2] if q + 2 < len(pw) else 0
    bits = (w0 << 64) | (w1 << 32) | w2
    code = (bits >> (96 - 2 * (c + 13) - 18)) & 0x3FFFF
    base, rsh = 0, 0
    if o <= 1:
        base = ENT_EL0 if o == 0 else ENT_EL1
    if o >= T - 2:
        base, rsh = (ENT_ER1, 2) if o == T - 2 else (ENT_ER0, 4)
    return base + (code >> rsh)


@pytest.mark.parametrize("L", [8, 9, 14, 20, 37, 100])
def test_table_rows_equal_conv2_activations(L):
    shp = fo.CNNShape(L, A, F, 100, K)
    ws = fo.trained_like_weights(shp.weight_shapes(), 4)
    T = shp.conv_len
    idx = np.random.default_rng(L).integers(0, A, size=(3, L), dtype=np.uint8)
    x = fo.one_hot(idx, A)
    h1 = fo.relu(fo.conv1d(x, ws[0].astype(np.float64), ws[1].astype(np.float64), "valid"))
    h2 = fo.relu(fo.conv1d(h1, ws[2].astype(np.float64), ws[3].astype(np.float64), "same"))
    assert h2.shape == (3, T, F)
    for s in range(3):
        ents = [row_entry(idx[s], o, T) for o in range(T)]
        assert min(ents) >= 0 and max(ents) < N_ENT
        assert sum(e >= N_MAIN for e in ents) == 4  # exactly the four truncated windows leave the interior table
        np.testing.assert_allclose(table_rows(ents, ws), h2[s], rtol=1e-12, atol=1e-12)
    assert row_entry(idx[0], -1, T) == -1 and row_entry(idx[0], T, T) == -1


def test_tile_row_mapping_and_conv3_taps():
    """The producer writes input row c of tile q (h2 position 16q + c - 1) into core matrix c; tap j of the MMA reads
    core matrix c + j for output row c, i.e. h2[o + j - 1] with o = 16q + c: conv3 with "same" padding of k3 = 3."""
    L = 41
    shp = fo.CNNShape(L, A, F, 100, K)
    ws = [w.astype(np.float64) for w in fo.trained_like_weights(shp.weight_shapes(), 8)]
    T = shp.conv_len
    nti = (T + 15) // 16
    idx = np.random.default_rng(1).integers(0, A, size=(1, L), dtype=np.uint8)
    x = fo.one_hot(idx, A)
    h1 = fo.relu(fo.conv1d(x, ws[0], ws[1], "valid"))
    h2 = fo.relu(fo.conv1d(h1, ws[2], ws[3], "same"))
    h3 = fo.relu(fo.conv1d(h2, ws[4], ws[5], "same"))[0]
    feat = np.zeros(F)
    for q in range(nti):
        cms = np.zeros((18, F))
        for c in range(18):
            e = row_entry(idx[0], 16 * q + c - 1, T)
            if e >= 0:
                cms[c] = table_rows([e], ws)[0]
        for c in range(16):
            o = 16 * q + c
            y = np.maximum(ws[5] + sum(cms[c + j] @ ws[4][j] for j in range(3)), 0.0)
            if o < T:
                np.testing.assert_allclose(y, h3[o], rtol=1e-12, atol=1e-12)
                feat = np.maximum(feat, y)
    np.testing.assert_allclose(feat, h3.max(axis=0), rtol=1e-12, atol=1e-12)
