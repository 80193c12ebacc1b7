"""CPU restatement of the reference's table landscapes.  TEST INFRASTRUCTURE ONLY — never imported by the product.

Pinned: tests/golden/ref_landscapes.json holds outputs of the reference's OWN classes (imported from /root/reference
by tests/golden/make_golden_landscapes.py) for these functions' inputs; tests/test_oracle_golden.py checks them.

  * additive_fitness  — AdditiveAAVPackaging (flexs/landscapes/additive_aav_packaging.py:79-118)
  * tfbinding_dict / tfbinding_fitness — TFBinding (flexs/landscapes/tf_binding.py:22-44)
"""
import numpy as np


def compute_max_possible(data, phenotype):
    """additive_aav_packaging.py:79-96: per position the best residue whose packaging log2 is above -6."""
    best_seq, max_fitness = "", 0
    for pos in data:
        current_max, current_best = -10, "M"
        for aa in data[pos]:
            fit = data[pos][aa][phenotype]
            if fit > current_max and data[pos][aa]["log2_packaging_v_wt"] > -6:
                current_best, current_max = aa, fit
        best_seq += current_best
        max_fitness += current_max
    return best_seq, max_fitness


def additive_fitness(sequences, data, phenotype, start, mfm, max_possible, noise_draws):
    """additive_aav_packaging.py:98-118 with the np.random.normal draws passed in (one per sequence, in order)."""
    out = []
    for seq, eps in zip(sequences, noise_draws):
        total = 0
        for i, s in enumerate(seq):
            if s in data[start + i]:
                total += data[start + i][s][phenotype]
        normed = (total + mfm * max_possible) / (max_possible * (mfm + 1))
        out.append(max(0, normed + eps))
    return np.array(out)


def tfbinding_dict(fwd, rev, escore):
    """tf_binding.py:33-41: min-max normalised E-score; the reverse-strand column is written last."""
    escore = np.asarray(escore, dtype=np.float64)
    norm = (escore - escore.min()) / (escore.max() - escore.min())
    d = dict(zip(fwd, norm))
    d.update(zip(rev, norm))
    return d


def tfbinding_fitness(sequences, table):
    """tf_binding.py:43-44 (KeyError for a sequence that is not in the file)."""
    return np.array([table[s] for s in sequences])
