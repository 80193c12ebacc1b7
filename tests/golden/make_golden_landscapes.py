"""Generates tests/golden/ref_landscapes.json, aav_450_540_subs.json and tfbind_4mers.txt.  AUTHORING container only:

    python tests/golden/make_golden_landscapes.py

Outputs of the reference's OWN ``AdditiveAAVPackaging`` and ``TFBinding`` classes (imported from /root/reference with
the hand-built ``flexs`` namespace of make_golden.py).  The AAV fixture is the window 450..539 of the reference's
AAV2_single_subs.json restricted to the keys the class reads (two phenotypes + packaging), so the class under test
can be built without the 5 MB file; the TF-binding fixture is a synthetic 4-mer file in the reference's format.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import REF, _load, reference_namespace  # noqa: E402


def fl(x):
    """JSON-safe exact float: hex string."""
    return float(x).hex()


def main():
    reference_namespace()
    aav = _load("flexs.landscapes.additive_aav_packaging", REF / "flexs/landscapes/additive_aav_packaging.py")
    tfb = _load("flexs.landscapes.tf_binding", REF / "flexs/landscapes/tf_binding.py")
    rng = np.random.default_rng(450540)

    # ---- AdditiveAAVPackaging ---------------------------------------------------------------
    full = json.load(open(REF / "flexs/landscapes/data/additive_aav_packaging/AAV2_single_subs.json"))
    keep = ("log2_heart_v_wt", "log2_liver_v_wt", "log2_packaging_v_wt")
    window = {pos: {aa: {k: v[k] for k in keep} for aa, v in full[pos].items()} for pos in full if 450 <= int(pos) < 540}
    json.dump(window, open(HERE / "aav_450_540_subs.json", "w"))

    cases = []
    residues = "ILVAGMFYWEDQNHCRKSPT"
    for phenotype, mfm, noise, seed in [("heart", 1, 0, 0), ("liver", 1, 0, 1), ("heart", 2, 0, 2), ("liver", 0.5, 0.1, 3)]:
        land = aav.AdditiveAAVPackaging(phenotype=phenotype, minimum_fitness_multiplier=mfm, start=450, end=540, noise=noise)
        wt = land.wild_type
        seqs = [wt, land.top_seq]
        for m in range(60):                                       # random mutants of the wild type (1..4 / 1..30 substitutions)
            s = list(wt)
            for p in rng.choice(90, size=int(rng.integers(1, 5 if m < 40 else 31)), replace=False):
                s[p] = residues[int(rng.integers(0, 20))]
            seqs.append("".join(s))
        seqs += ["".join(residues[i] for i in rng.integers(0, 20, size=90)) for _ in range(10)]   # far from wt -> clipped
        seqs += [wt[:10] + "*" + wt[11:], wt[:3] + "XBZ" + wt[6:], wt[:45], "M"]               # stop codon, unknown, short
        np.random.seed(seed)
        out = land.get_fitness(seqs)
        after = float(np.random.random())                         # where the global numpy stream stands afterwards
        cases.append({"phenotype": phenotype, "mfm": mfm, "noise": noise, "seed": seed, "sequences": seqs,
                      "fitness": [fl(v) for v in out], "dtype": str(out.dtype), "cost": land.cost,
                      "next_uniform": fl(after), "top_seq": land.top_seq, "max_possible": fl(land.max_possible),
                      "wild_type": wt, "name": land.name})
    land = aav.AdditiveAAVPackaging(phenotype="heart", start=450, end=540)
    allzero = land.get_fitness(["P" * 90, "W" * 90])
    too_long = None
    try:
        land.get_fitness([land.wild_type + "A"])
    except Exception as e:  # noqa: BLE001
        too_long = [type(e).__name__, str(e)]

    # ---- TFBinding on a synthetic 4-mer file in the reference's format ------------------------
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rows, seen = [], set()
    import itertools

    for tup in itertools.product("ACGT", repeat=4):
        s = "".join(tup)
        r = "".join(comp[c] for c in reversed(s))
        if s in seen or r in seen:
            continue
        seen.update((s, r))
        rows.append((s, r))
    esc = np.round(rng.uniform(-0.5, 0.5, size=len(rows)), 5)
    with open(HERE / "tfbind_4mers.txt", "w") as f:
        f.write("8-mer\t8-mer\tE-score\tMedian\tZ-score\n")
        for (s, r), e in zip(rows, esc):
            f.write(f"{s}\t{r}\t{e:.5f}\t{1000.0 + 1000 * e:.2f}\t{e * 3:.4f}\n")
    tf = tfb.TFBinding(str(HERE / "tfbind_4mers.txt"))
    keys = sorted(tf.sequences)
    query = [keys[i] for i in rng.integers(0, len(keys), size=64)]
    tf_out = tf.get_fitness(query)
    missing = None
    try:
        tf.get_fitness(["ACGT", "ACGN"])
    except Exception as e:  # noqa: BLE001
        missing = [type(e).__name__, str(e)]
    reg = aav.registry()

    # a real measurement file: a checksum of the whole dict (the file itself does not travel)
    real = tfb.TFBinding(str(REF / "flexs/landscapes/data/tf_binding/ARX_REF_R1_8mers.txt"))
    rk = sorted(real.sequences)
    real_sum = float(np.sum([real.sequences[k] for k in rk]))
    json.dump({
        "aav": cases, "aav_all_clipped": {"out": [float(v) for v in allzero], "dtype": str(allzero.dtype)},
        "aav_too_long": too_long, "aav_registry": reg,
        "tf": {"name": tf.name, "dict": {k: fl(tf.sequences[k]) for k in keys}, "query": query,
               "fitness": [fl(v) for v in tf_out], "dtype": str(tf_out.dtype), "cost": tf.cost, "missing": missing},
        "tf_real": {"file": "ARX_REF_R1_8mers.txt", "n_keys": len(rk), "sum": fl(real_sum),
                    "probe": {k: fl(real.sequences[k]) for k in ("GCTCGAGC", "AAAAAAAA", "TTTTTTTT", "ACGTACGT")}},
    }, open(HERE / "ref_landscapes.json", "w"))
    print("landscape fixtures written")


if __name__ == "__main__":
    main()
