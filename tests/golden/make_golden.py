"""Generates the committed fixtures of tests/golden/.

  reference_energies.json   published ensemble energies of the reference, copied out of
                            /root/reference/data/comp_energy_stats (bin size 32) + comp_energy_results.csv
                            — the only golden VALUES the reference holds for the sweep
  sweep_golden.json         sha256 of post-sweep fields / observables produced by the CPU oracle
                            (oracle/lqft_oracle.c) on seeded inputs.  They pin the oracle against drift
                            and let the GPU tests check against committed vectors.

Run from the repo root:  python tests/golden/make_golden.py   (needs /root/reference for the first file)
"""
import csv
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/data"

SWEEP_CASES = [
    # name, (X, Y, T), beta, seed, sweeps, wilson (w, h)
    ("plain_4x6x8", (4, 6, 8), 0.25, 11, 5, (0, 0)),
    ("plain_16", (16, 16, 16), 0.2, 0xC0FFEE, 4, (0, 0)),
    ("plain_6x10x14", (6, 10, 14), 0.31, 5, 3, (0, 0)),
    ("plain_2x2x2", (2, 2, 2), 0.5, 9, 6, (0, 0)),
    ("wilson_16_5x16", (16, 16, 16), 0.2, 77, 4, (5, 16)),
    ("wilson_8x4x6_3x2", (8, 4, 6), 0.26, 3, 5, (3, 2)),
    ("wilson_8x2x4_8x4", (8, 2, 4), 0.3, 4, 5, (8, 4)),
    ("plain_24", (24, 24, 24), 0.29, 123, 2, (0, 0)),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seeded_field(n, seed=20261019):
    return np.random.default_rng(seed).integers(-128, 128, size=n, dtype=np.int32)


def make_sweeps():
    out = {}
    for name, size, beta, seed, sweeps, (w, h) in SWEEP_CASES:
        lat = O.Lattice(size)
        wil = O.Wilson(lat, w, h) if w and h else None
        v = seeded_field(lat.n_sites)
        acc = [O.sweep(lat, v, beta, seed, s, order=1, wilson=wil) for s in range(sweeps)]
        e_int, e_wrapped = O.integer_energy(lat, v)
        sites = [O.pick_site(seed, 4, 7, r, lat.n_sites) for r in range(5)]
        out[name] = {
            "size": list(size), "beta": beta, "seed": seed, "sweeps": sweeps, "wilson": [w, h],
            "accepted": acc, "field_sha256": sha(v), "e_int": e_int, "s_float": O.float_action(lat, v),
            "corr_sites": sites, "corr": O.correlator(lat, v, sites).tolist(),
        }
    return out


def make_reference_energies():
    runs = {}
    with open(os.path.join(REF, "comp_energy_results.csv")) as fh:
        for row in csv.DictReader(fh):
            runs[int(row["index"])] = row
    out = []
    for idx, row in sorted(runs.items()):
        path = os.path.join(REF, "comp_energy_stats", f"energy_{idx}_stats.csv")
        if not os.path.exists(path):
            continue
        with open(path) as fh:
            stats = {int(r["bin_size"]): r for r in csv.DictReader(fh)}
        s = stats[max(stats)]
        comptype = row["comptype"]
        w = h = 0
        if "Wilson" in comptype:
            dims = [tok for tok in comptype.split() if "x" in tok and tok.replace("x", "").isdigit()][0]
            w, h = (int(v) for v in dims.split("x"))
        out.append({"index": idx, "x": int(row["x"]), "y": int(row["y"]), "t": int(row["t"]),
                    "temp": float(row["temp"]), "wilson": [w, h], "bin_size": max(stats),
                    "energy": float(s["energy"]), "energy_err": float(s["energy_err"])})
    return out


def make_reference_corr_lengths():
    """Second-moment correlation lengths (corr12 +- corr12_err, largest bin size) keyed by (L, temp)."""
    runs = {}
    with open(os.path.join(REF, "results_comp.csv")) as fh:
        for row in csv.DictReader(fh):
            runs[int(row["index"])] = row
    out = []
    for idx, row in sorted(runs.items()):
        path = os.path.join(REF, "comp_correlation_stats", f"correlation_{idx}_len.csv")
        if not os.path.exists(path):
            continue
        with open(path) as fh:
            stats = {int(r["bin_size"]): r for r in csv.DictReader(fh)}
        s = stats[max(stats)]
        if not s["corr12"] or not s["corr12_err"]:
            continue
        out.append({"index": idx, "x": int(row["x"]), "y": int(row["y"]), "t": int(row["t"]), "temp": float(row["temp"]),
                    "bin_size": max(stats), "corr12": float(s["corr12"]), "corr12_err": float(s["corr12_err"])})
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "sweep_golden.json"), "w") as fh:
        json.dump(make_sweeps(), fh, indent=1)
    if os.path.isdir(REF):
        with open(os.path.join(HERE, "reference_energies.json"), "w") as fh:
            json.dump(make_reference_energies(), fh, indent=1)
        with open(os.path.join(HERE, "reference_corr_lengths.json"), "w") as fh:
            json.dump(make_reference_corr_lengths(), fh, indent=1)
    print("ok")
