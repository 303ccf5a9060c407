#!/usr/bin/env python3
"""Writes tests/golden/oracle_vectors.npz: small input/output vectors of the CPU oracle.

The reference itself cannot be imported here (it needs TensorFlow), so these vectors pin the ORACLE
(against accidental change) and give the GPU tests fixed inputs; the vectors that come from the
reference proper are the README rows in readme_kstargamma.json.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle import genbod_numpy as orc  # noqa: E402
from oracle.philox import uniform_table  # noqa: E402
from tests.helpers import decays as D  # noqa: E402


def boost_rows(n, seed):
    rng = np.random.default_rng(seed)
    p = rng.normal(0.0, 20000.0, (n, 3))
    return np.concatenate([p, np.sqrt((p * p).sum(1, keepdims=True) + D.B0_MASS**2)], axis=1)


CASES = {
    "flat3": (D.flat_desc(3), 64, None),
    "flat5_mixed": (D.flat_desc(5, masses=[D.KAON_MASS, D.PION_MASS, 0.0, 105.6583755, D.PION_MASS]), 48, None),
    "flat4_boost": (D.flat_desc(4), 64, 7),
    "kstar_gamma_relbw": (D.kstar_gamma_desc("relbw"), 64, None),
    "kstar_gamma_gauss_boost": (D.kstar_gamma_desc("gauss"), 48, 9),
    "k1_gamma_bw_gauss": (D.k1_gamma_desc("bw", "gauss"), 48, None),
}


def build():
    out = {}
    for name, (desc, n, boost_seed) in CASES.items():
        tree = D.to_oracle(desc)
        draws = uniform_table(20240 + len(name), 0, n, orc.n_draws(tree))
        boost = None if boost_seed is None else boost_rows(n, boost_seed)
        w, wmax, parts = orc.generate(tree, n, draws, boost_to=boost, normalize_weights=False)
        out[f"{name}/draws"] = draws
        if boost is not None:
            out[f"{name}/boost"] = boost
        out[f"{name}/weights"] = w
        out[f"{name}/weights_max"] = wmax
        out[f"{name}/particles"] = np.stack([parts[k] for k in orc.output_names(tree)])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_vectors.npz"), **build())
    print("written")
