"""The decay trees behind tests/golden/reference_vectors.npz (written by tests/golden/make_reference_golden.py from the
reference's own source run on NumPy), as the (name, mass, kind, width, children) descriptions of helpers/decays.py.
Resonances whose mass came from a Python callable in the reference run are ``user`` nodes here: the oracle and the
kernel are fed the masses the callable returned (stored in the .npz as ``<case>/user_mass/<name>``)."""

import os

import numpy as np

from . import decays as D

GOLDEN = os.path.join(os.path.dirname(__file__), "..", "golden", "reference_vectors.npz")
MUON_MASS = 105.6583755


def _leaf(name, mass):
    return (name, mass, "fixed", 0.0, [])


def _kstar_gamma(kind, mass=D.KSTARZ_MASS):
    return ("B0", D.B0_MASS, "fixed", 0.0, [
        ("K*0", mass, kind, 0.0, [_leaf("K+", D.KAON_MASS), _leaf("pi-", D.PION_MASS)]), _leaf("gamma", 0.0)])


_K1_FIXED = ("B+", D.B0_MASS, "fixed", 0.0, [
    ("K1+", D.K1_MASS, "fixed", 0.0, [
        ("K*0", D.KSTARZ_MASS, "fixed", 0.0, [_leaf("K+", D.KAON_MASS), _leaf("pi-", D.PION_MASS)]), _leaf("pi+", D.PION_MASS)]),
    _leaf("gamma", 0.0)])
_NINE_BELOW = ("A", 12000.0, "fixed", 0.0, [
    ("X", 0.0, "user", 0.0, [_leaf(f"x{i}", D.PION_MASS) for i in range(9)]), _leaf("g", 0.0), _leaf("h", 938.272)])

CASES = {
    "flat2": D.flat_desc(2),
    "flat3": D.flat_desc(3),
    "flat4": D.flat_desc(4),
    "flat10": D.flat_desc(10),
    "flat5_mixed": D.flat_desc(5, masses=[D.KAON_MASS, D.PION_MASS, 0.0, MUON_MASS, D.PION_MASS]),
    "flat4_boost": D.flat_desc(4),
    "flat3_boost_one_row": D.flat_desc(3),
    "kstar_gamma_fixed": _kstar_gamma("fixed"),
    "kstar_gamma_callable": _kstar_gamma("user", 0.0),
    "kstar_gamma_callable_boost": _kstar_gamma("user", 0.0),
    "k1_gamma_callable": ("B+", D.B0_MASS, "fixed", 0.0, [
        ("K1+", 0.0, "user", 0.0, [
            ("K*0", 0.0, "user", 0.0, [_leaf("K+", D.KAON_MASS), _leaf("pi-", D.PION_MASS)]), _leaf("pi+", D.PION_MASS)]),
        _leaf("gamma", 0.0)]),
    "two_branches": ("B", D.B0_MASS, "fixed", 0.0, [
        ("X", 1500.0, "fixed", 0.0, [_leaf("a", D.PION_MASS), _leaf("b", D.KAON_MASS)]),
        ("Y", 0.0, "user", 0.0, [_leaf("c", D.PION_MASS), _leaf("d", D.PION_MASS), _leaf("e", 0.0)]),
        _leaf("f", MUON_MASS)]),
    # second batch
    "flat8": D.flat_desc(8),
    "flat9": D.flat_desc(9),
    "flat2_boost": D.flat_desc(2, masses=[D.PION_MASS, D.KAON_MASS]),
    "flat3_massless": D.flat_desc(3, top=100.0, masses=[0.0, 0.0, 0.0]),
    "kstar_gamma_fixed_boost_one_row": _kstar_gamma("fixed"),
    "k1_gamma_fixed": _K1_FIXED,
    "k1_gamma_fixed_boost": _K1_FIXED,
    "nine_body_subdecay": _NINE_BELOW,
    "nine_body_subdecay_boost": _NINE_BELOW,
    "three_levels_three_body": ("T", 9000.0, "fixed", 0.0, [
        ("R1", 0.0, "user", 0.0, [
            ("R2", 2000.0, "fixed", 0.0, [_leaf("a", D.PION_MASS), _leaf("b", D.KAON_MASS), _leaf("c", 0.0)]),
            _leaf("d", D.PION_MASS), _leaf("e", MUON_MASS)]),
        _leaf("f", 938.272)]),
}


def load(case):
    """dict with draws, boost (or None), weights, weights_max, names, particles (P, n, 4), user_masses {name: (n,)}."""
    data = np.load(GOLDEN)
    prefix = case + "/"
    out = {k[len(prefix):]: data[k] for k in data.files if k.startswith(prefix)}
    out["boost"] = out.get("boost")
    out["names"] = [str(s) for s in out["names"]]
    out["user_masses"] = {k[len("user_mass/"):]: v for k, v in out.items() if k.startswith("user_mass/")}
    return out
