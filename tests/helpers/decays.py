"""Test decays: the trees of the reference's tests/helpers/decays.py:14-76, built twice -- as
phasespace_b200 GenParticles and as oracle ``Part`` trees -- from one description."""

from oracle import genbod_numpy as orc

B0_MASS = 5279.58
PION_MASS = 139.57018
KAON_MASS = 493.677
K1_MASS = 1272.0
K1_WIDTH = 90.0
KSTARZ_MASS = 895.81
KSTARZ_WIDTH = 47.4

# description: (name, mass, kind, width, [children])
def flat_desc(n, top=B0_MASS, masses=None):
    masses = masses or [PION_MASS] * n
    return ("top", top, "fixed", 0.0, [(f"p_{i}", m, "fixed", 0.0, []) for i, m in enumerate(masses)])


def kstar_gamma_desc(kind="gauss", width=KSTARZ_WIDTH):
    return ("B0", B0_MASS, "fixed", 0.0, [
        ("K*0", KSTARZ_MASS, kind, width, [("K+", KAON_MASS, "fixed", 0.0, []), ("pi-", PION_MASS, "fixed", 0.0, [])]),
        ("gamma", 0.0, "fixed", 0.0, []),
    ])


def k1_gamma_desc(k1_kind="gauss", kst_kind="gauss"):
    return ("B+", B0_MASS, "fixed", 0.0, [
        ("K1+", K1_MASS, k1_kind, K1_WIDTH, [
            ("K*0", KSTARZ_MASS, kst_kind, KSTARZ_WIDTH,
             [("K+", KAON_MASS, "fixed", 0.0, []), ("pi-", PION_MASS, "fixed", 0.0, [])]),
            ("pi+", PION_MASS, "fixed", 0.0, []),
        ]),
        ("gamma", 0.0, "fixed", 0.0, []),
    ])


def to_oracle(desc):
    name, mass, kind, width, children = desc
    k = orc.FIXED if kind == "fixed" else (orc.USER if kind == "user" else orc.KIND_NAMES[kind])
    return orc.Part(name, mass, k, width, [to_oracle(c) for c in children])


def to_gen(desc, user_masses=None):
    """``user`` nodes become GenParticles with a Python mass callable that returns ``user_masses[name]``."""
    from phasespace_b200 import GenParticle
    from phasespace_b200.fromdecay import mass_functions as mf

    name, mass, kind, width, children = desc
    if kind == "fixed":
        part = GenParticle(name, mass)
    elif kind == "user":
        import torch

        table = torch.as_tensor(user_masses[name], dtype=torch.float64)
        part = GenParticle(name, lambda min_mass, max_mass, n_events, table=table: table.to(min_mass.device))
    else:
        part = GenParticle(name, mf.DEFAULT_CONVERTER[kind](mass, width))
    if children:
        part.set_children(*[to_gen(c, user_masses) for c in children])
    return part
