#!/usr/bin/env python3
"""Writes tests/golden/reference_fromdecay.json: what the REFERENCE'S OWN ``GenMultiDecay.from_dict``
(src/phasespace/fromdecay/genmultidecay.py:28-153, 200-340, run unmodified) builds for the example decay chains.

The reference's fromdecay package imports TensorFlow, ``particle``, ``zfit`` and ``zfit_physics``; none is installed.
from_dict itself only needs ``tf.cast``, ``tnp.prod`` and ``Particle.from_evtgen_name(name).mass/.width``, so this
script installs the TensorFlow stand-in of make_reference_golden.py plus: a ``particle`` module whose ``Particle`` reads
the repo's own table (phasespace_b200/fromdecay/pdg.py -- masses are therefore NOT an independent check, the
traversal is), and empty ``zfit`` / ``zfit_physics`` modules (only touched when a mass function is *called*).
Recorded per decay mode: probability, and for every particle its name, whether its mass is fixed, the mass value or
the mass function's name, and its children -- the naming scheme ("pi+ [0]"), the cartesian product of sub-modes with
multiplied probabilities, the precedence zfit key > particle_model_map > DEFAULT_MASS_FUNC and the width tolerance.

    python tests/golden/make_reference_fromdecay.py          (needs /root/reference)
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_reference_golden as g  # noqa: E402

CASES = [  # (chain attribute in tests/helpers/example_decay_chains.py, from_dict keyword arguments)
    ("dplus_single", {"tolerance": 1e-10}),
    ("dplus_single", {}),
    ("pi0_4branches", {"tolerance": 1e-10}),
    ("dplus_4grandbranches", {"tolerance": 1e-10}),
    ("dplus_4grandbranches", {"tolerance": 1e-10, "particle_model_map": {"pi0": "gauss"}}),
    ("dstarplus_big_decay", {}),
    ("dstarplus_big_decay", {"tolerance": 1e-10}),
    ("dstarplus_big_decay", {"tolerance": 1e-10, "particle_model_map": {"D0": "bw", "pi0": "gauss"}}),
]


def describe(part):
    """Structure of a GenParticle tree of either implementation."""
    fixed = bool(part.has_fixed_mass)
    if fixed:
        mass = float(np.asarray(part.get_mass() if not hasattr(part, "_mass_val") or callable(part._mass_val) else part._mass_val).reshape(-1)[0])
    else:
        mass = part._mass.__name__
    return {"name": part.name, "fixed": fixed, "mass": mass, "children": [describe(c) for c in part.children]}


def install_fromdecay_stand_ins():
    g.install_tensorflow_stand_in()
    tf = sys.modules["tensorflow"]
    tf.cast = lambda x, dtype: np.asarray(x, dtype=np.float64)
    from phasespace_b200.fromdecay import pdg  # the repo's particle table: plain Python, no CUDA needed

    class Particle:
        def __init__(self, mass, width):
            self.mass, self.width = mass, width

        @classmethod
        def from_evtgen_name(cls, name):
            return cls(*pdg._TABLE[name])  # (mass, width); pdg.lookup would find this very stand-in

    particle = types.ModuleType("particle")
    particle.Particle = Particle
    sys.modules["particle"] = particle
    for name in ("zfit", "zfit_physics"):
        sys.modules.setdefault(name, types.ModuleType(name))


def main():
    install_fromdecay_stand_ins()
    from tests.helpers import example_decay_chains as chains  # the DecayLanguage dicts (hand-written: see its docstring)

    sys.path.insert(0, g.REFERENCE_SRC)
    import phasespace  # noqa: F401  the reference
    from phasespace.fromdecay import GenMultiDecay

    assert os.path.realpath(phasespace.__file__).startswith(os.path.realpath(g.REFERENCE_SRC))
    out = []
    for attr, kwargs in CASES:
        multi = GenMultiDecay.from_dict(getattr(chains, attr), **kwargs)
        out.append({"chain": attr, "kwargs": kwargs,
                    "modes": [{"probability": float(p), "tree": describe(part)} for p, part in multi.gen_particles]})
    with open(os.path.join(HERE, "reference_fromdecay.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("written", sum(len(c["modes"]) for c in out), "decay modes in", len(out), "cases")


if __name__ == "__main__":
    main()
