"""GenMultiDecay.from_dict: ports of the construction checks of the reference's
tests/fromdecay/test_fromdecay.py:40-209 (the generate() halves are in tests/test_multidecay_gpu.py)."""

from copy import deepcopy

import pytest
from numpy.testing import assert_almost_equal

from phasespace_b200.fromdecay import GenMultiDecay
from phasespace_b200.fromdecay.mass_functions import DEFAULT_CONVERTER

from .helpers import example_decay_chains


def test_invalid_chain():
    dc = deepcopy(example_decay_chains.dplus_single)
    dc["D+"][0]["fs"][0] = 1
    with pytest.raises(TypeError):
        GenMultiDecay.from_dict(dc)


def test_invalid_mass_func_name():
    dc = deepcopy(example_decay_chains.dplus_single)
    dc["D+"][0]["fs"][-1]["pi0"][0]["zfit"] = "invalid name"
    with pytest.raises(KeyError):
        GenMultiDecay.from_dict(dc, tolerance=1e-10)


def test_single_chain():
    container = GenMultiDecay.from_dict(example_decay_chains.dplus_single, tolerance=1e-10)
    output_decay = container.gen_particles
    assert len(output_decay) == 1
    prob, gen = output_decay[0]
    assert_almost_equal(prob, 1)
    assert gen.name == "D+"
    assert {p.name for p in gen.children} == {"K-", "pi+", "pi+ [0]", "pi0"}
    for p in gen.children:
        if "pi0" == p.name[:3]:
            assert not p.has_fixed_mass
            assert p._mass.__name__ == "relbw"
            assert {c.name for c in p.children} == {"gamma", "gamma [0]"}
        else:
            assert p.has_fixed_mass


def test_branching_children():
    container = GenMultiDecay.from_dict(example_decay_chains.pi0_4branches, tolerance=1e-10)
    assert len(container.gen_particles) == 4
    assert_almost_equal(sum(d[0] for d in container.gen_particles), 1)


def test_branching_grandchilden():
    decay_dict = deepcopy(example_decay_chains.dplus_4grandbranches)
    for mass_function, decay_mode in zip(("relbw", "bw", "gauss"), decay_dict["D+"][0]["fs"][-1]["pi0"]):
        decay_mode["zfit"] = mass_function
    container = GenMultiDecay.from_dict(decay_dict, tolerance=1e-10)
    output_decays = container.gen_particles
    assert len(output_decays) == 4
    assert_almost_equal(sum(d[0] for d in output_decays), 1)
    for p, mass_func in zip(output_decays, ("relbw", "bw", "gauss", GenMultiDecay.DEFAULT_MASS_FUNC)):
        assert p[1].children[-1].name == "pi0"
        assert p[1].children[-1]._mass.__name__ == mass_func


def test_particle_model_map():
    container = GenMultiDecay.from_dict(example_decay_chains.dplus_4grandbranches, particle_model_map={"pi0": "bw"},
                                        tolerance=1e-10)
    assert len(container.gen_particles) == 4
    assert_almost_equal(sum(d[0] for d in container.gen_particles), 1)
    for p in container.gen_particles:
        assert p[1].children[-1].name[:3] == "pi0"
        assert p[1].children[-1]._mass.__name__ == "bw"


def test_mass_converter():
    dc = deepcopy(example_decay_chains.dplus_4grandbranches)
    dc["D+"][0]["fs"][-1]["pi0"][-1]["zfit"] = "rel-BW"
    container = GenMultiDecay.from_dict(dc, tolerance=1e-10, mass_converter={"rel-BW": DEFAULT_CONVERTER["relbw"]})
    assert len(container.gen_particles) == 4
    for decay in container.gen_particles:
        for child in decay[1].children:
            if "pi0" in child.name:
                assert not child.has_fixed_mass


def test_big_decay():
    container = GenMultiDecay.from_dict(example_decay_chains.dstarplus_big_decay)
    assert_almost_equal(sum(d[0] for d in container.gen_particles), 1)
    assert len(container.gen_particles) == 1 + 16 + 4
    # default tolerance 0.01 MeV: the pi0 (width ~8e-6 MeV) keeps a fixed mass
    for _, gen in container.gen_particles:
        assert gen.name == "D*+"


def test_mass_width_tolerance_and_default_mass_func():
    old_tol, old_func = GenMultiDecay.MASS_WIDTH_TOLERANCE, GenMultiDecay.DEFAULT_MASS_FUNC
    try:
        GenMultiDecay.MASS_WIDTH_TOLERANCE = 1e-10
        for p in GenMultiDecay.from_dict(example_decay_chains.dplus_4grandbranches).gen_particles:
            assert p[1].children[-1].name[:3] == "pi0" and not p[1].children[-1].has_fixed_mass
        GenMultiDecay.DEFAULT_MASS_FUNC = "bw"
        for p in GenMultiDecay.from_dict(example_decay_chains.dplus_4grandbranches, tolerance=1e-10).gen_particles:
            assert p[1].children[-1]._mass.__name__ == "bw"
    finally:
        GenMultiDecay.MASS_WIDTH_TOLERANCE, GenMultiDecay.DEFAULT_MASS_FUNC = old_tol, old_func


def test_unique_names():
    from phasespace_b200.fromdecay.genmultidecay import _unique_name

    seen = set()
    assert [_unique_name("pi+", seen) for _ in range(4)] == ["pi+", "pi+ [0]", "pi+ [1]", "pi+ [2]"]


def _describe(part):
    fixed = bool(part.has_fixed_mass)
    mass = float(part._mass_val) if fixed else part._mass.__name__
    return {"name": part.name, "fixed": fixed, "mass": mass, "children": [_describe(c) for c in part.children]}


def _load_reference_fromdecay():
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_fromdecay.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("case", range(8))
def test_from_dict_matches_the_reference_implementation(case):
    """tests/golden/reference_fromdecay.json holds what the reference's OWN from_dict / _recursively_traverse /
    _unique_name (run unmodified through stand-ins, tests/golden/make_reference_fromdecay.py) builds for these chains:
    same decay modes in the same order, same probabilities, same particle names ("pi+ [0]"), same fixed/resonant split
    under the width tolerance, same mass-function choice (zfit key > particle_model_map > default)."""
    ref = _load_reference_fromdecay()[case]
    multi = GenMultiDecay.from_dict(getattr(example_decay_chains, ref["chain"]), **ref["kwargs"])
    assert len(multi.gen_particles) == len(ref["modes"])
    for (prob, part), want in zip(multi.gen_particles, ref["modes"]):
        assert float(prob) == pytest.approx(want["probability"], rel=1e-12)
        assert _describe(part) == want["tree"]
