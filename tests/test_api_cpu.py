"""Object-model behaviour of the API mirror, without a GPU: ports of the reference's
tests/test_nbody_decay.py and the non-generating parts of tests/test_chain.py, tests/test_random.py."""

import os
import re

import numpy as np
import pytest
import torch

import phasespace_b200 as phsp
from phasespace_b200 import GenParticle, nbody_decay

from .helpers import decays as D

B0_MASS, PION_MASS = D.B0_MASS, D.PION_MASS


# ---- tests/test_nbody_decay.py:18-60
def test_no_names():
    decay = nbody_decay(B0_MASS, [PION_MASS] * 3)
    assert decay.name == "top"
    assert all(part.name == f"p_{i}" for i, part in enumerate(decay.children))


def test_top_name():
    decay = nbody_decay(B0_MASS, [PION_MASS] * 3, top_name="B0")
    assert decay.name == "B0"
    assert all(part.name == f"p_{i}" for i, part in enumerate(decay.children))


def test_children_names():
    names = [f"pion_{i}" for i in range(3)]
    decay = nbody_decay(B0_MASS, [PION_MASS] * 3, names=names)
    assert decay.name == "top" and names == [p.name for p in decay.children]


def test_all_names():
    names = [f"pion_{i}" for i in range(3)]
    decay = nbody_decay(B0_MASS, [PION_MASS] * 3, top_name="B0", names=names)
    assert decay.name == "B0" and names == [p.name for p in decay.children]


def test_mismatching_names():
    with pytest.raises(ValueError):
        nbody_decay(B0_MASS, [PION_MASS] * 3, names=[f"pion_{i}" for i in range(4)])


# ---- tests/test_chain.py:21-92, 117-128
def test_name_clashes():
    with pytest.raises(KeyError):
        GenParticle("Top", 0).set_children(GenParticle("Kstarz", D.KSTARZ_MASS), GenParticle("Kstarz", D.KSTARZ_MASS))
    with pytest.raises(KeyError):
        GenParticle("Top", 0).set_children(GenParticle("Top", D.KSTARZ_MASS), GenParticle("Kstarz", D.KSTARZ_MASS))
    with pytest.raises(KeyError):
        GenParticle("Top", 0).set_children(
            GenParticle("Kstarz0", D.KSTARZ_MASS).set_children(GenParticle("K+", D.KAON_MASS), GenParticle("pi-", PION_MASS)),
            GenParticle("Kstarz0", D.KSTARZ_MASS).set_children(GenParticle("K+", D.KAON_MASS), GenParticle("pi-_1", PION_MASS)),
        )


def test_wrong_children():
    with pytest.raises(ValueError):
        GenParticle("Top", 0).set_children(GenParticle("Kstarz0", D.KSTARZ_MASS))


def test_grandchildren():
    top = GenParticle("Top", 0)
    assert not top.has_children and not top.has_grandchildren
    assert not top.set_children(GenParticle("Child1", D.KSTARZ_MASS), GenParticle("Child2", D.KSTARZ_MASS)).has_grandchildren
    assert top.has_children
    assert D.to_gen(D.kstar_gamma_desc()).has_grandchildren


def test_reset_children():
    top = GenParticle("Top", 0).set_children(GenParticle("Child1", D.KSTARZ_MASS), GenParticle("Child2", D.KSTARZ_MASS))
    with pytest.raises(ValueError):
        top.set_children(GenParticle("Child3", D.KSTARZ_MASS), GenParticle("Child4", D.KSTARZ_MASS))


def test_no_children():
    with pytest.raises(ValueError):
        GenParticle("Top", 0).generate(n_events=1)


def test_resonance_top():
    kstar = D.to_gen(D.kstar_gamma_desc()).children[0]
    with pytest.raises(ValueError):
        kstar.generate(n_events=1)


def test_repr():
    b0 = D.to_gen(D.kstar_gamma_desc())
    assert str(b0) == "<phasespace.GenParticle: name='B0' mass=5279.58 children=[K*0, gamma]>"
    assert str(b0.children[0]) == "<phasespace.GenParticle: name='K*0' mass=variable children=[K+, pi-]>"


def test_mass_handling():
    p = GenParticle("x", 3)
    assert p.has_fixed_mass and p.get_mass().dtype == torch.float64 and tuple(p.get_mass().shape) == (1,)

    def fn(min_mass, max_mass, n_events):
        return (min_mass + max_mass) / 2

    r = GenParticle("r", fn)
    assert not r.has_fixed_mass and r._mass.__name__ == "fn"
    out = r.get_mass(np.zeros(4), np.full(4, 10.0), 4)
    assert tuple(out.shape) == (4,) and torch.allclose(out, torch.full((4,), 5.0, dtype=torch.float64))

    def fn_seed(min_mass, max_mass, n_events, seed=None):
        assert isinstance(seed, phsp.random.Generator)  # ps.py:176-181: the seed is turned into a generator
        return min_mass

    assert tuple(GenParticle("s", fn_seed).get_mass(np.ones(2), np.ones(2), 2, seed=5).shape) == (2,)


def test_removed_entry_points():
    from phasespace_b200 import phasespace as ps

    with pytest.raises(RuntimeError):
        nbody_decay(B0_MASS, [PION_MASS] * 3).generate_tensor(10)
    with pytest.raises(NameError):
        ps.Particle()
    with pytest.raises(NameError):
        ps.generate_decay()


def test_public_exports():
    for name in ("nbody_decay", "GenParticle", "random", "to_vectors", "numpy"):  # reference __init__.py:16
        assert hasattr(phsp, name)
    from phasespace_b200.fromdecay import GenMultiDecay  # noqa: F401


# ---- tests/test_random.py:7-26, re-expressed for the counter-based generator
@pytest.mark.parametrize("seed", [lambda: 15, lambda: phsp.random.Generator.from_seed(15)])
def test_get_rng(seed):
    rng1, rng2 = phsp.random.get_rng(seed()), phsp.random.get_rng(seed())
    rnd1, rnd2 = rng1.uniform_full_int(shape=(100,)), rng2.uniform_full_int(shape=(100,))
    rng3 = phsp.random.get_rng()
    rng4 = phsp.random.get_rng(seed())
    _ = rng4.split(1)  # advances the generator
    rnd3, rnd4 = rng3.uniform_full_int(shape=(100,)), rng4.uniform_full_int(shape=(100,))
    assert torch.equal(rnd1, rnd2)
    assert not torch.equal(rnd1, rnd3)
    assert not torch.equal(rnd4, rnd3)
    assert not torch.equal(rnd4, rnd1)


def test_generator_passthrough_and_global():
    g = phsp.random.Generator(3)
    assert phsp.random.get_rng(g) is g
    assert phsp.random.get_rng(None) is phsp.random.get_global_generator()
    assert g.reserve(10) == 0 and g.reserve(5) == 10 and g.offset == 15
    u = phsp.random.Generator(9).uniform((1000,))
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0


def test_host_philox_matches_the_oracle_stream():
    from oracle.philox import philox4x32_10

    words = phsp.random._philox_words(0x0123456789ABCDEF, (1 << 32) - 5, 16, stream=3)
    ev = np.arange((1 << 32) - 5, (1 << 32) + 11, dtype=np.uint64)
    ref = philox4x32_10((ev & 0xFFFFFFFF, ev >> np.uint64(32), 0, 3), (0x89ABCDEF, 0x01234567))
    for got, want in zip(words, ref):
        np.testing.assert_array_equal(got.numpy().astype(np.uint64), want)


# ---- the product path has no CPU fallback and never touches the oracle
def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        nbody_decay(B0_MASS, [PION_MASS] * 3).generate(10)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        phsp.kinematics.mass(np.ones((2, 4)))


def test_product_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "phasespace_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "genbod_ref" not in text and "genbod_numpy" not in text.replace("oracle/genbod_numpy.py", ""), f


def test_compiled_tree_follows_children_added_to_a_leaf():
    """The reference latches only decaying particles (ps.py:207-210, 318): a leaf can get children after the tree above
    it was compiled, and the next compile() must see them (host-only: ps_tree_create needs no GPU)."""
    kstar = GenParticle("K*0", 895.81)
    top = GenParticle("B0", 5279.58).set_children(kstar, GenParticle("gamma", 0.0))
    first = top.compile()
    assert first.names == ["K*0", "gamma"] and top.compile() is first
    other = nbody_decay(1000.0, [100.0, 100.0])  # a change elsewhere re-validates, the tree is kept
    assert other.compile().n_particles == 2 and top.compile() is first
    kstar.set_children(GenParticle("K+", 493.677), GenParticle("pi-", 139.57018))
    second = top.compile()
    assert second is not first and second.names == ["K*0", "gamma", "K+", "pi-"] and second.n_draws == 4


def test_resonant_generating_particle_is_described_as_a_fixed_top():
    """ps.py:534-543: with boost_to a resonance may be the particle generate() is called on; its descriptor is a fixed-mass
    top whose mass the kernel takes from boost_to, and without boost_to the reference's ValueError stays."""
    res = GenParticle("R", lambda min_mass, max_mass, n_events: min_mass).set_children(
        GenParticle("a", 100.0), GenParticle("b", 200.0))
    compiled = res.compile()
    assert compiled.top_is_resonance and compiled.n_user == 0 and not np.isfinite(compiled.w_max)
    with pytest.raises(ValueError, match="Cannot use resonance as top particle"):
        compiled.launch(10, seed=1)
    with pytest.raises(ValueError, match="Cannot use resonance as top particle"):
        res.generate(10)


def test_output_buffer_validation_rules():
    from phasespace_b200.phasespace import _check_out

    ok = torch.empty((3, 10, 4), dtype=torch.float64)
    _check_out(ok, "particles", (3, 10, 4), None, pinned_host=True)
    _check_out(torch.empty((3, 16, 4), dtype=torch.float64)[:, :10], "particles", (3, 10, 4), None, pinned_host=True, planes=True)
    for bad, kwargs in ((torch.empty((3, 9, 4), dtype=torch.float64), {}),
                        (torch.empty((3, 10, 4), dtype=torch.float32), {}),
                        (torch.empty((3, 4, 10), dtype=torch.float64).permute(0, 2, 1), {"planes": True}),
                        (torch.empty((3, 10, 8), dtype=torch.float64)[:, :, ::2], {"planes": True}),
                        (np.empty((3, 10, 4)), {})):
        with pytest.raises(ValueError, match="particles"):
            _check_out(bad, "particles", (3, 10, 4), None, pinned_host=True, **kwargs)
    with pytest.raises(ValueError, match="must be on"):
        _check_out(ok, "particles", (3, 10, 4), torch.device("cuda", 0))
