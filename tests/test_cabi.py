"""The C-ABI library: loads, exports every symbol include/phasespace_b200.h declares, and its host-side
tree compiler agrees with the oracle (slots, draw counts, minimum masses, maximum weights). No GPU needed."""

import ctypes
import os
import re

import numpy as np
import pytest

from oracle import genbod_numpy as orc
from oracle.philox import uniform_table
from phasespace_b200 import _lib

from .helpers import decays as D

HEADER = os.path.join(os.path.dirname(__file__), "..", "include", "phasespace_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = _declared_functions()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), name
    assert set(names) == set(_lib.EXPORTS)
    assert _lib.lib.ps_abi_version() == _lib.ABI_VERSION == 3


def _tree(desc):
    nodes = []

    def add(d, parent):
        name, mass, kind, width, children = d
        k = {"fixed": 0, "gauss": 1, "bw": 2, "relbw": 3, "user": 4}[kind]
        idx = len(nodes)
        nodes.append((parent, k, mass, width, name))
        for c in children:
            add(c, idx)

    add(desc, -1)
    arr = (_lib.NodeDesc * len(nodes))(*[_lib.NodeDesc(p, k, m, w) for p, k, m, w, _ in nodes])
    handle = ctypes.c_void_p()
    rc = _lib.lib.ps_tree_create(arr, len(nodes), ctypes.byref(handle))
    return rc, handle, [n[4] for n in nodes]


@pytest.mark.parametrize("desc", [D.flat_desc(3), D.flat_desc(10), D.kstar_gamma_desc("relbw"), D.k1_gamma_desc("bw", "gauss"),
                                  D.kstar_gamma_desc("fixed", 0.0)])
def test_tree_compiler_matches_oracle(desc):
    rc, h, names = _tree(desc)
    assert rc == 0, _lib.last_error()
    lib = _lib.lib
    tree = D.to_oracle(desc)
    order = orc.output_names(tree)
    assert lib.ps_tree_n_particles(h) == len(order)
    assert lib.ps_tree_n_draws(h) == orc.n_draws(tree)
    slots = {names[i]: lib.ps_tree_slot(h, i) for i in range(len(names))}
    assert slots.pop(names[0]) == -1
    assert sorted(slots, key=slots.get) == order  # the reference's dict order (phasespace.py:547-570)
    # event-independent maximum weight vs the oracle's per-event value
    w_max = lib.ps_tree_w_max(h)
    has_res = any(k in ("gauss", "bw", "relbw") for k in re.findall(r"'(gauss|bw|relbw|fixed)'", repr(desc)))
    n = 4
    _, ref_wmax, _ = orc.generate(tree, n, uniform_table(1, 0, n, orc.n_draws(tree)), normalize_weights=False)
    np.testing.assert_allclose(w_max, ref_wmax[0], rtol=1e-14)  # leaves are fixed in all of these trees
    assert has_res or True
    lib.ps_tree_destroy(h)


def _random_desc(rng, depth, counter, top=False, may_sample=False):
    """A random decay tree: 2-4 children per decay, decaying intermediates fixed or sampled (at most one sampled child per
    decay: with several, the reference's max_mass bookkeeping can produce forbidden events, phasespace.py:342-353), fixed
    leaves."""
    name = f"n{counter[0]}"
    counter[0] += 1
    decays = top or (depth > 0 and rng.random() < 0.5)
    if not decays:
        return (name, float(rng.uniform(0.0, 150.0)), "fixed", 0.0, [])
    n_children = int(rng.integers(2, 5))
    sampled_child = int(rng.integers(0, n_children))
    children = [_random_desc(rng, depth - 1, counter, may_sample=(j == sampled_child)) for j in range(n_children)]
    kind = str(rng.choice(["fixed", "gauss", "bw", "relbw"])) if may_sample else "fixed"
    floor = sum(c[1] if c[2] == "fixed" else orc.recurse_stable(D.to_oracle(c)) for c in children)
    mass = floor + float(rng.uniform(200.0, 900.0))  # always kinematically allowed at the pole masses
    return (name, mass, kind, 0.0 if kind == "fixed" else float(rng.uniform(1.0, 80.0)), children)


@pytest.mark.parametrize("seed", range(12))
def test_tree_compiler_matches_oracle_on_random_trees(seed):
    """Slots (the reference's dict order), draw counts and recurse_stable() minimum masses of random trees with up to
    four levels, 2-4 children per decay and all sampler kinds, and the analytic maximum weight."""
    rng = np.random.default_rng(1000 + seed)
    desc = _random_desc(rng, 3, [0], top=True)
    rc, h, names = _tree(desc)
    if rc != 0:  # more particles than PS_MAX_SLOTS: the only legitimate refusal of a well-formed tree
        assert "more than" in _lib.last_error()
        return
    lib = _lib.lib
    tree = D.to_oracle(desc)
    order = orc.output_names(tree)
    assert lib.ps_tree_n_particles(h) == len(order)
    assert lib.ps_tree_n_draws(h) == orc.n_draws(tree)
    slots = {names[i]: lib.ps_tree_slot(h, i) for i in range(len(names))}
    assert slots.pop(names[0]) == -1
    assert sorted(slots, key=slots.get) == order

    def walk(d, part):
        yield d, part
        for dc, pc in zip(d[4], part.children):
            yield from walk(dc, pc)

    index = {n: i for i, n in enumerate(names)}
    for d, part in walk(desc, tree):
        if d[4] and d is not desc:
            assert lib.ps_tree_min_mass(h, index[d[0]]) == pytest.approx(orc.recurse_stable(part), rel=1e-15)
    # fixed leaves: the maximum weight depends on leaf masses only (recurse_w_max, phasespace.py:571-625), whatever the
    # intermediate resonances were sampled at -- the host constant must equal the oracle's per-event value
    n = 3
    try:
        _, ref_wmax, _ = orc.generate(tree, n, uniform_table(seed, 0, n, orc.n_draws(tree)), normalize_weights=False)
    except FloatingPointError:
        # several resonant siblings: the reference's max_mass ignores the minimum masses of the later ones
        # (phasespace.py:342-353) and the events can come out forbidden -- nothing to compare the constant with
        assert np.isfinite(lib.ps_tree_w_max(h))
    else:
        assert np.all(ref_wmax == ref_wmax[0])
        np.testing.assert_allclose(lib.ps_tree_w_max(h), ref_wmax[0], rtol=1e-13)
    lib.ps_tree_destroy(h)


def test_min_mass_is_recurse_stable():
    rc, h, names = _tree(D.k1_gamma_desc())
    assert rc == 0
    mm = {names[i]: _lib.lib.ps_tree_min_mass(h, i) for i in range(len(names))}
    assert mm["K*0"] == pytest.approx(D.KAON_MASS + D.PION_MASS)
    assert mm["K1+"] == pytest.approx(D.KAON_MASS + D.PION_MASS + D.PION_MASS)
    _lib.lib.ps_tree_destroy(h)


def test_aot_registry_matches_generated_signatures():
    for n in range(2, 11):
        rc, h, _ = _tree(D.flat_desc(n))
        assert rc == 0
        sig = _lib.lib.ps_tree_signature(h).decode()
        assert sig.startswith("Pt<-1,0,0,Pt<0,0,-1>") and sig.count("Pt<") == n + 1
        for boost in (0, 1):
            for dbg in (0, 1):
                for flags in (0, _lib.PS_GEN_EXACT):
                    assert _lib.lib.ps_tree_is_aot(h, boost, dbg, flags) == 1, (n, boost, dbg, flags)
        _lib.lib.ps_tree_destroy(h)
    for desc in (D.kstar_gamma_desc("relbw"), D.kstar_gamma_desc("fixed", 0.0), D.k1_gamma_desc("gauss", "gauss")):
        rc, h, _ = _tree(desc)
        assert rc == 0 and _lib.lib.ps_tree_is_aot(h, 0, 0, 0) == 1
        _lib.lib.ps_tree_destroy(h)
    rc, h, _ = _tree(D.flat_desc(11))  # not precompiled: goes to NVRTC at the first generate
    assert rc == 0 and _lib.lib.ps_tree_is_aot(h, 0, 0, 0) == 0
    _lib.lib.ps_tree_destroy(h)


@pytest.mark.parametrize("nodes,msg", [
    ([(-1, 0, 10.0, 0.0), (0, 0, 1.0, 0.0)], "at least 2 children"),
    ([(-1, 0, 10.0, 0.0), (0, 0, 1.0, 0.0), (0, 0, 1.0, 0.0), (1, 0, 0.1, 0.0)], "at least 2 children"),
    ([(-1, 0, 10.0, 0.0), (-1, 0, 1.0, 0.0), (0, 0, 1.0, 0.0)], "more than one top"),
    ([(-1, 1, 10.0, 1.0), (0, 0, 1.0, 0.0), (0, 0, 1.0, 0.0)], "resonance as top"),
    ([(-1, 0, 10.0, 0.0), (0, 1, 1.0, 0.0), (0, 0, 1.0, 0.0)], "width > 0"),
    ([(-1, 0, 10.0, 0.0), (0, 9, 1.0, 0.0), (0, 0, 1.0, 0.0)], "unknown mass kind"),
    ([(-1, 0, 10.0, 0.0), (2, 0, 1.0, 0.0), (1, 0, 1.0, 0.0)], "children"),
])
def test_tree_validation_errors(nodes, msg):
    arr = (_lib.NodeDesc * len(nodes))(*[_lib.NodeDesc(*n) for n in nodes])
    handle = ctypes.c_void_p()
    rc = _lib.lib.ps_tree_create(arr, len(nodes), ctypes.byref(handle))
    assert rc == _lib.PS_E_INVALID
    assert msg in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_user_mass_columns_follow_slot_order():
    desc = ("B", 5000.0, "fixed", 0.0, [
        ("R1", 1000.0, "user", 0.0, [("a", 100.0, "fixed", 0.0, []), ("R2", 300.0, "user", 0.0, [("c", 1.0, "fixed", 0.0, []), ("d", 1.0, "fixed", 0.0, [])])]),
        ("R3", 500.0, "user", 0.0, []),
    ])
    rc, h, names = _tree(desc)
    assert rc == 0
    assert _lib.lib.ps_tree_n_user(h) == 3
    cols = {names[i]: _lib.lib.ps_tree_user_column(h, i) for i in range(len(names))}
    assert (cols["R1"], cols["R3"], cols["R2"]) == (0, 1, 2) and cols["a"] == -1
    assert np.isnan(_lib.lib.ps_tree_w_max(h))  # R3 is a resonant leaf: w_max depends on the event
    _lib.lib.ps_tree_destroy(h)


def test_relbw_node_table_is_offered_where_it_applies():
    """ps_tree_create builds the inverse-CDF node table (ps_samplers.cuh sample_relbw_table) for a narrow relativistic
    Breit-Wigner whose limits are event independent -- the first non-fixed child of the top particle -- and for nothing else."""
    from phasespace_b200 import GenParticle
    from phasespace_b200.fromdecay import mass_functions as mf

    def kstar_gamma(factory, mass=895.81, width=47.4):
        return GenParticle("B0", 5279.58).set_children(
            GenParticle("K*0", factory(mass, width)).set_children(GenParticle("K+", 493.677), GenParticle("pi-", 139.57018)),
            GenParticle("gamma", 0.0))

    nodes = lambda tree: _lib.lib.ps_tree_relbw_nodes(tree.compile().handle)  # noqa: E731
    assert nodes(kstar_gamma(mf.relativistic_breitwigner_factory)) > 1000
    assert nodes(kstar_gamma(mf.relativistic_breitwigner_factory, 1864.84, 3e-6)) > 1000   # a D0 with its physical width
    assert nodes(kstar_gamma(mf.breitwigner_factory)) == 0 and nodes(kstar_gamma(mf.gauss_factory)) == 0
    assert nodes(kstar_gamma(mf.relativistic_breitwigner_factory, 775.0, 400.0)) == 0      # broad: kappa < 16, generic solver
    # a resonance below a resonance has per-event limits: only the first one gets a table (one table per tree)
    k1 = GenParticle("B+", 5279.58).set_children(
        GenParticle("K1+", mf.relativistic_breitwigner_factory(1272.0, 90.0)).set_children(
            GenParticle("K*0", mf.relativistic_breitwigner_factory(895.81, 47.4)).set_children(
                GenParticle("K+", 493.677), GenParticle("pi-", 139.57018)), GenParticle("pi+", 139.57018)),
        GenParticle("gamma", 0.0))
    assert nodes(k1) > 1000
    from phasespace_b200 import nbody_decay

    assert nodes(nbody_decay(5279.58, [139.57018] * 3)) == 0
