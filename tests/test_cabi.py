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
    assert _lib.lib.ps_abi_version() == _lib.ABI_VERSION == 4


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


def _node_table(tree):
    handle = tree.compile().handle
    n = _lib.lib.ps_tree_node_table_size(handle)
    out = np.zeros(max(n, 1))
    meta, w_min = (ctypes.c_int32 * 4)(), ctypes.c_double()
    assert _lib.lib.ps_tree_node_table(handle, out.ctypes.data, n, ctypes.addressof(meta), ctypes.addressof(w_min)) == n
    return out[:n], list(meta), w_min.value


def _locate_node(u, mlog, j_lo, j_hi):
    """numpy statement of ps_samplers.cuh node_locate: index of the nearest node and u - u_node (exact)."""
    m_nodes = 1 << mlog
    upper = u >= 0.5
    w = np.where(upper, 1.0 - u, u)
    big_j = np.where(upper, j_hi, j_lo)
    j = np.where(w > 0, -(np.frexp(w)[1] - 1) - 2, 1 << 20)
    j = np.maximum(j, 0)
    last = j >= big_j
    j = np.where(last, big_j, j)
    sh = np.where(last, big_j + 1, j + 2)
    s = np.ldexp(np.where(last, 1.0, 2.0) - np.ldexp(w, sh), mlog)
    i = np.rint(s).astype(np.int64)
    dw = np.ldexp(s - i, -sh - mlog)
    return np.where(upper, (j_lo + 1) * (m_nodes + 1), 0) + j * (m_nodes + 1) + i, np.where(upper, dw, -dw), w


@pytest.mark.parametrize("case", [
    (5279.58, 895.81, 47.4, 493.677, 139.57018, 0.0),     # K*-like: lower limit 5.5 widths below the pole, upper 92 above
    (1500.0, 895.81, 47.4, 493.677, 139.57018, 620.0),    # window that ends below the pole
    (5279.58, 895.81, 47.4, 493.677, 493.677, 0.0),       # lower limit above the pole
    (2010.26, 1864.84, 0.05, 493.677, 139.57018, 139.57018),  # both ends thousands of widths away
], ids=["kstar", "below_pole", "above_pole", "narrow"])
def test_gauss_node_table_matches_multiprecision_inverse_cdf(case):
    """ps_tree_create's table for a truncated normal with event-independent limits (ps_samplers.cuh sample_gauss_table;
    the sampler of tests/helpers/decays.py:26-37 and mass_functions.py:16-40): the kernel's arithmetic, restated in numpy on
    the host copy of the table, against a 40-digit inverse CDF, for draws all over [0, 1) and hard against both ends."""
    import phasespace_b200 as ps
    from phasespace_b200.fromdecay import mass_functions as mf

    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    top, m, width, d1, d2, third = case
    tree = ps.GenParticle("X", top).set_children(
        ps.GenParticle("R", mf.gauss_factory(m, width)).set_children(ps.GenParticle("a", d1), ps.GenParticle("b", d2)),
        ps.GenParticle("c", third))
    tab, (kind, mlog, j_lo, j_hi), w_min = _node_table(tree)
    assert kind == 1 and tab.size == 2 * ((j_lo + 1) + (j_hi + 1)) * ((1 << mlog) + 1)
    lo, hi = d1 + d2, top - third
    a, b = mp.mpf(lo - m) / width, mp.mpf(hi - m) / width
    p_a, q_a, p_b, q_b = mp.ncdf(a), mp.ncdf(-a), mp.ncdf(b), mp.ncdf(-b)
    d_p = (1 - p_a - q_b) if a <= 0 <= b else (p_b - p_a if b < 0 else q_a - q_b)
    rng = np.random.default_rng(5)
    ends = np.ldexp(rng.random(300) + 0.5, -rng.integers(1, 52, 300))
    u = np.concatenate([rng.random(1500), ends, 1.0 - ends, [0.0, 0.5, 0.25, 0.75, 2.0 ** -51, 1 - 2.0 ** -51]])
    u = np.floor(u * 2.0 ** 51) / 2.0 ** 51  # the generator's own grid
    idx, du, w = _locate_node(u, mlog, j_lo, j_hi)
    keep = w >= w_min  # closer to an end than w_min: left to the library path by the kernel
    assert keep.sum() >= u.size - 8 and (w_min == 0.0 or w_min == 2.0 ** -51)
    z, r = tab[2 * idx], tab[2 * idx + 1]
    with np.errstate(all="ignore"):
        d = du * float(d_p) * r
        z2 = z * z
        c3, c4 = z2 / 3.0 + 1.0 / 6.0, z * (z2 * 0.25 + 7.0 / 24.0)
        c5, c6 = z2 * (z2 * 0.2 + 23.0 / 60.0) + 7.0 / 120.0, z * (z2 * (z2 / 6.0 + 163.0 / 360.0) + 127.0 / 720.0)
        z_new = d * (d * (d * (d * (d * (d * c6 + c5) + c4) + c3) + 0.5 * z) + 1.0) + z
    for uk, zk in zip(u[keep], z_new[keep]):
        p_lower, p_upper = p_a + mp.mpf(float(uk)) * d_p, q_b + (1 - mp.mpf(float(uk))) * d_p
        if p_lower <= p_upper:
            exact = -mp.sqrt(2) * mp.erfinv(1 - 2 * p_lower) if p_lower > mp.mpf("1e-25") else None
        else:
            exact = mp.sqrt(2) * mp.erfinv(1 - 2 * p_upper) if p_upper > mp.mpf("1e-25") else None
        if exact is None:  # erfinv loses its argument next to 1: compare the probabilities instead
            got_p = mp.ncdf(mp.mpf(float(zk))) if p_lower <= p_upper else mp.ncdf(-mp.mpf(float(zk)))
            want_p = min(p_lower, p_upper)
            assert abs(got_p - want_p) <= 1e-14 * want_p * (1 + zk * zk), (uk, zk)
        else:
            assert abs(zk - float(exact)) <= 4 * np.spacing(max(abs(zk), 1.0)), (uk, zk, float(exact))


def test_node_tables_are_offered_per_sampler():
    """Cauchy resonances and trees without a sampled resonance have no table; ps_tree_relbw_nodes counts only its own."""
    import phasespace_b200 as ps
    from phasespace_b200.fromdecay import mass_functions as mf

    def chain(factory):
        return ps.GenParticle("B0", 5279.58).set_children(
            ps.GenParticle("K*0", factory(895.81, 47.4)).set_children(ps.GenParticle("K+", 493.677), ps.GenParticle("pi-", 139.57018)),
            ps.GenParticle("gamma", 0.0))
    size = lambda tree: _lib.lib.ps_tree_node_table_size(tree.compile().handle)  # noqa: E731
    assert size(chain(mf.gauss_factory)) > 0 and size(chain(mf.relativistic_breitwigner_factory)) > 0
    assert size(chain(mf.breitwigner_factory)) == 0
    assert _lib.lib.ps_tree_relbw_nodes(chain(mf.gauss_factory).compile().handle) == 0
    assert size(ps.nbody_decay(5279.58, [139.57018] * 3)) == 0


@pytest.mark.parametrize("case", [
    (5279.58, 895.81, 47.4, 493.677, 139.57018, 0.0),                          # K*-like
    (2010.26, 1864.84, 1.6050035525481589e-09, 493.677, 139.57039, 139.57039),  # a D0 with its physical width: kappa 5e12
    (50000.0, 895.81, 47.4, 493.677, 139.57018, 0.0),                           # upper limit a thousand widths away
    (5279.58, 775.26, 155.0, 139.57018, 139.57018, 139.57018),                  # broad: kappa 20
], ids=["kstar", "d0_physical_width", "far_upper_limit", "kappa_20"])
def test_relbw_node_table_matches_multiprecision_inverse_cdf(case):
    """ps_tree_create's table for a relativistic Breit-Wigner with event-independent limits (ps_samplers.cuh
    sample_relbw_table; mass_functions.py:70-102): the kernel's arithmetic, restated in numpy on the host copy of the
    table, against a 50-digit inversion of the CDF, for draws all over [0, 1) and hard against both ends.  The nodes are
    solved relative to the end of the mass range their half of the uniform approaches, so the reference is too.  Also
    guards the cost of building it (this D0 once took 0.5 s, and was refused a table after 1.6 s below a D*+)."""
    import time

    import phasespace_b200 as ps
    from phasespace_b200.fromdecay import mass_functions as mf

    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    top, m, width, d1, d2, third = case
    t0 = time.perf_counter()
    tree = ps.GenParticle("X", top).set_children(
        ps.GenParticle("R", mf.relativistic_breitwigner_factory(m, width)).set_children(
            ps.GenParticle("a", d1), ps.GenParticle("b", d2)), ps.GenParticle("c", third))
    tab, (kind, mlog, j_lo, j_hi), w_min = _node_table(tree)
    assert time.perf_counter() - t0 < 0.4
    assert kind == 3 and w_min == 0.0 and tab.size == ((j_lo + 1) + (j_hi + 1)) * ((1 << mlog) + 1)
    # the constants as the host computes them in fp64 (ps_api.cu host_sampler_constants / host_sampler_limits)
    a, b = m * m, m * width
    p = np.sqrt(0.5 * (np.sqrt(a * a + b * b) + a))
    q = b / (2.0 * p)
    kappa = 2.0 * p / q
    lo, hi = d1 + d2, top - third

    def g_fp64(z):
        return np.arctan(z) + np.arctan(z + kappa) + np.log(((z + kappa) ** 2 + 1.0) / (z * z + 1.0)) / kappa

    d_g = g_fp64((hi - p) / q) - g_fp64((lo - p) / q)
    rng = np.random.default_rng(3)
    ends = np.ldexp(rng.random(40) + 0.5, -rng.integers(1, 52, 40))
    u = np.concatenate([rng.random(120), ends, 1.0 - ends, [0.0, 0.5, 2.0 ** -51, 1 - 2.0 ** -51]])
    u = np.floor(u * 2.0 ** 51) / 2.0 ** 51
    idx, du, _ = _locate_node(u, mlog, j_lo, j_hi)
    z = tab[idx]
    h = du * d_g * (1.0 / (kappa * kappa + 4.0))
    y = z + kappa
    aa, bb = z * z + 1.0, y * y + 1.0
    p1, p2, p3 = 2.0 * (z * bb + y * aa), 8.0 * z * y + 2.0 * (aa + bb), 12.0 * (y + z)
    d = h * (aa * bb)
    al, be, ga = h * p1, h * p2 * d, h * p3 * d * d
    x = np.clip(p + q * (z + d * (1 + al / 2 + (be + al * al) / 6 + (ga + 4 * al * be + al ** 3) / 24)), lo, hi)

    k_mp = mp.mpf(kappa)

    def g_mp(t):
        return mp.atan(t) + mp.atan(t + k_mp) + mp.log(((t + k_mp) ** 2 + 1) / (t * t + 1)) / k_mp

    z_lo, z_hi = (mp.mpf(lo) - mp.mpf(p)) / mp.mpf(q), (mp.mpf(hi) - mp.mpf(p)) / mp.mpf(q)
    g_lo, g_hi = g_mp(z_lo), g_mp(z_hi)
    for uk, xk in zip(u, x):
        if uk == 0.0:
            assert xk == lo
            continue
        target = g_lo + mp.mpf(float(uk)) * mp.mpf(d_g) if uk < 0.5 else g_hi - (1 - mp.mpf(float(uk))) * mp.mpf(d_g)
        z_true = mp.findroot(lambda t: g_mp(t) - target, mp.mpf(float((xk - p) / q)), tol=mp.mpf("1e-40"), maxsteps=200)
        x_true = min(max(float(mp.mpf(p) + mp.mpf(q) * z_true), lo), hi)
        zt = float(z_true)
        # what one rounding of the CDF costs at this mass (tests/helpers/conditioning.py sampler_noise), plus the series'
        # truncation at the table's design point (alpha <= 0.004: up to 2e-14 for a resonance as broad as the rho)
        conditioning = 2.0 ** -53 * (zt * zt + 1.0) * ((zt + kappa) ** 2 + 1.0) / (kappa * kappa + 4.0) * q / x_true
        assert abs(xk - x_true) / x_true <= 5e-14 + 8.0 * conditioning, (uk, xk, x_true)
