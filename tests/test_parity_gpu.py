"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerance (BASELINE.json north_star): 1e-12 relative.  Weights are compared relatively; momentum
components cross zero, so 4-vectors are compared as |delta| / M_top.
Both arithmetic variants are checked: ``exact`` (reference operation order) and the default fast one.
"""

import json
import os

import numpy as np
import pytest
import torch

from oracle import genbod_numpy as orc
from oracle.philox import uniform_table

from .helpers import conditioning
from .helpers import decays as D

pytestmark = pytest.mark.gpu

TOL = 1e-12
N = 20000


def _np(t):
    return t.detach().cpu().numpy()


_SOFT = bool(os.environ.get("PS_PARITY_NOASSERT"))  # calibration runs only: report the margins, assert nothing
C_AMP = 8.0  # events are held to max(tol, C_AMP * eps * A), A = tests/helpers/conditioning.py (measured margins: _report)


def _report(label, **fields):
    """PS_PARITY_REPORT=<file>: append the measured margins of every comparison (kept under profiles/)."""
    path = os.environ.get("PS_PARITY_REPORT")
    if path:
        with open(path, "a") as fh:
            fh.write(json.dumps({"case": label, **fields}) + "\n")


def _check_events(err, bound, amp, tol, label):
    """``err`` (per event: relative weight difference, or 4-vector difference on the event's scale) stays below the
    event's own ``bound`` = max(tol, c * A_i) (conditioning.py).  There is no blanket allowance: an event above ``tol``
    must be explained by its own amplification A_i -- a two-body step near its threshold, where the weight itself goes
    to zero, a mass re-derived from a boosted 4-vector, a sampled mass.  How many such events there are and how close
    the worst one comes to its bound is reported."""
    above = err > tol
    worst = int(np.argmax(err / bound))
    _report(label, n=int(err.size), tol=tol, n_above_tol=int(above.sum()), max_err=float(err.max()),
            max_err_over_bound=float((err / bound).max()),
            max_err_over_eps_amp=float((err / (conditioning.EPS * np.maximum(amp, 1.0))).max()),
            min_amp_of_events_above_tol=float(amp[above].min()) if above.any() else None, max_amp=float(amp.max()))
    assert _SOFT or err[worst] <= bound[worst], (
        f"{label}: event {worst}: error {err[worst]:.3e} > bound {bound[worst]:.3e} (A = {amp[worst]:.4g}); "
        f"{int(above.sum())} of {err.size} events above {tol:g}")


def _rel(got, ref):
    return np.where(ref == 0.0, np.abs(got), np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))


def _check_against(desc, got, ref, boost_np, label, tol=TOL):
    """``got`` / ``ref``: ``(weights, [weights_max,] {name: (n, 4)})`` of the kernel and of the reference / oracle."""
    parts_ref = ref[-1]
    n = ref[0].shape[0]
    amp, w_bound, v_bound = conditioning.event_bounds(desc, parts_ref, boost_np, tol=tol, c_amp=C_AMP)
    _check_events(_rel(_np(got[0]), ref[0]), w_bound, amp, tol, label + "/w")
    if len(ref) == 3:
        # the maximum weight is evaluated at the kinematic limits: no threshold factor of its own.  Under boost_to it
        # depends on the event through kin.mass(boost_to), the same expression in reference and kernel.
        wmax_err = _rel(_np(got[1]), ref[1])
        _check_events(wmax_err, np.full(n, tol), np.ones(n), tol, label + "/w_max")
    if boost_np is None:
        scale = np.full(n, float(desc[1]))
    else:
        scale = np.broadcast_to(np.asarray(boost_np, dtype=np.float64).reshape(-1, 4)[:, 3], (n,))
    v_err = np.zeros(n)
    for name, vec in parts_ref.items():
        v_err = np.maximum(v_err, np.max(np.abs(_np(got[-1][name]) - vec), axis=1) / scale)
    _check_events(v_err, v_bound, amp, tol, label + "/vec")


def _compare(desc, n, *, exact, use_rng, boost=None, normalize=True, seed=1234, tol=TOL):
    gen, tree = D.to_gen(desc), D.to_oracle(desc)
    draws = uniform_table(seed, 0, n, orc.n_draws(tree))
    boost_np = None if boost is None else np.asarray(boost, dtype=np.float64)
    ref = orc.generate(tree, n, draws, boost_to=boost_np, normalize_weights=normalize)
    kwargs = dict(boost_to=None if boost is None else torch.as_tensor(boost_np), normalize_weights=normalize, exact=exact)
    if use_rng:
        got = gen.generate(n, seed=seed, **kwargs)
    else:
        got = gen.generate(n, debug_uniforms=torch.as_tensor(draws), **kwargs)
    assert list(got[-1].keys()) == orc.output_names(tree)
    label = f"{desc[0]}/{len(ref[-1])}p/{'exact' if exact else 'fast'}/{'rng' if use_rng else 'dbg'}" + ("/boost" if boost is not None else "")
    _check_against(desc, got, ref, boost_np, label, tol)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("n_body", [2, 3, 4, 5, 7, 10])
def test_flat_debug_uniforms(n_body, exact):
    _compare(D.flat_desc(n_body), N, exact=exact, use_rng=False)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("n_body", [2, 3, 4, 10])
def test_flat_inkernel_philox(n_body, exact):
    """Normal mode: the in-kernel Philox stream equals the oracle's table bit for bit, so the events agree too."""
    _compare(D.flat_desc(n_body), N, exact=exact, use_rng=True, seed=77)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_mixed_masses_unnormalized(exact):
    desc = D.flat_desc(4, masses=[D.KAON_MASS, D.PION_MASS, 0.0, 105.6583755])
    _compare(desc, N, exact=exact, use_rng=False, normalize=False)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("kind", ["fixed", "gauss", "bw", "relbw"])
def test_kstar_gamma(kind, exact):
    desc = D.kstar_gamma_desc(kind) if kind != "fixed" else D.kstar_gamma_desc("fixed", 0.0)
    _compare(desc, N, exact=exact, use_rng=False)
    _compare(desc, N, exact=exact, use_rng=True, normalize=False)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("kinds", [("fixed", "fixed"), ("gauss", "gauss"), ("relbw", "relbw"), ("bw", "gauss")])
def test_k1_gamma(kinds, exact):
    _compare(D.k1_gamma_desc(*kinds), N, exact=exact, use_rng=False)


def _fonll_like_boost(n, mass, seed=1234):
    """Synthetic stand-in for tests/helpers/rapidsim.py:46-51 of the reference (MeV)."""
    rng = np.random.default_rng(seed)
    pt = 1000.0 * rng.exponential(5.0, n)
    eta = rng.uniform(2.0, 5.0, n)
    phi = rng.uniform(0.0, 2 * np.pi, n)
    px, py, pz = pt * np.cos(phi), pt * np.sin(phi), pt * np.sinh(eta)
    return np.stack([px, py, pz, np.sqrt(px * px + py * py + pz * pz + mass * mass)], axis=1)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_boost_to_per_event(exact):
    boost = _fonll_like_boost(N, D.B0_MASS)
    # flat tree: the weights depend on the boost only through kin.mass(boost_to), the same expression in oracle and
    # kernel; the boosted vectors carry the reference's own gamma = 1/sqrt(1 - beta^2) rounding (conditioning.py)
    _compare(D.flat_desc(4), N, exact=exact, use_rng=False, boost=boost)
    # chain: the reference re-derives each resonance mass as sqrt(E^2 - p^2) of its *boosted* 4-vector
    # (phasespace.py:322 via 564), an expression whose rounding error grows like eps * gamma^2 -- that is part of every
    # event's own bound (conditioning.py), not a blanket tolerance
    _compare(D.kstar_gamma_desc("gauss"), N, exact=exact, use_rng=True, boost=boost, normalize=False, seed=5)
    # moderate boosts (gamma <= 20, an LHC B meson)
    slow = boost.copy()
    slow[:, :3] *= 20.0 * D.B0_MASS / np.maximum(np.linalg.norm(slow[:, :3], axis=1, keepdims=True), 20.0 * D.B0_MASS)
    slow[:, 3] = np.sqrt((slow[:, :3] ** 2).sum(1) + D.B0_MASS**2)
    _compare(D.kstar_gamma_desc("gauss"), N, exact=exact, use_rng=True, boost=slow, normalize=False)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_boost_to_broadcast(exact):
    boost = _fonll_like_boost(1, D.B0_MASS, seed=5)
    _compare(D.flat_desc(3), N, exact=exact, use_rng=False, boost=boost)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_ten_body_boost_to(exact):
    """The top-down path (nine or more children) under boost_to: no early stores, every product boosted at the end."""
    n = 4000
    _compare(D.flat_desc(10), n, exact=exact, use_rng=True, boost=_fonll_like_boost(n, D.B0_MASS, seed=9), normalize=False)
    _compare(D.flat_desc(9), n, exact=exact, use_rng=False, boost=_fonll_like_boost(1, D.B0_MASS, seed=3))


def test_sharding_is_bit_identical():
    """Contiguous index ranges generated separately equal the single launch bit for bit."""
    gen = D.to_gen(D.flat_desc(3))
    comp = gen.compile()
    n = 100_003
    w, _, p, _ = comp.launch(n, seed=99, first_event=0)
    cuts = [0, 1, 255, 256, 257, 40_000, 77_777, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        w2, _, p2, _ = comp.launch(hi - lo, seed=99, first_event=lo)
        assert torch.equal(w2, w[lo:hi])
        assert torch.equal(p2, p[:, lo:hi])


def test_chain_sharding_is_bit_identical():
    comp = D.to_gen(D.kstar_gamma_desc("relbw")).compile()
    n = 30_001
    w, _, p, _ = comp.launch(n, seed=5)
    w2, _, p2, _ = comp.launch(n - 12_345, seed=5, first_event=12_345)
    assert torch.equal(w2, w[12_345:]) and torch.equal(p2, p[:, 12_345:])


def test_user_callable_split_path():
    """A plain Python callable as mass function: evaluated on the device first, then fed per event."""
    from phasespace_b200 import GenParticle

    n = 5000

    def kstar_mass(min_mass, max_mass, n_events):
        assert min_mass.shape == (n_events,) and max_mass.shape == (n_events,)
        assert torch.allclose(min_mass, torch.full_like(min_mass, D.KAON_MASS + D.PION_MASS))
        assert torch.allclose(max_mass, torch.full_like(max_mass, D.B0_MASS))
        return torch.linspace(700.0, 1100.0, n_events, dtype=torch.float64, device=min_mass.device)

    gen = GenParticle("B0", D.B0_MASS).set_children(
        GenParticle("K*0", kstar_mass).set_children(GenParticle("K+", D.KAON_MASS), GenParticle("pi-", D.PION_MASS)),
        GenParticle("gamma", 0.0),
    )
    tree = D.to_oracle(D.kstar_gamma_desc("fixed", 0.0))
    tree.children[0].kind = orc.USER
    masses = np.linspace(700.0, 1100.0, n)
    draws = uniform_table(3, 0, n, orc.n_draws(tree))
    ref_w, ref_p = orc.generate(tree, n, draws, user_masses={"K*0": masses})
    w, p = gen.generate(n, debug_uniforms=torch.as_tensor(draws))
    np.testing.assert_allclose(_np(w), ref_w, rtol=TOL)
    for name in ref_p:
        assert np.max(np.abs(_np(p[name]) - ref_p[name])) / D.B0_MASS <= TOL
    m_kst = np.sqrt(np.maximum(_np(p["K*0"])[:, 3] ** 2 - (_np(p["K*0"])[:, :3] ** 2).sum(1), 0))
    np.testing.assert_allclose(m_kst, masses, rtol=1e-9)


def test_forbidden_decay_raises():
    import phasespace_b200 as ps

    with pytest.raises(ps.ForbiddenDecayError):
        ps.nbody_decay(300.0, [139.57018] * 3).generate(10)
    # per-event: a boost_to whose invariant mass is too small for the children
    boost = np.array([[0.0, 0.0, 100.0, 250.0]])
    with pytest.raises(ps.ForbiddenDecayError):
        ps.nbody_decay(D.B0_MASS, [139.57018] * 3).generate(10, boost_to=boost)


def _random_tree(rng, depth=0, name="top"):
    """A random decay tree description: fan-out 2-4, depth <= 3, at most one sampled resonance per sibling group
    (a second one could be squeezed out by the Cauchy tail of the first: SURVEY.md quirk 4), fixed-mass
    intermediates otherwise.  Returns (desc, sum of leaf masses)."""
    n_children = int(rng.integers(2, 5 if depth < 2 else 3))
    children, leaf_sum, sampled_used = [], 0.0, False
    for j in range(n_children):
        cname = f"{name}_{j}"
        if depth < 2 and rng.random() < 0.45:
            sub, sub_leaves = _random_tree(rng, depth + 1, cname)
            kind = "fixed" if sampled_used else str(rng.choice(["fixed", "gauss", "bw", "relbw"]))
            sampled_used |= kind != "fixed"
            mass = sub_leaves * float(rng.uniform(1.3, 2.0))
            width = 0.0 if kind == "fixed" else mass * float(rng.uniform(0.005, 0.08))
            children.append((cname, mass, kind, width, sub[4]))
            leaf_sum += mass if kind == "fixed" else sub_leaves
        else:
            mass = float(rng.choice([0.0, 0.511, 105.66, 139.57018, 493.677, 938.272]))
            children.append((cname, mass, "fixed", 0.0, []))
            leaf_sum += mass
    return (name, 0.0, "fixed", 0.0, children), leaf_sum


def _node_floor(desc):
    """Smallest mass the node can take: its own mass if fixed, else the sum of the floors below it."""
    return desc[1] if desc[2] == "fixed" and desc[1] > 0 else sum(_node_floor(c) for c in desc[4])


@pytest.mark.parametrize("tree_seed", [1, 2, 3, 4, 5])
def test_random_trees_jit(tree_seed):
    """Randomly shaped trees (none of them compiled ahead of time) through NVRTC, in-kernel Philox, against the oracle."""
    rng = np.random.default_rng(1000 + tree_seed)
    desc, _ = _random_tree(rng)
    # room for the widest resonance: the top mass well above the sum of the pole masses of its children
    top_mass = 3.0 * sum(c[1] if c[4] else _node_floor(c) for c in desc[4]) + 1000.0
    desc = (desc[0], top_mass, "fixed", 0.0, desc[4])
    # Tolerance: the reference (and the oracle) re-derive the mass of every decaying particle as sqrt(E^2 - p^2) of
    # its boosted 4-vector, which loses eps * gamma^2 per level, and far Cauchy tails land within
    # ~1e-4 of a threshold now and then: both are in each event's own bound (conditioning.py), for the weights and for
    # the 4-vectors of the particles below such an intermediate.
    _compare(desc, 4000, exact=False, use_rng=True)
    _compare(desc, 500, exact=True, use_rng=False)


from .helpers import reference_cases as R  # noqa: E402


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("case", list(R.CASES))
def test_debug_mode_matches_reference_source(case, exact):
    """BASELINE.json north_star: "a debug mode feeds the reference's own uniform draws and must match its fp64 outputs
    within 1e-12 relative".  The fp64 outputs are those of the reference's own source code (run on NumPy through a
    TensorFlow stand-in, tests/golden/make_reference_golden.py); the kernel is fed the same uniform table in the
    reference's draw order, the same boosts and the same callable masses."""
    ref = R.load(case)
    gen = D.to_gen(R.CASES[case], ref["user_masses"])
    n = ref["draws"].shape[0]
    boost = None if ref["boost"] is None else torch.as_tensor(ref["boost"])
    w, w_max, parts = gen.generate(n, boost_to=boost, normalize_weights=False, exact=exact,
                                   debug_uniforms=torch.as_tensor(ref["draws"]))
    assert list(parts) == ref["names"]
    # 1e-12 for every case, boosted or not; an event beyond it must be explained by its own amplification (a boosted
    # resonance re-derives its mass as sqrt(E^2 - p^2): eps * gamma^2, conditioning.py) -- flat decays have none
    ref_parts = {name: ref["particles"][k] for k, name in enumerate(ref["names"])}
    _check_against(R.CASES[case], (w, w_max, parts), (ref["weights"], ref["weights_max"], ref_parts), ref["boost"],
                   f"refsrc/{case}/{'exact' if exact else 'fast'}")


def _leaf(name, mass):
    return (name, mass, "fixed", 0.0, [])


_EDGE_TREES = {
    # massless products allow arbitrarily light subsystems: the boosts then amplify rounding by gamma^2 (conditioning.py)
    "three_massless": (D.flat_desc(3, top=100.0, masses=[0.0, 0.0, 0.0]), 2000, 1e-12),
    "gev_units": (("B", 5.27958, "fixed", 0.0, [
        ("K*", 0.89581, "relbw", 0.0474, [_leaf("K", 0.493677), _leaf("pi", 0.13957)]), _leaf("g", 0.0)]), 2000, 1e-12),
    "sixteen_body": (D.flat_desc(16, top=20000.0, masses=[139.57] * 16), 1000, 1e-12),
    "forty_body_max_slots": (D.flat_desc(40, top=60000.0, masses=[139.57] * 40), 200, 1e-12),
    "depth_five_chain": (("A", 10000.0, "fixed", 0.0, [
        ("B", 6000.0, "gauss", 100.0, [
            ("C", 3500.0, "bw", 50.0, [
                ("D", 2000.0, "relbw", 30.0, [
                    ("E", 900.0, "fixed", 0.0, [_leaf("e1", 139.57), _leaf("e2", 493.677)]), _leaf("d2", 139.57)]),
                _leaf("c2", 105.66)]),
            _leaf("b2", 0.0)]),
        _leaf("a2", 938.272)]), 2000, 1e-12),
    # the top-down path (nine or more children) below the top particle: its products are boosted by the parent afterwards
    "nine_body_subdecay": (("A", 12000.0, "fixed", 0.0, [
        ("X", 6000.0, "gauss", 100.0, [_leaf(f"x{i}", 139.57) for i in range(9)]), _leaf("g", 0.0), _leaf("h", 938.272)]),
        2000, 1e-12),
    "one_kev_above_threshold": (D.flat_desc(3, top=3 * D.PION_MASS + 1e-3, masses=[D.PION_MASS] * 3), 2000, 1e-12),
}


@pytest.mark.parametrize("name", list(_EDGE_TREES))
def test_edge_trees(name):
    """Shapes and scales off the beaten path (all but the last are NVRTC-compiled): massless products, GeV units, the
    largest tree the parameter block holds, a depth-five chain through all three samplers, a decay at threshold."""
    desc, n, tol = _EDGE_TREES[name]
    _compare(desc, n, exact=False, use_rng=True, tol=tol)


_RELBW_TABLE_CASES = {
    # (top mass, resonance mass, width, daughters, third particle): all have event-independent limits -> node table
    "kstar_like": (5279.58, 895.81, 47.4, (493.677, 139.57018), 0.0),
    "narrow_1e-5": (5279.58, 1019.461, 0.0102, (493.677, 493.677), 139.57018),
    "d0_physical_width": (2010.26, 1864.84, 1.6e-9 * 1864.84, (493.677, 139.57018), 139.57018),
    "upper_limit_100_widths_per_gev": (50000.0, 895.81, 47.4, (493.677, 139.57018), 0.0),
    "kappa_20": (5279.58, 775.26, 155.0, (139.57018, 139.57018), 139.57018),
    "window_cuts_the_peak": (1500.0, 895.81, 47.4, (493.677, 139.57018), 620.0),
}


@pytest.mark.parametrize("name", list(_RELBW_TABLE_CASES))
def test_relbw_node_table_against_the_oracle(name):
    """The tabulated relativistic Breit-Wigner (event-independent limits: ps_samplers.cuh sample_relbw_table) against the
    oracle's inverse CDF fed the same uniforms, through the whole generator: weights and 4-vectors within the per-event
    bounds, for narrow and broad resonances, far upper limits and a window that cuts into the peak."""
    import phasespace_b200 as ps
    from phasespace_b200 import _lib

    top, m, w, (d1, d2), third = _RELBW_TABLE_CASES[name]
    desc = ("X", top, "fixed", 0.0, [("R", m, "relbw", w, [_leaf("a", d1), _leaf("b", d2)]), _leaf("c", third)])
    assert _lib.lib.ps_tree_relbw_nodes(D.to_gen(desc).compile().handle) > 0
    _compare(desc, 20000, exact=False, use_rng=True, seed=41)
    _compare(desc, 5000, exact=True, use_rng=False, seed=42, normalize=False)


_GAUSS_TABLE_CASES = {
    # (top mass, resonance mass, width, daughters, third particle): event-independent limits -> node table
    "kstar_like": (5279.58, 895.81, 47.4, (493.677, 139.57018), 0.0),
    "window_cuts_the_peak": (1500.0, 895.81, 47.4, (493.677, 139.57018), 620.0),
    "lower_limit_above_the_pole": (5279.58, 895.81, 47.4, (493.677, 493.677), 0.0),
    "narrow": (5279.58, 1019.461, 4.2, (493.677, 493.677), 139.57018),
    "both_ends_thousands_of_widths_away": (2010.26, 1864.84, 0.05, (493.677, 139.57018), 139.57018),
    "broad": (3000.0, 2500.0, 300.0, (139.57018, 139.57018), 139.57018),
}


@pytest.mark.parametrize("name", list(_GAUSS_TABLE_CASES))
def test_gauss_node_table_against_the_oracle(name):
    """The tabulated truncated normal (event-independent limits: ps_samplers.cuh sample_gauss_table) against the oracle's
    inverse CDF (scipy ndtr / ndtri) fed the same uniforms, through the whole generator, for windows around, below and
    above the pole, narrow and broad widths.  Also with the draws pushed against both ends of [0, 1): the last
    dyadic levels of the table and, where an end of the mass range lies beyond the reach of fp64, the library path."""
    from phasespace_b200 import _lib

    top, m, w, (d1, d2), third = _GAUSS_TABLE_CASES[name]
    desc = ("X", top, "fixed", 0.0, [("R", m, "gauss", w, [_leaf("a", d1), _leaf("b", d2)]), _leaf("c", third)])
    assert _lib.lib.ps_tree_node_table_size(D.to_gen(desc).compile().handle) > 0
    _compare(desc, 20000, exact=False, use_rng=True, seed=41)
    _compare(desc, 5000, exact=True, use_rng=False, seed=42, normalize=False)
    # debug uniforms with the resonance's draw (column 0) pushed towards 0 and 1, on and off the generator's 51-bit grid.
    # An end of the mass range that the draws can actually reach is approached only to 2^-37: a resonance mass within
    # rounding of its limit makes the reference itself fail ("Forbidden decay": the mass it re-derives from the boosted
    # 4-vector dips below the daughters' sum).  An end beyond the reach of fp64 is approached to the last bit.
    singular_lo, singular_hi = name == "both_ends_thousands_of_widths_away", name in (
        "kstar_like", "lower_limit_above_the_pole", "narrow", "both_ends_thousands_of_widths_away")
    gen, tree = D.to_gen(desc), D.to_oracle(desc)
    n = 4096
    draws = uniform_table(43, 0, n, orc.n_draws(tree))
    k = np.arange(n)
    lower = k % 2 == 0
    depth = np.where(lower, 51 if singular_lo else 37, 51 if singular_hi else 37)
    tiny = np.ldexp(draws[:, 0] + 0.5, -((k // 2) % depth) - 1)
    draws[:, 0] = np.where(lower, tiny, 1.0 - tiny)
    if singular_hi:
        draws[:4, 0] = [1.0 - 2.0 ** -51, 1.0 - 2.0 ** -50, 0.5, 0.25]
    if singular_lo:
        draws[4:6, 0] = [2.0 ** -51, 2.0 ** -50]
    assert draws[:, 0].min() > 0.0 and draws[:, 0].max() < 1.0
    ref = orc.generate(tree, n, draws)
    got = gen.generate(n, debug_uniforms=torch.as_tensor(draws), exact=False)
    _check_against(desc, got, ref, None, f"gauss_table/{name}/ends")
