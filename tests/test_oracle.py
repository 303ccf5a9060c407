"""CPU tests of the oracle: pins it against the reference's own numbers where they exist
(README output rows, closed forms) and against the invariants the reference's tests assert."""

import json
import os

import numpy as np
import pytest
from scipy import stats

from oracle import genbod_numpy as orc
from oracle import genbod_ref
from oracle.philox import philox4x32_10, uniform_table

from .helpers import decays as D

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ Philox
def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def run(ctr, key):
        return [int(x) for x in philox4x32_10(ctr, key)]

    assert run((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert run((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_table_is_a_pure_function_of_seed_event_column():
    full = uniform_table(42, 0, 1000, 7)
    assert full.min() >= 0.0 and full.max() < 1.0
    np.testing.assert_array_equal(uniform_table(42, 300, 200, 7), full[300:500])
    np.testing.assert_array_equal(uniform_table(42, 0, 1000, 4), full[:, :4])
    assert not np.array_equal(uniform_table(43, 0, 1000, 7), full)
    big = uniform_table(42, (1 << 32) - 2, 4, 2)  # 64-bit event counter
    assert len(np.unique(big)) == 8
    assert abs(full.mean() - 0.5) < 0.02


# ------------------------------------------------------------------ reference numbers
def test_readme_two_body_energies():
    """README.rst:156-176: E(K*) = 2715.78804872, E(gamma) = 2563.79195128 (closed form (M^2+m1^2-m2^2)/2M)."""
    tree = orc.Part("B0", 5279.58, children=[orc.Part("K*", 895.81), orc.Part("gamma", 0.0)])
    w, parts = orc.generate(tree, 10, uniform_table(1, 0, 10, 2))
    np.testing.assert_allclose(parts["K*"][:, 3], 2715.78804872, atol=5e-9)
    np.testing.assert_allclose(parts["gamma"][:, 3], 2563.79195128, atol=5e-9)
    np.testing.assert_allclose(w, 1.0, rtol=1e-14)  # quirk 3: 2-body normalised weight is 1 up to ulps


def _invert_two_body(p_first, pd):
    """Draws (a, b) that make the reference produce first child momentum ``p_first`` (rest frame of the parent):
    particle 0 = (0, +pd, 0) rotated about z by acos(2a-1) then about y by 2 pi b (phasespace.py:414-481)."""
    px, py, pz = p_first
    cz = py / pd
    sz = np.sqrt(max(0.0, 1.0 - cz * cz))
    ang = np.arctan2(-pz / (sz * pd), -px / (sz * pd)) % (2 * np.pi)
    return (cz + 1.0) / 2.0, ang / (2 * np.pi)


def test_readme_output_rows_are_reproduced():
    """The six printed events of README.rst:156-183 are reference outputs.  Recover the four draws of each
    event from (K*, K+) and feed them to the oracle: it must give back all four printed 4-vectors."""
    with open(os.path.join(GOLDEN, "readme_kstargamma.json")) as fh:
        g = json.load(fh)
    m = g["masses"]
    tree = orc.Part("B0", m["B0"], children=[
        orc.Part("K*", m["K*"], children=[orc.Part("K+", m["K+"]), orc.Part("pi-", m["pi-"])]),
        orc.Part("gamma", m["gamma"])])
    kst, kp = np.array(g["K*"]), np.array(g["K+"])
    draws = np.zeros((len(kst), 4))
    pd_top = float(orc.pdk(m["B0"], m["K*"], m["gamma"]))
    pd_kst = float(orc.pdk(m["K*"], m["K+"], m["pi-"]))
    for i in range(len(kst)):
        draws[i, 0:2] = _invert_two_body(kst[i, :3], pd_top)
        # K+ in the K* rest frame: boost by -p/E of the K* (plain textbook boost, independent of the oracle)
        b = -kst[i, :3] / kst[i, 3]
        b2 = b @ b
        gamma = 1.0 / np.sqrt(1.0 - b2)
        bp = kp[i, :3] @ b
        p_rest = kp[i, :3] + ((gamma - 1.0) / b2 * bp + gamma * kp[i, 3]) * b
        draws[i, 2:4] = _invert_two_body(p_rest, pd_kst)
    assert np.all((draws >= 0) & (draws < 1))
    w, parts = orc.generate(tree, len(kst), draws)
    for name in ("K*", "gamma", "K+", "pi-"):
        np.testing.assert_allclose(parts[name], np.array(g[name]), atol=2e-7, rtol=0, err_msg=name)
    assert list(parts) == ["K*", "gamma", "K+", "pi-"]  # tests/test_chain.py:102 of the reference
    assert np.all(w < 1)


# ------------------------------------------------------------------ invariants
@pytest.mark.parametrize("n_body", [2, 3, 4, 6, 10])
def test_flat_invariants(n_body):
    n = 4000
    tree = orc.flat(D.B0_MASS, [D.PION_MASS] * n_body)
    assert orc.n_draws(tree) == 3 * n_body - 4
    w, wmax, parts = orc.generate(tree, n, uniform_table(n_body, 0, n, orc.n_draws(tree)), normalize_weights=False)
    total = sum(parts.values())
    np.testing.assert_allclose(total, np.broadcast_to([0, 0, 0, D.B0_MASS], (n, 4)), atol=1e-9)
    for vec in parts.values():
        np.testing.assert_allclose(orc.kin_mass(vec)[:, 0], D.PION_MASS, atol=1e-7)
    assert np.all(w > 0)
    if n_body > 2:
        assert np.all(w / wmax < 1)  # tests/test_generate.py:28 of the reference
    np.testing.assert_allclose(wmax, genbod_ref.wtmax(D.B0_MASS, [D.PION_MASS] * n_body), rtol=1e-14)
    assert list(parts) == [f"p_{i}" for i in range(n_body)]


@pytest.mark.parametrize("kind", ["gauss", "bw", "relbw"])
def test_chain_invariants(kind):
    n = 3000
    tree = D.to_oracle(D.k1_gamma_desc(kind, kind))
    assert list(orc.output_names(tree)) == ["K1+", "gamma", "K*0", "pi+", "K+", "pi-"]  # test_chain.py:113
    draws = uniform_table(5, 0, n, orc.n_draws(tree))
    w, parts = orc.generate(tree, n, draws)
    assert np.all(w < 1) and np.all(w > 0)
    np.testing.assert_allclose(parts["K1+"] + parts["gamma"], np.broadcast_to([0, 0, 0, D.B0_MASS], (n, 4)), atol=1e-8)
    np.testing.assert_allclose(parts["K*0"] + parts["pi+"], parts["K1+"], atol=1e-8)
    np.testing.assert_allclose(parts["K+"] + parts["pi-"], parts["K*0"], atol=1e-8)
    m_kst = orc.kin_mass(parts["K*0"])[:, 0]
    assert m_kst.min() >= D.KAON_MASS + D.PION_MASS - 1e-6
    assert np.all(m_kst + D.PION_MASS <= orc.kin_mass(parts["K1+"])[:, 0] + 1e-6)


def test_tree_w_max_formula():
    """recurse_w_max for K* gamma = pdk(M_B - m_gamma, m_K, m_pi) * pdk(M_B, m_K + m_pi, m_gamma)."""
    tree = D.to_oracle(D.kstar_gamma_desc("fixed", 0.0))
    _, wmax, _ = orc.generate(tree, 3, uniform_table(1, 0, 3, orc.n_draws(tree)), normalize_weights=False)
    expect = orc.pdk(D.B0_MASS - 0.0, D.KAON_MASS, D.PION_MASS) * orc.pdk(D.B0_MASS, D.KAON_MASS + D.PION_MASS, 0.0)
    np.testing.assert_allclose(wmax, expect, rtol=1e-13)


def test_forbidden_and_argument_errors():
    with pytest.raises(FloatingPointError, match="Forbidden decay"):
        orc.generate(orc.flat(300.0, [D.PION_MASS] * 3), 4, np.full((4, 5), 0.5))
    with pytest.raises(ValueError, match="resonance as top"):
        orc.generate(D.to_oracle(D.kstar_gamma_desc("gauss")).children[0], 4, np.full((4, 2), 0.5))
    with pytest.raises(ValueError, match="boost_to"):
        orc.generate(orc.flat(D.B0_MASS, [D.PION_MASS] * 3), 4, np.full((4, 5), 0.5), boost_to=np.ones((3, 4)))


def test_boost_to_moves_the_system():
    n = 500
    tree = orc.flat(D.B0_MASS, [D.PION_MASS] * 4)
    rng = np.random.default_rng(3)
    p = rng.normal(0, 30000.0, (n, 3))
    boost = np.concatenate([p, np.sqrt((p * p).sum(1, keepdims=True) + D.B0_MASS**2)], axis=1)
    w, parts = orc.generate(tree, n, uniform_table(9, 0, n, 8), boost_to=boost)
    np.testing.assert_allclose(sum(parts.values()), boost, rtol=1e-11, atol=1e-7)
    w0, _ = orc.generate(tree, n, uniform_table(9, 0, n, 8))
    np.testing.assert_allclose(w, w0, rtol=1e-6)  # weights are frame independent


def test_oracle_golden_vectors_are_stable():
    from tests.golden import make_golden

    stored = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    fresh = make_golden.build()
    assert set(stored.files) == set(fresh)
    for key in stored.files:
        np.testing.assert_allclose(fresh[key], stored[key], rtol=1e-13, atol=1e-9, err_msg=key)


# ------------------------------------------------------------------ statistics
def _physics_histos(weights, parts):
    """tests/test_physics.py:96-107 of the reference: weighted 100-bin histograms of every coordinate + the weights."""
    from .helpers.plotting import make_norm_histo

    histos = []
    for vec in parts:
        for coord in range(4):
            histos.append(make_norm_histo(vec[:, coord], range_=(-3000 if coord != 3 else 0, 3000), weights=weights))
    return histos, make_norm_histo(weights, range_=(0, 1 + 1e-8))


@pytest.mark.parametrize("n_body", [2, 3, 4])
def test_oracle_agrees_with_c_genbod(n_body):
    """The reference's KS procedure (ks_2samp on bin heights, p > 0.05) between the numpy oracle and the
    independent scalar C GENBOD that stands in for the missing data/bto{n}pi.root (1e5 events each)."""
    n = 100_000
    masses = [D.PION_MASS] * n_body
    ref_w, ref_p, _ = genbod_ref.generate(D.B0_MASS, masses, n, seed=4, n_threads=2)
    tree = orc.flat(D.B0_MASS, masses)
    w, parts = orc.generate(tree, n, uniform_table(11, 0, n, orc.n_draws(tree)))
    h, hw = _physics_histos(w, [parts[f"p_{i}"] for i in range(n_body)])
    rh, rhw = _physics_histos(ref_w, [ref_p[i] for i in range(n_body)])
    p_values = [stats.ks_2samp(a, b)[1] for a, b in zip(h, rh)] + [stats.ks_2samp(hw, rhw)[1]]
    assert min(p_values) > 0.05, p_values
    if n_body > 2:  # a direct two-sample test on an invariant mass and on the weights
        m01 = orc.kin_mass(parts["p_0"] + parts["p_1"])[:, 0]
        r01 = orc.kin_mass(ref_p[0] + ref_p[1])[:, 0]
        assert stats.ks_2samp(m01, r01)[1] > 1e-3
        assert stats.ks_2samp(w, ref_w)[1] > 1e-3


@pytest.mark.parametrize("kind,m,width,lo,hi", [
    ("gauss", 895.81, 47.4, 633.247, 5279.58), ("gauss", 895.81, 47.4, 950.0, 1000.0), ("gauss", 100.0, 5.0, 140.0, 150.0),
    ("bw", 895.81, 47.4, 633.247, 5279.58), ("relbw", 895.81, 47.4, 633.247, 5279.58), ("relbw", 134.9768, 0.5, 0.0, 1200.0),
])
def test_mass_samplers_follow_their_densities(kind, m, width, lo, hi):
    n = 40_000
    u = uniform_table(77, 0, n, 1)[:, 0]
    x = orc.sample_mass(orc.KIND_NAMES[kind], m, width, np.full(n, lo), np.full(n, hi), u)
    assert x.min() >= lo and x.max() <= hi
    if kind == "gauss":
        a, b = (lo - m) / width, (hi - m) / width
        cdf = stats.truncnorm(a, b, loc=m, scale=width).cdf
    elif kind == "bw":
        c = stats.cauchy(loc=m, scale=width)
        cdf = lambda v: (c.cdf(v) - c.cdf(lo)) / (c.cdf(hi) - c.cdf(lo))  # noqa: E731
    else:
        grid = np.linspace(lo, hi, 400_001)
        dens = 1.0 / ((grid**2 - m * m) ** 2 + (m * width) ** 2)
        cum = np.concatenate([[0.0], np.cumsum((dens[1:] + dens[:-1]) / 2)])
        cum /= cum[-1]
        cdf = lambda v: np.interp(v, grid, cum)  # noqa: E731
    assert stats.kstest(x, cdf)[1] > 1e-3


def test_fonll_oracle_is_numpy_choice():
    """oracle.generate_fonll restates tests/helpers/rapidsim.py:23-51: fed the uniforms numpy's RandomState consumes,
    it reproduces np.random.choice / np.random.uniform exactly."""
    pt_edges = np.linspace(0.0, 40.0, 81)
    pt_centers = pt_edges[:-1] + 0.25
    pt_values = np.exp(-pt_centers / 5.0) * pt_centers
    eta_centers = np.linspace(1.9, 5.1, 33)
    eta_values = np.exp(-0.5 * ((eta_centers - 3.2) / 0.9) ** 2)
    n, mass = 5000, 5279.58
    rs = np.random.RandomState(7)
    pt_rand = 1000 * np.abs(rs.choice(pt_centers, size=n, p=pt_values / np.sum(pt_values)))
    eta_rand = rs.choice(eta_centers, size=n, p=eta_values / np.sum(eta_values))
    phi_rand = rs.uniform(0, 2 * np.pi, size=n)
    px, py, pz = pt_rand * np.cos(phi_rand), pt_rand * np.sin(phi_rand), pt_rand * np.sinh(eta_rand)
    ref = np.stack([px, py, pz, np.sqrt(px * px + py * py + pz * pz + mass * mass)])
    rs = np.random.RandomState(7)
    draws = np.stack([rs.random_sample(n), rs.random_sample(n), rs.random_sample(n)], axis=1)
    got = orc.generate_fonll(mass, pt_centers, pt_values, eta_centers, eta_values, draws)
    np.testing.assert_array_equal(got, ref.transpose())


NINE_CHILDREN_CASES = ("flat9", "nine_body_subdecay", "nine_body_subdecay_boost")


@pytest.mark.parametrize("case", list(__import__("tests.helpers.reference_cases", fromlist=["CASES"]).CASES))
def test_oracle_matches_reference_source(case, monkeypatch):
    """The oracle against the reference's OWN source code run on NumPy (tests/golden/make_reference_golden.py): same
    uniform draws in the reference's draw order, same boosts, same callable masses.  The two differ at most in the
    association of a few sums (numpy's pairwise reduction inside the reference run), hence 1e-13 and not bit equality."""
    from tests.helpers import reference_cases as R

    ref = R.load(case)
    tree = D.to_oracle(R.CASES[case])
    n = ref["draws"].shape[0]
    assert orc.n_draws(tree) == ref["draws"].shape[1]
    scale = D.B0_MASS if ref["boost"] is None else float(np.max(ref["boost"][:, 3]))

    def run(order):
        monkeypatch.setattr(orc, "ROW_SUM_ORDER", order)
        return orc.generate(tree, n, ref["draws"], boost_to=ref["boost"], normalize_weights=False,
                            user_masses=ref["user_masses"] or None)

    # the reference's expression tnp.sum(masses, axis=1) as NumPy evaluates it (pairwise from eight children on): every
    # case, the nine-body ones included, comes out as the reference's source computed it
    w, w_max, parts = run("numpy")
    assert list(parts) == ref["names"]  # the reference's dict order (phasespace.py:547-570)
    np.testing.assert_allclose(w, ref["weights"], rtol=1e-13)
    np.testing.assert_allclose(w_max, ref["weights_max"], rtol=1e-13)
    for k, name in enumerate(ref["names"]):
        assert np.max(np.abs(parts[name] - ref["particles"][k])) <= 1e-13 * scale, name
    # the default association (left to right: SURVEY.md appendix A, and what the CUDA kernel does) differs from it only
    # where a node has nine children, by the last bit of the mass sum: per event no more than that bit amplified by the
    # event's own conditioning (tests/helpers/conditioning.py), the bound the GPU parity tests use
    w, w_max, parts = run("left")
    if case in NINE_CHILDREN_CASES:
        from tests.helpers import conditioning

        ref_parts = {name: ref["particles"][k] for k, name in enumerate(ref["names"])}
        amp, w_bound, v_bound = conditioning.event_bounds(R.CASES[case], ref_parts, ref["boost"], tol=1e-13, c_amp=8.0)
        w_err = np.abs(w - ref["weights"]) / ref["weights"]
        assert np.all(w_err <= w_bound), float(np.max(w_err / w_bound))
        assert 1e-13 < np.max(w_err) < 1e-11  # the association does show, and stays near the north star's 1e-12
        v_err = np.max([np.max(np.abs(parts[name] - ref["particles"][k]), axis=1) for k, name in enumerate(ref["names"])], axis=0)
        v_scale = scale if ref["boost"] is None else ref["boost"][:, 3]
        assert np.all(v_err / v_scale <= v_bound)
        np.testing.assert_allclose(w_max, ref["weights_max"], rtol=1e-13)
    else:
        np.testing.assert_allclose(w, ref["weights"], rtol=1e-13)
        np.testing.assert_allclose(w_max, ref["weights_max"], rtol=1e-13)
        for k, name in enumerate(ref["names"]):
            assert np.max(np.abs(parts[name] - ref["particles"][k])) <= 1e-13 * scale, name


@pytest.mark.parametrize("width", [1e-3, 1e-6, 1.6e-9])
def test_relbw_narrow_resonance(width):
    """width / mass below 1e-8 (a D0 with its physical width in a DecayLanguage chain): the quartic's roots p + i q must
    come from 2 p q = m * width -- sqrt((rho - m^2) / 2) cancels to zero.  A narrow relativistic Breit-Wigner is a
    Lorentzian of half width ``width / 2`` around the pole."""
    n, m = 20001, 1864.84
    u = (np.arange(n) + 0.5) / n
    x = orc.sample_mass(orc.RELBW, m, width, np.full(n, 633.0), np.full(n, 5279.0), u)
    assert np.all(np.diff(x) >= 0) and np.all(np.isfinite(x))
    q25, q50, q75 = np.quantile(x, [0.25, 0.5, 0.75])
    assert abs(q50 - m) < 1e-3 * width + 1e-12 * m
    assert (q75 - q25) / width == pytest.approx(1.0, rel=2e-2)
