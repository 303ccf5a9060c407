"""Ports of the reference's generating tests (tests/test_generate.py, tests/test_chain.py:95-114) plus the
argument semantics of SURVEY.md appendix B, on the GPU."""

import os

import numpy as np
import pytest
import torch

import phasespace_b200 as phasespace
from phasespace_b200 import GenParticle

from .helpers import decays as D

pytestmark = pytest.mark.gpu
B0_MASS, PION_MASS = D.B0_MASS, D.PION_MASS


def _np(t):
    return t.detach().cpu().numpy()


def test_one_event():
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    norm_weights, particles = decay.generate(n_events=1)
    assert norm_weights.shape[0] == 1
    assert np.all(_np(norm_weights) < 1)
    assert len(particles) == 3
    assert all(part.shape == (1, 4) for part in particles.values())


@pytest.mark.parametrize("n_events", [5, 523])
def test_n_events(n_events):
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    norm_weights, particles = decay.generate(n_events=n_events)
    assert norm_weights.shape[0] == n_events and norm_weights.dtype == torch.float64 and norm_weights.is_cuda
    assert np.all(_np(norm_weights) < 1)
    assert len(particles) == 3
    assert all(part.shape == (n_events, 4) for part in particles.values())
    assert len(norm_weights) == n_events
    # n_events as a tensor, like the reference's tf.Variable use (tests/test_physics.py:86-95)
    w2, _ = decay.generate(torch.tensor(n_events))
    assert len(w2) == n_events


@pytest.fixture
def vector_module(monkeypatch):
    """The real ``vector`` package when it is installed, else the test stand-in (tests/helpers/vector_standin.py),
    so that the product's conversion code (ps.py:676-695, 817-835) always runs."""
    try:
        import vector
    except ImportError:
        import sys

        from .helpers import vector_standin as vector

        monkeypatch.setitem(sys.modules, "vector", vector)
    return vector


@pytest.mark.parametrize("n_events", [1, 5, 523])
def test_as_vectors(vector_module, n_events):
    """tests/test_generate.py:34-63 of the reference, as_vectors=True."""
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    w, particles = decay.generate(n_events=n_events, as_vectors=True, seed=4)
    assert all(isinstance(p, vector_module.Momentum) and isinstance(p, vector_module.Lorentz) for p in particles.values())
    stacked = {k: np.stack([p.px, p.py, p.pz, p.E], axis=-1) for k, p in particles.items()}
    assert len(stacked) == 3 and all(v.shape == (n_events, 4) for v in stacked.values())
    _, tensors = decay.generate(n_events=n_events, seed=4)
    assert all(np.array_equal(stacked[k], _np(tensors[k])) for k in tensors)
    again = phasespace.to_vectors(tensors)
    assert all(np.array_equal(again[k].energy, stacked[k][:, 3]) for k in tensors)


def test_boost_to_as_vector_object(vector_module):
    """ps.py:676-695 (tests/test_physics.py:192-199 of the reference): a ``vector`` momentum Lorentz array as boost_to
    gives the same events as the equivalent (n, 4) array; anything else that is a ``vector`` object is refused."""
    n = 257
    rng = np.random.default_rng(8)
    p = rng.normal(0.0, 20000.0, (n, 3))
    e = np.sqrt((p * p).sum(1) + B0_MASS**2)
    booster = vector_module.array(dict(px=p[:, 0], py=p[:, 1], pz=p[:, 2], e=e))  # the spelling of tests/test_physics.py:193-199
    decay = D.to_gen(D.kstar_gamma_desc("relbw"))
    w_v, parts_v = decay.generate(n, boost_to=booster, seed=6)
    w_a, parts_a = decay.generate(n, boost_to=np.concatenate([p, e[:, None]], axis=1), seed=6)
    assert torch.equal(w_v, w_a) and all(torch.equal(parts_v[k], parts_a[k]) for k in parts_a)
    total = _np(parts_v["K*0"] + parts_v["gamma"])
    assert np.max(np.abs(total - np.concatenate([p, e[:, None]], axis=1))) / e.max() < 1e-12
    with pytest.raises(ValueError, match="momentum Lorentz vector"):
        decay.generate(n, boost_to=vector_module.array({"x": p[:, 0], "y": p[:, 1], "z": p[:, 2], "t": e}))
    with pytest.raises(ValueError, match="momentum Lorentz vector"):
        decay.generate(n, boost_to=vector_module.array({"px": p[:, 0], "py": p[:, 1], "pz": p[:, 2]}))
    with pytest.raises(ValueError, match="doesn't match the boost_to input size"):
        decay.generate(n + 1, boost_to=booster)


def test_deterministic_events():
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    common_seed = 36
    w1, p1 = decay.generate(n_events=100, seed=common_seed)
    wg, pg = decay.generate(n_events=100)
    wr, pr = decay.generate(n_events=100, seed=152)
    w2, p2 = decay.generate(n_events=100, seed=common_seed)
    np.testing.assert_allclose(_np(w1), _np(w2))
    for a, b in zip(p1.values(), p2.values()):
        np.testing.assert_allclose(_np(a), _np(b))
    assert not np.allclose(_np(w1), _np(wr))
    for a, b in zip(p1.values(), pr.values()):
        assert not np.allclose(_np(a), _np(b))
    assert not np.allclose(_np(wg), _np(wr))
    for a, b in zip(pg.values(), pr.values()):
        assert not np.allclose(_np(a), _np(b))
    # the global generator advances between calls (docs/usage.rst:203-213)
    wg2, _ = decay.generate(n_events=100)
    assert not np.allclose(_np(wg), _np(wg2))
    # a Generator object is used and advanced
    rng = phasespace.random.Generator(7)
    a, _ = decay.generate(50, seed=rng)
    b, _ = decay.generate(50, seed=rng)
    both, _ = decay.generate(100, seed=7)
    assert torch.equal(torch.cat([a, b]), both)


def test_kstargamma():
    decay = D.to_gen(D.kstar_gamma_desc("gauss"))
    norm_weights, particles = decay.generate(n_events=1000)
    assert norm_weights.shape[0] == 1000
    assert np.all(_np(norm_weights) < 1)
    assert len(particles) == 4
    assert list(particles.keys()) == ["K*0", "gamma", "K+", "pi-"]
    assert all(part.shape == (1000, 4) for part in particles.values())


def test_k1gamma():
    decay = D.to_gen(D.k1_gamma_desc())
    norm_weights, particles = decay.generate(n_events=1000)
    assert norm_weights.shape[0] == 1000
    assert np.all(_np(norm_weights) < 1)
    assert set(particles.keys()) == {"K1+", "K*0", "gamma", "K+", "pi-", "pi+"}
    assert all(part.shape == (1000, 4) for part in particles.values())


def test_set_children_after_generate():
    top = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 2)
    top.generate(3)
    with pytest.raises(RuntimeError):
        top.set_children(GenParticle("a", 1.0), GenParticle("b", 1.0))


def test_latch_is_set_when_the_tree_was_compiled_beforehand():
    """ps.py:318: the latch is set by the first generate() -- also when compile() was called by hand before it (the
    plain single-call path is taken from the second call on)."""
    top = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    top.compile()
    assert not top._generate_called
    first = top.generate(5, seed=3)
    assert top._generate_called
    again = top.generate(5, seed=3)  # the plain path
    assert torch.equal(first[0], again[0]) and all(torch.equal(first[1][k], again[1][k]) for k in first[1])
    with pytest.raises(RuntimeError):
        top.set_children(GenParticle("a", 1.0), GenParticle("b", 1.0))


def test_generate_host_takes_vector_boosts(vector_module):
    """generate_host converts boost_to like generate (ps.py:676-695)."""
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    n = 64
    rng = np.random.default_rng(5)
    p = rng.normal(size=(n, 3)) * 3000.0
    e = np.sqrt((p * p).sum(axis=1) + B0_MASS**2)
    booster = vector_module.array(dict(px=p[:, 0], py=p[:, 1], pz=p[:, 2], e=e))
    w_v, parts_v = decay.generate_host(n, boost_to=booster, seed=9)
    w_a, parts_a = decay.generate_host(n, boost_to=np.column_stack([p, e]), seed=9)
    assert np.array_equal(w_v, w_a) and all(np.array_equal(parts_v[k], parts_a[k]) for k in parts_a)
    with pytest.raises(ValueError, match="Bad shape"):
        decay.generate_host(n, boost_to=np.zeros((n, 3)))


def test_boost_to_semantics():
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    n = 64
    row = np.array([1000.0, -2000.0, 30000.0, np.sqrt(1000.0**2 + 2000.0**2 + 30000.0**2 + B0_MASS**2)])
    boost = np.tile(row, (n, 1))
    w_arr, p_arr = decay.generate(n, boost_to=boost, seed=3)
    w_one, p_one = decay.generate(n, boost_to=row[None, :], seed=3)
    w_list, p_list = decay.generate(n, boost_to=[list(c) for c in boost.T], seed=3)  # a list is read transposed (ps.py:51-53)
    w_t, p_t = decay.generate(n, boost_to=torch.as_tensor(boost, device="cuda"), seed=3)
    for other in (p_one, p_list, p_t):
        for name in p_arr:
            assert torch.equal(p_arr[name], other[name])
    total = sum(_np(v) for v in p_arr.values())
    np.testing.assert_allclose(total, boost, rtol=1e-12)
    with pytest.raises(ValueError, match="doesn't match the boost_to input size"):
        decay.generate(n, boost_to=boost[:5])
    with pytest.raises(ValueError):
        decay.generate(n, boost_to=np.ones((n, 3)))


def test_normalize_false_returns_three():
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 4)
    w, wmax, parts = decay.generate(100, normalize_weights=False, seed=1)
    wn, _ = decay.generate(100, seed=1)
    np.testing.assert_allclose(_np(w) / _np(wmax), _np(wn), rtol=1e-14)
    assert torch.all(wmax == wmax[0])


def test_empty_and_odd_sizes():
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    w, p = decay.generate(0)
    assert w.shape == (0,) and all(v.shape == (0, 4) for v in p.values())
    for n in (1, 31, 33, 255, 257, 100_001):
        w, p = decay.generate(n, seed=4)
        assert w.shape == (n,) and torch.isfinite(w).all()
        assert all(torch.isfinite(v).all() for v in p.values())


def test_golden_vectors():
    """The committed oracle vectors (tests/golden/oracle_vectors.npz) through the CUDA path."""
    from .golden import make_golden

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))
    for name, (desc, n, boost_seed) in make_golden.CASES.items():
        gen = D.to_gen(desc)
        boost = torch.as_tensor(g[f"{name}/boost"]) if boost_seed is not None else None
        for exact in (True, False):
            w, wmax, parts = gen.generate(n, boost_to=boost, normalize_weights=False, exact=exact,
                                          debug_uniforms=torch.as_tensor(g[f"{name}/draws"]))
            scale = D.B0_MASS if boost is None else float(boost[:, 3].max())
            ref_w = g[f"{name}/weights"]
            assert np.max(np.abs(_np(w) - ref_w)) <= 1e-11 * np.max(ref_w), name
            np.testing.assert_allclose(_np(wmax), g[f"{name}/weights_max"], rtol=1e-11, err_msg=name)
            got = np.stack([_np(v) for v in parts.values()])
            assert np.max(np.abs(got - g[f"{name}/particles"])) / scale <= 1e-11, name


def test_jit_compiled_tree():
    """A tree with no ahead-of-time kernel (11 bodies; a 3 -> (2, 3) chain) is specialised by NVRTC."""
    from oracle import genbod_numpy as orc
    from oracle.philox import uniform_table

    n = 2000
    masses = [PION_MASS] * 11
    desc = D.flat_desc(11, masses=masses)
    tree = D.to_oracle(desc)
    ref_w, ref_p = orc.generate(tree, n, uniform_table(8, 0, n, orc.n_draws(tree)))
    for exact in (False, True):  # the exact variant is JIT-compiled with --fmad=false
        w, p = D.to_gen(desc).generate(n, seed=8, exact=exact)
        np.testing.assert_allclose(_np(w), ref_w, rtol=1e-11)
        for name in ref_p:
            assert np.max(np.abs(_np(p[name]) - ref_p[name])) / B0_MASS <= 1e-12
    # the weights-only mask kernel of the unweighting path is JIT-compiled for this tree as well
    from phasespace_b200.unweight import generate_unweighted

    w_fast, _ = D.to_gen(desc).generate(n, seed=8)
    parts, idx, w_acc = generate_unweighted(D.to_gen(desc), n, 8, w_bound=float(ref_w.max()) * 1.01)
    assert 0 < len(idx) < n and torch.equal(w_acc, w_fast[idx])
    chain = ("B", 5279.58, "fixed", 0.0, [
        ("X", 1500.0, "relbw", 100.0, [("a", 139.57, "fixed", 0.0, []), ("b", 493.677, "fixed", 0.0, [])]),
        ("Y", 2000.0, "gauss", 50.0, [("c", 139.57, "fixed", 0.0, []), ("d", 139.57, "fixed", 0.0, []), ("e", 0.0, "fixed", 0.0, [])]),
        ("f", 105.66, "fixed", 0.0, []),
    ])
    tree = D.to_oracle(chain)
    ref_w, ref_p = orc.generate(tree, n, uniform_table(9, 0, n, orc.n_draws(tree)))
    w, p = D.to_gen(chain).generate(n, seed=9)
    assert list(p) == orc.output_names(tree)
    assert np.max(np.abs(_np(w) - ref_w)) <= 1e-11 * ref_w.max()
    for name in ref_p:
        assert np.max(np.abs(_np(p[name]) - ref_p[name])) / B0_MASS <= 5e-12, name


def test_generate_host_matches_device():
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    n = 300_007
    w, p = decay.generate(n, seed=21)
    hw, hp = decay.generate_host(n, seed=21, chunk_events=65536)
    assert np.array_equal(_np(w), hw)
    for name in p:
        assert np.array_equal(_np(p[name]), hp[name])
    chain = D.to_gen(D.kstar_gamma_desc("relbw"))
    w, wmax, p = chain.generate(50_000, seed=2, normalize_weights=False)
    hw, hwmax, hp = chain.generate_host(50_000, seed=2, normalize_weights=False, chunk_events=8192)
    assert np.array_equal(_np(w), hw) and np.array_equal(_np(wmax), hwmax)
    assert all(np.array_equal(_np(p[k]), hp[k]) for k in p)
    # host-side boost_to (the H2D leg): per-event rows across chunk boundaries, and one broadcast row
    four = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 4)
    n = 70_001
    rng = np.random.default_rng(4)
    mom = rng.normal(0.0, 8000.0, (n, 3))
    boost = np.concatenate([mom, np.sqrt((mom * mom).sum(1, keepdims=True) + B0_MASS**2)], axis=1)
    for b in (boost, boost[:1]):
        w, p = four.generate(n, boost_to=torch.as_tensor(b), seed=6)
        hw, hp = four.generate_host(n, boost_to=b, seed=6, chunk_events=16384)
        assert np.array_equal(_np(w), hw) and all(np.array_equal(_np(p[k]), hp[k]) for k in p)
    with pytest.raises(ValueError):
        four.generate_host(n, boost_to=boost[:5], seed=6)
    # the staging arena grows when a later call asks for larger chunks (and for more particles per event)
    ten = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 10)
    w, p = ten.generate(300_000, seed=7)
    hw, hp = ten.generate_host(300_000, seed=7, chunk_events=1 << 18)
    assert np.array_equal(_np(w), hw) and all(np.array_equal(_np(p[k]), hp[k]) for k in p)


def test_kinematics_helpers():
    from phasespace_b200 import kinematics as kin

    rng = np.random.default_rng(1)
    p = rng.normal(0, 1000.0, (1000, 3))
    vec = np.concatenate([p, np.sqrt((p * p).sum(1, keepdims=True) + 139.57**2)], axis=1)
    np.testing.assert_allclose(_np(kin.mass(vec))[:, 0], 139.57, rtol=1e-8)
    b = rng.uniform(-0.5, 0.5, (1000, 3))
    out = _np(kin.lorentz_boost(vec, b))
    from oracle import genbod_numpy as orc

    np.testing.assert_allclose(out, orc.lorentz_boost(vec, b), rtol=1e-13, atol=1e-10)
    same = _np(kin.lorentz_boost(vec, np.zeros((1, 3))))
    np.testing.assert_array_equal(same, vec)


def test_launch_is_cuda_graph_capturable():
    """The launch is a plain asynchronous kernel launch on the current stream, so a loop of generations can be
    captured once in a CUDA graph and replayed (streams and graphs instead of a tracing compiler)."""
    comp = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3).compile()
    n = 4096
    dev = torch.device("cuda")
    outs = [{"weights": torch.empty((n,), dtype=torch.float64, device=dev),
             "particles": torch.empty((3, n, 4), dtype=torch.float64, device=dev)} for _ in range(3)]
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    comp.launch(n, 5, out=outs[0], err=err)  # warm-up outside the capture (occupancy query, module load)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for k, out in enumerate(outs):
            comp.launch(n, 5, first_event=k * n, out=out, err=err)
    for out in outs:
        out["weights"].zero_()
    graph.replay()
    torch.cuda.synchronize()
    w, _, p, _ = comp.launch(3 * n, 5)
    assert torch.equal(torch.cat([o["weights"] for o in outs]), w)
    assert torch.equal(torch.cat([o["particles"] for o in outs], dim=1), p)


def test_fonll_boost_generator():
    """generate_fonll (the reference tests' boost_to source, tests/helpers/rapidsim.py:23-51) on the device: equal to
    the numpy restatement fed the same Philox draws, usable as boost_to, independent of how the range is split."""
    from oracle import genbod_numpy as orc
    from oracle.philox import uniform_table
    from phasespace_b200.boost import generate_fonll
    from phasespace_b200.random import Generator

    pt_edges = np.linspace(0.0, 40.0, 81)
    pt_centers = pt_edges[:-1] + 0.25
    pt_values = np.exp(-pt_centers / 5.0) * pt_centers
    eta_centers = np.linspace(1.9, 5.1, 33)
    eta_values = np.exp(-0.5 * ((eta_centers - 3.2) / 0.9) ** 2)
    n, seed = 100_003, 17
    got = generate_fonll(B0_MASS, (pt_edges, pt_values), (eta_centers, eta_values), n, seed=seed)
    assert tuple(got.shape) == (n, 4) and got.is_cuda
    ref = orc.generate_fonll(B0_MASS, pt_centers, pt_values, eta_centers, eta_values, uniform_table(seed, 0, n, 3, stream=4))
    scale = ref[:, 3].max()
    assert np.max(np.abs(_np(got) - ref)) / scale < 1e-14
    # the masses of the rows are the requested one; the spectrum is the requested one
    m = np.sqrt(ref[:, 3] ** 2 - (ref[:, :3] ** 2).sum(1))
    np.testing.assert_allclose(m, B0_MASS, rtol=1e-9)
    pt = np.hypot(_np(got)[:, 0], _np(got)[:, 1]) / 1000.0
    counts = np.histogram(pt, bins=pt_edges)[0] / n
    assert np.max(np.abs(counts - pt_values / pt_values.sum())) < 5e-3
    # one generator, two consecutive calls == one call (draws are keyed on the event index)
    g = Generator(seed)
    a = generate_fonll(B0_MASS, (pt_edges, pt_values), (eta_centers, eta_values), 60_000, seed=g)
    b = generate_fonll(B0_MASS, (pt_edges, pt_values), (eta_centers, eta_values), n - 60_000, seed=g)
    assert torch.equal(torch.cat([a, b]), got)
    # as boost_to
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    w, parts = decay.generate(n, boost_to=got, seed=3)
    total = sum(parts.values())
    assert float((total - got).abs().max()) / scale < 1e-12
    with pytest.raises(ValueError):
        generate_fonll(B0_MASS, (pt_edges, pt_values[:-1]), (eta_centers, eta_values), 10)


@pytest.mark.parametrize("n", [1, 31, 255, 257, 100_003])
def test_no_writes_outside_the_output_buffers(n):
    """Canaries around every output buffer (compute-sanitizer is not available on the GPU pool): odd sizes, 3-body,
    chunked 2-body grid and a chain; nothing outside [0, n) may be touched and everything inside must be written."""
    pad, sentinel = 64, -12345.678
    dev = torch.device("cuda")
    for decay in (phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3), phasespace.nbody_decay(B0_MASS, [PION_MASS, D.KAON_MASS]),
                  D.to_gen(D.kstar_gamma_desc("relbw"))):
        comp = decay.compile()
        P = comp.n_particles
        wbuf = torch.full((n + 2 * pad,), sentinel, dtype=torch.float64, device=dev)
        mbuf = torch.full((n + 2 * pad,), sentinel, dtype=torch.float64, device=dev)
        pbuf = torch.full((P, n + 2 * pad, 4), sentinel, dtype=torch.float64, device=dev)
        out = {"weights": wbuf[pad:pad + n], "weights_max": mbuf[pad:pad + n], "particles": pbuf[:, pad:pad + n, :]}
        comp.launch(n, seed=3, normalize_weights=False, out=out)
        torch.cuda.synchronize()
        for buf in (wbuf, mbuf):
            assert bool((buf[:pad] == sentinel).all()) and bool((buf[pad + n:] == sentinel).all())
            assert not bool((buf[pad:pad + n] == sentinel).any())
        assert bool((pbuf[:, :pad] == sentinel).all()) and bool((pbuf[:, pad + n:] == sentinel).all())
        assert not bool((pbuf[:, pad:pad + n] == sentinel).any())


def test_jit_disk_cache(tmp_path):
    """NVRTC results are cached on disk (PS_B200_CACHE_DIR): a second process loads the cubin instead of compiling,
    with identical events; an empty PS_B200_CACHE_DIR disables the cache."""
    import subprocess
    import sys
    import time

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    script = ("import sys, time; sys.path.insert(0, %r)\n"
              "import torch, phasespace_b200 as ps\n"
              "ps.nbody_decay(5279.58, [139.57018] * 3).generate(10); torch.cuda.synchronize()\n"  # CUDA context, library
              "d = ps.nbody_decay(5279.58, [139.57018] * 12)\n"  # 12 bodies: never compiled ahead of time
              "t0 = time.perf_counter(); w, p = d.generate(1000, seed=4); torch.cuda.synchronize()\n"
              "print(time.perf_counter() - t0, float(w.sum()), float(p['p_11'].sum()))\n") % root

    def run(cache_dir):
        env = dict(os.environ, PS_B200_CACHE_DIR=cache_dir)
        out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        t, sw, sp = out.stdout.split()[-3:]
        return float(t), sw, sp

    cache = str(tmp_path / "jit")
    os.makedirs(cache)
    t_cold, sw0, sp0 = run(cache)
    files = [f for f in os.listdir(cache) if f.endswith(".cubin")]
    assert len(files) == 1, files
    t_warm, sw1, sp1 = run(cache)
    assert (sw0, sp0) == (sw1, sp1)
    assert t_warm < t_cold, (t_cold, t_warm)
    assert [f for f in os.listdir(cache) if f.endswith(".cubin")] == files
    _, sw2, sp2 = run("")  # disabled
    assert (sw2, sp2) == (sw0, sp0) and sorted(os.listdir(cache)) == sorted(files)


@pytest.mark.parametrize("n", [1, 1000, 300_001])
def test_kernel_epilogue_statistics(n):
    """The generator's own epilogue [max w, sum w, sum w^2, max w_max] (what the multi-GPU all-reduce combines)."""
    dev = torch.device("cuda")
    for decay, boost in ((phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3), None),
                         (D.to_gen(D.kstar_gamma_desc("gauss")), None),
                         (phasespace.nbody_decay(B0_MASS, [PION_MASS] * 4), True)):
        comp = decay.compile()
        b = None
        if boost:
            p = torch.randn((n, 3), dtype=torch.float64, device=dev) * 3000.0
            b = torch.cat([p, torch.sqrt((p * p).sum(1, keepdim=True) + B0_MASS**2)], dim=1).contiguous()
        stats = torch.zeros((4,), dtype=torch.float64, device=dev)
        w, w_max, _, _ = comp.launch(n, seed=5, boost_to=b, normalize_weights=False, stats=stats)
        s = stats.cpu().numpy()
        assert s[0] == float(w.max())
        np.testing.assert_allclose(s[1], float(w.sum()), rtol=1e-12)
        np.testing.assert_allclose(s[2], float((w * w).sum()), rtol=1e-12)
        if b is None:
            assert s[3] == float(w_max.max())  # event-independent maximum weight: reported exactly
        else:
            np.testing.assert_allclose(s[3], float(w_max.max()), rtol=1e-14)


def test_threaded_generation():
    """Four host threads drive the library at once (registry, JIT, error state and the host arena are shared): every
    thread gets exactly the events a single thread gets."""
    import threading

    def work(k):
        decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * (3 + k))  # 3..6 bodies
        chain = D.to_gen(D.kstar_gamma_desc(["fixed", "gauss", "bw", "relbw"][k]))
        out = []
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for rep in range(5):
                w, p = decay.generate(20_000, seed=100 + k)
                cw, cp = chain.generate(20_000, seed=200 + k)
                hw, hp = decay.generate_host(20_000, seed=100 + k, chunk_events=4096)
                out.append((w.cpu(), p["p_0"].cpu(), cw.cpu(), cp["K+"].cpu(), torch.as_tensor(hw.copy())))
        stream.synchronize()
        return out

    expected = [work(k) for k in range(4)]
    results = [None] * 4
    threads = [threading.Thread(target=lambda k=k: results.__setitem__(k, work(k))) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in range(4):
        assert results[k] is not None
        for got, want in zip(results[k], expected[k]):
            assert all(torch.equal(a, b) for a, b in zip(got, want))
            assert torch.equal(got[0], got[4])  # generate_host delivers the same weights


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices():
    """One process drives two GPUs (threading): the same tree, AOT and JIT kernels, the host arena and the
    histogram consumer on either device give identical events; two halves of an index range, one per device, concatenate
    to the single-device result."""
    from phasespace_b200.histogram import generate_histograms

    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    jit = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 11)
    chain = D.to_gen(D.kstar_gamma_desc("relbw"))
    n = 100_000
    results = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            w, p = decay.generate(n, seed=9)
            jw, jp = jit.generate(2000, seed=9)
            cw, cp = chain.generate(n, seed=9)
            hw, hp = decay.generate_host(n, seed=9, chunk_events=16384)
            hist, hist_w = generate_histograms(decay, n, seed=9, normalize=False, w_cap=1.0)
            assert w.device.index == dev and cp["K+"].device.index == dev and hist_w.device.index == dev
            results.append((w.cpu(), p["p_2"].cpu(), jw.cpu(), jp["p_10"].cpu(), cw.cpu(), cp["K+"].cpu(),
                            torch.as_tensor(hw.copy()), hist_w.cpu(), hist["p_0"].cpu()))
    for a, b in zip(*results):
        assert torch.equal(a, b) if a.dtype != torch.float64 or a.ndim < 2 or a.shape[0] != 4 else torch.allclose(a, b, rtol=1e-12)
    comp = decay.compile()
    with torch.cuda.device(0):
        w0, _, s0, _ = comp.launch(n // 2, seed=9, first_event=0)
    with torch.cuda.device(1):
        w1, _, s1, _ = comp.launch(n - n // 2, seed=9, first_event=n // 2)
    assert torch.equal(torch.cat([w0.cpu(), w1.cpu()]), results[0][0])


@pytest.mark.parametrize("which", ["two_body", "kstar_gamma"])
def test_covering_grid_with_several_tiles_per_cta(which):
    """Memory-bound trees use the covering grid, where a CTA owns K consecutive tiles (K = 4..16 once the launch is
    large enough): an odd event count that leaves the last CTA a partial chunk, compared piecewise with small launches
    (K = 1) of the same index ranges."""
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS, D.KAON_MASS]) if which == "two_body" else D.to_gen(D.kstar_gamma_desc("bw"))
    n, seed = 5_000_003, 12
    comp = decay.compile()
    w, _, slab, _ = comp.launch(n, seed)
    assert bool(torch.isfinite(w).all())
    for first, count in ((0, 70_001), (2_499_871, 99_999), (n - 65_537, 65_537)):
        wp, _, sp, _ = comp.launch(count, seed, first_event=first)
        assert torch.equal(wp, w[first:first + count])
        assert torch.equal(sp, slab[:, first:first + count])


def test_out_buffers_are_validated():
    """Caller-supplied output buffers reach the C ABI as bare pointers: wrong shape, dtype, device or layout must be
    refused on the host instead of letting the kernel (or the D2H copies) write out of bounds."""
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    comp = decay.compile()
    n, dev = 1000, torch.device("cuda", torch.cuda.current_device())
    good = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
            "particles": torch.empty((3, n, 4), dtype=torch.float64, device=dev)}
    comp.launch(n, seed=1, out=good)
    bad = [
        ("weights", torch.empty((n - 1,), dtype=torch.float64, device=dev)),                 # too small
        ("weights", torch.empty((n,), dtype=torch.float32, device=dev)),                     # dtype
        ("weights", torch.empty((2 * n,), dtype=torch.float64, device=dev)[::2]),            # strided
        ("weights", torch.empty((n,), dtype=torch.float64)),                                 # host memory
        ("particles", torch.empty((3, n - 1, 4), dtype=torch.float64, device=dev)),          # too small
        ("particles", torch.empty((2, n, 4), dtype=torch.float64, device=dev)),              # a plane short
        ("particles", torch.empty((3, 4, n), dtype=torch.float64, device=dev).permute(0, 2, 1)),  # layout
        ("particles", np.empty((3, n, 4))),                                                  # not a tensor
    ]
    for key, tensor in bad:
        with pytest.raises(ValueError, match=key):
            comp.launch(n, seed=1, out={**good, key: tensor})
    with pytest.raises(ValueError, match="weights_max"):
        comp.launch(n, seed=1, normalize_weights=False, out={**good, "weights_max": torch.empty((5,), dtype=torch.float64, device=dev)})
    with pytest.raises(ValueError, match="stats"):
        comp.launch(n, seed=1, out=good, stats=torch.zeros((3,), dtype=torch.float64, device=dev))
    with pytest.raises(ValueError, match="err"):
        comp.launch(n, seed=1, out=good, err=torch.zeros((1,), dtype=torch.int64, device=dev))
    # host delivery: pinned or not, the buffers must be CPU tensors of the exact shape
    host_good = {"weights": torch.empty((n,), dtype=torch.float64, pin_memory=True),
                 "particles": torch.empty((3, n, 4), dtype=torch.float64, pin_memory=True)}
    decay.generate_host(n, seed=1, out=host_good)
    with pytest.raises(ValueError, match="particles"):
        decay.generate_host(n, seed=1, out={**host_good, "particles": torch.empty((3, n // 2, 4), dtype=torch.float64)})
    with pytest.raises(ValueError, match="weights"):
        decay.generate_host(n, seed=1, out={**host_good, "weights": good["weights"]})  # a CUDA tensor


def test_sharded_generation_reports_forbidden_decays():
    """generate_sharded raises where generate() would (the reference asserts "Forbidden decay", ps.py:357-363)."""
    from phasespace_b200 import sharding

    with pytest.raises(phasespace.ForbiddenDecayError):  # statically forbidden masses
        sharding.generate_sharded(phasespace.nbody_decay(300.0, [PION_MASS] * 3), 100, seed=1)
    boost = np.array([[0.0, 0.0, 100.0, 250.0]])  # invariant mass too small for three pions
    with pytest.raises(phasespace.ForbiddenDecayError):
        sharding.generate_sharded(phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3), 100, seed=1, boost_to=boost)
    # a resonance squeezed out by its sibling: K* (min 633) next to a 4700 MeV particle in a 5279 MeV parent
    from phasespace_b200.fromdecay import mass_functions as mf

    squeezed = GenParticle("B", B0_MASS).set_children(
        GenParticle("K*", mf.gauss_factory(895.81, 47.4)).set_children(GenParticle("K", D.KAON_MASS), GenParticle("pi", PION_MASS)),
        GenParticle("X", 4700.0))
    with pytest.raises(phasespace.ForbiddenDecayError):
        squeezed.generate(100, seed=1)
    squeezed2 = GenParticle("B", B0_MASS).set_children(
        GenParticle("K*", mf.gauss_factory(895.81, 47.4)).set_children(GenParticle("K", D.KAON_MASS), GenParticle("pi", PION_MASS)),
        GenParticle("X", 4700.0))
    with pytest.raises(phasespace.ForbiddenDecayError):
        sharding.generate_sharded(squeezed2, 100, seed=1)
    # deferred check: the caller's flag is set, nothing raises
    err = torch.zeros((1,), dtype=torch.int32, device="cuda")
    sharding.generate_sharded(phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3), 100, seed=1, boost_to=boost, err=err)
    assert int(err.item()) == 1
    # and an allowed decay still comes back, identical to generate()
    decay = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 3)
    first, count, w, w_max, parts = sharding.generate_sharded(decay, 1000, seed=4)
    w_ref, p_ref = decay.generate(1000, seed=4)
    assert (first, count) == (0, 1000) and torch.equal(w, w_ref) and torch.equal(parts["p_1"], p_ref["p_1"])


def test_resonance_as_generating_particle_with_boost_to():
    """ps.py:534-543: a resonance may be the particle generate() is called on when boost_to is given -- its mass is then
    kin.mass(boost_to) and the mass function is never called (what the recursion does for every decaying child)."""
    from phasespace_b200.fromdecay import mass_functions as mf

    kstar = GenParticle("K*0", mf.relativistic_breitwigner_factory(D.KSTARZ_MASS, D.KSTARZ_WIDTH)).set_children(
        GenParticle("K+", D.KAON_MASS), GenParticle("pi-", PION_MASS))
    with pytest.raises(ValueError, match="Cannot use resonance as top particle"):
        kstar.generate(10)
    n = 5000
    m = np.random.default_rng(3).uniform(700.0, 1100.0, n)
    p = np.random.default_rng(4).normal(0.0, 2000.0, (n, 3))
    boost = np.concatenate([p, np.sqrt((p * p).sum(1) + m * m)[:, None]], axis=1)
    w, parts = kstar.generate(n, boost_to=boost, seed=2)
    assert list(parts) == ["K+", "pi-"]
    total = _np(parts["K+"] + parts["pi-"])
    assert np.max(np.abs(total - boost)) / np.max(boost[:, 3]) < 1e-12
    # the same events as a fixed-mass particle of any mass boosted to the same rows (the top mass comes from boost_to)
    fixed = GenParticle("K*0", 900.0).set_children(GenParticle("K+", D.KAON_MASS), GenParticle("pi-", PION_MASS))
    w2, parts2 = fixed.generate(n, boost_to=boost, seed=2)
    assert torch.equal(w, w2) and torch.equal(parts["K+"], parts2["K+"])
    # a Python callable as mass function of the generating particle is never evaluated either
    def never(min_mass, max_mass, n_events):
        raise AssertionError("the generating particle's mass function must not be called")

    top = GenParticle("R", never).set_children(GenParticle("a", D.KAON_MASS), GenParticle("b", PION_MASS))
    w3, parts3 = top.generate(n, boost_to=boost, seed=2)
    assert torch.equal(w3, w) and torch.equal(parts3["a"], parts["K+"])
    with pytest.raises(ValueError, match="Cannot use resonance as top particle"):
        top.compile().launch(10, seed=1)


def test_children_added_to_a_leaf_after_compilation():
    """The reference latches only decaying particles (ps.py:207-210, 318): a leaf can still get children after its
    parent has generated, and later generate() calls must include them (no stale compiled tree)."""
    kstar = GenParticle("K*0", D.KSTARZ_MASS)
    top = GenParticle("B0", B0_MASS).set_children(kstar, GenParticle("gamma", 0.0))
    w, parts = top.generate(100, seed=1)
    assert list(parts) == ["K*0", "gamma"]
    kstar.set_children(GenParticle("K+", D.KAON_MASS), GenParticle("pi-", PION_MASS))
    w, parts = top.generate(100, seed=1)
    assert list(parts) == ["K*0", "gamma", "K+", "pi-"]
    assert float((parts["K+"] + parts["pi-"] - parts["K*0"]).abs().max()) / B0_MASS < 1e-12
    with pytest.raises(RuntimeError):
        kstar.set_children(GenParticle("x", 1.0), GenParticle("y", 1.0))


def test_back_to_back_small_calls_keep_stream_order():
    """Small launches opt into programmatic dependent launch (ps_kernel.cuh pdl_prologue): the next call's CTAs may
    become resident while the previous grid drains, but nothing is read or written before `griddepcontrol.wait`.  What a
    caller relies on must hold: (a) calls that write the SAME buffers one after the other leave the last call's events
    there (output memory recycled by the allocator is exactly this case); (b) a boost_to array rewritten on the stream
    between the calls is read as it stands at each call; (c) a consumer kernel that follows sees complete results."""
    n, dev = 200_000, torch.device("cuda")
    comp = phasespace.nbody_decay(B0_MASS, [PION_MASS, D.KAON_MASS]).compile()
    out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
           "particles": torch.empty((2, n, 4), dtype=torch.float64, device=dev)}
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    sums = []
    for seed in range(1, 41):
        comp.launch(n, seed, out=out, err=err)
        sums.append(out["particles"][:, :, 0].sum())  # a torch kernel right behind ours
    torch.cuda.synchronize()
    w_last, _, p_last, _ = comp.launch(n, 40)
    assert torch.equal(out["weights"], w_last) and torch.equal(out["particles"], p_last)
    for seed in (1, 17, 39):
        _, _, p, _ = comp.launch(n, seed)
        assert float(sums[seed - 1]) == float(p[:, :, 0].sum())
    # (b) the input changes on the stream between the calls
    four = phasespace.nbody_decay(B0_MASS, [PION_MASS] * 4).compile()
    boost = torch.tensor([[0.0, 0.0, 1000.0, (1000.0 ** 2 + B0_MASS ** 2) ** 0.5]], dtype=torch.float64, device=dev).repeat(n, 1)
    got = []
    for k in range(12):
        boost[:, 2] += 500.0
        boost[:, 3] = torch.sqrt(boost[:, 2] ** 2 + B0_MASS ** 2)
        w, _, p, _ = four.launch(n, 7, boost_to=boost)
        got.append((w, p))
    torch.cuda.synchronize()
    for k in (0, 5, 11):
        pz = 1000.0 + 500.0 * (k + 1)
        ref_boost = torch.tensor([[0.0, 0.0, pz, (pz ** 2 + B0_MASS ** 2) ** 0.5]], dtype=torch.float64, device=dev).repeat(n, 1)
        w, _, p, _ = four.launch(n, 7, boost_to=ref_boost)
        torch.cuda.synchronize()
        assert torch.equal(got[k][0], w) and torch.equal(got[k][1], p)
