"""GenMultiDecay.generate: the check_norm helper and the generating halves of the reference's
tests/fromdecay/test_fromdecay.py:14-37, 59-188."""

import numpy as np
import pytest
import torch

import phasespace_b200 as phasespace
from phasespace_b200.fromdecay import GenMultiDecay

from .helpers import decays as D
from .helpers import example_decay_chains

pytestmark = pytest.mark.gpu


def check_norm(full_decay, **kwargs):
    all_return_args = []
    for norm in (True, False):
        return_args = full_decay.generate(normalize_weights=norm, **kwargs)
        assert len(return_args) == (2 if norm else 3)
        assert sum(len(w) for w in return_args[0]) == kwargs["n_events"]
        if not norm:
            assert all(len(w) == len(mw) for w, mw in zip(return_args[0], return_args[1]))
        all_return_args.append(return_args)
    return all_return_args


def test_single_chain():
    container = GenMultiDecay.from_dict(example_decay_chains.dplus_single, tolerance=1e-10)
    check_norm(container, n_events=1)
    (_, decay_list), _ = check_norm(container, n_events=100)
    assert len(decay_list) == 1
    events = decay_list[0]
    assert set(events.keys()) == {"K-", "pi+", "pi+ [0]", "pi0", "gamma", "gamma [0]"}
    assert all(len(p) == 100 for p in events.values())
    # the pi0 is a relativistic Breit-Wigner with its physical width (7.8e-6 MeV, width/mass = 6e-8) at this tolerance:
    # the two photons reconstruct it within a few widths of the pole (regression for the narrow-resonance constants)
    (weights, decay_list) = container.generate(n_events=20_000, seed=2)
    ev = decay_list[0]
    gg = ev["gamma"] + ev["gamma [0]"]
    m_gg = torch.sqrt(gg[:, 3] ** 2 - (gg[:, :3] ** 2).sum(1))
    assert float((m_gg - 134.9768).abs().median()) < 1e-4
    total = sum(ev[k] for k in ("K-", "pi+", "pi+ [0]", "gamma", "gamma [0]"))
    assert float(total[:, :3].abs().max()) < 1e-8 and bool(torch.isfinite(weights[0]).all())


@pytest.mark.parametrize("chain", ["pi0_4branches", "dplus_4grandbranches", "dstarplus_big_decay"])
def test_branching(chain):
    tol = None if chain == "dstarplus_big_decay" else 1e-10
    container = GenMultiDecay.from_dict(getattr(example_decay_chains, chain), tolerance=tol)
    check_norm(container, n_events=1)
    check_norm(container, n_events=100)
    (weights, events), _ = check_norm(container, n_events=20_000, seed=5)
    assert all(torch.isfinite(w).all() and (w <= 1 + 1e-12).all() for w in weights)
    assert max(float(w.max()) for w in weights) > 0.0


def test_list_container_and_global_max():
    modes = [(0.7, phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3)),
             (0.3, D.to_gen(D.kstar_gamma_desc("gauss")))]
    multi = GenMultiDecay(modes)
    n = 50_000
    weights, max_weights, events = multi.generate(n, normalize_weights=False, seed=3)
    counts = [len(w) for w in weights]
    assert sum(counts) == n and abs(counts[0] / n - 0.7) < 0.02
    normed, events2 = multi.generate(n, normalize_weights=True, seed=3)
    total_max = max(float(mw.max()) for mw in max_weights)  # genmultidecay.py:192-195
    for w, wn in zip(weights, normed):
        np.testing.assert_allclose(wn.cpu().numpy(), w.cpu().numpy() / total_max, rtol=1e-14)
    assert list(events[1]) == ["K*0", "gamma", "K+", "pi-"]
    again, _ = multi.generate(n, normalize_weights=True, seed=3)
    assert all(torch.equal(a, b) for a, b in zip(normed, again))


def test_global_max_with_event_dependent_maximum_weights():
    """A mode whose maximum weight depends on the event (a resonant LEAF, or boost_to) cannot be normalised inside
    the kernel by a number known in advance: the reference's two-pass recipe (genmultidecay.py:185-195) runs, and the
    results still equal w / max over all modes of max(w_max)."""
    from phasespace_b200 import GenParticle
    from phasespace_b200.fromdecay import mass_functions as mf

    leafy = GenParticle("B0", D.B0_MASS).set_children(
        GenParticle("rho", mf.breitwigner_factory(775.26, 149.1)), GenParticle("pi", D.PION_MASS), GenParticle("K", D.KAON_MASS))
    multi = GenMultiDecay([(0.5, phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 4)), (0.5, leafy)])
    n = 40_000
    weights, max_weights, events = multi.generate(n, normalize_weights=False, seed=11)
    assert float(max_weights[1].std()) > 0.0  # really event dependent
    normed, _ = multi.generate(n, normalize_weights=True, seed=11)
    total_max = max(float(mw.max()) for mw in max_weights)
    for w, wn in zip(weights, normed):
        np.testing.assert_allclose(wn.cpu().numpy(), w.cpu().numpy() / total_max, rtol=1e-14)
    # boost_to: the same fallback
    boost = np.array([[1000.0, -2000.0, 30000.0, np.sqrt(1000.0**2 + 2000.0**2 + 30000.0**2 + D.B0_MASS**2)]])
    flat = GenMultiDecay([(0.6, phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3)),
                          (0.4, phasespace.nbody_decay(D.B0_MASS, [D.KAON_MASS, D.PION_MASS]))])
    w_b, mw_b, _ = flat.generate(n, normalize_weights=False, seed=2, boost_to=boost)
    n_b, _ = flat.generate(n, normalize_weights=True, seed=2, boost_to=boost)
    t_max = max(float(m.max()) for m in mw_b)
    for w, wn in zip(w_b, n_b):
        np.testing.assert_allclose(wn.cpu().numpy(), w.cpu().numpy() / t_max, rtol=1e-14)


def test_weight_scale_is_refused_where_it_cannot_be_folded():
    comp = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3).compile()
    w1, _, _, _ = comp.launch(1000, seed=1)
    w2, _, _, _ = comp.launch(1000, seed=1, weight_scale=0.25)
    np.testing.assert_allclose(w2.cpu().numpy(), 0.25 * w1.cpu().numpy(), rtol=1e-15)
    with pytest.raises(ValueError, match="weight_scale"):
        comp.launch(1000, seed=1, weight_scale=0.5, normalize_weights=False)
    with pytest.raises(ValueError, match="weight_scale"):
        comp.launch(1000, seed=1, weight_scale=0.5, boost_to=torch.tensor([[0.0, 0.0, 100.0, 6000.0]], dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError, match="weight_scale"):
        comp.launch(1000, seed=1, weight_scale=float("nan"))


def test_modes_are_compiled_side_by_side(tmp_path, monkeypatch):
    """Kernels of decay modes that have no ahead-of-time kernel are compiled by NVRTC (1-3 s each).  GenMultiDecay
    prepares all populated modes concurrently before the first launch (ps_tree_prepare; the reference traces one
    tf.function per mode, one after the other, genmultidecay.py:185-190): with an empty on-disk cache the first
    generate() of six new trees must cost well under six single compilations, every mode ends up prepared, and the
    events do not depend on who compiled the kernel."""
    import time

    from phasespace_b200.phasespace import prepare_all

    monkeypatch.setenv("PS_B200_CACHE_DIR", str(tmp_path))

    def modes(tag):
        out = []
        for k in range(6):  # six tree shapes no other test uses: 3 + k bodies below a two-body decay
            sub = phasespace.GenParticle(f"X{tag}", 3000.0 + k).set_children(
                *[phasespace.GenParticle(f"x{tag}{j}", 100.0 + j) for j in range(3 + k)])
            top = phasespace.GenParticle(f"T{tag}", 9000.0).set_children(
                sub, phasespace.GenParticle(f"g{tag}", 0.0), phasespace.GenParticle(f"h{tag}", 500.0 + k))
            out.append((1.0 / 6.0, top))
        return out

    gmd = GenMultiDecay(modes("a"))
    t0 = time.perf_counter()
    w_a, ev_a = gmd.generate(6000, seed=5)
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    # one more tree of the size of the largest mode (another shape: the six above are in the process by now), alone
    single = phasespace.GenParticle("Ts", 9000.0).set_children(
        phasespace.GenParticle("Xs", 3000.0).set_children(*[phasespace.GenParticle(f"xs{j}", 100.0 + j) for j in range(8)]),
        phasespace.GenParticle("gs", 0.0), phasespace.GenParticle("hs", 500.0), phasespace.GenParticle("is", 139.57)).compile()
    t0 = time.perf_counter()
    single.prepare(device_index=torch.cuda.current_device())
    one = time.perf_counter() - t0
    assert all(gen.compile()._prepared for _, gen in gmd.gen_particles)
    import os

    if (os.cpu_count() or 1) >= 4:
        # six trees of 0.6-1.0 x `one` each: about 5 x `one` one after the other, about `one` side by side
        assert first < 2.5 * one + 2.0, (first, one)
    # the same modes again, now from the registry of the process and prepared one by one
    gmd_b = GenMultiDecay(modes("b"))
    for _, gen in gmd_b.gen_particles:
        prepare_all([gen.compile()])
    w_b, ev_b = gmd_b.generate(6000, seed=5)
    assert len(w_a) == len(w_b) == 6
    for wa, wb, ea, eb in zip(w_a, w_b, ev_a, ev_b):
        assert torch.equal(wa, wb)
        assert all(torch.equal(ea[ka], eb[kb]) for ka, kb in zip(ea, eb))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_multidecay_is_bit_identical(world):
    """the mode counts come from the seed, every rank generates a contiguous slice of every mode at the
    event indices the single-GPU call uses -- the concatenation over ranks equals the single-GPU result bit for bit, with
    no collective (the global maximum weight of these chains is known on the host)."""
    container = GenMultiDecay.from_dict(example_decay_chains.dstarplus_big_decay, tolerance=1e-10)
    n = 50_001
    w_ref, ev_ref = container.generate(n, seed=11)
    shares = [container.generate_sharded(n, seed=11, rank=r, world=world) for r in range(world)]
    assert all(len(s[0]) == len(w_ref) for s in shares)
    for k, (w, ev) in enumerate(zip(w_ref, ev_ref)):
        assert torch.equal(torch.cat([s[0][k] for s in shares]), w)
        for name, vec in ev.items():
            assert torch.equal(torch.cat([s[1][k][name] for s in shares]), vec)
    # unnormalised weights come with their maximum weights, sharded the same way
    w3, mw3, _ = container.generate(n, normalize_weights=False, seed=12)
    parts = [container.generate_sharded(n, normalize_weights=False, seed=12, rank=r, world=world) for r in range(world)]
    for k in range(len(w3)):
        assert torch.equal(torch.cat([p[0][k] for p in parts]), w3[k])
        assert torch.equal(torch.cat([p[1][k] for p in parts]), mw3[k])


def test_sharded_multidecay_needs_a_process_group_for_a_global_maximum():
    """Modes whose maximum weight depends on the event (boost_to) are normalised by the maximum over ALL ranks: one scalar
    MAX all-reduce (tests/test_sharding_gloo.py runs it on two ranks).  One rank is the identity; several ranks without a
    process group are refused instead of normalising each share by its own maximum."""
    container = GenMultiDecay.from_dict(example_decay_chains.dplus_4grandbranches, tolerance=1e-10)
    boost = torch.tensor([[0.0, 0.0, 3000.0, (3000.0 ** 2 + 1869.66 ** 2) ** 0.5]], dtype=torch.float64)
    w_ref, _ = container.generate(2000, seed=3, boost_to=boost)
    w_one, _ = container.generate_sharded(2000, seed=3, boost_to=boost, rank=0, world=1)
    assert all(torch.equal(a, b) for a, b in zip(w_ref, w_one))
    with pytest.raises(RuntimeError):
        container.generate_sharded(2000, seed=3, boost_to=boost, rank=0, world=2)
    with pytest.raises(ValueError):
        container.generate_sharded(2000, seed=3, boost_to=boost.expand(2000, 4), rank=0, world=2)


def test_fast_path_and_side_streams_give_the_same_events():
    """GenMultiDecay issues its modes through CompiledDecay.generate_fast (one allocation, one C-ABI call, one shared
    forbidden-decay flag) on the current stream for small calls and fanned out over side streams for large ones; the
    general per-mode path (any extra keyword) and both stream layouts must give identical events.  A forbidden mode is
    still reported through the shared flag."""
    container = GenMultiDecay.from_dict(example_decay_chains.dstarplus_big_decay, tolerance=1e-10)
    n = 30_011
    w_fast, ev_fast = container.generate(n, seed=21)
    w_gen, ev_gen = container.generate(n, seed=21, exact=False)  # a keyword: the general path
    streamed = GenMultiDecay.from_dict(example_decay_chains.dstarplus_big_decay, tolerance=1e-10)
    streamed.MIN_EVENTS_FOR_STREAMS = 0
    w_st, ev_st = streamed.generate(n, seed=21)
    w3_fast, mw_fast, _ = container.generate(n, normalize_weights=False, seed=22)
    w3_gen, mw_gen, _ = container.generate(n, normalize_weights=False, seed=22, exact=False)
    assert len(w_fast) == len(w_gen) == len(w_st) > 3
    for k in range(len(w_fast)):
        assert torch.equal(w_fast[k], w_gen[k]) and torch.equal(w_fast[k], w_st[k])
        for name in ev_gen[k]:
            assert torch.equal(ev_fast[k][name], ev_gen[k][name]) and torch.equal(ev_fast[k][name], ev_st[k][name])
    assert len(w3_fast) == len(w3_gen) == len(mw_fast) == len(mw_gen) > 3
    for k in range(len(w3_fast)):
        assert torch.equal(w3_fast[k], w3_gen[k]) and torch.equal(mw_fast[k], mw_gen[k])
    # a resonance that leaves no room for its sibling: forbidden in (nearly) every event, flagged by the kernel
    from phasespace_b200.fromdecay import mass_functions as mf
    from phasespace_b200.phasespace import ForbiddenDecayError

    bad = phasespace.GenParticle("A", 1000.0).set_children(
        phasespace.GenParticle("R1", mf.gauss_factory(990.0, 1.0)).set_children(
            phasespace.GenParticle("a", 100.0), phasespace.GenParticle("b", 100.0)),
        phasespace.GenParticle("R2", mf.gauss_factory(500.0, 1.0)).set_children(
            phasespace.GenParticle("c", 100.0), phasespace.GenParticle("d", 100.0)))
    good = phasespace.GenParticle("B", 1000.0).set_children(phasespace.GenParticle("e", 100.0), phasespace.GenParticle("f", 100.0))
    with pytest.raises(ForbiddenDecayError):
        GenMultiDecay([(0.5, good), (0.5, bad)]).generate(1000, seed=1)
