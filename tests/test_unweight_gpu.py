"""Accept-reject unweighting fused with the generator (BASELINE config 5 scaled down): the accepted set is
exactly {i : u_i * w_bound < w_i / w_max} with u from Philox stream 1, ordered, and regenerated bit for bit."""

import numpy as np
import pytest
import torch

import phasespace_b200 as phasespace
from oracle import genbod_numpy as orc
from oracle.philox import uniform_table
from phasespace_b200.unweight import UnweightingOverflowError, generate_unweighted

from .helpers import decays as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_body,w_bound,n", [(3, 1.0, 100_003), (3, 0.4, 50_000), (10, 1e-6, 200_000), (4, 1.0, 255)])
def test_unweighted_set_matches_oracle(n_body, w_bound, n):
    seed, first = 17, 1000
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * n_body)
    parts, index, weights = generate_unweighted(decay, n, seed, w_bound=w_bound, first_event=first, on_overflow="ignore")
    tree = orc.flat(D.B0_MASS, [D.PION_MASS] * n_body)
    ref_w, ref_p = orc.generate(tree, n, uniform_table(seed, first, n, orc.n_draws(tree)))
    u = uniform_table(seed, first, n, 1, stream=1)[:, 0]
    margin = np.abs(u * w_bound - ref_w) > 1e-12 * ref_w  # events decided by rounding are not compared
    accept = u * w_bound < ref_w
    idx = index.cpu().numpy() - first
    assert np.all(np.diff(idx) > 0)
    got = np.zeros(n, dtype=bool)
    got[idx] = True
    assert np.array_equal(got[margin], accept[margin])
    np.testing.assert_allclose(weights.cpu().numpy(), ref_w[idx], rtol=1e-11)
    assert len(idx) > 0
    for name in ref_p:
        assert np.max(np.abs(parts[name].cpu().numpy() - ref_p[name][idx])) / D.B0_MASS <= 1e-12
    # regenerated events are the same bits the plain generator makes for those indices
    w_all, p_all = decay.generate(n, seed=phasespace.random.Generator(seed, first))
    assert torch.equal(weights, w_all[idx])
    assert all(torch.equal(parts[k], p_all[k][idx]) for k in parts)
    if w_bound == 1.0 and n > 1000:
        assert abs(len(idx) / n - ref_w.mean()) < 5 * np.sqrt(ref_w.mean() / n)


def test_unweighting_is_independent_of_chunking():
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 4)
    _, full, _ = generate_unweighted(decay, 70_000, 5, w_bound=0.2, on_overflow="ignore")
    a = generate_unweighted(decay, 30_001, 5, w_bound=0.2, on_overflow="ignore")[1]
    b = generate_unweighted(decay, 70_000 - 30_001, 5, w_bound=0.2, first_event=30_001, on_overflow="ignore")[1]
    assert torch.equal(torch.cat([a, b]), full)


def test_unweighted_chain():
    chain = D.to_gen(D.kstar_gamma_desc("relbw"))
    parts, index, weights = generate_unweighted(chain, 40_000, 8)
    assert list(parts) == ["K*0", "gamma", "K+", "pi-"]
    assert 0 < len(index) < 40_000 and torch.all(weights <= 1)
    w_all, _ = chain.generate(40_000, seed=8)
    assert torch.equal(weights, w_all[index])


@pytest.mark.parametrize("n", [0, 1, 2048 * 256, 2048 * 256 + 1, 1_300_003])
def test_multi_segment_scan(n):
    """The tile-count scan works in segments of 2048 tiles (524288 events): sizes of zero, one, exactly one segment,
    one event more, and two and a half segments, against the accept rule applied to the plain generator's weights."""
    seed, w_bound = 23, 0.5
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3)
    parts, index, weights = generate_unweighted(decay, n, seed, w_bound=w_bound, on_overflow="ignore")
    if n == 0:
        assert index.numel() == 0 and weights.numel() == 0
        return
    w_all, p_all = decay.generate(n, seed=seed)
    u = torch.as_tensor(uniform_table(seed, 0, n, 1, stream=1)[:, 0], device=w_all.device)
    decided = (u * w_bound - w_all).abs() > 1e-12  # events decided by rounding are not compared
    expect = u * w_bound < w_all
    got = torch.zeros(n, dtype=torch.bool, device=w_all.device)
    got[index] = True
    assert bool((index[1:] > index[:-1]).all())
    assert torch.equal(got[decided], expect[decided])
    assert torch.equal(weights, w_all[index]) and torch.equal(parts["p_1"], p_all["p_1"][index])


@pytest.mark.parametrize("n_body,w_bound", [(3, 0.4), (3, 1.0), (4, 0.25), (10, 2e-7)])
def test_overflow_is_counted(n_body, w_bound):
    """A candidate with w / w_max > w_bound is accepted with probability one instead of > 1: the mask pass counts
    them, and the default is to refuse the biased sample."""
    n, seed = 300_000, 29
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * n_body)
    w_all, _ = decay.generate(n, seed=seed)
    decided = (w_all - w_bound).abs() > 1e-12 * w_bound  # weights within rounding of the bound may fall either way
    lo = int(((w_all > w_bound) & decided).sum())
    hi = lo + int((~decided).sum())
    parts, index, weights, n_over = generate_unweighted(decay, n, seed, w_bound=w_bound, return_overflow=True)
    assert lo <= n_over <= hi
    # every overflowing candidate is in the accepted set (u < 1 <= w / w_bound)
    got = torch.zeros(n, dtype=torch.bool, device=w_all.device)
    got[index] = True
    assert bool(got[(w_all > w_bound) & decided].all())
    if lo > 0:
        with pytest.raises(UnweightingOverflowError):
            generate_unweighted(decay, n, seed, w_bound=w_bound)
        with pytest.warns(RuntimeWarning):
            generate_unweighted(decay, n, seed, w_bound=w_bound, on_overflow="warn")
    else:
        generate_unweighted(decay, n, seed, w_bound=w_bound)  # does not raise


def test_unweighting_parity_at_1e8_candidates():
    """Full-size parity of the accept-reject pass (BASELINE config 5's chunk size, 2**27 = 1.3e8 candidates of the
    10-body decay): the accepted set is exactly {i : u_i w_bound < w_i} with w from the plain generator (a different
    kernel, 4-vectors and all) and u recomputed from the (seed, event) counters; it is independent of how the range
    is chunked; the overflow count equals the number of weights above the bound."""
    n, seed, first = 1 << 27, 5, 7_000_000_000
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 10)
    comp = decay.compile()
    dev = torch.device("cuda")
    w_all = torch.empty((n,), dtype=torch.float64, device=dev)
    piece = 1 << 24  # the plain generator in pieces: only the weights are kept (each piece's slab is 5 GB)
    slab = torch.empty((10, piece, 4), dtype=torch.float64, device=dev)
    for k in range(0, n, piece):
        comp.launch(piece, seed, first_event=first + k, out={"weights": w_all[k:k + piece], "particles": slab})
    del slab
    w_bound = float(w_all.max()) * 0.8  # below the maximum on purpose: some candidates overflow
    parts, index, weights, n_over = generate_unweighted(decay, n, seed, w_bound=w_bound, first_event=first,
                                                        return_overflow=True)
    assert n_over == int((w_all > w_bound).sum()) > 0
    assert bool((index[1:] > index[:-1]).all())
    local = index - first
    assert torch.equal(weights, w_all[local])
    # acceptance rule on a random subset of candidates, accepted or not (u from the numpy Philox restatement)
    rng = np.random.default_rng(0)
    probe = np.unique(rng.integers(0, n, 200_000))
    # uniform_table works on contiguous ranges: probe 200 short runs instead of single indices
    starts = np.unique(rng.integers(0, n - 1024, 200))
    got = torch.zeros(n, dtype=torch.bool, device=dev)
    got[local] = True
    checked = 0
    for s0 in starts:
        u = torch.as_tensor(uniform_table(seed, first + int(s0), 1024, 1, stream=1)[:, 0], device=dev)
        w = w_all[s0:s0 + 1024]
        decided = (u * w_bound - w).abs() > 1e-12 * w
        expect = u * w_bound < w
        assert torch.equal(got[s0:s0 + 1024][decided], expect[decided])
        checked += int(decided.sum())
    assert checked > 150_000
    del probe
    # the accepted fraction is the mean of min(w / w_bound, 1) within binomial noise
    p_acc = float(torch.clamp(w_all / w_bound, max=1.0).mean())
    assert abs(index.numel() / n - p_acc) < 6 * np.sqrt(p_acc / n)
    # chunking: two unequal chunks give the same ordered list
    cut = 50_000_017
    a = generate_unweighted(decay, cut, seed, w_bound=w_bound, first_event=first, on_overflow="ignore")[1]
    b = generate_unweighted(decay, n - cut, seed, w_bound=w_bound, first_event=first + cut, on_overflow="ignore")[1]
    assert torch.equal(torch.cat([a, b]), index)
    # accepted events are on shell and conserve four-momentum
    total = sum(parts.values())
    assert float(total[:, :3].abs().max()) / D.B0_MASS < 1e-12
    assert float((total[:, 3] - D.B0_MASS).abs().max()) / D.B0_MASS < 1e-12


def test_more_than_2_to_the_32_candidates_in_one_call():
    """Maximum sizes: one call over 2**32 + 12345 candidates (event counts and indices are int64 throughout: 512 MB of
    mask bits, 1.7e7 tiles, 8193 scan segments).  A loose bound keeps the accepted sample at a few GB.  The ordered
    list reaches past 2**32 and equals the lists of two unequal chunks of the same range, weights and events included."""
    n, seed, first, w_bound = (1 << 32) + 12345, 11, 3, 40.0
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3)
    parts, index, weights = generate_unweighted(decay, n, seed, w_bound=w_bound, first_event=first)
    n_acc = index.numel()
    assert abs(n_acc / n - 0.27 / w_bound) < 0.02 / w_bound  # mean normalised weight of the flat three-body decay
    assert bool((index[1:] > index[:-1]).all())
    assert int(index[0]) >= first and int(index[-1]) < first + n and int(index[-1]) - first > (1 << 32) - 100_000
    assert int((index - first >= (1 << 32)).sum()) > 0  # the tail beyond the 32-bit range is there
    assert float(weights.max()) <= 1.0 and float(weights.min()) > 0.0
    cut = (1 << 31) + 777
    pa, ia, wa = generate_unweighted(decay, cut, seed, w_bound=w_bound, first_event=first)
    pb, ib, wb = generate_unweighted(decay, n - cut, seed, w_bound=w_bound, first_event=first + cut)
    assert ia.numel() + ib.numel() == n_acc
    assert torch.equal(index[:ia.numel()], ia) and torch.equal(index[ia.numel():], ib)
    assert torch.equal(weights[:ia.numel()], wa) and torch.equal(weights[ia.numel():], wb)
    for name in parts:
        assert torch.equal(parts[name][:ia.numel()], pa[name]) and torch.equal(parts[name][ia.numel():], pb[name])
