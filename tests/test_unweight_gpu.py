"""Accept-reject unweighting fused with the generator (BASELINE config 5 scaled down): the accepted set is
exactly {i : u_i * w_bound < w_i / w_max} with u from Philox stream 1, ordered, and regenerated bit for bit."""

import numpy as np
import pytest
import torch

import phasespace_b200 as phasespace
from oracle import genbod_numpy as orc
from oracle.philox import uniform_table
from phasespace_b200.unweight import generate_unweighted

from .helpers import decays as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_body,w_bound,n", [(3, 1.0, 100_003), (3, 0.4, 50_000), (10, 1e-6, 200_000), (4, 1.0, 255)])
def test_unweighted_set_matches_oracle(n_body, w_bound, n):
    seed, first = 17, 1000
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * n_body)
    parts, index, weights = generate_unweighted(decay, n, seed, w_bound=w_bound, first_event=first)
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
    _, full, _ = generate_unweighted(decay, 70_000, 5, w_bound=0.2)
    a = generate_unweighted(decay, 30_001, 5, w_bound=0.2)[1]
    b = generate_unweighted(decay, 70_000 - 30_001, 5, w_bound=0.2, first_event=30_001)[1]
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
    parts, index, weights = generate_unweighted(decay, n, seed, w_bound=w_bound)
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
