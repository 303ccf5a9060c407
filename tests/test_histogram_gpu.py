"""Fused weighted histograms (ps_histogram) equal the reference-test recipe applied to the stored events:
make_norm_histo(coordinate, weights=weights) for every coordinate, and the weight histogram on (0, 1 + 1e-8)."""

import numpy as np
import pytest
import torch

import phasespace_b200 as phasespace
from phasespace_b200.histogram import generate_histograms

from .helpers import decays as D
from .helpers.plotting import make_norm_histo

pytestmark = pytest.mark.gpu


def _reference_recipe(weights, parts):
    w = weights.cpu().numpy()
    out = {}
    for name, vec in parts.items():
        v = vec.cpu().numpy()
        out[name] = np.stack([make_norm_histo(v[:, c], range_=(-3000 if c != 3 else 0, 3000), weights=w) for c in range(4)])
    return out, make_norm_histo(w, range_=(0, 1 + 1e-8))


@pytest.mark.parametrize("desc", [D.flat_desc(3), D.flat_desc(4), D.kstar_gamma_desc("relbw"), D.flat_desc(11)],
                         ids=["3body", "4body", "kstar_gamma", "11body_jit"])
def test_fused_histograms_match_stored_events(desc):
    n, seed = 300_000, 21
    gen = D.to_gen(desc)
    weights, parts = gen.generate(n, seed=seed)
    ref_parts, ref_w = _reference_recipe(weights, parts)
    got_parts, got_w = generate_histograms(gen, n, seed=seed)
    assert list(got_parts) == list(parts)
    for name in parts:
        np.testing.assert_allclose(got_parts[name].cpu().numpy(), ref_parts[name], rtol=1e-9, atol=1e-15, err_msg=name)
    np.testing.assert_allclose(got_w.cpu().numpy(), ref_w, rtol=1e-12, atol=0)
    # un-normalised: every in-range event is counted once in the weight histogram, chunking does not matter
    raw_parts, raw_w = generate_histograms(gen, n, seed=seed, normalize=False, chunk_events=70_001)
    assert float(raw_w.sum()) == n
    raw2_parts, raw2_w = generate_histograms(gen, n, seed=seed, normalize=False)
    assert torch.equal(raw_w, raw2_w)
    for name in parts:
        np.testing.assert_allclose(raw_parts[name].cpu().numpy(), raw2_parts[name].cpu().numpy(), rtol=1e-11)


def test_fused_histograms_with_boost():
    n, seed = 100_000, 4
    rng = np.random.default_rng(2)
    p = rng.normal(0, 1500.0, (n, 3))
    boost = np.concatenate([p, np.sqrt((p * p).sum(1, keepdims=True) + D.B0_MASS**2)], axis=1)
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3)
    weights, parts = decay.generate(n, boost_to=boost, seed=seed)
    ref_parts, ref_w = _reference_recipe(weights, parts)
    got_parts, got_w = generate_histograms(decay, n, seed=seed, boost_to=boost)
    for name in parts:
        np.testing.assert_allclose(got_parts[name].cpu().numpy(), ref_parts[name], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(got_w.cpu().numpy(), ref_w, rtol=1e-12)
    with pytest.raises(ValueError):
        generate_histograms(decay, n, seed=seed, bins=5000)


def test_sharded_histograms_single_process_and_ranges():
    """generate_histograms_sharded in one process is generate_histograms; histogramming two index ranges separately
    (what two ranks do) and adding them gives the histograms of the whole range."""
    from phasespace_b200 import sharding

    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3)
    n, seed = 200_000, 5
    parts, w = generate_histograms(decay, n, seed=seed, normalize=False, first_event=0)
    sh_parts, sh_w = sharding.generate_histograms_sharded(decay, n, seed, normalize=False)
    assert torch.equal(w, sh_w) and all(torch.allclose(parts[k], sh_parts[k], rtol=1e-12) for k in parts)
    (f0, c0), (f1, c1) = sharding.shard_range(n, 0, 2), sharding.shard_range(n, 1, 2)
    pa, wa = generate_histograms(decay, c0, seed=seed, normalize=False, first_event=f0)
    pb, wb = generate_histograms(decay, c1, seed=seed, normalize=False, first_event=f1)
    assert torch.equal(wa + wb, w)
    for k in parts:
        np.testing.assert_allclose((pa[k] + pb[k]).cpu().numpy(), parts[k].cpu().numpy(), rtol=1e-11)


def test_histogram_folds_and_single_bin():
    """Enough events per CTA for several fixed-point folds (one every 16 tiles), and the worst case for the 32-bit
    limbs: a single bin that takes every event of every tile."""
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3)
    n, seed = 3_000_000, 31
    weights, parts = decay.generate(n, seed=seed)
    ref_parts, ref_w = _reference_recipe(weights, parts)
    got_parts, got_w = generate_histograms(decay, n, seed=seed)
    for name in parts:
        np.testing.assert_allclose(got_parts[name].cpu().numpy(), ref_parts[name], rtol=1e-9, atol=1e-15, err_msg=name)
    np.testing.assert_allclose(got_w.cpu().numpy(), ref_w, rtol=1e-12, atol=0)
    one_parts, one_w = generate_histograms(decay, n, seed=seed, bins=1, normalize=False)
    assert float(one_w[0]) == n
    total = float(weights.sum())
    for name in parts:  # |p| < 3000 and 0 <= E < 3000 for every pion of a B0 decay: each row holds the sum of all weights
        np.testing.assert_allclose(one_parts[name].cpu().numpy()[:, 0], total, rtol=1e-12)
