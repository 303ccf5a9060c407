"""BASELINE.json's full sizes (1e8 events) through size-independent properties: 4-momentum conservation, on-shell
masses, weight bounds, agreement of weight moments with the oracle, and chunk-by-chunk bit identity
(a contiguous index range regenerated separately equals the slice of the big launch)."""

import numpy as np
import pytest
import torch

import phasespace_b200 as phasespace
from oracle import genbod_numpy as orc
from oracle.philox import uniform_table

from .helpers import decays as D

pytestmark = pytest.mark.gpu
N_FULL = 100_000_000


def _mass(v):
    return torch.sqrt(torch.clamp(v[:, 3] ** 2 - (v[:, :3] ** 2).sum(1), min=0.0))


def test_config2_three_body_1e8():
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 3)
    comp = decay.compile()
    w, _, p, err = comp.launch(N_FULL, seed=1)
    assert int(err.item()) == 0
    total = p.sum(0)
    assert float(total[:, :3].abs().max()) / D.B0_MASS < 1e-12
    assert float((total[:, 3] - D.B0_MASS).abs().max()) / D.B0_MASS < 1e-12
    for s in range(3):
        assert float((_mass(p[s]) - D.PION_MASS).abs().max()) < 1e-6
    assert float(w.min()) > 0.0 and float(w.max()) < 1.0
    # moments of the normalised weight against an oracle sample (statistical: 5 sigma of the oracle sample)
    n_ref = 400_000
    tree = orc.flat(D.B0_MASS, [D.PION_MASS] * 3)
    ref_w, _ = orc.generate(tree, n_ref, uniform_table(123, 0, n_ref, 5))
    assert abs(float(w.mean()) - ref_w.mean()) < 5 * ref_w.std() / np.sqrt(n_ref)
    assert abs(float(w.max()) - ref_w.max()) < 2e-3
    # the head of the big launch equals the oracle fed the same (seed, event, column) draws
    head_w, head_p = orc.generate(tree, 1000, uniform_table(1, 0, 1000, 5))
    np.testing.assert_allclose(w[:1000].cpu().numpy(), head_w, rtol=1e-12)
    tail_first = N_FULL - 1000
    tail_w, _ = orc.generate(tree, 1000, uniform_table(1, tail_first, 1000, 5))
    np.testing.assert_allclose(w[tail_first:].cpu().numpy(), tail_w, rtol=1e-12)
    # eight contiguous shards (what eight GPUs would generate) are bit-identical slices
    per = N_FULL // 8
    for r in range(8):
        w2, _, p2, _ = comp.launch(per, seed=1, first_event=r * per)
        assert torch.equal(w2, w[r * per:(r + 1) * per]) and torch.equal(p2, p[:, r * per:(r + 1) * per])


def test_config3_four_body_boosted_1e8():
    n = N_FULL
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1234)
    pt = 1000.0 * torch.empty(n, dtype=torch.float64, device=dev).exponential_(0.2, generator=g)
    eta = 2 + 3 * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    phi = 2 * np.pi * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    boost = torch.stack([pt * torch.cos(phi), pt * torch.sin(phi), pt * torch.sinh(eta), torch.zeros_like(pt)], dim=1)
    boost[:, 3] = torch.sqrt((boost[:, :3] ** 2).sum(1) + D.B0_MASS**2)
    del pt, eta, phi
    decay = phasespace.nbody_decay(D.B0_MASS, [D.PION_MASS] * 4)
    w, p = decay.generate(n, boost_to=boost, seed=2)
    total = sum(p.values())
    assert float(((total - boost).abs() / boost[:, 3:4]).max()) < 1e-12
    del total
    assert float(w.min()) > 0.0 and float(w.max()) < 1.0
    w0, _ = decay.generate(n, seed=2)  # same draws at rest: weights are frame independent
    assert float(((w - w0).abs() / w0).max()) < 1e-6
    k = 2000
    tree = orc.flat(D.B0_MASS, [D.PION_MASS] * 4)
    ref_w, ref_p = orc.generate(tree, k, uniform_table(2, 0, k, 8), boost_to=boost[:k].cpu().numpy())
    np.testing.assert_allclose(w[:k].cpu().numpy(), ref_w, rtol=1e-9)


def test_config4_kstar_gamma_relbw_1e8():
    n = N_FULL
    chain = D.to_gen(D.kstar_gamma_desc("relbw"))
    w, p = chain.generate(n, seed=3)
    assert list(p) == ["K*0", "gamma", "K+", "pi-"]
    top = p["K*0"] + p["gamma"]
    assert float(top[:, :3].abs().max()) / D.B0_MASS < 1e-12 and float((top[:, 3] - D.B0_MASS).abs().max()) / D.B0_MASS < 1e-12
    del top
    assert float(((p["K+"] + p["pi-"] - p["K*0"]).abs()).max()) / D.B0_MASS < 1e-12
    m_kst = _mass(p["K*0"])
    assert float(m_kst.min()) >= D.KAON_MASS + D.PION_MASS - 1e-6 and float(m_kst.max()) <= D.B0_MASS + 1e-6
    assert float(w.min()) >= 0.0 and float(w.max()) < 1.0
    # the K* line shape: compare quantiles with an oracle sample of the same density
    n_ref = 400_000
    ref = orc.sample_mass(orc.RELBW, D.KSTARZ_MASS, D.KSTARZ_WIDTH, np.full(n_ref, D.KAON_MASS + D.PION_MASS),
                          np.full(n_ref, D.B0_MASS), np.random.default_rng(1).random(n_ref))
    qs = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9], dtype=torch.float64, device=m_kst.device)
    got = torch.quantile(m_kst[:10_000_000], qs).cpu().numpy()
    np.testing.assert_allclose(got, np.quantile(ref, qs.cpu().numpy()), rtol=2e-3)


def test_more_than_2_to_the_31_events_in_one_launch():
    """Maximum sizes: the largest single launch a B200 holds -- 2**31 + 4096 two-body events, 155 GB of outputs in one
    allocation (event counts, row offsets and plane strides are int64 throughout; 8.4e6 tiles on the covering grid).
    Checked chunk by chunk: momentum conservation, on-shell masses, unit weights, and
    head / middle / tail against the oracle and against separate launches at the same event indices."""
    n = (1 << 31) + 4096
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 72 * n + (6 << 30):
        pytest.skip(f"needs {72 * n / 1e9:.0f} GB of free HBM, the device has {free / 1e9:.0f} GB free")
    m1, m2 = D.PION_MASS, D.KAON_MASS
    decay = phasespace.nbody_decay(D.B0_MASS, [m1, m2])
    comp = decay.compile()
    w, _, p, err = comp.launch(n, seed=17)
    assert int(err.item()) == 0
    step = 1 << 25
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        a, b = p[0, lo:hi], p[1, lo:hi]
        total = a + b
        assert float(total[:, :3].abs().max()) / D.B0_MASS < 1e-12
        assert float((total[:, 3] - D.B0_MASS).abs().max()) / D.B0_MASS < 1e-12
        assert float((_mass(a) - m1).abs().max()) < 1e-6 and float((_mass(b) - m2).abs().max()) < 1e-6
        assert float((w[lo:hi] - 1.0).abs().max()) < 1e-14
        del a, b, total
    tree = orc.flat(D.B0_MASS, [m1, m2])
    for first in (0, (1 << 31) - 500, n - 1000):  # the middle window straddles the 32-bit boundary
        ref_w, ref_p = orc.generate(tree, 1000, uniform_table(17, first, 1000, 2))
        np.testing.assert_allclose(w[first:first + 1000].cpu().numpy(), ref_w, rtol=1e-12)
        for k, name in enumerate(orc.output_names(tree)):
            np.testing.assert_allclose(p[k, first:first + 1000].cpu().numpy(), ref_p[name], rtol=0, atol=1e-12 * D.B0_MASS)
        w2, _, p2, _ = comp.launch(1000, seed=17, first_event=first)
        assert torch.equal(w2, w[first:first + 1000]) and torch.equal(p2, p[:, first:first + 1000])
    del w, p
    torch.cuda.empty_cache()
