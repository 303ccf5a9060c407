"""How ill-conditioned is an event?  A per-event amplification factor computed from the REFERENCE outputs alone.

The weight of an event is a product of two-body momenta pdk(a, b, c) = sqrt((a-b-c)(a+b+c)(a-b+c)(a+b-c)) / 2a
(phasespace.py:56-71).  Near a threshold the factor (a - b - c) cancels: a rounding error eps * a in the invariant
masses becomes a relative error eps * a / (a - b - c) of the weight.  A decaying particle that is not the top particle
has its mass re-derived as sqrt(E^2 - p^2) of its *boosted* 4-vector (phasespace.py:322 via 564), which adds a relative
error eps * gamma^2 (gamma = E / m in the output frame) to its a.  So two correct fp64 evaluations of the same event
(the reference's and the kernel's differ in the last bits of sin/cos/sqrt and in FMA contraction) agree in the weight to

        |dw| / w  <~  eps * A,      A = sum over decaying nodes (1 + gamma_node^2) * sum_k a_k / (a_k - b_k - c_k)

and no better.  Two more terms of the same kind:

* a resonance mass sampled in the kernel comes out of a different inverse-CDF algorithm than the oracle's (CUDA
  normcdfinv / a table of long-double nodes and a series for the relativistic Breit-Wigner, against scipy and libm); the
  two masses agree to ``MASS_NOISE`` in the bulk (tests/test_samplers_gpu.py), in the far tails only to what an fp64
  evaluation of the CDF allows there (``sampler_noise``: dx = dG / G'(x), the density G' vanishes), and that
  difference is amplified by the same A;
* the reference boosts with gamma = 1 / sqrt(1 - beta^2) (kinematics.py:141), which itself loses eps * gamma^2 for a fast
  subsystem (1 - beta^2 cancels); the kernel's fast variant boosts with gamma = E / M and is closer to the exact result
  than the reference, so boosted 4-VECTORS (not the weights, which are frame independent) differ by the reference's
  own eps * G, G = sum of gamma^2 over the boosts applied to the event (inside each node and by each parent).

The parity tests hold every event to max(1e-12, c * A_i) -- weights relatively, 4-vectors on the scale of the event's own
energy -- instead of allowing a blanket fraction of events a blanket looser tolerance.
"""

import numpy as np

EPS = 2.0 ** -53
MASS_NOISE = 2e-13  # relative agreement of in-kernel sampled resonance masses with the oracle's


def _mass(v):
    return np.sqrt(np.maximum(v[:, 3] ** 2 - (v[:, 0] ** 2 + v[:, 1] ** 2 + v[:, 2] ** 2), 0.0))


def amplification(desc, particles, boost=None, with_boosts=False):
    """``A`` per event for the decay tree ``desc`` (tests/helpers/decays.py description) whose reference output
    4-vectors are ``particles`` (name -> (N, 4) numpy arrays).  ``boost``: the ``boost_to`` rows, if any.
    ``with_boosts``: also return ``G`` = sum of gamma^2 over every Lorentz boost the reference applies to the event (the
    n - 2 boosts inside each node, phasespace.py:484-493, and each node's final boost, 389-391)."""
    n_events = next(iter(particles.values())).shape[0]
    total = np.zeros(n_events)
    boosts = np.zeros(n_events)

    def node(d, gamma2):
        _, _, _, _, children = d
        if not children:
            return
        acc = np.zeros((n_events, 4))
        inv_prev = None
        steps = np.zeros(n_events)
        for k, child in enumerate(children):
            vec = particles[child[0]]
            acc = acc + vec
            inv = _mass(acc)
            if k > 0:
                a, b, c = inv, inv_prev, _mass(vec)
                steps += a / np.maximum(a - b - c, 1e-300)
                if k > 1:  # the subsystem of particles 0..k-1 (mass b) was boosted into the frame of 0..k (mass a)
                    boosts[:] += ((a * a + b * b - c * c) / np.maximum(2.0 * a * b, 1e-300)) ** 2
            inv_prev = inv
        total[:] += (1.0 + gamma2) * steps
        boosts[:] += gamma2
        for child in children:
            if child[4]:
                vec = particles[child[0]]
                m = np.maximum(_mass(vec), 1e-300)
                node(child, (vec[:, 3] / m) ** 2)

    if boost is not None:
        b = np.broadcast_to(np.asarray(boost, dtype=np.float64).reshape(-1, 4), (n_events, 4))
        g2 = (b[:, 3] / np.maximum(_mass(b), 1e-300)) ** 2
    else:
        g2 = np.zeros(n_events)
    node(desc, g2)
    return (total, boosts) if with_boosts else total


def sampler_noise(desc, particles):
    """Relative noise of the sampled resonance masses that no fp64 inverse CDF can avoid, per event, summed over the
    sampled resonances of the tree: the CDF value is known to ~4 eps (a sum of arctangents of order pi), and the inverse
    maps that to dx = dG / G'(x), which explodes in the far tails where the density G' vanishes.

        relbw:  G'(z) = (kappa^2 + 4) / ((z^2 + 1)((z + kappa)^2 + 1)) per unit of z = (x - p) / q
        bw:     x = m + w tan(t):  dx = w (1 + z^2) dt

        gauss:  the fp64 expression Phi(a) + u (Phi(b) - Phi(a)) (oracle, and the kernel's library path) is rounded on
                the scale of its own value p: dz = eps p / phi(z).  Kernel and oracle both work in the tail nearest to
                the lower limit (mirrored when the lower limit lies above the pole: lo = the sum of the fixed masses
                below the resonance, event independent), where p / phi -> 1 / |z| stays small; towards the OTHER end
                p -> 1 and the digits run out (z = 5: dz ~ 1e-10).  The kernel's node table for event-independent limits
                (ps_samplers.cuh sample_gauss_table) does not lose them, the oracle it is compared with does.

    The mass of each resonance is read off its reference 4-vector."""
    n_events = next(iter(particles.values())).shape[0]
    noise = np.zeros(n_events)

    def min_mass(d):  # recurse_stable (phasespace.py:326-333)
        return sum(c[1] if c[2] == "fixed" else min_mass(c) for c in d[4])

    def walk(d):
        name, mass, kind, width, children = d
        if kind == "gauss" and name in particles:
            from scipy.special import ndtr

            x = np.maximum(_mass(particles[name]), 1e-300)
            z = (x - mass) / width
            if min_mass(d) > mass:
                z = -z
            zc = np.minimum(np.abs(z), 37.0) * np.sign(z)
            noise[:] += 4.0 * EPS * ndtr(zc) * np.sqrt(2.0 * np.pi) * np.exp(0.5 * zc * zc) * width / x
        if kind in ("relbw", "bw") and name in particles:
            x = np.maximum(_mass(particles[name]), 1e-300)
            if kind == "relbw":
                a, b = mass * mass, mass * width
                rho = np.sqrt(a * a + b * b)
                p = np.sqrt(0.5 * (rho + a))
                q = b / (2.0 * p)
                kappa = 2.0 * p / q
                z = (x - p) / q
                noise[:] += 4.0 * EPS * (z * z + 1.0) * ((z + kappa) ** 2 + 1.0) / (kappa * kappa + 4.0) * q / x
            else:
                z = (x - mass) / width
                noise[:] += 4.0 * EPS * (z * z + 1.0) * width / x
        for c in children:
            walk(c)

    walk(desc)
    return noise


def has_sampled_resonance(desc):
    return desc[2] in ("gauss", "bw", "relbw") or any(has_sampled_resonance(c) for c in desc[4])


def top_gamma2(n_events, boost):
    if boost is None:
        return np.zeros(n_events)
    b = np.broadcast_to(np.asarray(boost, dtype=np.float64).reshape(-1, 4), (n_events, 4))
    return (b[:, 3] / np.maximum(_mass(b), 1e-300)) ** 2


def event_bounds(desc, particles, boost=None, tol=1e-12, c_amp=8.0):
    """``(weight_bound, vector_bound)`` per event: relative for the weights, on the scale of the event's energy (the top
    mass at rest) for the 4-vectors."""
    amp, boosts = amplification(desc, particles, boost, with_boosts=True)
    unit = max(c_amp * EPS, MASS_NOISE if has_sampled_resonance(desc) else 0.0)
    if has_sampled_resonance(desc):  # plus what the inverse CDF's conditioning adds for this event's own masses
        unit = unit + c_amp * sampler_noise(desc, particles)
    w_bound = np.maximum(tol, unit * amp)
    v_bound = np.maximum(tol, unit * amp + c_amp * EPS * boosts)
    return amp, w_bound, v_bound
