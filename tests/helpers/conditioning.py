"""How ill-conditioned is an event?  A per-event amplification factor computed from the REFERENCE outputs alone.

The weight of an event is a product of two-body momenta pdk(a, b, c) = sqrt((a-b-c)(a+b+c)(a-b+c)(a+b-c)) / 2a
(phasespace.py:56-71).  Near a threshold the factor (a - b - c) cancels: a rounding error eps * a in the invariant
masses becomes a relative error eps * a / (a - b - c) of the weight.  A decaying particle that is not the top particle
has its mass re-derived as sqrt(E^2 - p^2) of its *boosted* 4-vector (phasespace.py:322 via 564), which adds a relative
error eps * gamma^2 (gamma = E / m in the output frame) to its a.  So two correct fp64 evaluations of the same event
(the reference's and the kernel's differ in the last bits of sin/cos/sqrt and in FMA contraction) agree in the weight to

        |dw| / w  <~  eps * A,      A = sum over decaying nodes (1 + gamma_node^2) * sum_k a_k / (a_k - b_k - c_k)

and no better.  The parity tests hold every event to max(1e-12, C * eps * A) instead of allowing a blanket fraction of
events a blanket looser tolerance.
"""

import numpy as np

EPS = 2.0 ** -53


def _mass(v):
    return np.sqrt(np.maximum(v[:, 3] ** 2 - (v[:, 0] ** 2 + v[:, 1] ** 2 + v[:, 2] ** 2), 0.0))


def amplification(desc, particles, boost=None, top_mass=None):
    """``A`` per event for the decay tree ``desc`` (tests/helpers/decays.py description) whose reference output
    4-vectors are ``particles`` (name -> (N, 4) numpy arrays).  ``boost``: the ``boost_to`` rows, if any."""
    n_events = next(iter(particles.values())).shape[0]
    total = np.zeros(n_events)

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
            inv_prev = inv
        total[:] += (1.0 + gamma2) * steps
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
    return total


def weight_tolerance(desc, particles, boost=None, tol=1e-12, c=8.0):
    """Per-event relative tolerance for the weights: ``max(tol, c * eps * A)``."""
    return np.maximum(tol, c * EPS * amplification(desc, particles, boost))
