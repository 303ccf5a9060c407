"""CPU oracle: numpy restatement of the reference's GENBOD event generator.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product path (phasespace_b200) never does.

PARITY STATUS: PINNED against the reference's own source code.  The reference (/root/reference/src/phasespace)
needs TensorFlow >= 2.19 + tensorflow_probability, neither installed nor installable offline; but the only TensorFlow
surface its generator uses is ``tensorflow.experimental.numpy`` (a NumPy-compatible API), ``tf.function`` and
``tf.random.Generator.uniform``.  tests/golden/make_reference_golden.py installs a stand-in ``tensorflow`` module that
maps those to NumPy and to a table-driven generator, imports the reference UNMODIFIED and runs its
``GenParticle.generate`` on fixed uniform tables for 22 trees (flat 2/3/4/5/8/9/10 bodies, massless products, per-event
and one-row ``boost_to``, K* gamma and K1 gamma chains with fixed and callable resonance masses, two decaying branches
side by side, a nine-body decay below the top particle, three levels of three-body decays); the outputs are committed
as tests/golden/reference_vectors.npz.  This oracle, fed the same tables, reproduces every weight, maximum weight and
4-vector of those runs BIT FOR BIT (tests/test_oracle.py::test_oracle_matches_reference_source holds it to 1e-13 so that
a different NumPy build may differ in the last bit of sin/cos) -- for the three nine-body cases once the one reduction
whose association the reference leaves to its backend, ``tnp.sum(masses, axis=1)``, is taken in NumPy's pairwise order
(``ROW_SUM_ORDER`` below; left to right, the default and the kernel's order, differs there in the last bit of the mass
sum).  It is also pinned by
  * the six events printed in the reference's README (README.rst:156-183), which are real outputs of
    ``generate()`` under TensorFlow for B0 -> K*(-> K+ pi-) gamma: tests/test_oracle.py recovers the four draws of
    each event from the printed K* and K+ vectors, feeds them to this oracle and gets all four printed
    4-vectors back to the printed precision (tests/golden/readme_kstargamma.json),
  * the closed-form two-body energies of the same README block (E(K*)=2715.78804872, E(gamma)=2563.79195128),
  * the invariants the reference's own tests assert (shapes, w/w_max < 1, key order, error types),
  * 4-momentum conservation / on-shell masses, which hold by construction of the algorithm,
  * an independent scalar C implementation (oracle/genbod_ref.c) of the same published GENBOD
    algorithm (F. James, CERN 68-15), through the reference's own KS procedure.
What stays UNPINNED is the arithmetic the reference delegates to third parties: the TensorFlow Philox
stream and the tfp / zfit / zfit_physics mass samplers (upstream tests check their shapes only).

Every function cites the reference lines it follows ("ps.py" = src/phasespace/phasespace.py,
"kin.py" = src/phasespace/kinematics.py).  Random numbers are an *argument* (a
``(n_events, n_draws)`` table in the tree-wide draw order below) so that the oracle stands in
for "the reference fed these uniform draws".

Draw order (ps.py:344-353, 367, 444-451, 555-567): depth-first over decaying particles, a node
completely before any of its children, siblings in declaration order; inside one node:
  [one column per resonant child, in child order]  ->  u[0..n-3]  ->  (a_1, b_1) ... (a_{n-1}, b_{n-1})
The reference's own mass callables draw from TF/tfp/zfit generators (third party, unpinned); the
built-in samplers here are the inverse-CDF restatements documented in ``sample_mass``.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

try:  # scipy is test-side only
    from scipy.special import ndtr, ndtri
except Exception:  # pragma: no cover
    ndtr = ndtri = None

FIXED, GAUSS, BW, RELBW, USER = 0, 1, 2, 3, 4
KIND_NAMES = {"gauss": GAUSS, "bw": BW, "relbw": RELBW}


@dataclass
class Part:
    """Decay-tree node: the oracle's stand-in for ``GenParticle`` (ps.py:74-226)."""

    name: str
    mass: float = 0.0  # fixed mass, or the pole mass of a resonance
    kind: int = FIXED
    width: float = 0.0
    children: list = field(default_factory=list)

    @property
    def has_fixed_mass(self):
        return self.kind == FIXED

    @property
    def has_children(self):
        return bool(self.children)

    @property
    def has_grandchildren(self):  # ps.py:228-233
        return any(c.has_children for c in self.children)


def flat(mass_top, masses, names=None, top_name="top"):
    """``nbody_decay`` (ps.py:778-806)."""
    names = names or [f"p_{i}" for i in range(len(masses))]
    return Part(top_name, mass_top, children=[Part(n, m) for n, m in zip(names, masses)])


def n_draws(part: Part) -> int:
    """Number of uniform columns the whole tree consumes per event."""
    if not part.children:
        return 0
    n = len(part.children)
    total = sum(1 for c in part.children if c.kind in (GAUSS, BW, RELBW)) + (n - 2) + 2 * (n - 1)
    return total + sum(n_draws(c) for c in part.children)


def output_names(part: Part) -> list:
    """Key order of the output dict (ps.py:547-570)."""
    out = [c.name for c in part.children]
    for c in part.children:
        if c.children:
            out.extend(output_names(c))
    return out


# ----------------------------------------------------------------------------- kinematics
def pdk(a, b, c):
    """Two-body break-up momentum, CERN 68-15 eq. 9.17 (ps.py:56-71, same factor order)."""
    x = (a - b - c) * (a + b + c) * (a - b + c) * (a + b - c)
    return np.sqrt(x) / (2.0 * a)


def kin_mass(vector):
    """kin.py:93-104 with the metric (-1,-1,-1,+1) summed left to right; ``(N,4) -> (N,1)``."""
    sq = vector * vector
    s = -sq[:, 0:1]
    s = s + -sq[:, 1:2]
    s = s + -sq[:, 2:3]
    s = s + sq[:, 3:4]
    return np.sqrt(s)


def boost_components(vector):
    """kin.py:166-175: ``p / E`` -> ``(N,3)``."""
    return vector[:, 0:3] / vector[:, 3:4]


def lorentz_boost(vector, boost):
    """kin.py:122-149, including the reference's *global* zero-boost test (kin.py:147-148)."""
    boost = np.asarray(boost)[:, 0:3]
    b2 = boost[:, 0:1] * boost[:, 0:1]
    b2 = b2 + boost[:, 1:2] * boost[:, 1:2]
    b2 = b2 + boost[:, 2:3] * boost[:, 2:3]
    if np.all(b2 == 0.0):
        return vector
    with np.errstate(divide="ignore", invalid="ignore"):
        gamma = 1.0 / np.sqrt(1.0 - b2)
        gamma2 = (gamma - 1.0) / b2
    ve = vector[:, 3:4]
    vp = vector[:, 0:3]
    bp = vp[:, 0:1] * boost[:, 0:1]
    bp = bp + vp[:, 1:2] * boost[:, 1:2]
    bp = bp + vp[:, 2:3] * boost[:, 2:3]
    vp2 = vp + (gamma2 * bp + gamma * ve) * boost
    ve2 = gamma * (ve + bp)
    return np.concatenate([vp2, ve2], axis=1)


def get_w_max(available_mass, masses):
    """GENBOD maximum weight (ps.py:288-298). ``(N,1),(N,n) -> (N,1)``."""
    emmax = available_mass + masses[:, 0:1]
    emmin = np.zeros_like(emmax)
    w_max = np.ones_like(emmax)
    for i in range(1, masses.shape[1]):
        emmin = emmin + masses[:, i - 1 : i]
        emmax = emmax + masses[:, i : i + 1]
        w_max = w_max * pdk(emmax, emmin, masses[:, i : i + 1])
    return w_max


# ----------------------------------------------------------------------------- resonance masses
def relbw_cdf(x, m, gamma):
    """Antiderivative of ``1/((x^2-m^2)^2 + m^2 gamma^2)`` (the zfit_physics
    RelativisticBreitWigner shape as a density in the mass x, fromdecay/mass_functions.py:89-100).

    With a = m^2, b = m*gamma, rho = sqrt(a^2+b^2), p = sqrt((rho+a)/2), q = sqrt((rho-a)/2) the
    quartic factors as (x^2-2px+rho)(x^2+2px+rho) and partial fractions give
        F(x) = [atan((x-p)/q) + atan((x+p)/q)] / (4 rho q) + ln((x^2+2px+rho)/(x^2-2px+rho)) / (8 rho p).
    """
    a, b = m * m, m * gamma
    rho = math.sqrt(a * a + b * b)
    p = math.sqrt(0.5 * (rho + a))
    q = b / (2.0 * p)  # = sqrt((rho - a) / 2) without the cancellation for narrow resonances
    c_atan, c_log = 1.0 / (4.0 * rho * q), 1.0 / (8.0 * rho * p)
    x2r = x * x + rho
    tpx = 2.0 * p * x
    return c_atan * (np.arctan((x - p) / q) + np.arctan((x + p) / q)) + c_log * np.log((x2r + tpx) / (x2r - tpx))


RELBW_MAX_ITERS = 40
RELBW_LAST_STEP = 1e-9


def _relbw_consts(m, gamma):
    a, b = m * m, m * gamma
    rho = math.sqrt(a * a + b * b)
    p = math.sqrt(0.5 * (rho + a))
    q = b / (2.0 * p)  # 2 p q = b: stable for width / mass down to anything representable
    return p, q, 1.0 / (4.0 * rho * q), 1.0 / (8.0 * rho * p)


def _relbw_cdf_t(consts, t, x, tt):
    """``relbw_cdf`` written in the Newton variable: atan((x-p)/q) = t and (x-p)^2 + q^2 = q^2 (1 + tan^2 t)."""
    p, q, c_atan, c_log = consts
    big_a = (x + p) * (x + p) + q * q
    big_b = q * q * (1.0 + tt * tt)
    return c_atan * (t + np.arctan((x + p) / q)) + c_log * np.log(big_a / big_b)


def sample_mass(kind, m, width, lo, hi, u):
    """One mass per event inside the per-event limits ``[lo, hi]`` from one uniform ``u``.

    Upstream these are zfit / zfit_physics / tfp samplers (mass_functions.py:31-38,58-65,89-100;
    tests/helpers/decays.py:26-37) whose arithmetic is third party and *unpinned*; the shapes are
    restated from their documented densities and sampled by inverse CDF:
      GAUSS  truncated normal(mu=m, sigma=width)  (zfit Gauss / tfp TruncatedNormal)
      BW     Cauchy  ~ 1/(1+((x-m)/width)^2)     (zfit Cauchy, gamma=width)
      RELBW  ~ 1/((x^2-m^2)^2 + m^2 width^2)      (zfit_physics RelativisticBreitWigner)
    The kernel (phasespace_b200/csrc/ps_samplers.cuh) solves the same equations
    ``x = F^-1(F(lo) + u (F(hi) - F(lo)))`` with solvers of its own.
    """
    lo, hi, u = (np.asarray(v, dtype=np.float64) for v in (lo, hi, u))
    if kind == GAUSS:
        alpha, beta = (lo - m) / width, (hi - m) / width
        # work in the tail that keeps precision: lower tail if alpha < 0 else mirrored
        flip = alpha > 0.0
        a_ = np.where(flip, -beta, alpha)
        b_ = np.where(flip, -alpha, beta)
        u_ = np.where(flip, 1.0 - u, u)
        ca, cb = ndtr(a_), ndtr(b_)
        z = ndtri(ca + u_ * (cb - ca))
        z = np.where(flip, -z, z)
        x = m + width * z
    elif kind == BW:
        ta, tb = np.arctan((lo - m) / width), np.arctan((hi - m) / width)
        x = m + width * np.tan(ta + u * (tb - ta))
    elif kind == RELBW:
        consts = _relbw_consts(m, width)
        p, q = consts[0], consts[1]
        t_lo, t_hi = np.arctan((lo - p) / q), np.arctan((hi - p) / q)
        f_lo = _relbw_cdf_t(consts, t_lo, lo, (lo - p) / q)
        f_hi = _relbw_cdf_t(consts, t_hi, hi, (hi - p) / q)
        target = f_lo + u * (f_hi - f_lo)
        t = t_lo + u * (t_hi - t_lo)
        # Newton in t = atan((x-p)/q), in which the CDF is nearly linear, run to full convergence: every event
        # takes one more step after its first step below RELBW_LAST_STEP (the error left after a step s is
        # ~C s^2 with C up to a few hundred in the far tail, so that last step lands on the rounding floor).
        active = np.ones(np.shape(t), dtype=bool)
        last = np.zeros(np.shape(t), dtype=bool)
        for _ in range(RELBW_MAX_ITERS):
            tt = np.tan(t)
            x = p + q * tt
            g = _relbw_cdf_t(consts, t, x, tt) - target
            step = g * q * ((x + p) * (x + p) + q * q)
            t = np.where(active, np.minimum(np.maximum(t - step, t_lo), t_hi), t)
            active = active & ~last
            last = last | (np.abs(step) < RELBW_LAST_STEP)
            if not active.any():
                break
        x = p + q * np.tan(t)
    else:
        raise ValueError(f"unknown sampler kind {kind}")
    return np.minimum(np.maximum(x, lo), hi)


def recurse_stable(part: Part):
    """Minimum mass of a resonance = sum of the fixed masses below it (ps.py:326-333)."""
    out = 0
    for child in part.children:
        out = out + (child.mass if child.has_fixed_mass else recurse_stable(child))
    return out


# ----------------------------------------------------------------------------- one decay node
# Association of ``tnp.sum(masses, axis=1, keepdims=True)`` (ps.py:357) -- the one reduction of the path whose order the
# reference leaves to its backend.  "left": m0 + m1 + ... left to right (what the CUDA kernel does).
# "numpy": the expression itself on the NumPy API, i.e. NumPy's pairwise summation along the contiguous axis, which for
# eight or more children adds eight strided partial sums pairwise -- what the reference's source computes when it runs
# on NumPy (tests/golden/make_reference_golden.py).  With "numpy" the oracle reproduces those runs bit for bit in all
# 22 cases; with "left" 19 are bit-identical and the three nine-body cases differ by the last bit of the mass sum,
# amplified into at most 2.7e-12 of the weight next to a threshold (tests/test_oracle.py).  TensorFlow's own order
# (Eigen / XLA packet reductions) is a third one; all are within the north star's 1e-12 after conditioning.
ROW_SUM_ORDER = "left"


def row_sum(masses):
    """``tnp.sum(masses, axis=1, keepdims=True)`` (ps.py:357) in the association ``ROW_SUM_ORDER`` names."""
    if ROW_SUM_ORDER == "numpy":
        return np.sum(np.ascontiguousarray(masses), axis=1, keepdims=True)
    msum = masses[:, 0:1]
    for i in range(1, masses.shape[1]):
        msum = msum + masses[:, i : i + 1]
    return msum


def generate_node(part: Part, momentum, n_events, draws, col, user_masses=None):
    """``GenParticle._generate`` + ``_generate_part2`` (ps.py:300-397, 399-495).

    ``draws[:, col:]`` supplies this node's uniforms; returns the updated column.
    """
    if not part.children:
        raise ValueError("No children have been configured")
    p_top = np.asarray(momentum, dtype=np.float64).reshape(-1, 4)
    top_mass = np.broadcast_to(kin_mass(p_top), (n_events, 1))  # ps.py:322
    n = len(part.children)
    fixed = [c.mass for c in part.children if c.has_fixed_mass]
    mass_from_stable = 0.0
    for f in fixed:  # tnp.sum over the list, left to right (ps.py:335-341)
        mass_from_stable = mass_from_stable + f
    max_mass = top_mass - mass_from_stable  # ps.py:342
    masses = []
    for child in part.children:  # ps.py:344-353
        if child.has_fixed_mass:
            masses.append(np.full((n_events, 1), float(child.mass)))
        else:
            min_mass = np.full((n_events, 1), float(recurse_stable(child)))
            if child.kind == USER:
                mass = np.asarray(user_masses[child.name], dtype=np.float64).reshape(n_events, 1)
            else:
                mass = sample_mass(child.kind, child.mass, child.width, min_mass, max_mass, draws[:, col : col + 1])
                col += 1
            max_mass = max_mass - mass
            masses.append(mass)
    masses = np.concatenate(masses, axis=1)
    msum = row_sum(masses)  # tnp.sum(masses, axis=1, keepdims=True) (ps.py:357), left to right unless ROW_SUM_ORDER says otherwise
    available_mass = top_mass - msum
    if not np.all(available_mass >= 0.0):  # ps.py:358-363 (NaN fails this too)
        raise FloatingPointError("Forbidden decay")
    w_max = get_w_max(available_mass, masses)  # ps.py:364
    p_top_boost = boost_components(np.broadcast_to(p_top, (n_events, 4)))  # ps.py:365
    rnd = np.sort(draws[:, col : col + n - 2], axis=1)  # ps.py:367-375
    col += n - 2
    rnd = np.concatenate([np.zeros((n_events, 1)), rnd, np.ones((n_events, 1))], axis=1)
    sum_ = np.zeros((n_events, 1))
    inv = []
    for i in range(n):  # ps.py:379-384
        sum_ = sum_ + masses[:, i : i + 1]
        inv.append(rnd[:, i : i + 1] * available_mass + sum_)
    # ---- _generate_part2 (ps.py:399-495)
    pds = [pdk(inv[i + 1], inv[i], masses[:, i + 1 : i + 2]) for i in range(n - 1)]
    weights = pds[0]
    for i in range(1, n - 1):  # tnp.prod over the list axis (ps.py:412)
        weights = weights * pds[i]
    zero = np.zeros_like(pds[0])
    parts = [np.concatenate([zero, pds[0], zero, np.sqrt(pds[0] * pds[0] + masses[:, 0:1] * masses[:, 0:1])], axis=1)]
    k = 1
    while True:
        mk = masses[:, k : k + 1]
        parts.append(np.concatenate([zero, -pds[k - 1], zero, np.sqrt(pds[k - 1] * pds[k - 1] + mk * mk)], axis=1))
        cos_z = 2.0 * draws[:, col : col + 1] - 1.0  # ps.py:444-447
        sin_z = np.sqrt(1.0 - cos_z * cos_z)
        ang_y = 2.0 * math.pi * draws[:, col + 1 : col + 2]  # ps.py:448-454
        col += 2
        cos_y, sin_y = np.cos(ang_y), np.sin(ang_y)
        for j in range(k + 1):  # ps.py:456-481
            px, py = parts[j][:, 0:1], parts[j][:, 1:2]
            parts[j] = np.concatenate(
                [cos_z * px - sin_z * py, sin_z * px + cos_z * py, parts[j][:, 2:3], parts[j][:, 3:4]], axis=1
            )
            px, pz = parts[j][:, 0:1], parts[j][:, 2:3]
            parts[j] = np.concatenate(
                [cos_y * px - sin_y * pz, parts[j][:, 1:2], sin_y * px + cos_y * pz, parts[j][:, 3:4]], axis=1
            )
        if k == n - 1:
            break
        betas = pds[k] / np.sqrt(pds[k] * pds[k] + inv[k] * inv[k])  # ps.py:484-486
        bvec = np.concatenate([zero, betas, zero], axis=1)
        parts = [lorentz_boost(p, bvec) for p in parts]  # ps.py:487-493
        k += 1
    parts = [lorentz_boost(p, p_top_boost) for p in parts]  # ps.py:389-391
    return weights.reshape(n_events), w_max.reshape(n_events), parts, masses, col


# ----------------------------------------------------------------------------- decay tree
def recursive_generate(part: Part, n_events, draws, col=0, boost_to=None, recalculate_max_weights=False,
                       user_masses=None):
    """``GenParticle._recursive_generate`` (ps.py:497-626)."""
    if boost_to is not None:
        momentum = np.asarray(boost_to, dtype=np.float64).reshape(-1, 4)
    else:
        if not part.has_fixed_mass:
            raise ValueError("Cannot use resonance as top particle")
        momentum = np.broadcast_to(np.array([0.0, 0.0, 0.0, float(part.mass)]), (n_events, 4))
    weights, weights_max, parts, child_masses, col = generate_node(part, momentum, n_events, draws, col, user_masses)
    out_parts = {c.name: parts[i] for i, c in enumerate(part.children)}
    out_masses = {c.name: child_masses[:, i : i + 1] for i, c in enumerate(part.children)}
    for i, child in enumerate(part.children):  # ps.py:555-570
        if child.has_children:
            cw, _, cparts, cmasses, col = recursive_generate(
                child, n_events, draws, col, boost_to=parts[i], recalculate_max_weights=False, user_masses=user_masses
            )
            weights = weights * cw
            out_parts.update(cparts)
            out_masses.update(cmasses)
    if recalculate_max_weights:  # ps.py:571-625

        def build_mass_tree(p, leaf):
            if p.has_children:
                leaf[p.name] = {}
                for ch in p.children:
                    build_mass_tree(ch, leaf[p.name])
            else:
                leaf[p.name] = out_masses[p.name]

        def flattened(d):
            out = []
            for val in d.values():
                out.extend(flattened(val) if isinstance(val, dict) else [val])
            return out

        def recurse_w_max(parent_mass, tree):
            available_mass = parent_mass - sum(flattened(tree))
            masses = []
            w_max = np.ones_like(available_mass)
            for child, child_mass in tree.items():
                if isinstance(child_mass, dict):
                    others = {k: v for k, v in tree.items() if k != child}
                    w_max = w_max * recurse_w_max(parent_mass - sum(flattened(others)), child_mass)
                    masses.append(sum(flattened(child_mass)))
                else:
                    masses.append(child_mass)
            masses = np.concatenate([np.broadcast_to(m, available_mass.shape) for m in masses], axis=1)
            return w_max * get_w_max(available_mass, masses)

        mass_tree = {}
        build_mass_tree(part, mass_tree)
        wm = recurse_w_max(kin_mass(np.asarray(momentum, dtype=np.float64).reshape(-1, 4)), mass_tree[part.name])
        weights_max = np.broadcast_to(wm, (n_events, 1)).reshape(n_events).copy()
    return weights, weights_max, out_parts, out_masses, col


def generate(part: Part, n_events, draws, boost_to=None, normalize_weights=True, user_masses=None):
    """``GenParticle.generate`` (ps.py:628-717) with the uniform table as an argument."""
    draws = np.asarray(draws, dtype=np.float64).reshape(n_events, -1)
    if boost_to is not None:
        boost_to = np.asarray(boost_to, dtype=np.float64)
        if boost_to.shape[0] not in (n_events, 1):  # ps.py:697-703
            raise ValueError(f"The number of events requested ({n_events}) doesn't match the boost_to input size")
    weights, weights_max, parts, _, col = recursive_generate(
        part, n_events, draws, 0, boost_to=boost_to, recalculate_max_weights=part.has_grandchildren,
        user_masses=user_masses,
    )
    assert col == n_draws(part), (col, n_draws(part))
    if normalize_weights:
        return weights / weights_max, parts
    return weights, weights_max, parts


# ----------------------------------------------------------------------------- FONLL-style boosts
def generate_fonll(mass, pt_centers, pt_values, eta_centers, eta_values, draws, pt_scale=1000.0):
    """``tests/helpers/rapidsim.py:23-51`` (``generate_fonll``) with the random numbers as an argument.

    ``draws`` is ``(n_events, 3)``: the uniforms behind ``np.random.choice`` for pT, for eta, and the one scaled to
    phi.  ``np.random.choice(a, p=p)`` is ``a[cumsum(p).searchsorted(u, side="right")]`` with the cumulative sum
    renormalised to end at 1.  Returns ``(n_events, 4)`` rows ``[px, py, pz, E]`` (the reference stacks them as
    ``(4, n_events)`` and transposes at the call site, tests/test_physics.py:190-191).
    """
    def choice(centers, values, u):
        p = np.asarray(values, dtype=np.float64) / np.sum(values)
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        idx = np.minimum(cdf.searchsorted(u, side="right"), len(cdf) - 1)
        return np.asarray(centers, dtype=np.float64)[idx]

    draws = np.asarray(draws, dtype=np.float64)
    pt_rand = pt_scale * np.abs(choice(pt_centers, pt_values, draws[:, 0]))
    eta_rand = choice(eta_centers, eta_values, draws[:, 1])
    phi_rand = draws[:, 2] * (2 * np.pi)
    px = pt_rand * np.cos(phi_rand)
    py = pt_rand * np.sin(phi_rand)
    pz = pt_rand * np.sinh(eta_rand)
    e = np.sqrt(px * px + py * py + pz * pz + mass * mass)
    return np.stack([px, py, pz, e], axis=1)
