// Resonance mass samplers: one mass inside per-event limits [lo, hi] from ONE uniform (inverse CDF).
//
// Upstream these are Python callables built on third-party libraries (zfit Gauss / Cauchy,
// zfit_physics RelativisticBreitWigner: fromdecay/mass_functions.py:16-102; tfp TruncatedNormal:
// tests/helpers/decays.py:26-37) sampled one value per event with per-event limits.  Their
// arithmetic is not in the reference tree; the densities are restated here and sampled by inverse
// CDF so that a resonance costs exactly one draw (deterministic draw count, no rejection loop, no
// warp divergence).  oracle/genbod_numpy.py::sample_mass is the CPU statement of the same formulas.
#ifndef PS_SAMPLERS_CUH_
#define PS_SAMPLERS_CUH_

#include "ps_params.h"

namespace psk {

__device__ __forceinline__ double clamp_mass(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// truncated normal(mu = m, sigma = w) on [lo, hi]
__device__ __forceinline__ double sample_gauss(double m, double w, double lo, double hi, double u) {
    const double alpha = (lo - m) / w, beta = (hi - m) / w;
    const bool flip = alpha > 0.0;  // work in the tail that keeps precision
    const double a_ = flip ? -beta : alpha, b_ = flip ? -alpha : beta, u_ = flip ? 1.0 - u : u;
    const double ca = normcdf(a_), cb = normcdf(b_);
    double z = normcdfinv(ca + u_ * (cb - ca));
    z = flip ? -z : z;
    return clamp_mass(m + w * z, lo, hi);
}

// Cauchy ~ 1 / (1 + ((x-m)/w)^2) on [lo, hi]
__device__ __forceinline__ double sample_cauchy(double m, double w, double lo, double hi, double u) {
    const double ta = atan((lo - m) / w), tb = atan((hi - m) / w);
    return clamp_mass(m + w * tan(ta + u * (tb - ta)), lo, hi);
}

struct RelBwConst {
    double a, b, rho, p, q, c_atan, c_log;
};

__device__ __forceinline__ RelBwConst relbw_const(double m, double w) {
    RelBwConst k;
    k.a = m * m;
    k.b = m * w;
    k.rho = sqrt(k.a * k.a + k.b * k.b);
    k.p = sqrt(0.5 * (k.rho + k.a));
    k.q = sqrt(0.5 * (k.rho - k.a));
    k.c_atan = 1.0 / (4.0 * k.rho * k.q);
    k.c_log = 1.0 / (8.0 * k.rho * k.p);
    return k;
}

// The antiderivative of 1/((x^2-m^2)^2 + m^2 w^2): the quartic factors as (x^2-2px+rho)(x^2+2px+rho) and
// partial fractions give  F(x) = [atan((x-p)/q) + atan((x+p)/q)] / (4 rho q) + ln((x^2+2px+rho)/(x^2-2px+rho)) / (8 rho p).
#define PS_RELBW_MAX_ITERS 8
#define PS_RELBW_LAST_STEP 1e-7

// ~ 1/((x^2-m^2)^2 + m^2 w^2) on [lo, hi].  Newton in t = atan((x-p)/q), in which the CDF is nearly linear
// (its dominant term is t / (4 rho q)).  With x = p + q tan t the first arctangent of the CDF is t itself,
// (x-p)^2 + q^2 = q^2 sec^2 t, and the Newton step  g * D(x) / (q sec^2 t)  collapses to  g * q * A  with
// A = (x+p)^2 + q^2: one tan, one atan, one log per iteration.  Convergence is quadratic (measured: the error
// left after a step s is ~0.03 s^2), so the iteration stops right after applying a step below 1e-7 rad
// (typically the third), or after PS_RELBW_MAX_ITERS.
__device__ __forceinline__ double relbw_cdf_t(const RelBwConst& k, double t, double x, double tt) {
    const double A = (x + k.p) * (x + k.p) + k.q * k.q;
    const double B = k.q * k.q * (1.0 + tt * tt);
    return k.c_atan * (t + atan((x + k.p) / k.q)) + k.c_log * log(A / B);
}
__device__ __forceinline__ double sample_relbw(double m, double w, double lo, double hi, double u) {
    const RelBwConst k = relbw_const(m, w);
    const double t_lo = atan((lo - k.p) / k.q), t_hi = atan((hi - k.p) / k.q);
    const double f_lo = relbw_cdf_t(k, t_lo, lo, (lo - k.p) / k.q), f_hi = relbw_cdf_t(k, t_hi, hi, (hi - k.p) / k.q);
    const double target = f_lo + u * (f_hi - f_lo);
    double t = t_lo + u * (t_hi - t_lo);
#pragma unroll 1
    for (int it = 0; it < PS_RELBW_MAX_ITERS; ++it) {
        const double tt = tan(t);
        const double x = k.p + k.q * tt;
        const double g = relbw_cdf_t(k, t, x, tt) - target;
        const double A = (x + k.p) * (x + k.p) + k.q * k.q;
        const double step = g * k.q * A;
        t = fmin(fmax(t - step, t_lo), t_hi);
        if (fabs(step) < PS_RELBW_LAST_STEP) break;  // the error left after a step s is ~0.03 s^2
    }
    return clamp_mass(k.p + k.q * tan(t), lo, hi);
}

__device__ __forceinline__ double sample_mass(int kind, double m, double w, double lo, double hi, double u) {
    switch (kind) {
        case PS_KIND_GAUSS: return sample_gauss(m, w, lo, hi, u);
        case PS_KIND_BW: return sample_cauchy(m, w, lo, hi, u);
        default: return sample_relbw(m, w, lo, hi, u);
    }
}

}  // namespace psk
#endif
