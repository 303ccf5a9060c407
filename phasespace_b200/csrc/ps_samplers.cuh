// Resonance mass samplers: one mass inside per-event limits [lo, hi] from ONE uniform (inverse CDF).
//
// Upstream these are Python callables built on third-party libraries (zfit Gauss / Cauchy,
// zfit_physics RelativisticBreitWigner: fromdecay/mass_functions.py:16-102; tfp TruncatedNormal:
// tests/helpers/decays.py:26-37) sampled one value per event with per-event limits.  Their
// arithmetic is not in the reference tree; the densities are restated here and sampled by inverse
// CDF so that a resonance costs exactly one draw (deterministic draw count, no rejection loop).
// oracle/genbod_numpy.py::sample_mass is the CPU statement of the same equations
// (x = F^-1(F(lo) + u (F(hi) - F(lo)))); the solvers here are built for the B200's fp64 pipe:
//   * everything that depends only on (mass, width) comes precomputed in the parameter block (PsParams::sc);
//   * when the limits are event independent (CONST_LIM: the first resonance below a top particle at rest,
//     i.e. B -> K* gamma) so does everything that depends only on the limits (PsParams::lim);
//   * tan / far-field atan are range-specialised (ps_fastmath.cuh), and the relativistic Breit-Wigner CDF
//     is inverted by two fp32 Halley steps on the special-function unit followed by one fp64
//     fourth-order inverse-series step instead of a data-dependent fp64 Newton loop.
#ifndef PS_SAMPLERS_CUH_
#define PS_SAMPLERS_CUH_

#include "ps_params.h"
#include "ps_fastmath.cuh"

namespace psk {

struct SamplerConst {
    double m, w;                // pole mass and width
    double c0, c1, c2, c3, c4;  // PsParams::sc
};

__device__ __forceinline__ SamplerConst sampler_const(const PsParams& prm, int s1) {
    return SamplerConst{prm.mass[s1], prm.width[s1], prm.sc[0][s1], prm.sc[1][s1], prm.sc[2][s1], prm.sc[3][s1], prm.sc[4][s1]};
}

// fmin(fmax(x, lo), hi) -- NaN maps to lo -- as two compares and selects (fmin / fmax are seven instructions each)
__device__ __forceinline__ double clamp_mass(double x, double lo, double hi) {
    const double y = x >= lo ? x : lo;
    return y > hi ? hi : y;
}

// truncated normal(mu = m, sigma = w) on [lo, hi]
template <bool CONST_LIM>
__device__ __forceinline__ double sample_gauss(const SamplerConst& k, const double* lim, double lo, double hi, double u) {
    double ca, dc;
    bool flip;
    if constexpr (CONST_LIM) {
        ca = lim[0], dc = lim[1], flip = lim[2] != 0.0;
    } else {
        const double alpha = (lo - k.m) * k.c0, beta = (hi - k.m) * k.c0;
        flip = alpha > 0.0;  // work in the tail that keeps precision
        const double a_ = flip ? -beta : alpha, b_ = flip ? -alpha : beta;
        ca = normcdf(a_);
        dc = normcdf(b_) - ca;
    }
    const double u_ = flip ? 1.0 - u : u;
    double z = normcdfinv(fma(u_, dc, ca));
    z = flip ? -z : z;
    return clamp_mass(fma(k.w, z, k.m), lo, hi);
}

// Cauchy ~ 1 / (1 + ((x-m)/w)^2) on [lo, hi]
template <bool CONST_LIM>
__device__ __forceinline__ double sample_cauchy(const SamplerConst& k, const double* lim, double lo, double hi, double u) {
    double ta, dt;
    if constexpr (CONST_LIM) {
        ta = lim[0], dt = lim[1];
    } else {
        ta = atan((lo - k.m) * k.c0);
        dt = atan((hi - k.m) * k.c0) - ta;
    }
    double z, c2;
    tan_pio2(fma(u, dt, ta), z, c2);
    return clamp_mass(fma(k.w, z, k.m), lo, hi);
}

// ------------------------------------------------------------------ relativistic Breit-Wigner
// Density ~ 1/((x^2-m^2)^2 + m^2 w^2).  The quartic factors as ((x-p)^2+q^2)((x+p)^2+q^2) with
// rho = sqrt(m^4 + m^2 w^2), p = sqrt((rho+m^2)/2), q = sqrt((rho-m^2)/2).  In z = (x-p)/q, kappa = 2p/q the CDF is,
// up to a constant factor,
//     G(z) = atan(z) + atan(z+kappa) + ln(((z+kappa)^2+1)/(z^2+1)) / kappa ,   G'(z) = (kappa^2+4) / ((z^2+1)((z+kappa)^2+1)),
// and the sample is the root of G(z) = G(z_lo) + u (G(z_hi) - G(z_lo)).  Solver (narrow resonances, kappa >= 16):
//   1. t = atan z linearises the core (G = t + slowly varying terms): start from the linear interpolation in t and
//      take two Halley steps in fp32 -- sin, cos, log, reciprocals on the special-function unit, ~30 instructions
//      each.  That leaves a relative error of 1e-6 .. 1e-2 in z (the far tail, where G'(t) -> 0, is the worst).
//   2. one evaluation of G in fp64 (range-specialised tan, far-field atan series, one log) and the inverse-function
//      Taylor series of z(G) to fourth order, whose coefficients are polynomials in z because 1/G' is one:
//      the truncation error is ~alpha^4/100 relative to the step, alpha = h P'(z)/(kappa^2+4), h = target - G.
//   3. if |alpha| > 0.01 (not observed with limits inside [0, 10 m]; percent-level when the upper limit is 100 m,
//      where step 1 is poor) repeat step 2 once or twice while |alpha| <= 0.2, else fall back to the generic solver.
// Broad resonances (kappa < 16) use the generic Newton iteration with library functions below.
#define PS_RELBW_MAX_ITERS 8
#define PS_RELBW_LAST_STEP 1e-9
#define PS_RELBW_MIN_KAPPA 16.0
#define PS_RELBW_ALPHA_OK 0.01
#define PS_RELBW_ALPHA_RETRY 0.2

// fp32 special-function unit, flush-to-zero forms: the default (non-ftz) forms of the CUDA fast intrinsics wrap every
// MUFU in denormal scaling code, three to six extra instructions each
__device__ __forceinline__ float f32_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f32_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f32_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f32_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// slowly varying part of G:  atan(z+kappa) + ln(((z+kappa)^2+1) inv_zz1)/kappa,   inv_zz1 = 1/(z^2+1)
template <bool FAR>
__device__ __forceinline__ double relbw_slow(const SamplerConst& k, double z, double inv_zz1) {
    const double y = z + k.c2;
    const double at = FAR ? atan_far(y) : atan(y);
    return fma(log(fma(y, y, 1.0) * inv_zz1), k.c3, at);
}

__device__ __forceinline__ void relbw_limits(const SamplerConst& k, double lo, double hi, bool far, double& t_lo,
                                             double& dt, double& g_lo, double& dg) {
    const double iq = rcp_pos(k.c1);
    const double z_lo = (lo - k.c0) * iq, z_hi = (hi - k.c0) * iq;
    t_lo = atan(z_lo);
    const double t_hi = atan(z_hi);
    const double i_lo = rcp_pos(fma(z_lo, z_lo, 1.0)), i_hi = rcp_pos(fma(z_hi, z_hi, 1.0));
    double s_lo, s_hi;
    if (far) {
        s_lo = relbw_slow<true>(k, z_lo, i_lo), s_hi = relbw_slow<true>(k, z_hi, i_hi);
    } else {
        s_lo = relbw_slow<false>(k, z_lo, i_lo), s_hi = relbw_slow<false>(k, z_hi, i_hi);
    }
    dt = t_hi - t_lo;
    g_lo = t_lo + s_lo;
    dg = (t_hi + s_hi) - g_lo;
}

// generic solver: Newton in t with library functions, to full convergence
// (scalars by value: a struct reference would make every caller spill the constants to its stack frame)
static __device__ __noinline__ double relbw_solve_generic(double kappa, double inv_kappa, double ik2, double t_lo, double t_hi,
                                                          double t, double target) {
    bool last = false;
#pragma unroll 1
    for (int it = 0; it < 4 * PS_RELBW_MAX_ITERS; ++it) {
        const double z = tan(t);
        const double y = z + kappa;
        const double b = fma(y, y, 1.0);
        const double g = t + atan(y) + inv_kappa * log(b / fma(z, z, 1.0)) - target;
        const double step = g * b * ik2;
        t = fmin(fmax(t - step, t_lo), t_hi);
        if (last) break;
        last = fabs(step) < PS_RELBW_LAST_STEP;  // one more step after the first tiny one
    }
    return tan(t);
}

template <bool CONST_LIM>
__device__ __forceinline__ double sample_relbw(const SamplerConst& k, const double* lim, double lo, double hi, double u) {
    const bool far = k.c2 >= PS_RELBW_MIN_KAPPA;
    double t_lo, dt, g_lo, dg;
    if constexpr (CONST_LIM) {
        t_lo = lim[0], dt = lim[1], g_lo = lim[2], dg = lim[3];
    } else {
        relbw_limits(k, lo, hi, far, t_lo, dt, g_lo, dg);
    }
    const double target = fma(u, dg, g_lo);
    const double t0 = fma(u, dt, t_lo);
    const double t_hi = t_lo + dt;
    double z;
    if (far) {
        // 1. fp32 Halley steps in t
        const float kf = (float)k.c2, ikf = (float)k.c3, ik2f = (float)k.c4;
        const float tlof = (float)t_lo, thif = (float)t_hi, tgf = (float)target;
        float tf = (float)t0;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const float s = f32_sin(tf);
            const float c = fmaxf(f32_cos(tf), 1e-6f);
            const float c2 = c * c;
            const float zf = s * f32_rcp(c);
            const float y = zf + kf;
            const float iy = f32_rcp(y);
            const float iy2 = iy * iy;
            const float at = 1.5707964f - iy * (1.0f - iy2 * (0.33333334f - 0.2f * iy2));
            const float b = fmaf(y, y, 1.0f);
            const float bc = b * c2;
            const float g = (tf - tgf) + fmaf(f32_lg2(bc), ikf * 0.6931472f, at);
            const float n = g * b * ik2f;        // Newton step
            const float r = -y * f32_rcp(bc);    // G''/(2 G') in t
            const float den = fmaf(-n, r, 1.0f);
            const float step = den > 0.5f ? n * f32_rcp(den) : n;
            tf = fminf(fmaxf(tf - step, tlof), thif);
        }
        // 2./3. fp64: evaluate G once, fourth-order inverse series in z
        double t = (double)tf;
        double inv_zz1;
        tan_pio2(t, z, inv_zz1);
        bool done = false;
#pragma unroll 1
        for (int it = 0; it < 3; ++it) {
            const double y = z + k.c2;
            const double A = fma(z, z, 1.0), B = fma(y, y, 1.0);
            const double G = t + fma(log(B * inv_zz1), k.c3, atan_far(y));
            const double h = (target - G) * k.c4;   // h / (kappa^2 + 4)
            const double zy = z * y;
            const double P1 = 2.0 * fma(z, B, y * A);        // P = A B and its derivatives
            const double P2 = fma(8.0, zy, 2.0 * (A + B));
            const double P3 = 12.0 * (y + z);
            const double d = h * (A * B);                     // Newton step in z
            const double al = h * P1, be = h * P2 * d, ga = h * P3 * d * d;
            const double al2 = al * al;
            // dz = d (1 + al/2 + (be + al^2)/6 + (ga + 4 al be + al^3)/24)
            const double s4 = fma(al, fma(4.0, be, al2), ga);
            const double corr = fma(s4, 1.0 / 24.0, fma(be + al2, 1.0 / 6.0, fma(al, 0.5, 1.0)));
            const double zn = fma(d, corr, z);
            if (fabs(al) <= PS_RELBW_ALPHA_OK) {
                z = zn;
                done = true;
                break;
            }
            if (!(fabs(al) <= PS_RELBW_ALPHA_RETRY)) break;  // series useless (or NaN): generic solver
            // rare: move (t, z) along consistently and evaluate again
            t += atan((zn - z) / fma(z, zn, 1.0));
            z = zn;
            inv_zz1 = 1.0 / fma(z, z, 1.0);
        }
        if (!done) z = relbw_solve_generic(k.c2, k.c3, k.c4, t_lo, t_hi, t0, target);
    } else {
        z = relbw_solve_generic(k.c2, k.c3, k.c4, t_lo, t_hi, t0, target);
    }
    return clamp_mass(fma(k.c1, z, k.c0), lo, hi);
}

// ------------------------------------------------------------------ inverse CDFs from a table of nodes
// Event-independent limits (the first resonance below a top particle at rest, e.g. B -> K* gamma): the target value of the
// CDF is F_lo + u dF with the same F_lo, dF for every event, so the host can solve F(z_k) = F_lo + u_k dF once for a set of
// nodes u_k (ps_api.cu build_relbw_table / build_gauss_table, long double) and the kernel needs no transcendental function
// at all: it takes the nearest node, knows h = (u - u_k) dF EXACTLY (no evaluation of F), and applies the fourth-order
// inverse-function series around z_k.  The nodes are uniform in w = min(u, 1 - u) within dyadic levels
// [2^-(j+2), 2^-(j+1)), which puts them where z(u) bends -- the tails; the host checks on every level that the truncation
// error of the series stays below the rounding floor or does not offer a table.
//
// Layout (PsParams::node_tab): lower half of the uniform (u < 1/2: w = u) then upper half (w = 1 - u); per half J+1 levels
// (level j < J: w in [2^-(j+2), 2^-(j+1)), level J: w in [0, 2^-(J+1))), per level M+1 = 2^mlog + 1 nodes, node 0 at the
// upper edge of the level.  node_locate returns the index of the nearest node and du = u - u_node (exact).
__device__ __forceinline__ int node_locate(double u, int mlog, int J_lo, int J_hi, double& du, double& w_out) {
    const int M = 1 << mlog;  // nodes per level: M + 1
    const bool upper = u >= 0.5;
    const double w = upper ? 1.0 - u : u;  // exact
    const int J = upper ? J_hi : J_lo;
    // level: w in [2^e, 2^(e+1)) -> j = -e - 2, clamped to [0, J] (w = 0 and w = 1/2 land on the clamps)
    int j = 1021 - ((__double2hiint(w) >> 20) & 0x7ff);
    j = j < 0 ? 0 : j;
    const bool last = j >= J;
    j = last ? J : j;
    const int sh = last ? J + 1 : j + 2;                       // w 2^sh in [1, 2) (level j < J) or [0, 1) (level J)
    const double x = w * __hiloint2double((1023 + sh) << 20, 0);  // exact
    const double s = ((last ? 1.0 : 2.0) - x) * __hiloint2double((1023 + mlog) << 20, 0);  // exact: position in the level, 0 .. M
    const int i = __double2int_rn(s);
    const double dw = (s - (double)i) * __hiloint2double((1023 - sh - mlog) << 20, 0);       // w_node - w, exact
    du = upper ? dw : -dw;
    w_out = w;
    return (upper ? (J_lo + 1) * (M + 1) : 0) + j * (M + 1) + i;
}

// Relativistic Breit-Wigner: one z per node; the series parameter |alpha| = h |P'(z)| / (kappa^2 + 4) stays below 0.004
// (truncation ~ alpha^4 / 100 of the step).  ~45 instructions (30 fp64, one 8-byte load) instead of ~350 (140 fp64, 12 MUFU).
__device__ __forceinline__ double sample_relbw_table(const SamplerConst& k, const double* lim, const double* tab, int mlog,
                                                     int J_lo, int J_hi, double lo, double hi, double u) {
    double du, w;
    const int idx = node_locate(u, mlog, J_lo, J_hi, du, w);
    const double h = du * lim[3] * k.c4;       // (u - u_node) dG / (kappa^2 + 4)
    const double z = __ldg(tab + idx);
    const double y = z + k.c2;
    const double A = fma(z, z, 1.0), B = fma(y, y, 1.0);
    const double P1 = 2.0 * fma(z, B, y * A);  // P = A B and its derivatives
    const double P2 = fma(8.0, z * y, 2.0 * (A + B));
    const double P3 = 12.0 * (y + z);
    const double d = h * (A * B);              // Newton step in z
    const double al = h * P1, be = h * P2 * d, ga = h * P3 * d * d;
    const double al2 = al * al;
    const double s4 = fma(al, fma(4.0, be, al2), ga);
    const double corr = fma(s4, 1.0 / 24.0, fma(be + al2, 1.0 / 6.0, fma(al, 0.5, 1.0)));
    return clamp_mass(fma(k.c1, fma(d, corr, z), k.c0), lo, hi);
}

// Truncated normal: per node z_k = Phi^-1(p_k) and r_k = 1 / phi(z_k) (one 16-byte load).  The derivatives of the inverse
// are d^n z / dp^n = P_(n-1)(z) r^n with P_0 = 1, P_n = P'_(n-1) + n z P_(n-1), so with d = h r_k
//     z(p_k + h) = z + d + z d^2/2 + (2 z^2 + 1) d^3/6 + (6 z^3 + 7 z) d^4/24 + (24 z^4 + 46 z^2 + 7) d^5/120
//                  + (120 z^5 + 326 z^3 + 127 z) d^6/720 + O((720 z^6 + 2556 z^4 + 1740 z^2 + 127) d^7/5040);
// the host keeps that last term below 2e-16 at every node.  Sixth order rather than fourth because it lets the table be
// four times smaller (65 or 129 nodes per level): the nodes are read with one divergent load per warp, and with 513 nodes
// per level nearly every warp had a lane that missed L1 (measured: 2.21 ms per 1e8 K* gamma events, 2.4 with normcdfinv).
// Draws closer to an end of the range than w_min (only where an end lies so far out that Phi there is 0 or 1 in fp64, and
// then only u = 0 from the generator's own 51-bit stream) are left to the library path by the caller.
// ~40 instructions (21 fp64) instead of ~180 (normcdfinv).
// (coefficients in __constant__ memory: read as constant-bank operands of the DFMA instead of being materialised with
// two moves each, as in ps_fastmath.cuh)
static __constant__ double kGaussSeries[9] = {1.0 / 3.0, 1.0 / 6.0, 7.0 / 24.0, 0.2, 23.0 / 60.0, 7.0 / 120.0,
                                              1.0 / 6.0, 163.0 / 360.0, 127.0 / 720.0};
__device__ __forceinline__ double sample_gauss_table(const SamplerConst& k, const double* lim, const double* tab, int idx,
                                                     double du, double lo, double hi) {
    const double2 node = __ldg(reinterpret_cast<const double2*>(tab) + idx);
    const double z = node.x;
    const double d = du * lim[1] * node.y;     // (u - u_node) (Phi(b) - Phi(a)) / phi(z)
    const double z2 = z * z;
    const double c3 = fma(z2, kGaussSeries[0], kGaussSeries[1]);
    const double c4 = z * fma(z2, 0.25, kGaussSeries[2]);
    const double c5 = fma(z2, fma(z2, kGaussSeries[3], kGaussSeries[4]), kGaussSeries[5]);
    const double c6 = z * fma(z2, fma(z2, kGaussSeries[6], kGaussSeries[7]), kGaussSeries[8]);
    const double corr = fma(d, fma(d, fma(d, fma(d, fma(d, c6, c5), c4), c3), 0.5 * z), 1.0);
    return clamp_mass(fma(k.w, fma(d, corr, z), k.m), lo, hi);
}

template <bool CONST_LIM>
__device__ __forceinline__ double sample_mass(int kind, const SamplerConst& k, const double* lim, double lo, double hi,
                                              double u) {
    switch (kind) {
        case PS_KIND_GAUSS: return sample_gauss<CONST_LIM>(k, lim, lo, hi, u);
        case PS_KIND_BW: return sample_cauchy<CONST_LIM>(k, lim, lo, hi, u);
        default: return sample_relbw<CONST_LIM>(k, lim, lo, hi, u);
    }
}

}  // namespace psk
#endif
