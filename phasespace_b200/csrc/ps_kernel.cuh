// Fused GENBOD / Raubold-Lynch event generator for one decay tree: one event per thread.
//
// Replaces, in ONE kernel, everything below GenParticle._recursive_generate in the reference
// (src/phasespace/phasespace.py:497-626 -> _generate 300-397 -> _generate_part2 399-495 -> pdk 56-71,
// _get_w_max 288-298; kinematics.py:93-175; random.py) for a whole decay tree.  The tree is a
// compile-time type (Pt<...> below) so that every particle 4-vector lives in registers; masses and
// widths are runtime constants in the kernel-parameter block.  The same header is instantiated ahead
// of time for common trees (ps_aot.cu) and by NVRTC for any other tree (ps_api.cu).
//
// Two arithmetic variants, selected per translation unit:
//   PS_EXACT=1 (compiled with -fmad=false): the reference's operation order, op for op (SURVEY.md
//              appendix A); differs from the numpy oracle only through sin/cos (<= 2 ulp).
//   PS_EXACT=0 (default, the product path): algebraically identical but restructured for the B200
//              fp64 pipe -- reciprocals shared between pdk and the boosts, two-body energies from
//              invariants instead of square roots, gamma/gamma*beta boosts, sincospi, FMA
//              contraction.  Agrees with the oracle to ~1e-15 of the parent mass scale.
//
// This file must stay NVRTC-clean: no system headers.
#ifndef PS_KERNEL_CUH_
#define PS_KERNEL_CUH_

#include "ps_params.h"
#include "ps_philox.cuh"
#include "ps_samplers.cuh"
#include "ps_fastmath.cuh"

#ifndef PS_EXACT
#define PS_EXACT 0
#endif
#ifndef PS_BLOCK
#define PS_BLOCK 256
#endif
#ifndef PS_MIN_BLOCKS
#define PS_MIN_BLOCKS 2
#endif

namespace psk {

// ------------------------------------------------------------------ tiny metaprogramming kit
template <int... I>
struct Seq {};
template <int N, int... I>
struct MakeSeqT : MakeSeqT<N - 1, N - 1, I...> {};
template <int... I>
struct MakeSeqT<0, I...> {
    using type = Seq<I...>;
};
template <int N>
using MakeSeq = typename MakeSeqT<N>::type;

template <int J, class... Ts>
struct Nth;
template <class T, class... Ts>
struct Nth<0, T, Ts...> {
    using type = T;
};
template <int J, class T, class... Ts>
struct Nth<J, T, Ts...> : Nth<J - 1, Ts...> {};

// Decay-tree node.  SLOT: output slot of this particle (-1 for the top), in the reference's dict
// order (phasespace.py:547-570).  KIND: mass model (PS_KIND_*).  UCOL: column of this node's first
// uniform in the tree-wide draw order (decaying nodes only).  Ch...: children, in declaration order.
template <int SLOT, int KIND, int UCOL, class... Ch>
struct Pt {
    static constexpr int slot = SLOT, kind = KIND, ucol = UCOL, nch = sizeof...(Ch);
    static constexpr bool sampled = (KIND == PS_KIND_GAUSS || KIND == PS_KIND_BW || KIND == PS_KIND_RELBW);
    static constexpr int n_below = (0 + ... + (1 + Ch::n_below));
    static constexpr int n_sampled_children = (0 + ... + (Ch::sampled ? 1 : 0));
    static constexpr int n_sampled_below = (0 + ... + ((Ch::sampled ? 1 : 0) + Ch::n_sampled_below));
    // uniform draws of this node and of everything below it (ps_tree_n_draws): 3n-4 per decay + one per sampled child
    static constexpr int n_draws =
        (sizeof...(Ch) == 0 ? 0 : (int)(3 * sizeof...(Ch)) - 4 + n_sampled_children) + (0 + ... + Ch::n_draws);
    static constexpr bool resonant_leaf_below =
        (false || ... || ((Ch::nch == 0 && Ch::kind != PS_KIND_FIXED) || Ch::resonant_leaf_below));
    static constexpr bool has_grandchildren = (false || ... || (Ch::nch > 0));
    // every decay in the tree is a two-body decay (few instructions per byte written: memory-bound, see genbod_body)
    static constexpr bool two_body_only = (sizeof...(Ch) == 0 || sizeof...(Ch) == 2) && (true && ... && Ch::two_body_only);
    template <int J>
    using child = typename Nth<J, Ch...>::type;
    __host__ __device__ static constexpr int sampled_before(int j) {
        const bool s[] = {Ch::sampled..., false};
        int n = 0;
        for (int i = 0; i < j; ++i) n += s[i] ? 1 : 0;
        return n;
    }
    __host__ __device__ static constexpr int nonfixed_before(int j) {
        const bool s[] = {(Ch::kind != PS_KIND_FIXED)..., false};
        int n = 0;
        for (int i = 0; i < j; ++i) n += s[i] ? 1 : 0;
        return n;
    }
};

// ------------------------------------------------------------------ products of two-body momenta
// Fast variant: the maximum weight of a tree is a product of two-body momenta sqrt(x_i)/(2 a_i) over all decaying
// nodes; only its reciprocal is needed to normalise the weight, so the factors are multiplied up first --
// prod(2 a_i) * rsqrt(prod x_i) -- with one reciprocal square root per four factors (x_i ~ M^4: four of them stay
// far from overflow for any mass scale) instead of a square root and a reciprocal per factor.
struct WmaxAcc {
    double rs = 1.0, px = 1.0, pe = 1.0;
    int cnt = 0, n = 0;
    __device__ __forceinline__ void add(double x, double two_a) {
        px = cnt == 0 ? x : px * x;
        pe = n == 0 ? two_a : pe * two_a;
        ++n;
        if (++cnt == 4) flush();
    }
    __device__ __forceinline__ void flush() {
        if (cnt) rs = rs * rsqrt_pos(px);
        cnt = 0;
    }
    __device__ __forceinline__ double inverse() {  // 1 / w_max
        flush();
        return pe * rs;
    }
};

// ------------------------------------------------------------------ per-event state
// WONLY: only the event weight is wanted (the accept-reject mask pass): no kinematics, and the weight -- like the
// maximum weight -- is accumulated as a product in wacc instead of being formed factor by factor.
// KEYED: the Philox round keys are read from the parameter block (PsParams::philox_keys) instead of being re-derived from
// the seed with twenty uniform-datapath additions per tile.  Measured: flat four-body decay 2.61 -> 2.54 ms per 1e8 events
// (578 -> 545 instructions per loop iteration), but the memory-bound three-body decay 1.562 -> 1.577 ms (four more
// registers, a different store schedule), so the small trees keep the seed form (use_keyed_philox).
// EARLY: the children of a top particle that decays at rest through the top-down path (genbod_topdown) are final as soon
// as they are formed and are stored right away (out_row / out_stride) instead of being held until the end of the event.
// TABLE: the azimuths use the table-driven sincos (ps_fastmath.cuh sincos_turn, use_turn_table below).
template <int P, bool DBG, bool WONLY = false, bool KEYED = true, bool EARLY = false, bool TABLE = false>
struct Ctx {
    static constexpr bool wonly = WONLY && !PS_EXACT;
    static constexpr bool early_store = EARLY && !PS_EXACT;
    static constexpr bool turn_table = TABLE;
    static constexpr double turn_s = turn_scale<TABLE>();
    double px[P], py[P], pz[P], en[P], ms[P];
    double* out_row;        // EARLY: row of this event in particle plane 0
    long long out_stride;   // EARLY: doubles between planes
    unsigned int wd[4];     // descending access (mantissa_desc): the first block of the column consumed last
    WmaxAcc wacc;  // 1 / (product of the two-body momenta of the event), wonly
    unsigned long long seed, event;
    const unsigned int* keys;  // PsParams::philox_keys
    const double* dbg_row;
    unsigned int w[4];  // Philox block currently being consumed
    bool bad;

    // Draw number COL of this event: the 51-bit field at bit offset 51*COL of the event's Philox bit string
    // (ps_philox.cuh), as the two words of the double d = 1 + u/2 in [1, 1.5).  Draws are consumed strictly in column
    // order, so at most the block that holds the first bit and the one that holds the last bit are needed; which
    // blocks are fresh is known at compile time.
    template <int COL>
    __device__ __forceinline__ double mantissa() {
        constexpr int bit0 = PS_UNIFORM_BITS * COL;
        constexpr int FB = bit0 / 128, LB = (bit0 + PS_UNIFORM_BITS - 1) / 128;
        constexpr int PREV = COL == 0 ? -1 : (bit0 - 1) / 128;  // last block touched by the previous column
        if constexpr (FB > PREV) {
            const Philox4 r = KEYED ? philox4x32_10_keyed(keys, event, FB, 0u) : philox4x32_10(seed, event, FB, 0u);
            w[0] = r.w[0], w[1] = r.w[1], w[2] = r.w[2], w[3] = r.w[3];
        }
        unsigned int x[8] = {w[0], w[1], w[2], w[3], 0u, 0u, 0u, 0u};
        if constexpr (LB > FB) {
            const Philox4 r = KEYED ? philox4x32_10_keyed(keys, event, LB, 0u) : philox4x32_10(seed, event, LB, 0u);
            x[4] = r.w[0], x[5] = r.w[1], x[6] = r.w[2], x[7] = r.w[3];
            w[0] = r.w[0], w[1] = r.w[1], w[2] = r.w[2], w[3] = r.w[3];
        }
        unsigned int lo, hiword;
        field_words<bit0 - 128 * FB>(x, lo, hiword);
        return __hiloint2double((int)hiword, (int)lo);
    }
    // The same for draws consumed in DESCENDING column order (the top-down path forms the last particle first), from
    // column TOPCOL down: the block kept (wd) is the one that holds the first bit of the column consumed last, and the
    // block the ascending phase stopped in (index ASC, still in w) is reused when the descent reaches it.
    template <int COL, int TOPCOL, int ASC>
    __device__ __forceinline__ double mantissa_desc() {
        constexpr int bit0 = PS_UNIFORM_BITS * COL;
        constexpr int FB = bit0 / 128, LB = (bit0 + PS_UNIFORM_BITS - 1) / 128;
        constexpr int HELD = COL == TOPCOL ? -1 : (PS_UNIFORM_BITS * (COL + 1)) / 128;
        unsigned int x[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if constexpr (LB > FB) {
            if constexpr (LB == HELD) {
                x[4] = wd[0], x[5] = wd[1], x[6] = wd[2], x[7] = wd[3];
            } else {
                const Philox4 r = KEYED ? philox4x32_10_keyed(keys, event, LB, 0u) : philox4x32_10(seed, event, LB, 0u);
                x[4] = r.w[0], x[5] = r.w[1], x[6] = r.w[2], x[7] = r.w[3];
            }
        }
        if constexpr (FB == HELD) {
            x[0] = wd[0], x[1] = wd[1], x[2] = wd[2], x[3] = wd[3];
        } else if constexpr (FB == ASC) {
            x[0] = wd[0] = w[0], x[1] = wd[1] = w[1], x[2] = wd[2] = w[2], x[3] = wd[3] = w[3];
        } else {
            const Philox4 r = KEYED ? philox4x32_10_keyed(keys, event, FB, 0u) : philox4x32_10(seed, event, FB, 0u);
            x[0] = wd[0] = r.w[0], x[1] = wd[1] = r.w[1], x[2] = wd[2] = r.w[2], x[3] = wd[3] = r.w[3];
        }
        unsigned int lo, hiword;
        field_words<bit0 - 128 * FB>(x, lo, hiword);
        return __hiloint2double((int)hiword, (int)lo);
    }
    template <int COL, int TOPCOL, int ASC>
    __device__ __forceinline__ double next_cos_desc() {
        if constexpr (DBG)
            return 2.0 * dbg_row[COL] - 1.0;
        else
            return fma(mantissa_desc<COL, TOPCOL, ASC>(), 4.0, -5.0);
    }
    template <int COL, int TOPCOL, int ASC>
    __device__ __forceinline__ void next_turn_desc(double& y, double& t) {
        if constexpr (DBG) {
            const double u = dbg_row[COL];
            y = fma(u, turn_s, PS_TURN_MAGIC);
            t = u * turn_s;
        } else {
            const double d = mantissa_desc<COL, TOPCOL, ASC>();
            y = fma(d, 2.0 * turn_s, PS_TURN_MAGIC - 2.0 * turn_s);
            t = fma(d, 2.0 * turn_s, -2.0 * turn_s);
        }
    }
    // the uniform u = 2 d - 2 itself
    template <int COL>
    __device__ __forceinline__ double next() {
        if constexpr (DBG)
            return dbg_row[COL];
        else
            return fma(mantissa<COL>(), 2.0, -2.0);
    }
    // 2 u - 1 (the cosine of a polar angle) = 4 d - 5: exact either way, one instruction from d
    template <int COL>
    __device__ __forceinline__ double next_cos() {
        if constexpr (DBG || PS_EXACT)
            return 2.0 * next<COL>() - 1.0;
        else
            return fma(mantissa<COL>(), 4.0, -5.0);
    }
    // the arguments of sincos_turn for the azimuth 2 pi u: y = S u + PS_TURN_MAGIC (one rounding of the same real number
    // either way) and t = S u (exact: S = turn_scale<TABLE>() is a power of two)
    template <int COL>
    __device__ __forceinline__ void next_turn(double& y, double& t) {
        if constexpr (DBG) {
            const double u = dbg_row[COL];
            y = fma(u, turn_s, PS_TURN_MAGIC);
            t = u * turn_s;
        } else {
            const double d = mantissa<COL>();
            y = fma(d, 2.0 * turn_s, PS_TURN_MAGIC - 2.0 * turn_s);
            t = fma(d, 2.0 * turn_s, -2.0 * turn_s);
        }
    }
};

// ------------------------------------------------------------------ kinematics (kinematics.py)
// kin.mass, kinematics.py:93-104, metric (-,-,-,+) summed left to right
// (never FMA-contracted: E^2 - p^2 of a boosted vector cancels, so the last bit of each product matters)
__device__ __forceinline__ double kin_mass(double x, double y, double z, double e) {
    double s = -__dmul_rn(x, x);
    s = __dadd_rn(s, -__dmul_rn(y, y));
    s = __dadd_rn(s, -__dmul_rn(z, z));
    s = __dadd_rn(s, __dmul_rn(e, e));
    return sqrt(s);
}

// phasespace.py:56-71 -- CERN 68-15 eq. 9.17, factors in the reference's order
__device__ __forceinline__ double pdk_x(double a, double b, double c) {
    return (a - b - c) * (a + b + c) * (a - b + c) * (a + b - c);
}
__device__ __forceinline__ double pdk(double a, double b, double c) { return sqrt(pdk_x(a, b, c)) / (2.0 * a); }

// phasespace.py:288-298
template <int N>
__device__ __forceinline__ double get_w_max(double avail, const double (&m)[N]) {
    double emmax = avail + m[0], emmin = 0.0, w_max = 1.0;
#pragma unroll
    for (int i = 1; i < N; ++i) {
        emmin = emmin + m[i - 1];
        emmax = emmax + m[i];
        w_max = w_max * pdk(emmax, emmin, m[i]);
    }
    return w_max;
}
template <int N>
__device__ __forceinline__ void get_w_max(double avail, const double (&m)[N], WmaxAcc& acc) {
    double emmax = avail + m[0], emmin = 0.0;
#pragma unroll
    for (int i = 1; i < N; ++i) {
        emmin = emmin + m[i - 1];
        emmax = emmax + m[i];
        acc.add(pdk_x(emmax, emmin, m[i]), emmax + emmax);
    }
}

// kinematics.py:122-149, reference operation order.  The reference tests b2 == 0 over the whole
// batch (kinematics.py:147); here it is per event, a superset of the reference's finite behaviour.
__device__ __forceinline__ void lorentz_boost_exact(double& x, double& y, double& z, double& e, double bx, double by,
                                                    double bz) {
    double b2 = bx * bx;
    b2 = b2 + by * by;
    b2 = b2 + bz * bz;
    if (b2 == 0.0) return;
    const double gamma = 1.0 / sqrt(1.0 - b2);
    const double gamma2 = (gamma - 1.0) / b2;
    double bp = x * bx;
    bp = bp + y * by;
    bp = bp + z * bz;
    const double t = gamma2 * bp + gamma * e;
    x = x + t * bx;
    y = y + t * by;
    z = z + t * bz;
    e = gamma * (e + bp);
}

// One compare and four 32-bit selects.  (fmin / fmax cost fourteen instructions per exchange on sm_100a: there is no
// fp64 min/max instruction, and their NaN handling compiles to DSETP.MIN/MAX + SEL + FSEL + eight register moves.
// The sorted values are uniforms, never NaN.)
__device__ __forceinline__ void cmp_exchange(double& a, double& b) {
    const bool swap = b < a;
    const double lo = swap ? b : a, hi = swap ? a : b;
    a = lo, b = hi;
}
// Sorting networks: optimal (fewest comparators) up to 8 elements -- 1, 3, 5, 9, 12, 16, 19 instead of the K(K-1)/2
// of insertion order.
template <int K>
__device__ __forceinline__ void sort_network(double (&r)[K]) {
#define PS_CE(i, j) cmp_exchange(r[i], r[j])
    if constexpr (K == 2) {
        PS_CE(0, 1);
    } else if constexpr (K == 3) {
        PS_CE(0, 2); PS_CE(0, 1); PS_CE(1, 2);
    } else if constexpr (K == 4) {
        PS_CE(0, 2); PS_CE(1, 3); PS_CE(0, 1); PS_CE(2, 3); PS_CE(1, 2);
    } else if constexpr (K == 5) {
        PS_CE(0, 3); PS_CE(1, 4); PS_CE(0, 2); PS_CE(1, 3); PS_CE(0, 1); PS_CE(2, 4); PS_CE(1, 2); PS_CE(3, 4); PS_CE(2, 3);
    } else if constexpr (K == 6) {
        PS_CE(0, 5); PS_CE(1, 3); PS_CE(2, 4); PS_CE(1, 2); PS_CE(3, 4); PS_CE(0, 3); PS_CE(2, 5); PS_CE(0, 1); PS_CE(2, 3);
        PS_CE(4, 5); PS_CE(1, 2); PS_CE(3, 4);
    } else if constexpr (K == 7) {
        PS_CE(0, 6); PS_CE(2, 3); PS_CE(4, 5); PS_CE(0, 2); PS_CE(1, 4); PS_CE(3, 6); PS_CE(0, 1); PS_CE(2, 5); PS_CE(3, 4);
        PS_CE(1, 2); PS_CE(4, 6); PS_CE(2, 3); PS_CE(4, 5); PS_CE(1, 2); PS_CE(3, 4); PS_CE(5, 6);
    } else if constexpr (K == 8) {
        PS_CE(0, 2); PS_CE(1, 3); PS_CE(4, 6); PS_CE(5, 7); PS_CE(0, 4); PS_CE(1, 5); PS_CE(2, 6); PS_CE(3, 7); PS_CE(0, 1);
        PS_CE(2, 3); PS_CE(4, 5); PS_CE(6, 7); PS_CE(2, 4); PS_CE(3, 5); PS_CE(1, 4); PS_CE(3, 6); PS_CE(1, 2); PS_CE(3, 4);
        PS_CE(5, 6);
    } else {
#pragma unroll
        for (int i = 1; i < K; ++i) {
#pragma unroll
            for (int j = i; j > 0; --j) PS_CE(j - 1, j);
        }
    }
#undef PS_CE
}

// ------------------------------------------------------------------ child masses (phasespace.py:335-354)
template <class Node, int J, class C>
__device__ __forceinline__ void fixed_sum_one(const PsParams& prm, double& s, bool& first) {
    using Child = typename Node::template child<J>;
    if constexpr (Child::kind == PS_KIND_FIXED) {
        s = first ? prm.mass[Child::slot + 1] : s + prm.mass[Child::slot + 1];
        first = false;
    }
}

// TOP_AT_REST: Node is the top particle decaying at rest, so the limits of its first non-fixed child are event
// independent and come precomputed (PsParams::lim).
template <class Node, int J, bool TOP_AT_REST, class C>
__device__ __forceinline__ void assign_mass_one(C& c, const PsParams& prm, long long i, double* m, double& max_mass) {
    using Child = typename Node::template child<J>;
    constexpr int s1 = Child::slot + 1;
    if constexpr (Child::kind == PS_KIND_FIXED) {
        m[J] = prm.mass[s1];
    } else {
        double mass;
        if constexpr (Child::kind == PS_KIND_USER) {
            mass = prm.user_masses[i * prm.n_user + prm.user_col[s1]];
        } else {
            const double u = c.template next<Node::ucol + Node::sampled_before(J)>();
            constexpr bool CONST_LIM = TOP_AT_REST && Node::nonfixed_before(J) == 0;
            if (CONST_LIM && Child::kind == PS_KIND_RELBW && prm.node_tab) {
                mass = sample_relbw_table(sampler_const(prm, s1), prm.lim, prm.node_tab, prm.tab_mlog, prm.tab_J_lo,
                                          prm.tab_J_hi, prm.min_mass[s1], max_mass, u);
            } else if (CONST_LIM && Child::kind == PS_KIND_GAUSS && prm.node_tab) {
                double du, w;
                const int idx = node_locate(u, prm.tab_mlog, prm.tab_J_lo, prm.tab_J_hi, du, w);
                if (w >= prm.tab_wmin)
                    mass = sample_gauss_table(sampler_const(prm, s1), prm.lim, prm.node_tab, idx, du, prm.min_mass[s1], max_mass);
                else  // next to an end of the range that lies where Phi is 0 or 1 in fp64 (PsParams::tab_wmin)
                    mass = sample_gauss<CONST_LIM>(sampler_const(prm, s1), prm.lim, prm.min_mass[s1], max_mass, u);
            } else {
                mass = sample_mass<CONST_LIM>(Child::kind, sampler_const(prm, s1), prm.lim, prm.min_mass[s1], max_mass, u);
            }
        }
        max_mass = max_mass - mass;  // phasespace.py:352
        m[J] = mass;
    }
}

template <class Node, bool TOP_AT_REST, class C, int... J>
__device__ __forceinline__ void assign_masses(C& c, const PsParams& prm, long long i, double M, double* m, Seq<J...>) {
    double fs = 0.0;
    bool first = true;
    (fixed_sum_one<Node, J, C>(prm, fs, first), ...);
    double max_mass = M - fs;  // phasespace.py:342
    (assign_mass_one<Node, J, TOP_AT_REST, C>(c, prm, i, m, max_mass), ...);
}

template <int BASE, class C, int... I>
__device__ __forceinline__ void fetch_draws(C& c, double* s, Seq<I...>) {
    ((s[I] = c.template next<BASE + I>()), ...);
}

// ------------------------------------------------------------------ one GENBOD step (phasespace.py:428-494)
// K is the index of the particle created in this step; AB is the column of its (a, b) pair.
template <int N, int K, int AB, class C>
__device__ __forceinline__ void genbod_step_exact(C& c, const double* pd, const double* inv, const double* m, double* x,
                                                  double* y, double* z, double* e) {
    x[K] = 0.0;
    y[K] = -pd[K - 1];
    z[K] = 0.0;
    e[K] = sqrt(pd[K - 1] * pd[K - 1] + m[K] * m[K]);
    const double ua = c.template next<AB>();
    const double ub = c.template next<AB + 1>();
    const double cz = 2.0 * ua - 1.0;
    const double sz = sqrt(1.0 - cz * cz);
    const double ang = (2.0 * 3.141592653589793) * ub;
    double sy, cy;
    sincos(ang, &sy, &cy);
#pragma unroll
    for (int j = 0; j <= K; ++j) {
        const double x0 = x[j], y0 = y[j];
        const double x1 = cz * x0 - sz * y0;
        y[j] = sz * x0 + cz * y0;
        const double z0 = z[j];
        x[j] = cy * x1 - sy * z0;
        z[j] = sy * x1 + cy * z0;
    }
    if constexpr (K < N - 1) {
        const double beta = pd[K] / sqrt(pd[K] * pd[K] + inv[K] * inv[K]);
#pragma unroll
        for (int j = 0; j <= K; ++j) lorentz_boost_exact(x[j], y[j], z[j], e[j], 0.0, beta, 0.0);
    }
}

// Fast variant.  ra[i] = 1/inv[i+1]; esub[k] = energy of the (0..k) subsystem in the next frame.
template <int N, int K, int AB, class C>
__device__ __forceinline__ void genbod_step_fast(C& c, const double* pd, const double* ra, const double* enew,
                                                 const double* esub, double* x, double* y, double* z, double* e) {
    const double cz = c.template next_cos<AB>();
    const double sz = sqrt_unit1(fma(-cz, cz, 1.0));
    double ty, tu, sy, cy;
    c.template next_turn<AB + 1>(ty, tu);
    sincos_turn<C::turn_table>(ty, tu, sy, cy);
    const double t = sz * pd[K - 1];
    if constexpr (K == 1) {
        // particles 0 and 1 are back to back: (0, +-pd0, 0) rotated
        x[0] = -(cy * t);
        y[0] = cz * pd[0];
        z[0] = -(sy * t);
        e[0] = esub[0];
    } else {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double x0 = x[j], y0 = y[j], z0 = z[j];
            const double x1 = cz * x0 - sz * y0;
            y[j] = sz * x0 + cz * y0;
            x[j] = cy * x1 - sy * z0;
            z[j] = sy * x1 + cy * z0;
        }
    }
    x[K] = cy * t;
    y[K] = -(cz * pd[K - 1]);
    z[K] = sy * t;
    e[K] = enew[K];
    if constexpr (K < N - 1) {
        // boost along +y into the rest frame of subsystem K+1: gamma = E_sub/inv[K], gamma*beta = pd[K]/inv[K]
        const double g = esub[K] * ra[K - 1], gb = pd[K] * ra[K - 1];
#pragma unroll
        for (int j = 0; j <= K; ++j) {
            const double y0 = y[j], e0 = e[j];
            y[j] = fma(g, y0, gb * e0);
            e[j] = fma(g, e0, gb * y0);
        }
    }
}

template <int N, int AB0, class C, int... KM1>
__device__ __forceinline__ void genbod_steps_exact(C& c, const double* pd, const double* inv, const double* m, double* x,
                                                   double* y, double* z, double* e, Seq<KM1...>) {
    (genbod_step_exact<N, KM1 + 1, AB0 + 2 * KM1>(c, pd, inv, m, x, y, z, e), ...);
}
template <int N, int AB0, class C, int... KM1>
__device__ __forceinline__ void genbod_steps_fast(C& c, const double* pd, const double* ra, const double* enew,
                                                  const double* esub, double* x, double* y, double* z, double* e,
                                                  Seq<KM1...>) {
    (genbod_step_fast<N, KM1 + 1, AB0 + 2 * KM1>(c, pd, ra, enew, esub, x, y, z, e), ...);
}

// ------------------------------------------------------------------ stores
// One 256-bit store per particle per thread: a warp writes 32 consecutive 32-byte rows = 1 KiB
// contiguous, i.e. eight full 128-byte lines per instruction (STG.E.ENL2.256 on sm_100a; PTX ISA 8.8,
// CUDA >= 12.9).  Older NVRTC versions (PTX 8.7) fall back to two 128-bit halves.
#if defined(__CUDACC_VER_MAJOR__) && (__CUDACC_VER_MAJOR__ * 100 + __CUDACC_VER_MINOR__ >= 1209)
#define PS_HAVE_V4F64 1
#else
#define PS_HAVE_V4F64 0
#endif
__device__ __forceinline__ void store_row(double* p, double a, double b, double c, double d) {
#if PS_HAVE_V4F64
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p + 2), "d"(c), "d"(d) : "memory");
#endif
}
__device__ __forceinline__ void load_row(const double* p, double& a, double& b, double& c, double& d) {
#if PS_HAVE_V4F64
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#else
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(c), "=d"(d) : "l"(p + 2));
#endif
}

// ------------------------------------------------------------------ top-down formulation for many-body nodes
// The steps above touch every particle formed so far in every later step: (n(n+1)/2 - 1) rotations and (n(n-1)/2 - 1)
// boosts of a particle, 12 fp64 instructions per particle and step.  The same events come out of ONE accumulated
// Lorentz matrix when the steps are taken from the top down: with R_K the rotation and B_K the boost of step K,
//     p_K = M_K v_K,   M_K = R_{N-1} B_{N-2} R_{N-2} ... B_K R_K = M_{K+1} B_K R_K,   v_K = (0, -pd[K-1], 0, E_K),
// so each step costs one update of the matrix (three mixes of two columns, 48 instructions) and one product with a
// vector that has two non-zero components (8), whatever K -- 56 per step against 12 K + 8.  By the count that pays from
// eight bodies on, by the clock from nine (ten bodies: 434 instead of 568 instructions for the rotations and boosts), and a particle is final --
// stored, its registers free -- as soon as it is formed, instead of living through every later step.  The draws are
// the reference's, column for column (phasespace.py:444-451); they are merely consumed last pair first.
// Measured on a B200 (5e7 events, ms per launch, bottom-up -> top-down): 8 bodies 3.46 -> 3.47, 9 bodies 4.39 -> 4.07,
// 10 bodies 5.15 -> 4.67; six bodies lose 3 %.
#ifndef PS_TOPDOWN_MIN_N
#define PS_TOPDOWN_MIN_N 9
#endif

// quantities of the two-body step a -> b + cc: momentum, 1/a, energy of cc and of the subsystem b in the rest frame of a
__device__ __forceinline__ void two_body(double a, double b, double cc, double& pd, double& ra, double& enew, double& esub) {
    const double amb = a - b, apb = a + b;
    const double xx = (amb - cc) * (apb + cc) * (amb + cc) * (apb - cc);
    ra = rcp_pos1(a);
    const double hra = 0.5 * ra;
    pd = sqrt_pos1(xx) * hra;
    enew = fma(amb, apb, cc * cc) * hra;  // (a^2 + c^2 - b^2) / 2a
    esub = a - enew;
}

struct TdState {
    double X[4], Y[4], Z[4], T[4];  // columns of M_K: images of the x, y, z and E unit vectors; rows x, y, z, E
    double pd_hi, esub_hi;          // pd[K] and esub[K] of the step above
    double wn;                      // running product of the two-body momenta
};

template <class Node, int K, class C>
__device__ __forceinline__ void td_emit(C& c, double* x, double* y, double* z, double* e, double vx, double vy, double vz,
                                        double ve) {
    if constexpr (C::early_store && Node::slot < 0)
        store_row(c.out_row + Node::template child<K>::slot * c.out_stride, vx, vy, vz, ve);
    // (dead for early-stored leaves; decaying children are read again by the recursion)
    x[K] = vx, y[K] = vy, z[K] = vz, e[K] = ve;
}

template <int K, int AB0, int TOPCOL, int ASC, class C>
__device__ __forceinline__ void td_angles(C& c, double& cz, double& sz, double& cy, double& sy) {
    double ty, tu;
    c.template next_turn_desc<AB0 + 2 * (K - 1) + 1, TOPCOL, ASC>(ty, tu);
    sincos_turn<C::turn_table>(ty, tu, sy, cy);
    cz = c.template next_cos_desc<AB0 + 2 * (K - 1), TOPCOL, ASC>();
    sz = sqrt_unit1(fma(-cz, cz, 1.0));
}

// step K, N-2 >= K >= 2.  FIRST: the matrix still has the structure R_{N-1} left it with (T = e_E, no E row elsewhere).
template <class Node, int N, int K, int AB0, int TOPCOL, int ASC, bool FIRST, class C>
__device__ __forceinline__ void td_step(C& c, TdState& s, const double* inv, const double* m, double* x, double* y,
                                        double* z, double* e) {
    double pd, ra, enew, esub;
    two_body(inv[K], inv[K - 1], m[K], pd, ra, enew, esub);
    const double g = s.esub_hi * ra, gb = s.pd_hi * ra;  // boost of subsystem K into the rest frame of K+1
    double cz, sz, cy, sy;
    td_angles<K, AB0, TOPCOL, ASC>(c, cz, sz, cy, sy);
    if constexpr (FIRST) {
        // M B: Y <- g Y + gb T, T <- gb Y + g T with T = (0,0,0,1), Y[3] = 0
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double y0 = s.Y[r];
            s.Y[r] = g * y0;
            s.T[r] = gb * y0;
        }
        s.Y[3] = gb;
        s.T[3] = g;
        // M Ry: X <- cy X + sy Z, Z <- cy Z - sy X with X[3] = Z[3] = Z[1] = 0
#pragma unroll
        for (int r = 0; r < 3; r += 2) {
            const double x0 = s.X[r], z0 = s.Z[r];
            s.X[r] = fma(cy, x0, sy * z0);
            s.Z[r] = fma(cy, z0, -(sy * x0));
        }
        {
            const double x0 = s.X[1];
            s.X[1] = cy * x0;
            s.Z[1] = -(sy * x0);
        }
        // M Rz: X <- cz X + sz Y, Y <- cz Y - sz X with X[3] = 0
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double x0 = s.X[r], y0 = s.Y[r];
            s.X[r] = fma(cz, x0, sz * y0);
            s.Y[r] = fma(cz, y0, -(sz * x0));
        }
        s.X[3] = sz * gb;
        s.Y[3] = cz * gb;
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double y0 = s.Y[r], t0 = s.T[r];
            s.Y[r] = fma(g, y0, gb * t0);
            s.T[r] = fma(g, t0, gb * y0);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double x0 = s.X[r], z0 = s.Z[r];
            s.X[r] = fma(cy, x0, sy * z0);
            s.Z[r] = fma(cy, z0, -(sy * x0));
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double x0 = s.X[r], y0 = s.Y[r];
            s.X[r] = fma(cz, x0, sz * y0);
            s.Y[r] = fma(cz, y0, -(sz * x0));
        }
    }
    // p_K = M_K (0, -pd, 0, enew)
    td_emit<Node, K>(c, x, y, z, e, fma(enew, s.T[0], -(pd * s.Y[0])), fma(enew, s.T[1], -(pd * s.Y[1])),
                     fma(enew, s.T[2], -(pd * s.Y[2])), fma(enew, s.T[3], -(pd * s.Y[3])));
    s.pd_hi = pd;
    s.esub_hi = esub;
    s.wn = s.wn * pd;
}

template <class Node, int N, int AB0, int TOPCOL, int ASC, class C, int... I>
__device__ __forceinline__ void td_steps(C& c, TdState& s, const double* inv, const double* m, double* x, double* y,
                                         double* z, double* e, Seq<I...>) {
    // I = 0 .. N-4  ->  K = N-2 .. 2
    (td_step<Node, N, N - 2 - I, AB0, TOPCOL, ASC, I == 0>(c, s, inv, m, x, y, z, e), ...);
}

// Returns the product of the two-body momenta of the node; the children's 4-vectors in the node's rest frame go to
// x, y, z, e (or, for a top particle at rest in a generating kernel, straight to the output planes).
template <class Node, int N, int AB0, int ASC, class C>
__device__ __forceinline__ double genbod_topdown(C& c, const double* inv, const double* m, double* x, double* y, double* z,
                                                 double* e) {
    static_assert(N >= 4, "the top-down path needs at least one middle step");
    constexpr int TOPCOL = AB0 + 2 * (N - 2) + 1;
    TdState s;
    {  // K = N-1: M = R_{N-1}
        constexpr int K = N - 1;
        double pd, ra, enew, esub;
        two_body(inv[K], inv[K - 1], m[K], pd, ra, enew, esub);
        double cz, sz, cy, sy;
        td_angles<K, AB0, TOPCOL, ASC>(c, cz, sz, cy, sy);
        const double t = sz * pd;
        td_emit<Node, K>(c, x, y, z, e, cy * t, -(cz * pd), sy * t, enew);
        s.X[0] = cy * cz, s.X[1] = sz, s.X[2] = sy * cz, s.X[3] = 0.0;
        s.Y[0] = -(cy * sz), s.Y[1] = cz, s.Y[2] = -(sy * sz), s.Y[3] = 0.0;
        s.Z[0] = -sy, s.Z[1] = 0.0, s.Z[2] = cy, s.Z[3] = 0.0;
        s.T[0] = s.T[1] = s.T[2] = 0.0, s.T[3] = 1.0;
        s.pd_hi = pd, s.esub_hi = esub, s.wn = pd;
    }
    td_steps<Node, N, AB0, TOPCOL, ASC>(c, s, inv, m, x, y, z, e, MakeSeq<N - 3>{});
    {  // K = 1: only the y and E columns of M_1 are needed, for particles 1 and 0 (back to back)
        double pd, ra, enew, esub;
        two_body(inv[1], inv[0], m[1], pd, ra, enew, esub);
        const double g = s.esub_hi * ra, gb = s.pd_hi * ra;
        double cz, sz, cy, sy;
        td_angles<1, AB0, TOPCOL, ASC>(c, cz, sz, cy, sy);
        double v1[4], v0[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double y0 = s.Y[r], t0 = s.T[r];
            const double yb = fma(g, y0, gb * t0), tb = fma(g, t0, gb * y0);  // M B
            const double xr = fma(cy, s.X[r], sy * s.Z[r]);                    // x column of M B Ry
            const double yr = fma(cz, yb, -(sz * xr));                          // y column of M_1
            const double a = pd * yr;
            v1[r] = fma(enew, tb, -a);  // M_1 (0, -pd, 0, enew[1])
            v0[r] = fma(esub, tb, a);   // M_1 (0, +pd, 0, esub[0])
        }
        td_emit<Node, 1>(c, x, y, z, e, v1[0], v1[1], v1[2], v1[3]);
        td_emit<Node, 0>(c, x, y, z, e, v0[0], v0[1], v0[2], v0[3]);
        s.wn = s.wn * pd;
    }
    return s.wn;
}

template <class Node, class C, int... J>
__device__ __forceinline__ void store_children(C& c, const double* m, const double* x, const double* y, const double* z,
                                               const double* e, Seq<J...>) {
    ((c.px[Node::template child<J>::slot] = x[J], c.py[Node::template child<J>::slot] = y[J],
      c.pz[Node::template child<J>::slot] = z[J], c.en[Node::template child<J>::slot] = e[J],
      c.ms[Node::template child<J>::slot] = m[J]),
     ...);
}

template <class Node, bool REST, class C>
__device__ __forceinline__ double gen_node(C& c, const PsParams& prm, long long i, double ppx, double ppy, double ppz,
                                           double pe);

template <class Node, int J, class C>
__device__ __forceinline__ void recurse_child(C& c, const PsParams& prm, long long i, double& w) {
    using Child = typename Node::template child<J>;
    if constexpr (Child::nch > 0) {
        constexpr int s = Child::slot;
        w = w * gen_node<Child, false>(c, prm, i, c.px[s], c.py[s], c.pz[s], c.en[s]);  // phasespace.py:555-568
    }
}
template <class Node, class C, int... J>
__device__ __forceinline__ void recurse_children(C& c, const PsParams& prm, long long i, double& w, Seq<J...>) {
    (recurse_child<Node, J>(c, prm, i, w), ...);
}

// ------------------------------------------------------------------ one decay node
// GenParticle._generate + _generate_part2 (phasespace.py:300-495) for the node, then the recursion of
// _recursive_generate (phasespace.py:555-570).  (ppx,ppy,ppz,pe) is the parent's 4-vector in the
// output frame; REST says it is (0,0,0,M) so the final boost is the identity.  Returns the product
// of the two-body momenta of this node and of every decaying node below it.
template <class Node, bool REST, class C>
__device__ __forceinline__ double gen_node(C& c, const PsParams& prm, long long i, double ppx, double ppy, double ppz,
                                           double pe) {
    constexpr int N = Node::nch;
    constexpr int NS = Node::n_sampled_children;
    constexpr int SORT0 = Node::ucol + NS;       // first of the N-2 sorted uniforms
    constexpr int AB0 = Node::ucol + NS + N - 2;  // first (a,b) pair
    static_assert(N >= 2, "a decaying particle needs at least two children");

    double M;
    if constexpr (REST) {
        M = prm.mass[0];
    } else if constexpr (PS_EXACT || Node::slot < 0) {
        M = kin_mass(ppx, ppy, ppz, pe);  // phasespace.py:322: re-derived from the (boosted) 4-vector
    } else {
        // fast: the mass this particle was generated with.  Cheaper, and better conditioned than the
        // reference's E^2 - p^2 of a boosted vector (whose error ~ eps*gamma^2 is amplified near
        // threshold and can even turn an allowed decay into a "forbidden" one).
        M = c.ms[Node::slot];
    }

    double m[N];
    assign_masses<Node, REST && Node::slot < 0>(c, prm, i, M, m, MakeSeq<N>{});
    double msum = m[0];
#pragma unroll
    for (int j = 1; j < N; ++j) msum = msum + m[j];
    const double avail = M - msum;    // phasespace.py:357
    if (!(avail >= 0.0)) c.bad = true;  // "Forbidden decay", phasespace.py:358-363

    double r[N];
    r[0] = 0.0;
    r[N - 1] = 1.0;
    if constexpr (N > 2) {  // phasespace.py:367-375
        double s[N - 2];
        fetch_draws<SORT0>(c, s, MakeSeq<N - 2>{});
        sort_network<N - 2>(s);
#pragma unroll
        for (int j = 0; j < N - 2; ++j) r[j + 1] = s[j];
    }
    double inv[N];
    {
        double sum_ = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {  // phasespace.py:379-384
            sum_ = sum_ + m[j];
#if PS_EXACT
            inv[j] = r[j] * avail + sum_;
#else
            // r[0] = 0 and r[N-1] = 1 are exact; the middle ones are never contracted to an FMA because
            // near threshold (a - b - c) amplifies the last bit of inv[]
            inv[j] = j == 0 ? sum_ : (j == N - 1 ? avail + sum_ : __dadd_rn(__dmul_rn(r[j], avail), sum_));
#endif
        }
    }

    double x[N], y[N], z[N], e[N];
    double wn;
#if PS_EXACT
    {
        double pd[N - 1];
#pragma unroll
        for (int j = 0; j < N - 1; ++j) pd[j] = pdk(inv[j + 1], inv[j], m[j + 1]);  // phasespace.py:404-411
        wn = pd[0];
#pragma unroll
        for (int j = 1; j < N - 1; ++j) wn = wn * pd[j];  // phasespace.py:412
        x[0] = 0.0;
        y[0] = pd[0];
        z[0] = 0.0;
        e[0] = sqrt(pd[0] * pd[0] + m[0] * m[0]);  // phasespace.py:414-426
        genbod_steps_exact<N, AB0>(c, pd, inv, m, x, y, z, e, MakeSeq<N - 1>{});
        if constexpr (!REST) {  // phasespace.py:365, 389-391
            const double bx = ppx / pe, by = ppy / pe, bz = ppz / pe;
#pragma unroll
            for (int j = 0; j < N; ++j) lorentz_boost_exact(x[j], y[j], z[j], e[j], bx, by, bz);
        }
    }
#else
    if constexpr (C::wonly) {
#pragma unroll
        for (int j = 0; j < N - 1; ++j) {
            const double a = inv[j + 1], b = inv[j], cc = m[j + 1];
            const double amb = a - b, apb = a + b;
            c.wacc.add((amb - cc) * (apb + cc) * (amb + cc) * (apb - cc), a + a);
        }
        wn = 1.0;
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = y[j] = z[j] = e[j] = 0.0;
    } else {
        if constexpr (N >= PS_TOPDOWN_MIN_N) {
            // block index the ascending draws (sampled children, sorted uniforms) stopped in: still held in c.w
            constexpr int ASC = AB0 > 0 ? (PS_UNIFORM_BITS * AB0 - 1) / 128 : -1;
            wn = genbod_topdown<Node, N, AB0, ASC>(c, inv, m, x, y, z, e);
        } else {
            double pd[N - 1], ra[N - 1], enew[N], esub[N - 1];
#pragma unroll
            for (int j = 0; j < N - 1; ++j) {
                const double a = inv[j + 1], b = inv[j], cc = m[j + 1];
                const double amb = a - b, apb = a + b;
                const double xx = (amb - cc) * (apb + cc) * (amb + cc) * (apb - cc);
                ra[j] = rcp_pos1(a);
                const double hra = 0.5 * ra[j];
                pd[j] = sqrt_pos1(xx) * hra;
                enew[j + 1] = fma(amb, apb, cc * cc) * hra;  // (a^2 + c^2 - b^2) / 2a
                esub[j] = a - enew[j + 1];
            }
            wn = pd[0];
#pragma unroll
            for (int j = 1; j < N - 1; ++j) wn = wn * pd[j];
            genbod_steps_fast<N, AB0>(c, pd, ra, enew, esub, x, y, z, e, MakeSeq<N - 1>{});
        }
        if constexpr (!REST) {
            // boost by the parent: u = gamma*beta = p/M, gamma = E/M;  p' = p + (u.p/(gamma+1) + E) u ; E' = gamma E + u.p
            const double rM = rcp_pos1(M);
            const double ux = ppx * rM, uy = ppy * rM, uz = ppz * rM, g = pe * rM;
            const double g1 = rcp_pos1(g + 1.0);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double up = fma(ux, x[j], fma(uy, y[j], uz * z[j]));
                const double t = fma(up, g1, e[j]);
                x[j] = fma(t, ux, x[j]);
                y[j] = fma(t, uy, y[j]);
                z[j] = fma(t, uz, z[j]);
                e[j] = fma(g, e[j], up);
            }
        }
    }
#endif
    store_children<Node>(c, m, x, y, z, e, MakeSeq<N>{});
    recurse_children<Node>(c, prm, i, wn, MakeSeq<N>{});
    return wn;
}

// ------------------------------------------------------------------ tree-level maximum weight
// recurse_w_max, phasespace.py:571-625: leaf masses only, flattened depth first.
template <class Node, class C>
__device__ __forceinline__ void acc_leaves(const C& c, double& s);
template <class Node, int J, class C>
__device__ __forceinline__ void acc_leaves_child(const C& c, double& s) {
    using Child = typename Node::template child<J>;
    if constexpr (Child::nch == 0) {
        s = s + c.ms[Child::slot];
    } else {
        acc_leaves<Child>(c, s);
    }
}
template <class Node, class C, int... J>
__device__ __forceinline__ void acc_leaves_seq(const C& c, double& s, Seq<J...>) {
    (acc_leaves_child<Node, J>(c, s), ...);
}
template <class Node, class C>
__device__ __forceinline__ void acc_leaves(const C& c, double& s) {
    acc_leaves_seq<Node>(c, s, MakeSeq<Node::nch>{});
}
template <class Node, int SKIP, class C, int... J>
__device__ __forceinline__ void acc_leaves_except(const C& c, double& s, Seq<J...>) {
    ((J != SKIP ? acc_leaves_child<Node, J>(c, s) : (void)0), ...);
}

// ACC is double& (exact: the running product w_max, reference order) or WmaxAcc& (fast).
template <int N>
__device__ __forceinline__ void w_max_node(double avail, const double (&m)[N], double& w_max) {
    w_max = w_max * get_w_max<N>(avail, m);
}
template <int N>
__device__ __forceinline__ void w_max_node(double avail, const double (&m)[N], WmaxAcc& acc) {
    get_w_max<N>(avail, m, acc);
}

template <class Node, class C, class ACC>
__device__ __forceinline__ void recurse_w_max(const C& c, double parent_mass, ACC& acc);

template <class Node, int J, class C, class ACC>
__device__ __forceinline__ void w_max_child(const C& c, double parent_mass, double* masses, ACC& acc) {
    using Child = typename Node::template child<J>;
    if constexpr (Child::nch > 0) {
        double others = 0.0;
        acc_leaves_except<Node, J>(c, others, MakeSeq<Node::nch>{});
        recurse_w_max<Child>(c, parent_mass - others, acc);
        double own = 0.0;
        acc_leaves<Child>(c, own);
        masses[J] = own;
    } else {
        masses[J] = c.ms[Child::slot];
    }
}
template <class Node, class C, class ACC, int... J>
__device__ __forceinline__ void w_max_children(const C& c, double parent_mass, double* masses, ACC& acc, Seq<J...>) {
    (w_max_child<Node, J>(c, parent_mass, masses, acc), ...);
}
template <class Node, class C, class ACC>
__device__ __forceinline__ void recurse_w_max(const C& c, double parent_mass, ACC& acc) {
    constexpr int N = Node::nch;
    double flat = 0.0;
    acc_leaves<Node>(c, flat);
    const double avail = parent_mass - flat;
    double masses[N];
    w_max_children<Node>(c, parent_mass, masses, acc, MakeSeq<N>{});
    w_max_node<N>(avail, masses, acc);
}

template <class Node, class C, class ACC, int... J>
__device__ __forceinline__ void node_w_max(const C& c, double M, ACC& acc, Seq<J...>) {
    constexpr int N = Node::nch;
    const double m[N] = {c.ms[Node::template child<J>::slot]...};
    double msum = m[0];
#pragma unroll
    for (int j = 1; j < N; ++j) msum = msum + m[j];
    w_max_node<N>(M - msum, m, acc);
}

// Maximum weight of the event (phasespace.py:364 for a flat decay, 571-625 for a chain) as the pair the callers
// need: INV = 1 / w_max always; W_MAX itself only in the exact variant or when WANT_WMAX asks for it.
// Event independent (rest frame, fixed leaf masses): the host's constants.
template <class Tree, bool BOOST, class C>
__device__ __forceinline__ void event_w_max(const C& c, const PsParams& prm, double bx, double by, double bz, double be,
                                            bool want_w_max, double& w_max, double& inv) {
    constexpr bool WMAX_PER_EVENT = BOOST || Tree::resonant_leaf_below;
    if constexpr (WMAX_PER_EVENT) {
        const double M = BOOST ? kin_mass(bx, by, bz, be) : prm.mass[0];
#if PS_EXACT
        w_max = 1.0;
        if constexpr (Tree::has_grandchildren)
            recurse_w_max<Tree>(c, M, w_max);
        else
            node_w_max<Tree>(c, M, w_max, MakeSeq<Tree::nch>{});
        inv = 0.0;  // unused: the exact variant divides by w_max
#else
        WmaxAcc acc;
        if constexpr (Tree::has_grandchildren)
            recurse_w_max<Tree>(c, M, acc);
        else
            node_w_max<Tree>(c, M, acc, MakeSeq<Tree::nch>{});
        inv = acc.inverse();
        w_max = want_w_max ? rcp_pos(inv) : 0.0;
#endif
    } else {
        w_max = prm.wmax_const;
        inv = prm.inv_wmax_const;
    }
}

// ------------------------------------------------------------------ stores (store_row / load_row: see above the top-down path)
// Per-event inputs (the boost_to row, the event index of the regeneration pass) are the only long-latency loads of the
// loop, and nothing of an event can start before they return.  PS_PREFETCH_BOOST = 1: each thread prefetches the input of
// its NEXT event into L1 while it computes the current one (config 3, four bodies + per-event boost_to: 3.58 -> 3.23 ms
// per 1e8 events); = 2 (generating kernel): the next boost_to row is loaded into registers during the current event, a
// software pipeline one event deep (3.22 -> 3.09 ms; eight registers, no spills at the 128-register cap).
#ifndef PS_PREFETCH_BOOST
#define PS_PREFETCH_BOOST 2
#endif
__device__ __forceinline__ void prefetch_row(const double* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// slots below FIRST were stored early (Ctx::early_store)
template <int FIRST, class C, int... S>
__device__ __forceinline__ void store_particles(const C& c, double* base, long long stride, Seq<S...>) {
    ((S >= FIRST ? store_row(base + S * stride, c.px[S], c.py[S], c.pz[S], c.en[S]) : (void)0), ...);
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------ the kernel
// Work distribution over tiles of PS_BLOCK events, two shapes (compile-time, by tree size):
//   compute-heavy trees (a decay into three or more bodies anywhere): persistent grid (SM count x resident CTAs),
//       CTA b takes tiles b, b + grid, b + 2 grid, ...  The per-CTA prologue (Philox round keys, mass constants) and
//       the statistics epilogue are paid once.
//   memory-bound trees (at most PS_CHUNKED_MAX_P particles, or two-body decays only, e.g. B -> K* gamma, K* -> K pi:
//       measured 1.86 instead of 2.22 ms per 1e8 events): the grid covers all tiles and CTA b takes the K = prm.tiles_per_cta consecutive tiles from b*K.  The hardware
//       dispatches CTAs in index order, so chip-wide the stores sweep HBM in address order like a
//       memset (tools/store_bench.cu: 7.3-7.5 TB/s against 6.3 TB/s for the persistent stride, whose
//       CTAs drift apart); best for memory-bound trees (2 bodies).
// trees with more than PS_KEYED_MIN_DRAWS uniform draws per event read precomputed Philox round keys (see Ctx)
#ifndef PS_KEYED_MIN_DRAWS
#define PS_KEYED_MIN_DRAWS 5
#endif
// launch shape (see genbod_body): must match pick_grid in ps_api.cu.  (Chains of two-body decays with a sampled resonance
// were tried on the persistent grid in round 2: K* gamma Cauchy 1.98 -> 2.10 ms, normal and relativistic BW equal.)
#ifndef PS_CHUNKED_SAMPLED
#define PS_CHUNKED_SAMPLED 1
#endif
template <class Tree>
__host__ __device__ constexpr bool use_covering_grid() {
    return Tree::n_below <= PS_CHUNKED_MAX_P || (Tree::two_body_only && (PS_CHUNKED_SAMPLED || Tree::n_sampled_below == 0));
}
template <class Tree>
__host__ __device__ constexpr bool use_keyed_philox() {
    return Tree::n_draws > PS_KEYED_MIN_DRAWS;
}
// trees that are bound by instruction issue rather than by HBM use the table-driven sincos (ps_fastmath.cuh): more than
// PS_TABLE_MIN_DRAWS - 1 draws per event, or (PS_TABLE_SAMPLED) a resonance sampled in the kernel (K* gamma with a
// relativistic Breit-Wigner K*: 3.31 -> 3.22 ms per 1e8 events, Cauchy 2.11 -> 2.05)
#ifndef PS_TABLE_MIN_DRAWS
#define PS_TABLE_MIN_DRAWS 6
#endif
#ifndef PS_TABLE_SAMPLED
#define PS_TABLE_SAMPLED 1
#endif
template <class Tree>
__host__ __device__ constexpr bool use_turn_table() {
    return Tree::n_draws >= PS_TABLE_MIN_DRAWS || (PS_TABLE_SAMPLED && Tree::n_sampled_below > 0);
}

// PLAIN: the call every user makes -- consecutive event indices, normalised weights (no event_index list, no weights_max
// output): the three run-time questions about them leave the loop.  STATS: weight statistics requested (prm.stats).
// A kernel holds the plain body with and without statistics and the general one (genbod_dispatch); per event the plain
// body without statistics is 23 instructions shorter for three bodies (322 instead of 345), which is what counts once the
// board sits at its power limit and the kernel is bound by issue slots rather than by HBM.
#ifndef PS_SPECIALIZE
#define PS_SPECIALIZE 1
#endif
#ifndef PS_PDL
#define PS_PDL 1
#endif
template <class Tree, bool BOOST, bool DBG, bool PLAIN, bool STATS>
__device__ __forceinline__ void genbod_body_impl(const PsParams& prm) {
    constexpr int P = Tree::n_below;
    static_assert(P <= PS_MAX_SLOTS, "decay tree too large");

    double st_max = 0.0, st_sum = 0.0, st_sum2 = 0.0;
#if PS_EXACT
    double st_wmax = 0.0;
#else
    double st_inv_min = __longlong_as_double(0x7ff0000000000000LL);
#endif
    bool bad_any = false;
    const long long n = prm.n_events;
    // the shape is a property of the tree (PS_CHUNKED_MAX_P): the host launches accordingly
    constexpr bool CHUNKED = use_covering_grid<Tree>();
    const long long n_tiles = (n + PS_BLOCK - 1) / PS_BLOCK;
    long long tile = CHUNKED ? (long long)blockIdx.x * prm.tiles_per_cta : (long long)blockIdx.x;
    const long long tile_end = CHUNKED ? (tile + prm.tiles_per_cta < n_tiles ? tile + prm.tiles_per_cta : n_tiles) : n_tiles;
    const long long tile_step = CHUNKED ? 1 : (long long)gridDim.x;
#if PS_PREFETCH_BOOST == 2
    double nbx = 0.0, nby = 0.0, nbz = 0.0, nbe = 0.0;
    if constexpr (BOOST) {
        const long long i0 = tile * PS_BLOCK + threadIdx.x;
        if (tile < tile_end && i0 < n) load_row(prm.boost_to + i0 * prm.boost_stride, nbx, nby, nbz, nbe);
    }
#endif
    for (; tile < tile_end; tile += tile_step) {
        {
        const long long i = tile * PS_BLOCK + threadIdx.x;
        if (i < n) {
            // the children of a top particle at rest come out of the top-down path final: stored as they are formed
            constexpr bool EARLY = !BOOST && !PS_EXACT && Tree::nch >= PS_TOPDOWN_MIN_N;
            Ctx<P, DBG, false, use_keyed_philox<Tree>(), EARLY, use_turn_table<Tree>()> c;
            c.seed = prm.seed;
            c.keys = prm.philox_keys;
            if constexpr (PLAIN) {
                c.event = prm.first_event + (unsigned long long)i;
            } else if (prm.event_index) {  // the regeneration pass of the unweighting: this thread's next index into L1
                c.event = (unsigned long long)prm.event_index[i];
#if PS_PREFETCH_BOOST
                if (i + tile_step * PS_BLOCK < n) prefetch_row((const double*)(prm.event_index + i + tile_step * PS_BLOCK));
#endif
            } else {
                c.event = prm.first_event + (unsigned long long)i;
            }
            c.dbg_row = DBG ? prm.dbg_u + i * prm.dbg_ncols : nullptr;
            c.bad = false;
            c.out_row = prm.particles + i * 4;
            c.out_stride = prm.particle_stride;
            double bx = 0.0, by = 0.0, bz = 0.0, be = 0.0;
            if constexpr (BOOST) {
#if PS_PREFETCH_BOOST == 2
                // software pipeline in registers: this event's row was loaded during the previous event
                bx = nbx, by = nby, bz = nbz, be = nbe;
                const long long i_next = i + tile_step * PS_BLOCK;
                if (i_next < n) load_row(prm.boost_to + i_next * prm.boost_stride, nbx, nby, nbz, nbe);
#else
                load_row(prm.boost_to + i * prm.boost_stride, bx, by, bz, be);
#if PS_PREFETCH_BOOST
                // the row of this thread's NEXT event, into L1 while this event is computed: the load above is the only
                // long-latency instruction of the loop and nothing can start before it returns
                const long long i_next = i + tile_step * PS_BLOCK;
                if (i_next < n) prefetch_row(prm.boost_to + i_next * prm.boost_stride);
#endif
#endif
            }
            const double w = gen_node<Tree, !BOOST>(c, prm, i, bx, by, bz, be);
            double w_max, inv_w_max;
            event_w_max<Tree, BOOST>(c, prm, bx, by, bz, be, !PLAIN && !prm.normalize && prm.weights_max, w_max, inv_w_max);
            double w_out = w;
            if (PLAIN || prm.normalize) {  // phasespace.py:713-717
#if PS_EXACT
                w_out = w / w_max;
#else
                w_out = w * inv_w_max;
#endif
            } else if (prm.weights_max) {
                prm.weights_max[i] = w_max;
            }
            prm.weights[i] = w_out;
            store_particles<EARLY ? Tree::nch : 0>(c, prm.particles + i * 4, prm.particle_stride, MakeSeq<P>{});
            bad_any |= c.bad;
            if constexpr (STATS) {
                st_max = w_out > st_max ? w_out : st_max;  // (fmax / fmin: seven instructions each on sm_100a, see cmp_exchange)
                st_sum += w_out;
                st_sum2 += w_out * w_out;
#if PS_EXACT
                st_wmax = w_max > st_wmax ? w_max : st_wmax;
#else
                st_inv_min = inv_w_max < st_inv_min ? inv_w_max : st_inv_min;  // max w_max = 1 / min (1 / w_max)
#endif
            }
        }
        }
    }
    if (bad_any) *prm.err_flag = PS_ERR_FORBIDDEN;
    if (STATS && prm.stats) {  // epilogue: warp shuffle -> shared -> one atomic per CTA
        __shared__ double red[4][PS_BLOCK / 32];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        st_max = warp_max(st_max);
        st_sum = warp_sum(st_sum);
        st_sum2 = warp_sum(st_sum2);
#if PS_EXACT
        st_wmax = warp_max(st_wmax);
#else
        // no events: 1/inf = 0; an event-independent w_max is reported exactly, not as 1 / (1 / w_max)
        constexpr bool WMAX_CONST = !(BOOST || Tree::resonant_leaf_below);
        double st_wmax = warp_max(WMAX_CONST ? (st_inv_min == prm.inv_wmax_const ? prm.wmax_const : 0.0) : 1.0 / st_inv_min);
#endif
        if (lane == 0) {
            red[0][wid] = st_max;
            red[1][wid] = st_sum;
            red[2][wid] = st_sum2;
            red[3][wid] = st_wmax;
        }
        __syncthreads();
        if (wid == 0) {
            const bool in = lane < PS_BLOCK / 32;
            st_max = warp_max(in ? red[0][lane] : 0.0);
            st_sum = warp_sum(in ? red[1][lane] : 0.0);
            st_sum2 = warp_sum(in ? red[2][lane] : 0.0);
            st_wmax = warp_max(in ? red[3][lane] : 0.0);
            if (lane == 0) {
                // weights are >= 0, so their bit patterns order like unsigned integers
                atomicMax((unsigned long long*)&prm.stats[0], (unsigned long long)__double_as_longlong(st_max));
                atomicAdd(&prm.stats[1], st_sum);
                atomicAdd(&prm.stats[2], st_sum2);
                atomicMax((unsigned long long*)&prm.stats[3], (unsigned long long)__double_as_longlong(st_wmax));
            }
        }
    }
}

// Programmatic dependent launch (launch_generate in ps_api.cu sets the stream-serialisation attribute): as soon as every
// CTA of this grid has started, the CTAs of the NEXT launch in the stream may take the SM slots that this grid's last wave
// frees, and sit behind their own `wait` until this grid has completed and its stores are visible.  Nothing is read or
// written before the wait, so the ordering every caller relies on (inputs produced by, and output memory recycled from,
// earlier work in the stream) is unchanged; what is saved is the launch latency and the ramp of back-to-back calls --
// the 1e6-event calls of the reference's benchmark protocol, the per-mode launches of a GenMultiDecay, the chunks of
// generate_host and of the unweighting.  Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_prologue() {
#if PS_PDL
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

template <class Tree, bool BOOST, bool DBG>
__device__ __forceinline__ void genbod_body(const PsParams& prm) {
    pdl_prologue();
#if PS_SPECIALIZE
    if constexpr (!DBG) {  // the debug kernels are for parity, not for speed: one body
        if (!prm.event_index && prm.normalize) {
            if (prm.stats)
                genbod_body_impl<Tree, BOOST, DBG, true, true>(prm);
            else
                genbod_body_impl<Tree, BOOST, DBG, true, false>(prm);
            return;
        }
    }
#endif
    genbod_body_impl<Tree, BOOST, DBG, false, true>(prm);
}

// EXACT only distinguishes the symbol names of the two arithmetic variants (it must equal PS_EXACT).
template <class Tree, bool BOOST, bool DBG, int EXACT>
__global__ void __launch_bounds__(PS_BLOCK, PS_MIN_BLOCKS) genbod_kernel(const __grid_constant__ PsParams prm) {
    static_assert(EXACT == PS_EXACT, "variant tag must match the translation unit");
    genbod_body<Tree, BOOST, DBG>(prm);
}

// ------------------------------------------------------------------ unweighting, pass 1
// Recomputes the event weights only (the rotations, boosts and their draws are dead code here and
// are eliminated), draws one acceptance uniform per event from Philox stream 1 and records
// accept = (u * w_bound < w / w_max) as one bit per event.  Together with
// ps_unweight_scan / ps_unweight_index this yields the accepted events ordered by event index,
// independent of grid size or GPU count; ps_generate_indexed then regenerates only those.
// overflow (optional): number of candidates whose normalised weight exceeds w_bound -- such an event is accepted with
// probability 1 instead of w / w_bound > 1, i.e. the bound was too small and the sample is biased by exactly those
// events.  One compare and one ballot per event; the atomic only fires for warps that saw one.
template <class Tree, bool BOOST>
__device__ __forceinline__ void genbod_mask_body(const PsParams& prm, double w_bound, unsigned int* mask,
                                                 unsigned long long* overflow) {
    constexpr int P = Tree::n_below;
    const long long n = prm.n_events;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool bad_any = false;
    // no CTA-wide synchronisation: the per-tile counts the scan needs are the population counts of the mask words
    for (long long tile = blockIdx.x; tile * PS_BLOCK < n; tile += gridDim.x) {
        const long long i = tile * PS_BLOCK + threadIdx.x;
        bool accept = false, over = false;
        if (i < n) {
            Ctx<P, false, true, use_keyed_philox<Tree>()> c;
            c.seed = prm.seed;
            c.keys = prm.philox_keys;
            c.event = prm.first_event + (unsigned long long)i;
            c.dbg_row = nullptr;
            c.bad = false;
            double bx = 0.0, by = 0.0, bz = 0.0, be = 0.0;
            if constexpr (BOOST) load_row(prm.boost_to + i * prm.boost_stride, bx, by, bz, be);
            const double w = gen_node<Tree, !BOOST>(c, prm, i, bx, by, bz, be);
            double w_max, inv_w_max;
            event_w_max<Tree, BOOST>(c, prm, bx, by, bz, be, false, w_max, inv_w_max);
            const double u = use_keyed_philox<Tree>() ? philox_uniform0_keyed(prm.philox_keys, c.event, 1u)
                                                      : philox_uniform0(prm.seed, c.event, 1u);
#if PS_EXACT
            accept = u * w_bound * w_max < w;
            over = w_bound * w_max < w;
#else
            // w = 1 / c.wacc.inverse(): u w_bound < w / w_max  <=>  u w_bound / w < 1 / w_max  (w = 0: inf, rejected)
            const double bound_over_w = w_bound * c.wacc.inverse();
            accept = u * bound_over_w < inv_w_max;
            over = bound_over_w < inv_w_max;
            (void)w;
#endif
            bad_any |= c.bad;
        }
        const unsigned int word = __ballot_sync(0xffffffffu, accept);
        const unsigned int over_word = __ballot_sync(0xffffffffu, over);
        const long long wi = tile * (PS_BLOCK / 32) + wid;
        if (lane == 0 && wi * 32 < n) mask[wi] = word;
        if (lane == 0 && over_word && overflow) atomicAdd(overflow, (unsigned long long)__popc(over_word));
    }
    if (bad_any) *prm.err_flag = PS_ERR_FORBIDDEN;
}

// The weights-only pass holds no 4-vectors: at three resident CTAs (<= 80 registers, no spills) the ten-body mask pass
// takes 2.81 instead of 2.97 ms per 2^27 candidates on a B200 (four CTAs, 62 registers: 2.88 ms).
#ifndef PS_MASK_MIN_BLOCKS
#define PS_MASK_MIN_BLOCKS 3
#endif
template <class Tree, bool BOOST, int EXACT>
__global__ void __launch_bounds__(PS_BLOCK, PS_MASK_MIN_BLOCKS)
    genbod_mask_kernel(const __grid_constant__ PsParams prm, double w_bound, unsigned int* mask,
                       unsigned long long* overflow) {
    static_assert(EXACT == PS_EXACT, "variant tag must match the translation unit");
    genbod_mask_body<Tree, BOOST>(prm, w_bound, mask, overflow);
}

// ------------------------------------------------------------------ fused weighted histograms
// The consumer of the reference's own physics tests (tests/test_physics.py:96-107): a weighted histogram of every
// coordinate of every particle (momentum components on [p_lo, p_hi), energies on [e_lo, e_hi), weight = the
// normalised event weight) plus an unweighted histogram of the weights on [0, 1 + 1e-8).  Events are generated
// exactly as by genbod_kernel but never stored.
//
// Accumulation: shared-memory atomics exist natively only for 32-bit integers on sm_100a (fp64 and 64-bit integer
// adds compile to compare-and-swap retry loops, ATOMS.CAST.SPIN.64, which made the first version of this kernel
// slower than generating and storing the events).  So the weight is turned into a 60-bit fixed-point integer once
// per event (w * 2^57, weights below 8) and added as three 20-bit limbs with native ATOMS.ADD; a 32-bit cell takes
// 4096 such adds before it can overflow, so every PS_HIST_FOLD_TILES tiles the CTA folds the limbs into fp64
// accumulators (one owner thread per cell, no atomics) and clears them.  Integer sums are exact and order
// independent; the only roundings are the fixed-point conversion (2^-57 absolute) and one per fold.
struct PsHistParams {
    int n_bins;
    int pad_;
    double p_lo, p_inv_width, e_lo, e_inv_width, w_inv_width;
    double w_cap;        // weights in [w_cap 2^-PS_HIST_MIN_EXP, w_cap) take the fixed-point path (a power of two)
    double fixed_scale;  // 2^(PS_HIST_LIMBS * PS_HIST_LIMB_BITS) / w_cap
    double* hist;  // (4 P + 1, n_bins): rows px,py,pz,E of slot 0, ... , then the weight histogram
};
// a 32-bit cell takes 2^(32 - LIMB_BITS) adds of < 2^LIMB_BITS each; one tile adds at most PS_BLOCK per cell
#define PS_HIST_FOLD_TILES ((1 << (32 - PS_HIST_LIMB_BITS)) / PS_BLOCK)
// weights below w_cap * 2^-PS_HIST_MIN_EXP would lose relative precision in fixed point: they take the fp64 path
#define PS_HIST_MIN_EXP (PS_HIST_LIMBS * PS_HIST_LIMB_BITS - 34)

// bin of one coordinate, or an out-of-range index (as unsigned) if v is outside [lo, lo + n_bins / inv_width)
__device__ __forceinline__ unsigned int hist_bin(double v, double lo, double inv_width) {
    return (unsigned int)__double2int_rd((v - lo) * inv_width);  // saturates: far out-of-range values stay out of range
}

// shared layout: double acc[n_cells] | unsigned limbs[n_cells][PS_HIST_LIMBS]   (n_cells = (4 P + 1) n_bins)
// FAST: the event's weight as fixed-point limbs, native 32-bit atomics, no branches; else fp64 compare-and-swap loops.
template <bool FAST, class C, int S>
__device__ __forceinline__ void hist_particle(const C& c, const PsHistParams& h, double* acc, unsigned int* limbs,
                                              const unsigned int (&l)[PS_HIST_LIMBS], double w) {
    const double v[4] = {c.px[S], c.py[S], c.pz[S], c.en[S]};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned int bin = hist_bin(v[k], k == 3 ? h.e_lo : h.p_lo, k == 3 ? h.e_inv_width : h.p_inv_width);
        if (bin < (unsigned int)h.n_bins) {
            const unsigned int cell = (4 * S + k) * h.n_bins + bin;
            if constexpr (FAST) {
#pragma unroll
                for (int j = 0; j < PS_HIST_LIMBS; ++j) atomicAdd(limbs + cell * PS_HIST_LIMBS + j, l[j]);
            } else {
                atomicAdd(&acc[cell], w);
            }
        }
    }
}
template <bool FAST, class C, int... S>
__device__ __forceinline__ void hist_particles(const C& c, const PsHistParams& h, double* acc, unsigned int* limbs,
                                               const unsigned int (&l)[PS_HIST_LIMBS], double w, Seq<S...>) {
    (hist_particle<FAST, C, S>(c, h, acc, limbs, l, w), ...);
}

// value of the limbs of one cell (each limb sum is < 2^32, exact in fp64; the combination rounds once), then cleared
__device__ __forceinline__ double hist_take_cell(unsigned int* q) {
    double v = 0.0;
#pragma unroll
    for (int j = PS_HIST_LIMBS - 1; j >= 0; --j) {
        v = fma(v, (double)(1u << PS_HIST_LIMB_BITS), (double)q[j]);
        q[j] = 0u;
    }
    return v;
}

__device__ __forceinline__ void hist_fold(const PsHistParams& h, double* acc, unsigned int* limbs, int n_cells) {
    __syncthreads();
    const double inv_scale = 1.0 / h.fixed_scale;
    for (int k = threadIdx.x; k < n_cells; k += PS_BLOCK)
        acc[k] = fma(hist_take_cell(limbs + k * PS_HIST_LIMBS), inv_scale, acc[k]);
    __syncthreads();
}

template <class Tree, bool BOOST>
__device__ __forceinline__ void genbod_hist_body(const PsParams& prm, const PsHistParams& h) {
    constexpr int P = Tree::n_below;
    extern __shared__ double sh[];
    const int n_cells = (4 * P + 1) * h.n_bins;
    double* acc = sh;
    unsigned int* limbs = reinterpret_cast<unsigned int*>(sh + n_cells);
    for (int k = threadIdx.x; k < n_cells; k += PS_BLOCK) acc[k] = 0.0;
    for (int k = threadIdx.x; k < n_cells * PS_HIST_LIMBS; k += PS_BLOCK) limbs[k] = 0u;
    __syncthreads();
    bool bad_any = false;
    const long long n = prm.n_events;
    const double w_min = h.w_cap * (1.0 / (double)(1ull << PS_HIST_MIN_EXP));
    int since_fold = 0;
    for (long long tile = blockIdx.x; tile * PS_BLOCK < n; tile += gridDim.x) {
        const long long i = tile * PS_BLOCK + threadIdx.x;
        if (i < n) {
            Ctx<P, false, false, use_keyed_philox<Tree>(), false, use_turn_table<Tree>()> c;
            c.seed = prm.seed;
            c.keys = prm.philox_keys;
            c.event = prm.first_event + (unsigned long long)i;
            c.dbg_row = nullptr;
            c.bad = false;
            double bx = 0.0, by = 0.0, bz = 0.0, be = 0.0;
            if constexpr (BOOST) {
                load_row(prm.boost_to + i * prm.boost_stride, bx, by, bz, be);
#if PS_PREFETCH_BOOST
                if (i + (long long)gridDim.x * PS_BLOCK < n) prefetch_row(prm.boost_to + (i + (long long)gridDim.x * PS_BLOCK) * prm.boost_stride);
#endif
            }
            const double w = gen_node<Tree, !BOOST>(c, prm, i, bx, by, bz, be);
            double w_max, inv_w_max;
            event_w_max<Tree, BOOST>(c, prm, bx, by, bz, be, false, w_max, inv_w_max);
#if PS_EXACT
            const double w_norm = w / w_max;
#else
            const double w_norm = w * inv_w_max;
#endif
            bool fixed_ok = w_norm >= w_min && w_norm < h.w_cap;
            const unsigned long long fx = fixed_ok ? (unsigned long long)__double2ll_rn(w_norm * h.fixed_scale) : 0ull;
            unsigned int l[PS_HIST_LIMBS];
#pragma unroll
            for (int j = 0; j < PS_HIST_LIMBS; ++j)
                l[j] = (unsigned int)(fx >> (j * PS_HIST_LIMB_BITS)) & ((1u << PS_HIST_LIMB_BITS) - 1u);
            fixed_ok = fixed_ok && (fx >> (PS_HIST_LIMBS * PS_HIST_LIMB_BITS)) == 0ull;  // rounded up to w_cap itself
            if (fixed_ok)
                hist_particles<true>(c, h, acc, limbs, l, w_norm, MakeSeq<P>{});
            else  // far outside the expected range of weights (rare): fp64 atomics
                hist_particles<false>(c, h, acc, limbs, l, w_norm, MakeSeq<P>{});
            const unsigned int wb = hist_bin(w_norm, 0.0, h.w_inv_width);
            if (wb < (unsigned int)h.n_bins)  // a count: limb 0 of the weight row
                atomicAdd(limbs + (4 * P * h.n_bins + wb) * PS_HIST_LIMBS, 1u);
            bad_any |= c.bad;
        }
        if (++since_fold == PS_HIST_FOLD_TILES) {
            hist_fold(h, acc, limbs, n_cells);
            since_fold = 0;
        }
    }
    if (bad_any) *prm.err_flag = PS_ERR_FORBIDDEN;
    hist_fold(h, acc, limbs, n_cells);
    for (int k = threadIdx.x; k < n_cells; k += PS_BLOCK) {
        // the weight row holds counts in units of 1/fixed_scale (a power of two: exact)
        const double v = k >= 4 * P * h.n_bins ? acc[k] * h.fixed_scale : acc[k];
        if (v != 0.0) atomicAdd(&h.hist[k], v);
    }
}

template <class Tree, bool BOOST, int EXACT>
__global__ void __launch_bounds__(PS_BLOCK, PS_MIN_BLOCKS)
    genbod_hist_kernel(const __grid_constant__ PsParams prm, const __grid_constant__ PsHistParams h) {
    static_assert(EXACT == PS_EXACT, "variant tag must match the translation unit");
    genbod_hist_body<Tree, BOOST>(prm, h);
}

}  // namespace psk
#endif
