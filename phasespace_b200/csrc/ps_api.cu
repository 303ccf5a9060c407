// C-ABI implementation (include/phasespace_b200.h): decay-tree compiler, kernel registry (ahead of
// time + NVRTC), launchers, and the small tree-independent kernels.
//
// Host-side counterpart of the bookkeeping the reference does in Python at trace time:
// GenParticle.set_children validation (phasespace.py:191-221), output-dict order (547-570), draw
// order (344-353, 367, 444-451), recurse_stable (326-333), _get_w_max / recurse_w_max (288-298,
// 571-625).  There is deliberately no CPU implementation of the event generation here: without a
// CUDA device every compute entry point fails with PS_E_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/phasespace_b200.h"
#include "ps_params.h"
#include "ps_philox.cuh"
#include "ps_registry.h"
#include "ps_samplers.cuh"

#ifndef PS_BLOCK
#define PS_BLOCK 256
#endif

namespace {

thread_local std::string g_error;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define PS_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t err__ = (expr);                                                                \
        if (err__ != cudaSuccess) return fail(PS_E_CUDA, "%s: %s", #expr, cudaGetErrorString(err__)); \
    } while (0)

// ---------------------------------------------------------------------------------- registry
struct KernelKey {
    std::string sig;
    int variant, boost, dbg, exact;
    bool operator<(const KernelKey& o) const {
        if (sig != o.sig) return sig < o.sig;
        if (variant != o.variant) return variant < o.variant;
        if (boost != o.boost) return boost < o.boost;
        if (dbg != o.dbg) return dbg < o.dbg;
        return exact < o.exact;
    }
};

struct Registry {
    std::mutex mu;
    std::map<KernelKey, const void*> aot;
    std::map<KernelKey, const void*> jit;           // cudaKernel_t handles of NVRTC-built kernels
};
Registry& registry() {
    static Registry r;
    return r;
}

}  // namespace

extern "C" void ps_aot_register(const PsAotEntry* table, int n) {
    Registry& r = registry();
    std::lock_guard<std::mutex> lock(r.mu);
    for (int i = 0; i < n; ++i)
        r.aot[KernelKey{table[i].signature, table[i].variant, table[i].boost, table[i].dbg, table[i].exact}] = table[i].fn;
}

// ---------------------------------------------------------------------------------- decay tree
struct ps_tree {
    int n_nodes = 0;
    std::vector<ps_node_desc> nodes;
    std::vector<std::vector<int>> children;
    std::vector<int> slot, ucol, user_col;
    std::vector<double> min_mass;
    int top = -1, n_particles = 0, n_draws = 0, n_user = 0;
    bool has_grandchildren = false, resonant_leaf = false;
    bool two_body_only = true;  // every decaying node has exactly two children (Pt::two_body_only)
    double w_max = NAN;
    std::string signature;
    PsParams base;  // constant part of the parameter block
    // resolved kernels of this tree, [variant][boost][dbg][exact]: a small call (1e6 events last ~15 us on a B200) should
    // not pay for a mutex and a map lookup keyed on a hundred-character string every time
    mutable std::atomic<const void*> kernel_cache[3][2][2][2] = {};
    mutable std::atomic<int> blocks_per_sm_cache[3][2][2][2] = {};
    // nodes of the inverse CDF of a relativistic Breit-Wigner or a truncated normal with event-independent limits
    // (PsParams::node_tab): built once on the host, copied to each device at the first launch there
    std::vector<double> tab_nodes;
    int tab_mlog = 0, tab_J_lo = 0, tab_J_hi = 0;
    double tab_wmin = 0.0;
    int tab_kind = -1;  // PS_MASS_RELBW or PS_MASS_GAUSS when tab_nodes is not empty
    mutable std::mutex tab_mu;
    mutable std::map<int, double*> tab_dev;
    mutable std::atomic<unsigned long long> one_cta_ready[2][2] = {};  // [dbg][exact]: devices (bit mask) on which the shared-memory attribute is set (pick_grid)
};

namespace {

bool is_sampled(int kind) { return kind == PS_MASS_GAUSS || kind == PS_MASS_BW || kind == PS_MASS_RELBW; }

// phasespace.py:547-570 -- a node's children first, then each decaying child's subtree
void assign_slots(ps_tree& t, int node, int& next_slot, int& next_col) {
    const std::vector<int>& ch = t.children[node];
    for (int c : ch) t.slot[c] = next_slot++;
    const int n = (int)ch.size();
    t.ucol[node] = next_col;
    int sampled = 0;
    for (int c : ch) sampled += is_sampled(t.nodes[c].kind) ? 1 : 0;
    next_col += sampled + (n - 2) + 2 * (n - 1);
    for (int c : ch)
        if (!t.children[c].empty()) assign_slots(t, c, next_slot, next_col);
}

// phasespace.py:326-333
double recurse_stable(const ps_tree& t, int node) {
    double out = 0.0;
    for (int c : t.children[node])
        out = out + (t.nodes[c].kind == PS_MASS_FIXED ? t.nodes[c].mass : recurse_stable(t, c));
    return out;
}

std::string emit_signature(const ps_tree& t, int node) {
    std::string s = "Pt<" + std::to_string(t.slot[node]) + "," + std::to_string(t.nodes[node].kind) + "," +
                    std::to_string(t.children[node].empty() ? -1 : t.ucol[node]);
    for (int c : t.children[node]) s += "," + emit_signature(t, c);
    return s + ">";
}

double host_pdk(double a, double b, double c) {
    volatile double x = (a - b - c) * (a + b + c) * (a - b + c) * (a + b - c);
    return std::sqrt(x) / (2.0 * a);
}
// phasespace.py:288-298
double host_get_w_max(double avail, const std::vector<double>& m) {
    double emmax = avail + m[0], emmin = 0.0, w = 1.0;
    for (size_t i = 1; i < m.size(); ++i) {
        emmin = emmin + m[i - 1];
        emmax = emmax + m[i];
        w = w * host_pdk(emmax, emmin, m[i]);
    }
    return w;
}
void host_acc_leaves(const ps_tree& t, int node, double& s) {
    for (int c : t.children[node]) {
        if (t.children[c].empty())
            s = s + t.nodes[c].mass;
        else
            host_acc_leaves(t, c, s);
    }
}
// phasespace.py:590-616
double host_recurse_w_max(const ps_tree& t, int node, double parent_mass) {
    double flat = 0.0;
    host_acc_leaves(t, node, flat);
    const double avail = parent_mass - flat;
    std::vector<double> masses;
    double w = 1.0;
    for (int c : t.children[node]) {
        if (!t.children[c].empty()) {
            double others = 0.0;
            for (int o : t.children[node]) {
                if (o == c) continue;
                if (t.children[o].empty())
                    others = others + t.nodes[o].mass;
                else
                    host_acc_leaves(t, o, others);
            }
            w = w * host_recurse_w_max(t, c, parent_mass - others);
            double own = 0.0;
            host_acc_leaves(t, c, own);
            masses.push_back(own);
        } else {
            masses.push_back(t.nodes[c].mass);
        }
    }
    return w * host_get_w_max(avail, masses);
}

// Sampler constants that depend only on (mass, width): PsParams::sc.
void host_sampler_constants(int kind, double mass, double width, double out[5]) {
    for (int k = 0; k < 5; ++k) out[k] = 0.0;
    if (kind == PS_MASS_GAUSS || kind == PS_MASS_BW) {
        out[0] = 1.0 / width;
    } else if (kind == PS_MASS_RELBW) {
        const double a = mass * mass, b = mass * width;
        const double rho = std::sqrt(a * a + b * b);
        // p + i q = sqrt(a + i b): q from 2 p q = b, not from sqrt((rho - a) / 2), which cancels to zero for narrow
        // resonances (width / mass < 1e-8, e.g. the D0 in a DecayLanguage chain with a tiny width tolerance)
        const double p = std::sqrt(0.5 * (rho + a)), q = b / (2.0 * p);
        out[0] = p;
        out[1] = q;
        out[2] = 2.0 * p / q;
        out[3] = q / (2.0 * p);
        out[4] = 1.0 / (out[2] * out[2] + 4.0);
    }
}

// Sampler constants that depend on the limits [lo, hi] as well: PsParams::lim (same equations as
// ps_samplers.cuh: sample_gauss / sample_cauchy / relbw_limits).
void host_sampler_limits(int kind, double mass, double width, const double sc[5], double lo, double hi, double out[4]) {
    for (int k = 0; k < 4; ++k) out[k] = 0.0;
    if (kind == PS_MASS_BW) {
        out[0] = std::atan((lo - mass) / width);
        out[1] = std::atan((hi - mass) / width) - out[0];
    } else if (kind == PS_MASS_GAUSS) {
        const double alpha = (lo - mass) / width, beta = (hi - mass) / width;
        const bool flip = alpha > 0.0;
        const double a_ = flip ? -beta : alpha, b_ = flip ? -alpha : beta;
        const double ca = 0.5 * std::erfc(-a_ * M_SQRT1_2), cb = 0.5 * std::erfc(-b_ * M_SQRT1_2);
        out[0] = ca;
        out[1] = cb - ca;
        out[2] = flip ? 1.0 : 0.0;
    } else if (kind == PS_MASS_RELBW) {
        const double p = sc[0], q = sc[1], kappa = sc[2];
        const double z_lo = (lo - p) / q, z_hi = (hi - p) / q;
        auto G = [&](double z) {
            const double y = z + kappa;
            return std::atan(z) + std::atan(y) + std::log((y * y + 1.0) / (z * z + 1.0)) / kappa;
        };
        out[0] = std::atan(z_lo);
        out[1] = std::atan(z_hi) - out[0];
        out[2] = G(z_lo);
        out[3] = G(z_hi) - out[2];
    }
}

}  // namespace

// ---------------------------------------------------------------------------------- relBW node table
// ps_samplers.cuh sample_relbw_table: for a relativistic Breit-Wigner whose limits are the same for every event, solve
// G(z) = G_lo + u dG at the nodes u_k once (long double Newton from the neighbouring node, bisection as a safeguard).
namespace {

// G(z) - G(z_r) without cancellation: differences of arctangents as one atan2, of logarithms as log1p.  In the far tails
// G is flat to within the resolution of G itself (a D meson with its physical width: G changes by 1e-19, the long double's
// last bit, over 5e4 units of z at the lower limit), so nodes there are solved relative to the END of the range they
// approach; the targets w dG (lower half of the uniform) and -w dG (upper half) are exact.  As a side effect u -> 0 and
// u -> 1 map to the mass limits exactly.  (z, z_r >= -kappa/2 because masses are >= 0, so y = z + kappa >= kappa/2 > 0.)
long double relbw_dG(long double z, long double zr, long double kappa) {
    const long double y = z + kappa, yr = zr + kappa, dz = z - zr;
    const long double ay = dz * (y + yr) / (yr * yr + 1.0L), az = dz * (z + zr) / (zr * zr + 1.0L);
    long double logs;
    if (ay > -0.999999L && az > -0.999999L)
        logs = log1pl(ay) - log1pl(az);
    else  // a ratio next to zero: the plain form is as good there
        logs = logl((y * y + 1.0L) / (yr * yr + 1.0L)) - logl((z * z + 1.0L) / (zr * zr + 1.0L));
    return atan2l(dz, 1.0L + z * zr) + atan2l(dz, 1.0L + y * yr) + logs / kappa;
}

// z with G(z) - G(z_r) = target: Newton from a close start (the series around the neighbouring node), bisection as a safeguard
long double relbw_root(long double target, long double zr, long double kappa, long double z_lo, long double z_hi, long double z) {
    const long double k2 = kappa * kappa + 4.0L;
    z = z < z_lo ? z_lo : (z > z_hi ? z_hi : z);
    long double prev_step = HUGE_VALL;
    for (int it = 0; it < 40; ++it) {
        const long double y = z + kappa;
        const long double step = (relbw_dG(z, zr, kappa) - target) * (z * z + 1.0L) * (y * y + 1.0L) / k2;
        long double zn = z - step;
        zn = zn < z_lo ? z_lo : (zn > z_hi ? z_hi : zn);
        // quadratic convergence: a step below 1e-10 leaves an error below 1e-19, the noise floor of the 64-bit mantissa
        if (fabsl(zn - z) <= 1e-10L * (1.0L + fabsl(zn))) return zn;
        // steps that stopped shrinking below 1e-6: the rounding noise of G itself, amplified by 1 / G' (far tails)
        if (it >= 2 && fabsl(step) >= 0.5L * fabsl(prev_step) && fabsl(step) <= 1e-6L * (1.0L + fabsl(zn))) return zn;
        prev_step = step;
        z = zn;
    }
    long double a = z_lo, b = z_hi;  // Newton did not settle (never seen): bisection
    for (int it = 0; it < 200; ++it) {
        const long double mid = 0.5L * (a + b);
        (relbw_dG(mid, zr, kappa) < target ? a : b) = mid;
    }
    return 0.5L * (a + b);
}

// Fills t->tab_* when a table can be offered: narrow resonance (the solver's fast branch), proper limits, and at every node
// |alpha| = (half the node spacing in G) |P'(z)| / (kappa^2 + 4) <= PS_RELBW_TABLE_ALPHA.
//
// Cost: a tree with a D meson of its physical width (kappa ~ 5e12: 80 levels of 1025 nodes) took 0.5 s, a 21-mode
// DecayLanguage chain 12 s before its first event.  Now (a) every node starts from the kernel's own fourth-order series around
// its neighbour, so one Newton step confirms it instead of four or five finding it; (b) the levels are independent and are
// solved on the host's cores in parallel; (c) |P'| peaks INSIDE the last (uniform) level when the lower limit lies beyond
// the pole of the second Lorentzian, where the estimate of the level count from the end of the range alone is a few per cent
// short: a failure there adds a level instead of doubling the nodes of every level (which cannot help, and ended with no
// table at all after four full attempts).
#define PS_RELBW_TABLE_ALPHA 0.004
#define PS_RELBW_TABLE_MAX_J 40
struct RelbwTableSpec {
    long double kappa, k2, z_lo, z_hi, g_lo, dg;
};

// nodes of level j of one side (0: u = w, 1: u = 1 - w) into out[0 .. M]; false if the series parameter is too large somewhere
bool relbw_table_level(const RelbwTableSpec& s, int side, int j, int J, int mlog, double* out) {
    const int M = 1 << mlog;
    const bool last = j == J;
    const long double edge = ldexpl(1.0L, last ? -(J + 1) : -(j + 1));  // w of node 0
    const long double width = ldexpl(1.0L, last ? -(J + 1) : -(j + 2));
    const long double half_h = 0.5L * s.dg * width / M;
    const long double zr = side ? s.z_hi : s.z_lo;  // the end of the range this half of the uniform approaches
    long double z = 0.0L, g_prev = 0.0L;
    for (int i = 0; i <= M; ++i) {
        const long double w = edge - width * i / M;
        const long double target = (side ? -w : w) * s.dg;  // G(z) - G(z_r)
        long double z0;
        if (i == 0) {  // no neighbour yet: the core is linear in t = atan z
            const long double t_lo = atanl(s.z_lo), t_hi = atanl(s.z_hi);
            z0 = tanl(t_lo + (side ? 1.0L - w : w) * (t_hi - t_lo));
        } else {  // ps_samplers.cuh sample_relbw_table, around the previous node
            const long double y = z + s.kappa, A = z * z + 1.0L, B = y * y + 1.0L, h = (target - g_prev) / s.k2;
            const long double P1 = 2.0L * (z * B + y * A), P2 = 8.0L * z * y + 2.0L * (A + B), P3 = 12.0L * (y + z);
            const long double d = h * A * B, al = h * P1, be = h * P2 * d, ga = h * P3 * d * d;
            z0 = z + d * (1.0L + al / 2.0L + (be + al * al) / 6.0L + (ga + 4.0L * al * be + al * al * al) / 24.0L);
            if (!std::isfinite((double)z0)) z0 = z;
        }
        z = w == 0.0L ? zr : relbw_root(target, zr, s.kappa, s.z_lo, s.z_hi, z0);
        g_prev = target;
        const long double y = z + s.kappa;
        if (half_h * fabsl(2.0L * (z * (y * y + 1.0L) + y * (z * z + 1.0L))) / s.k2 > PS_RELBW_TABLE_ALPHA) return false;
        out[i] = (double)z;
    }
    return true;
}

void build_relbw_table(ps_tree* t, const double sc[5], const double lim[4], double lo, double hi) {
    const long double p = sc[0], q = sc[1], kappa = sc[2];
    if (!(kappa >= 16.0) || !(hi > lo) || !(lim[3] > 0.0)) return;
    RelbwTableSpec s;
    s.kappa = kappa, s.k2 = kappa * kappa + 4.0L, s.z_lo = (lo - p) / q, s.z_hi = (hi - p) / q;
    s.g_lo = lim[2], s.dg = lim[3];  // exactly the doubles the kernel uses
    auto abs_P1 = [&](long double z) {
        const long double y = z + kappa;
        return fabsl(2.0L * (z * (y * y + 1.0L) + y * (z * z + 1.0L)));
    };
    const bool debug = getenv("PS_B200_DEBUG_TABLE") != nullptr;
    for (int mlog = 9; mlog <= 12; ++mlog) {
        const int M = 1 << mlog;
        int J_side[2];
        for (int side = 0; side < 2; ++side) {
            // number of levels: the last one (uniform down to w = 0) must resolve the end of the range
            const long double z_end = side ? s.z_hi : s.z_lo;
            int J = 0;
            while (J < PS_RELBW_TABLE_MAX_J && 0.5L * s.dg * ldexpl(1.0L, -(J + 1) - mlog) * abs_P1(z_end) / s.k2 > PS_RELBW_TABLE_ALPHA) ++J;
            J_side[side] = J;
        }
        for (int attempt = 0; attempt < 4; ++attempt) {
            if (J_side[0] >= PS_RELBW_TABLE_MAX_J || J_side[1] >= PS_RELBW_TABLE_MAX_J) break;
            const int n_levels = J_side[0] + J_side[1] + 2;
            std::vector<double> nodes((size_t)n_levels * (M + 1));
            std::vector<char> level_ok(n_levels, 0);
            std::atomic<int> next{0};
            auto worker = [&]() {
                for (int l = next.fetch_add(1); l < n_levels; l = next.fetch_add(1)) {
                    const int side = l > J_side[0] ? 1 : 0, j = side ? l - J_side[0] - 1 : l;
                    level_ok[l] = relbw_table_level(s, side, j, J_side[side], mlog, nodes.data() + (size_t)l * (M + 1)) ? 1 : 0;
                }
            };
            unsigned n_threads = std::thread::hardware_concurrency();
            n_threads = n_threads < 1 ? 1 : (n_threads > 16 ? 16 : n_threads);
            if ((int)n_threads > n_levels) n_threads = (unsigned)n_levels;
            std::vector<std::thread> pool;
            for (unsigned k = 1; k < n_threads; ++k) pool.emplace_back(worker);
            worker();
            for (std::thread& th : pool) th.join();
            bool ok = true, grow = false;
            for (int l = 0; l < n_levels; ++l) {
                if (level_ok[l]) continue;
                ok = false;
                const int side = l > J_side[0] ? 1 : 0, j = side ? l - J_side[0] - 1 : l;
                if (debug) fprintf(stderr, "relbw node table: 2^%d nodes per level, side %d, level %d of %d refused\n", mlog, side, j, J_side[side]);
                if (j == J_side[side]) {  // the last level is too long: one more dyadic level
                    ++J_side[side];
                    grow = true;
                } else {
                    grow = false;  // a dyadic level needs more nodes: next mlog
                    break;
                }
            }
            if (ok) {
                t->tab_nodes.swap(nodes);
                t->tab_mlog = mlog;
                t->tab_J_lo = J_side[0];
                t->tab_J_hi = J_side[1];
                t->tab_kind = PS_MASS_RELBW;
                return;
            }
            if (!grow) break;
        }
    }
}

// ---------------------------------------------------------------------------------- truncated-normal node table
// ps_samplers.cuh sample_gauss_table: z_k = Phi^-1(Phi(a) + u_k (Phi(b) - Phi(a))) and 1 / phi(z_k) at the same nodes u_k,
// a = (lo - m) / w, b = (hi - m) / w.  Every probability is carried as a LOWER-tail or an UPPER-tail value, whichever is
// the smaller, so that nothing cancels: the table is accurate to the long double's 64 bits even where the fp64 expression
// Phi(a) + u dPhi of the library path has run out of digits (targets next to 1).
const long double PS_SQRT1_2L = 0.707106781186547524400844362104849039L;
const long double PS_SQRT_2PIL = 2.506628274631000502415765284811045253L;

long double norm_lower(long double z) { return 0.5L * erfcl(-z * PS_SQRT1_2L); }  // Phi(z); Phi(-z) is the upper tail

// z with Phi(z) = P for 0 < P <= 1/2: Newton on log Phi (concave and increasing: monotone convergence from the left,
// and from the right after one step), started from a neighbouring node.
long double norm_inv_lower(long double P, long double z) {
    if (!(P > 0.0L)) return -HUGE_VALL;
    if (!(z < 0.5L)) z = 0.0L;
    for (int it = 0; it < 200; ++it) {
        const long double F = norm_lower(z);
        const long double phi = expl(-0.5L * z * z) / PS_SQRT_2PIL;
        if (!(F > 0.0L) || !(phi > 0.0L)) { z *= 0.5L; continue; }  // started beyond the long double's range
        long double step = F / phi * logl(F / P);
        if (step > 4.0L) step = 4.0L;
        if (step < -4.0L) step = -4.0L;
        z -= step;
        // quadratic convergence: a step of 1e-9 leaves an error of ~|z| step^2 / 2, below the 64-bit floor
        if (fabsl(step) <= 1e-9L * (1.0L + fabsl(z))) break;
    }
    return z;
}

// Fills t->tab_* when a table can be offered: the seventh-order term of the inverse-function series stays below
// PS_GAUSS_TABLE_ERR (absolute, in z) at half a node spacing from every node that the kernel may use.
#define PS_GAUSS_TABLE_ERR 2e-16L
#define PS_GAUSS_TABLE_MAX_J 50 /* level 50 is w < 2^-51: below the resolution of the generator's own uniforms */
void build_gauss_table(ps_tree* t, double mass, double width, const double lim[4], double lo, double hi) {
    if (!(hi > lo) || !(lim[1] > 0.0) || !(width > 0.0)) return;
    const long double a = ((long double)lo - mass) / width, b = ((long double)hi - mass) / width;
    const long double P_a = norm_lower(a), Q_a = norm_lower(-a), P_b = norm_lower(b), Q_b = norm_lower(-b);
    // Phi(b) - Phi(a) without cancellation
    const long double dP = (a <= 0.0L && b >= 0.0L) ? (1.0L - P_a - Q_b) : (b < 0.0L ? P_b - P_a : Q_a - Q_b);
    if (!(dP > 0.0L)) return;
    // (z, 1/phi) of the node at distance w from the lower (side 0) or upper (side 1) end of the range, in units of u
    auto node = [&](int side, long double w, long double z_start, long double& z, long double& r) {
        const long double P = side ? P_b - w * dP : P_a + w * dP;  // lower-tail probability of the target
        const long double Q = side ? Q_b + w * dP : Q_a - w * dP;  // upper-tail probability
        if (P <= Q) z = norm_inv_lower(P, z_start);
        else z = -norm_inv_lower(Q, -z_start);
        r = PS_SQRT_2PIL * expl(0.5L * z * z);
    };
    auto series_err = [](long double z, long double d) {  // the seventh-order term of the series in sample_gauss_table
        const long double z2 = z * z, d2 = d * d;
        return (127.0L + z2 * (1740.0L + z2 * (2556.0L + z2 * 720.0L))) / 5040.0L * fabsl(d) * d2 * d2 * d2;
    };
    for (int mlog = 6; mlog <= 12; ++mlog) {
        const int M = 1 << mlog;
        std::vector<double> nodes;
        int J_side[2] = {0, 0};
        long double wmin = 0.0L;
        bool ok = true;
        for (int side = 0; side < 2 && ok; ++side) {
            // the last level is uniform down to w = 0: it must be so short that its end node (z = a or b) serves all of
            // it.  Where that end lies beyond the reach of fp64 (Phi = 0 or 1: z -> infinity as w -> 0) no level does; the
            // levels then go down to the generator's resolution and the last one is left to the library path (tab_wmin).
            int J = 0;
            bool regular_end = false;
            long double z_end, r_end;
            node(side, 0.0L, 0.0L, z_end, r_end);
            if (std::isfinite((double)z_end) && std::isfinite((double)r_end)) {
                for (; J < PS_GAUSS_TABLE_MAX_J; ++J) {
                    const long double half_h = 0.5L * dP * ldexpl(1.0L, -(J + 1) - mlog);
                    if (series_err(z_end, half_h * r_end) <= PS_GAUSS_TABLE_ERR) { regular_end = true; break; }
                }
            }
            if (!regular_end) {
                J = PS_GAUSS_TABLE_MAX_J;
                wmin = std::max(wmin, ldexpl(1.0L, -(J + 1)));
            }
            J_side[side] = J;
            long double z = 0.0L, r = 0.0L, w_prev = 0.5L;
            bool have_prev = false;
            for (int j = 0; j <= J && ok; ++j) {
                const bool last = j == J;
                const long double edge = ldexpl(1.0L, last ? -(J + 1) : -(j + 1));  // w of node 0
                const long double lw = ldexpl(1.0L, last ? -(J + 1) : -(j + 2));    // width of the level in w
                const long double half_h = 0.5L * dP * lw / M;
                for (int i = 0; i <= M; ++i) {
                    const long double w = edge - lw * i / M;
                    if (last && !regular_end) {  // never read (tab_wmin)
                        nodes.push_back(side ? 1e300 : -1e300);
                        nodes.push_back(0.0);
                        continue;
                    }
                    // Newton start: the kernel's own series from the previous node (one Newton step then confirms it)
                    long double z0 = z;
                    if (have_prev) {
                        const long double d = (side ? w_prev - w : w - w_prev) * dP * r, z2 = z * z;
                        z0 = z + d * (1.0L + d * (0.5L * z + d * ((2.0L * z2 + 1.0L) / 6.0L + d * (z * (6.0L * z2 + 7.0L) / 24.0L +
                                 d * ((z2 * (24.0L * z2 + 46.0L) + 7.0L) / 120.0L)))));
                        if (!std::isfinite((double)z0)) z0 = z;
                    }
                    node(side, w, z0, z, r);
                    w_prev = w;
                    have_prev = true;
                    if (!(series_err(z, half_h * r) <= PS_GAUSS_TABLE_ERR)) { ok = false; break; }
                    nodes.push_back((double)z);
                    nodes.push_back((double)r);
                }
            }
        }
        if (ok) {
            t->tab_nodes.swap(nodes);
            t->tab_mlog = mlog;
            t->tab_J_lo = J_side[0];
            t->tab_J_hi = J_side[1];
            t->tab_wmin = (double)wmin;
            t->tab_kind = PS_MASS_GAUSS;
            return;
        }
    }
}

}  // namespace

extern "C" int ps_tree_create(const ps_node_desc* nodes, int32_t n_nodes, ps_tree** out) {
    if (!nodes || !out || n_nodes < 3) return fail(PS_E_INVALID, "a decay tree needs a top particle and at least 2 children");
    if (n_nodes - 1 > PS_MAX_SLOTS) return fail(PS_E_INVALID, "decay tree has more than %d particles", PS_MAX_SLOTS);
    ps_tree* t = new ps_tree();
    t->n_nodes = n_nodes;
    t->nodes.assign(nodes, nodes + n_nodes);
    t->children.resize(n_nodes);
    t->slot.assign(n_nodes, -1);
    t->ucol.assign(n_nodes, -1);
    t->user_col.assign(n_nodes, -1);
    t->min_mass.assign(n_nodes, 0.0);
    for (int i = 0; i < n_nodes; ++i) {
        const int p = nodes[i].parent;
        if (p == -1) {
            if (t->top != -1) { delete t; return fail(PS_E_INVALID, "more than one top particle (parent == -1)"); }
            t->top = i;
        } else if (p < 0 || p >= n_nodes || p == i) {
            delete t;
            return fail(PS_E_INVALID, "node %d has invalid parent %d", i, p);
        } else {
            t->children[p].push_back(i);
        }
        if (nodes[i].kind < PS_MASS_FIXED || nodes[i].kind > PS_MASS_USER) { delete t; return fail(PS_E_INVALID, "node %d has unknown mass kind %d", i, nodes[i].kind); }
        if (is_sampled(nodes[i].kind) && !(nodes[i].width > 0.0)) { delete t; return fail(PS_E_INVALID, "node %d: a sampled resonance needs width > 0", i); }
    }
    if (t->top < 0) { delete t; return fail(PS_E_INVALID, "no top particle (parent == -1)"); }
    if (nodes[t->top].kind != PS_MASS_FIXED) { delete t; return fail(PS_E_INVALID, "Cannot use resonance as top particle"); }
    for (int i = 0; i < n_nodes; ++i)
        if (t->children[i].size() == 1) { delete t; return fail(PS_E_INVALID, "Have to set at least 2 children, not 1 for a particle to decay"); }
    if (t->children[t->top].empty()) { delete t; return fail(PS_E_INVALID, "No children have been configured"); }
    // reachability (rejects cycles / orphans)
    {
        std::vector<int> stack{t->top};
        int seen = 0;
        std::vector<char> mark(n_nodes, 0);
        while (!stack.empty()) {
            int v = stack.back();
            stack.pop_back();
            if (mark[v]) continue;
            mark[v] = 1;
            ++seen;
            for (int c : t->children[v]) stack.push_back(c);
        }
        if (seen != n_nodes) { delete t; return fail(PS_E_INVALID, "decay tree is not connected"); }
    }
    int next_slot = 0, next_col = 0;
    assign_slots(*t, t->top, next_slot, next_col);
    t->n_particles = next_slot;
    t->n_draws = next_col;
    memset(&t->base, 0, sizeof(PsParams));
    t->base.mass[0] = nodes[t->top].mass;
    if (t->children[t->top].size() != 2) t->two_body_only = false;
    for (int i = 0; i < n_nodes; ++i) {
        if (i == t->top) continue;
        const int s1 = t->slot[i] + 1;
        t->base.mass[s1] = nodes[i].mass;
        t->base.width[s1] = nodes[i].width;
        t->min_mass[i] = recurse_stable(*t, i);
        t->base.min_mass[s1] = t->min_mass[i];
        {
            double sc[5];
            host_sampler_constants(nodes[i].kind, nodes[i].mass, nodes[i].width, sc);
            for (int k = 0; k < 5; ++k) t->base.sc[k][s1] = sc[k];
        }
        if (nodes[i].kind == PS_MASS_USER) {
            t->user_col[i] = t->n_user++;
            t->base.user_col[s1] = t->user_col[i];
        }
        if (!t->children[i].empty()) t->has_grandchildren = true;
        if (!t->children[i].empty() && t->children[i].size() != 2) t->two_body_only = false;
        if (t->children[i].empty() && nodes[i].kind != PS_MASS_FIXED) t->resonant_leaf = true;
    }
    // user columns follow slot order so that the Python layer can fill them in generation order
    {
        std::vector<int> by_slot(t->n_particles, -1);
        for (int i = 0; i < n_nodes; ++i)
            if (i != t->top) by_slot[t->slot[i]] = i;
        int col = 0;
        for (int s = 0; s < t->n_particles; ++s) {
            const int i = by_slot[s];
            if (nodes[i].kind == PS_MASS_USER) {
                t->user_col[i] = col;
                t->base.user_col[s + 1] = col++;
            }
        }
    }
    // Event-independent limits of the first non-fixed child of a top particle at rest (phasespace.py:342-352:
    // max_mass = M - sum of the fixed children, min_mass = recurse_stable): see PsParams::lim.
    {
        double fs = 0.0;
        bool first = true;
        int first_nonfixed = -1;
        for (int c : t->children[t->top]) {
            if (nodes[c].kind == PS_MASS_FIXED) {
                fs = first ? nodes[c].mass : fs + nodes[c].mass;
                first = false;
            } else if (first_nonfixed < 0) {
                first_nonfixed = c;
            }
        }
        if (first_nonfixed >= 0 && is_sampled(nodes[first_nonfixed].kind)) {
            const int c = first_nonfixed, s1 = t->slot[c] + 1;
            double sc[5];
            for (int k = 0; k < 5; ++k) sc[k] = t->base.sc[k][s1];
            t->base.lim_slot1 = s1;
            host_sampler_limits(nodes[c].kind, nodes[c].mass, nodes[c].width, sc, t->min_mass[c],
                                nodes[t->top].mass - fs, t->base.lim);
            // PS_B200_NO_NODE_TABLES=1 (A/B measurements): sample with the solver / library inverse instead
            const char* no_tab = getenv("PS_B200_NO_NODE_TABLES");
            const bool tables = !(no_tab && *no_tab && *no_tab != '0');
            if (tables && nodes[c].kind == PS_MASS_RELBW)
                build_relbw_table(t, sc, t->base.lim, t->min_mass[c], nodes[t->top].mass - fs);
            if (tables && nodes[c].kind == PS_MASS_GAUSS)
                build_gauss_table(t, nodes[c].mass, nodes[c].width, t->base.lim, t->min_mass[c], nodes[t->top].mass - fs);
        }
    }
    t->signature = emit_signature(*t, t->top);
    if (!t->resonant_leaf) {
        const double M = nodes[t->top].mass;
        if (t->has_grandchildren) {
            t->w_max = host_recurse_w_max(*t, t->top, M);
        } else {
            std::vector<double> m;
            double msum = 0.0;
            bool first = true;
            for (int c : t->children[t->top]) {
                m.push_back(nodes[c].mass);
                msum = first ? nodes[c].mass : msum + nodes[c].mass;
                first = false;
            }
            t->w_max = host_get_w_max(M - msum, m);
        }
        t->base.wmax_const = t->w_max;
        t->base.inv_wmax_const = 1.0 / t->w_max;
    }
    *out = t;
    return 0;
}

extern "C" void ps_tree_destroy(ps_tree* tree) {
    if (!tree) return;
    if (!tree->tab_dev.empty()) {  // device copies of the node table
        int cur = 0;
        const bool have = cudaGetDevice(&cur) == cudaSuccess;
        for (auto& kv : tree->tab_dev) {
            if (cudaSetDevice(kv.first) == cudaSuccess) cudaFree(kv.second);
        }
        if (have) cudaSetDevice(cur);
        (void)cudaGetLastError();
    }
    delete tree;
}
extern "C" int32_t ps_tree_n_particles(const ps_tree* t) { return t ? t->n_particles : 0; }
extern "C" int32_t ps_tree_slot(const ps_tree* t, int32_t node) { return (t && node >= 0 && node < t->n_nodes) ? t->slot[node] : -1; }
extern "C" int32_t ps_tree_n_draws(const ps_tree* t) { return t ? t->n_draws : 0; }
extern "C" int32_t ps_tree_n_user(const ps_tree* t) { return t ? t->n_user : 0; }
extern "C" int32_t ps_tree_user_column(const ps_tree* t, int32_t node) { return (t && node >= 0 && node < t->n_nodes) ? t->user_col[node] : -1; }
extern "C" double ps_tree_min_mass(const ps_tree* t, int32_t node) { return (t && node >= 0 && node < t->n_nodes) ? t->min_mass[node] : NAN; }
extern "C" double ps_tree_w_max(const ps_tree* t) { return t ? t->w_max : NAN; }
extern "C" const char* ps_tree_signature(const ps_tree* t) { return t ? t->signature.c_str() : ""; }
extern "C" int32_t ps_tree_node_table_size(const ps_tree* t) { return t ? (int32_t)t->tab_nodes.size() : 0; }
extern "C" int32_t ps_tree_node_table(const ps_tree* t, double* out, int32_t capacity, int32_t* meta, double* w_min) {
    if (!t) return 0;
    const int32_t n = (int32_t)t->tab_nodes.size();
    if (meta) {
        meta[0] = n ? t->tab_kind : -1;
        meta[1] = t->tab_mlog;
        meta[2] = t->tab_J_lo;
        meta[3] = t->tab_J_hi;
    }
    if (w_min) *w_min = t->tab_wmin;
    if (out) memcpy(out, t->tab_nodes.data(), sizeof(double) * (size_t)(n < capacity ? n : (capacity > 0 ? capacity : 0)));
    return n;
}
extern "C" int32_t ps_tree_relbw_nodes(const ps_tree* t) {
    return t && t->tab_kind == PS_MASS_RELBW ? (int32_t)t->tab_nodes.size() : 0;
}
extern "C" const char* ps_last_error(void) { return g_error.c_str(); }
extern "C" int32_t ps_abi_version(void) { return PS_ABI_VERSION; }
extern "C" int64_t ps_launch_count(void) { return g_launches.load(); }

extern "C" int32_t ps_tree_is_aot(const ps_tree* t, int32_t with_boost, int32_t with_dbg, uint32_t flags) {
    if (!t) return 0;
    Registry& r = registry();
    std::lock_guard<std::mutex> lock(r.mu);
    return r.aot.count(KernelKey{t->signature, PS_VARIANT_GENERATE, with_boost ? 1 : 0, with_dbg ? 1 : 0, (flags & PS_GEN_EXACT) ? 1 : 0}) ? 1 : 0;
}

// ---------------------------------------------------------------------------------- NVRTC JIT
namespace {

// The kernel headers, embedded at build time (ps_embedded_src.inc is written by build()).
struct EmbeddedHeader {
    const char* name;
    const char* text;
};
#include "ps_embedded_src.inc"

typedef int nvrtcResult_t;
typedef void* nvrtcProgram_t;
struct Nvrtc {
    void* handle = nullptr;
    nvrtcResult_t (*CreateProgram)(nvrtcProgram_t*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult_t (*CompileProgram)(nvrtcProgram_t, int, const char* const*) = nullptr;
    nvrtcResult_t (*GetProgramLogSize)(nvrtcProgram_t, size_t*) = nullptr;
    nvrtcResult_t (*GetProgramLog)(nvrtcProgram_t, char*) = nullptr;
    nvrtcResult_t (*GetCUBINSize)(nvrtcProgram_t, size_t*) = nullptr;
    nvrtcResult_t (*GetCUBIN)(nvrtcProgram_t, char*) = nullptr;
    nvrtcResult_t (*DestroyProgram)(nvrtcProgram_t*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult_t) = nullptr;
    nvrtcResult_t (*Version)(int*, int*) = nullptr;
};

Nvrtc* load_nvrtc() {
    static Nvrtc n;
    static std::once_flag once;  // several threads may ask for their first NVRTC compilation at once (ps_tree_prepare)
    std::call_once(once, [] {
        // The toolkit's own NVRTC first (CUDA 12.9 emits PTX ISA 8.8, needed for 256-bit stores); a process
        // that already holds an older libnvrtc.so.12 (e.g. the one bundled with torch) is found last.
        const char* names[] = {getenv("PS_NVRTC_PATH"), "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so",
                               "libnvrtc.so.12", "libnvrtc.so"};
        void* handle = nullptr;
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (handle) break;
        }
        if (!handle) return;
#define PS_SYM(field, sym) *(void**)(&n.field) = dlsym(handle, sym)
        PS_SYM(CreateProgram, "nvrtcCreateProgram");
        PS_SYM(CompileProgram, "nvrtcCompileProgram");
        PS_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        PS_SYM(GetProgramLog, "nvrtcGetProgramLog");
        PS_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
        PS_SYM(GetCUBIN, "nvrtcGetCUBIN");
        PS_SYM(DestroyProgram, "nvrtcDestroyProgram");
        PS_SYM(GetErrorString, "nvrtcGetErrorString");
        PS_SYM(Version, "nvrtcVersion");
#undef PS_SYM
        if (n.CreateProgram && n.CompileProgram && n.GetCUBIN) n.handle = handle;
    });
    return n.handle ? &n : nullptr;
}

// On-disk cache of NVRTC results: $PS_B200_CACHE_DIR, else $XDG_CACHE_HOME/phasespace_b200, else
// ~/.cache/phasespace_b200; PS_B200_CACHE_DIR="" (empty) disables it.  The file name hashes everything the cubin
// depends on: the generated source, the embedded headers, the compile options and the NVRTC version.
uint64_t fnv1a(const void* data, size_t n, uint64_t h) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

std::string jit_cache_dir() {
    const char* env = getenv("PS_B200_CACHE_DIR");
    std::string dir;
    if (env) {
        if (!*env) return "";
        dir = env;
    } else if (const char* xdg = getenv("XDG_CACHE_HOME"); xdg && *xdg) {
        dir = std::string(xdg) + "/phasespace_b200";
    } else if (const char* home = getenv("HOME"); home && *home) {
        dir = std::string(home) + "/.cache";
        mkdir(dir.c_str(), 0755);
        dir += "/phasespace_b200";
    } else {
        return "";
    }
    mkdir(dir.c_str(), 0755);
    struct stat st;
    return (stat(dir.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) ? dir : "";
}

bool read_file(const std::string& path, std::vector<char>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    bool ok = n > 0;
    if (ok) {
        out.resize((size_t)n);
        ok = fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    }
    fclose(f);
    return ok;
}

void write_file_atomic(const std::string& path, const std::vector<char>& data) {
    static std::atomic<unsigned> serial{0};  // two threads may write (different or, rarely, the same) files at once
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid()) + "_" + std::to_string(serial.fetch_add(1));
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());
}

int load_cubin(const std::vector<char>& cubin, const void** fn_out) {
    cudaLibrary_t lib;
    PS_CUDA(cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaKernel_t kern;
    PS_CUDA(cudaLibraryGetKernel(&kern, lib, "ps_jit_kernel"));
    *fn_out = (const void*)kern;
    return 0;
}

// Specialise ps_kernel.cuh for one tree: returns a cudaKernel_t (usable with cudaLaunchKernel).
int jit_compile(const KernelKey& key, const void** fn_out) {
    Nvrtc* rt = load_nvrtc();
    if (!rt) return fail(PS_E_COMPILE, "tree %s has no ahead-of-time kernel and libnvrtc could not be loaded", key.sig.c_str());
    std::string src = "#define PS_EXACT " + std::to_string(key.exact) + "\n#include \"ps_kernel.cuh\"\nusing psk::Pt;\n";
    if (key.variant == PS_VARIANT_GENERATE) {
        src += "extern \"C\" __global__ void __launch_bounds__(PS_BLOCK, PS_MIN_BLOCKS) ps_jit_kernel(const __grid_constant__ PsParams prm) {\n"
               "  psk::genbod_body<" + key.sig + ", " + (key.boost ? "true" : "false") + ", " + (key.dbg ? "true" : "false") + ">(prm);\n}\n";
    } else if (key.variant == PS_VARIANT_HIST) {
        src += "extern \"C\" __global__ void __launch_bounds__(PS_BLOCK, PS_MIN_BLOCKS) ps_jit_kernel(const __grid_constant__ PsParams prm, const __grid_constant__ psk::PsHistParams h) {\n"
               "  psk::genbod_hist_body<" + key.sig + ", " + (key.boost ? "true" : "false") + ">(prm, h);\n}\n";
    } else {
        src += "extern \"C\" __global__ void __launch_bounds__(PS_BLOCK, PS_MIN_BLOCKS) ps_jit_kernel(const __grid_constant__ PsParams prm, double w_bound, unsigned int* mask, unsigned long long* overflow) {\n"
               "  psk::genbod_mask_body<" + key.sig + ", " + (key.boost ? "true" : "false") + ">(prm, w_bound, mask, overflow);\n}\n";
    }
    std::vector<const char*> hdr_text, hdr_name;
    for (const EmbeddedHeader* h = kEmbeddedHeaders; h->name; ++h) {
        hdr_text.push_back(h->text);
        hdr_name.push_back(h->name);
    }
    std::string cache_path;
    if (const std::string dir = jit_cache_dir(); !dir.empty()) {
        uint64_t h = fnv1a(src.data(), src.size(), 1469598103934665603ull);
        for (const char* t : hdr_text) h = fnv1a(t, strlen(t), h);
        int major = 0, minor = 0;
        if (rt->Version) rt->Version(&major, &minor);
        const std::string extra = "sm_100a;c++17;exact=" + std::to_string(key.exact) + ";nvrtc=" + std::to_string(major) + "." + std::to_string(minor);
        h = fnv1a(extra.data(), extra.size(), h);
        char name[64];
        snprintf(name, sizeof(name), "/ps_%016llx.cubin", (unsigned long long)h);
        cache_path = dir + name;
        std::vector<char> cached;
        if (read_file(cache_path, cached) && load_cubin(cached, fn_out) == 0) return 0;
        (void)cudaGetLastError();  // a stale or truncated file: compile again and overwrite it
    }
    nvrtcProgram_t prog = nullptr;
    nvrtcResult_t rc = rt->CreateProgram(&prog, src.c_str(), "ps_jit.cu", (int)hdr_text.size(), hdr_text.data(), hdr_name.data());
    if (rc != 0) return fail(PS_E_COMPILE, "nvrtcCreateProgram failed: %s", rt->GetErrorString ? rt->GetErrorString(rc) : "?");
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    if (key.exact) opts.push_back("--fmad=false");
    rc = rt->CompileProgram(prog, (int)opts.size(), opts.data());
    if (rc != 0) {
        size_t n = 0;
        std::string log;
        if (rt->GetProgramLogSize && rt->GetProgramLogSize(prog, &n) == 0 && n > 1) {
            log.resize(n);
            rt->GetProgramLog(prog, &log[0]);
        }
        rt->DestroyProgram(&prog);
        return fail(PS_E_COMPILE, "NVRTC failed for tree %s:\n%.1500s", key.sig.c_str(), log.c_str());
    }
    size_t n = 0;
    rt->GetCUBINSize(prog, &n);
    std::vector<char> cubin(n);
    rt->GetCUBIN(prog, cubin.data());
    rt->DestroyProgram(&prog);
    const int rc_load = load_cubin(cubin, fn_out);
    if (rc_load == 0 && !cache_path.empty()) write_file_atomic(cache_path, cubin);
    return rc_load;
}

int sm_count(int* out) {
    static std::atomic<int> cache[64] = {};
    int dev = 0;
    PS_CUDA(cudaGetDevice(&dev));
    int sms = (dev >= 0 && dev < 64) ? cache[dev].load(std::memory_order_relaxed) : 0;
    if (sms == 0) {
        PS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) cache[dev].store(sms, std::memory_order_relaxed);
    }
    *out = sms;
    return 0;
}

int get_kernel_slow(const ps_tree* t, int variant, int boost, int dbg, int exact, const void** fn);

int get_kernel(const ps_tree* t, int variant, int boost, int dbg, int exact, const void** fn) {
    std::atomic<const void*>& slot = t->kernel_cache[variant][boost ? 1 : 0][dbg ? 1 : 0][exact ? 1 : 0];
    const void* cached = slot.load(std::memory_order_acquire);
    if (cached) {
        *fn = cached;
        return 0;
    }
    int rc = get_kernel_slow(t, variant, boost ? 1 : 0, dbg ? 1 : 0, exact ? 1 : 0, fn);
    if (rc == 0) slot.store(*fn, std::memory_order_release);
    return rc;
}

int get_kernel_slow(const ps_tree* t, int variant, int boost, int dbg, int exact, const void** fn) {
    KernelKey key{t->signature, variant, boost, dbg, exact};
    Registry& r = registry();
    {
        std::lock_guard<std::mutex> lock(r.mu);
        auto it = r.aot.find(key);
        if (it != r.aot.end()) { *fn = it->second; return 0; }
        it = r.jit.find(key);
        if (it != r.jit.end()) { *fn = it->second; return 0; }
    }
    // NVRTC takes one to three seconds per tree: compiled OUTSIDE the registry lock, so that the modes of a GenMultiDecay
    // can be prepared by several host threads at once (ps_tree_prepare).  Two threads asking for the same new tree at the
    // same moment both compile it; the first result is kept.
    const void* compiled = nullptr;
    int rc = jit_compile(key, &compiled);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(r.mu);
    *fn = r.jit.emplace(key, compiled).first->second;
    return 0;
}

// Launch shape (see genbod_body): a persistent grid of SM count x resident CTAs with a static tile stride;
// for trees with at most PS_CHUNKED_MAX_P particles (memory-bound) a covering grid with up to 16
// consecutive tiles per CTA.
#ifndef PS_PDL_MAX_EVENTS
#define PS_PDL_MAX_EVENTS (1LL << 21) /* launches of at most this many events opt into programmatic dependent launch */
#endif
// Dynamic shared memory (unused by the kernel) that makes a second CTA of the same kernel not fit on an SM.
#define PS_ONE_CTA_SMEM (116 * 1024)

int pick_grid(const ps_tree* t, const void* fn, int boost, int dbg, int exact, long long n_events, int* grid,
              int* tiles_per_cta, size_t* dyn_smem) {
    *dyn_smem = 0;
    std::atomic<int>& slot = t->blocks_per_sm_cache[PS_VARIANT_GENERATE][boost ? 1 : 0][dbg ? 1 : 0][exact ? 1 : 0];
    int per_sm = slot.load(std::memory_order_relaxed);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PS_BLOCK, 0) != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            per_sm = 2;
        }
        slot.store(per_sm, std::memory_order_relaxed);
    }
    const long long tiles = (n_events + PS_BLOCK - 1) / PS_BLOCK;
    int k = 0;
    long long g;
#ifndef PS_CHUNKED_SAMPLED
#define PS_CHUNKED_SAMPLED 1
#endif
    bool sampled = false;
    for (const ps_node_desc& nd : t->nodes) sampled = sampled || is_sampled(nd.kind);
    if (t->n_particles <= PS_CHUNKED_MAX_P || (t->two_body_only && (PS_CHUNKED_SAMPLED || !sampled))) {  // = use_covering_grid<Tree>()
        k = 16;
        while (k > 1 && tiles / k < (long long)sms * per_sm * 4) k >>= 1;  // keep several waves of CTAs
        while ((tiles + k - 1) / k > 2147483647LL) k <<= 1;
        g = (tiles + k - 1) / k;
    } else {
        // Three particles at rest (the flat three-body decay): HBM-bound, and not short of warps -- with ONE CTA per SM
        // (148 x 256 threads) the stores of the whole chip stay closer together in address and stream 2-3 % faster than
        // with two (same box, same call: 1.52-1.53 against 1.56 ms per 1e8 events in a burst, 1.84 against 1.85 ms at the
        // power limit).  Four bodies and more are issue-bound and lose 3-18 % with one (profiles/r02_ab_resident.log).
        // A grid of one CTA per SM is only that if the hardware cannot double them up (it does when a launch meets the
        // tail of the previous one: 1.69 instead of 1.46 ms), so the launch asks for more than half an SM's shared memory.
        const bool one = t->n_particles <= 3 && !boost && per_sm > 1;
        if (one) {
            std::atomic<unsigned long long>& ready = t->one_cta_ready[dbg ? 1 : 0][exact ? 1 : 0];
            int dev = 0;
            PS_CUDA(cudaGetDevice(&dev));
            const unsigned long long bit = 1ull << (dev & 63);
            if (dev >= 64 || !(ready.load(std::memory_order_acquire) & bit)) {  // a function attribute is per device
                PS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_ONE_CTA_SMEM));
                ready.fetch_or(bit, std::memory_order_release);
            }
            *dyn_smem = PS_ONE_CTA_SMEM;
        }
        g = (long long)sms * (one ? 1 : per_sm);
        if (g > tiles) g = tiles;
    }
    *grid = (int)(g < 1 ? 1 : g);
    *tiles_per_cta = k;
    return 0;
}

bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

int fill_params(const ps_tree* t, PsParams& p, uint64_t seed, uint64_t first_event, const int64_t* event_index,
                int64_t n_events, const double* boost_to, int64_t boost_rows, const double* dbg_u,
                const double* user_masses, double* weights, double* weights_max, double* particles,
                int64_t particle_stride, double* stats, int32_t* err_flag, uint32_t flags) {
    if (!t) return fail(PS_E_INVALID, "null tree");
    if (n_events < 0) return fail(PS_E_INVALID, "n_events must be >= 0");
    if (!err_flag) return fail(PS_E_INVALID, "err_flag must not be NULL");
    if (boost_to && boost_rows != 1 && boost_rows != n_events)
        return fail(PS_E_INVALID, "The number of events requested (%lld) doesn't match the boost_to input size of (%lld, 4)", (long long)n_events, (long long)boost_rows);
    if (boost_to && !aligned32(boost_to)) return fail(PS_E_INVALID, "boost_to must be 32-byte aligned");
    if (t->n_user > 0 && !user_masses) return fail(PS_E_INVALID, "tree has %d user-mass particles but user_masses is NULL", t->n_user);
    p = t->base;
    p.seed = seed;
    for (int r = 0; r < 10; ++r) {
        p.philox_keys[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        p.philox_keys[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    p.first_event = first_event;
    p.n_events = n_events;
    p.event_index = (const long long*)event_index;
    p.boost_to = boost_to;
    p.boost_stride = (boost_to && boost_rows == n_events && n_events != 1) ? 4 : 0;
    p.dbg_u = dbg_u;
    p.dbg_ncols = t->n_draws;
    p.n_user = t->n_user;
    p.user_masses = user_masses;
    p.weights = weights;
    p.weights_max = weights_max;
    p.particles = particles;
    p.particle_stride = particle_stride;
    p.stats = stats;
    p.err_flag = err_flag;
    p.normalize = (flags & PS_GEN_NORMALIZE) ? 1 : 0;
    p.tiles_per_cta = 0;
    p.node_tab = nullptr;
    if (!t->tab_nodes.empty() && !boost_to && n_events > 0) {  // the limits are event independent at rest only
        int dev = 0;
        PS_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(t->tab_mu);
        double*& d = t->tab_dev[dev];
        if (!d) {  // first launch of this tree on this device (not inside a stream capture: warm up first, as for NVRTC)
            const size_t bytes = t->tab_nodes.size() * sizeof(double);
            PS_CUDA(cudaMalloc(&d, bytes));
            PS_CUDA(cudaMemcpy(d, t->tab_nodes.data(), bytes, cudaMemcpyHostToDevice));
        }
        p.node_tab = d;
        p.tab_mlog = t->tab_mlog;
        p.tab_J_lo = t->tab_J_lo;
        p.tab_J_hi = t->tab_J_hi;
        p.tab_wmin = t->tab_wmin;
    }
    return 0;
}

int launch_generate(const ps_tree* t, PsParams& p, uint32_t flags, void* stream) {
    if (p.n_events == 0) return 0;
    if (!p.weights || !p.particles) return fail(PS_E_INVALID, "weights and particles must not be NULL");
    if (!aligned32(p.particles) || (p.particle_stride & 3) || p.particle_stride < 4 * p.n_events)
        return fail(PS_E_INVALID, "particles must be 32-byte aligned with particle_stride a multiple of 4 and >= 4*n_events");
    const void* fn = nullptr;
    const int boost = p.boost_to ? 1 : 0, dbg = p.dbg_u ? 1 : 0, exact = (flags & PS_GEN_EXACT) ? 1 : 0;
    int rc = get_kernel(t, PS_VARIANT_GENERATE, boost, dbg, exact, &fn);
    if (rc) return rc;
    int grid = 1;
    size_t dyn_smem = 0;
    rc = pick_grid(t, fn, boost, dbg, exact, p.n_events, &grid, &p.tiles_per_cta, &dyn_smem);
    if (rc) return rc;
    void* args[] = {&p};
    // Programmatic stream serialisation (ps_kernel.cuh pdl_prologue): this launch's CTAs may become resident while the
    // previous grid's last wave drains.  Only for small calls: measured on a B200 (profiles/r02_ab_pdl.log), 1e6 two-body
    // events 13.7 -> 11.4 us per call (5.2 -> 6.3 TB/s), K* gamma 22.8 -> 21.3, 3e5 events 10.2 -> 8.7; but a memory-bound
    // grid of many waves whose first wave was placed as slots happened to free up streams 2-6 % slower than one dispatched
    // onto an empty chip (K* gamma, 3e6 and 1e7 events), and at 1e8 events there is nothing to gain.
    // PS_B200_NO_PDL=1 launches everything plainly (A/B measurements).
    static const bool pdl = [] {
        const char* e = getenv("PS_B200_NO_PDL");
        return !(e && *e && *e != '0');
    }();
    if (pdl && p.n_events <= PS_PDL_MAX_EVENTS) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(PS_BLOCK);
        cfg.dynamicSmemBytes = dyn_smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PS_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    } else {
        PS_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(PS_BLOCK), args, dyn_smem, (cudaStream_t)stream));
    }
    g_launches.fetch_add(1);
    return 0;
}

}  // namespace

extern "C" int ps_generate(const ps_tree* tree, uint64_t seed, uint64_t first_event, int64_t n_events,
                           const double* boost_to, int64_t boost_rows, const double* debug_uniforms,
                           const double* user_masses, double* weights, double* weights_max, double* particles,
                           int64_t particle_stride, double* stats, int32_t* err_flag, uint32_t flags, void* stream,
                           double weight_scale) {
    PsParams p;
    int rc = fill_params(tree, p, seed, first_event, nullptr, n_events, boost_to, boost_rows, debug_uniforms, user_masses,
                         weights, weights_max, particles, particle_stride, stats, err_flag, flags);
    if (rc) return rc;
    if (weight_scale != 1.0) {
        if (!(weight_scale > 0.0) || !std::isfinite(weight_scale)) return fail(PS_E_INVALID, "weight_scale must be a positive finite number");
        if (!(flags & PS_GEN_NORMALIZE) || boost_to || tree->resonant_leaf)
            return fail(PS_E_INVALID, "weight_scale needs PS_GEN_NORMALIZE and an event-independent maximum weight (no boost_to, no resonant leaves)");
        p.inv_wmax_const = tree->base.inv_wmax_const * weight_scale;  // folded into the constant the kernel multiplies by
        p.wmax_const = tree->base.wmax_const / weight_scale;
    }
    return launch_generate(tree, p, flags, stream);
}

extern "C" int ps_generate_indexed(const ps_tree* tree, uint64_t seed, const int64_t* event_index, int64_t n_events,
                                   const double* boost_to, int64_t boost_rows, const double* user_masses,
                                   double* weights, double* weights_max, double* particles, int64_t particle_stride,
                                   double* stats, int32_t* err_flag, uint32_t flags, void* stream) {
    if (!event_index && n_events > 0) return fail(PS_E_INVALID, "event_index must not be NULL");
    PsParams p;
    int rc = fill_params(tree, p, seed, 0, event_index, n_events, boost_to, boost_rows, nullptr, user_masses, weights,
                         weights_max, particles, particle_stride, stats, err_flag, flags);
    if (rc) return rc;
    return launch_generate(tree, p, flags, stream);
}

extern "C" int ps_tree_prepare(const ps_tree* tree, int32_t with_boost, int32_t with_debug_uniforms, uint32_t flags,
                               int32_t device) {
    if (!tree) return fail(PS_E_INVALID, "null tree");
    if (device >= 0) PS_CUDA(cudaSetDevice(device));
    const void* fn = nullptr;
    return get_kernel(tree, PS_VARIANT_GENERATE, with_boost ? 1 : 0, with_debug_uniforms ? 1 : 0, (flags & PS_GEN_EXACT) ? 1 : 0, &fn);
}

// ---------------------------------------------------------------------------------- host-buffer path
namespace {

// Staging arena of the host-buffer path, cached per device: two chunk buffers, two streams and an error
// flag.  cudaMalloc / cudaFree cost tens of milliseconds per call, as much as generating 1e9 events.
struct HostArena {
    std::mutex mu;  // one ps_generate_host at a time per device
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    double* dev[2] = {nullptr, nullptr};
    size_t capacity = 0;  // doubles per buffer
    int32_t* d_err = nullptr;
};

HostArena* host_arena() {
    static std::mutex mu;
    static std::map<int, HostArena*> arenas;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    HostArena*& a = arenas[dev];
    if (!a) a = new HostArena();
    return a;
}

}  // namespace

extern "C" int ps_generate_host(const ps_tree* tree, uint64_t seed, uint64_t first_event, int64_t n_events,
                                const double* host_boost_to, int64_t boost_rows, double* host_weights,
                                double* host_weights_max, double* host_particles, int64_t chunk_events, uint32_t flags) {
    if (!tree) return fail(PS_E_INVALID, "null tree");
    if (tree->n_user > 0) return fail(PS_E_INVALID, "ps_generate_host does not take user-mass trees");
    if (n_events < 0 || !host_weights || !host_particles) return fail(PS_E_INVALID, "bad host buffers");
    if (host_boost_to && boost_rows != 1 && boost_rows != n_events)
        return fail(PS_E_INVALID, "The number of events requested (%lld) doesn't match the boost_to input size of (%lld, 4)", (long long)n_events, (long long)boost_rows);
    if (n_events == 0) return 0;
    if (chunk_events <= 0) chunk_events = 1 << 23;
    if (chunk_events > n_events) chunk_events = n_events;
    chunk_events = (chunk_events + 7) & ~7LL;
    const int P = tree->n_particles;
    const bool want_wmax = !(flags & PS_GEN_NORMALIZE) && host_weights_max;
    const bool per_event_boost = host_boost_to && boost_rows == n_events && n_events != 1;
    HostArena* arena = host_arena();
    if (!arena) return fail(PS_E_CUDA, "no CUDA device");
    std::lock_guard<std::mutex> lock(arena->mu);
    // per buffer: weights | weights_max | particles (P planes) | boost rows
    const size_t per_slot = (size_t)chunk_events * (size_t)(2 + 4 * P + 4);
    if (!arena->d_err) {
        PS_CUDA(cudaMalloc(&arena->d_err, sizeof(int32_t)));
        for (int b = 0; b < 2; ++b) {
            PS_CUDA(cudaStreamCreateWithFlags(&arena->streams[b], cudaStreamNonBlocking));
            PS_CUDA(cudaEventCreateWithFlags(&arena->done[b], cudaEventDisableTiming));
        }
    }
    if (arena->capacity < per_slot) {
        for (int b = 0; b < 2; ++b) {
            if (arena->dev[b]) PS_CUDA(cudaFree(arena->dev[b]));
            arena->dev[b] = nullptr;
        }
        arena->capacity = 0;
        for (int b = 0; b < 2; ++b) PS_CUDA(cudaMalloc(&arena->dev[b], per_slot * sizeof(double)));
        arena->capacity = per_slot;
    }
    cudaStream_t* streams = arena->streams;
    PS_CUDA(cudaMemsetAsync(arena->d_err, 0, sizeof(int32_t), streams[0]));
    PS_CUDA(cudaEventRecord(arena->done[1], streams[0]));
    PS_CUDA(cudaStreamWaitEvent(streams[1], arena->done[1], 0));  // the flag is cleared before either stream writes it
    int rc = 0;
    int64_t done = 0;
    for (int it = 0; done < n_events; ++it) {
        const int b = it % 2;
        const int64_t n = (n_events - done < chunk_events) ? n_events - done : chunk_events;
        // buffer b was last used two chunks ago: wait (on the host) until its copies have drained
        if (it >= 2) PS_CUDA(cudaEventSynchronize(arena->done[b]));
        double* d_w = arena->dev[b];
        double* d_wmax = d_w + chunk_events;
        double* d_p = d_wmax + chunk_events;
        double* d_boost = d_p + (size_t)4 * P * chunk_events;
        const double* boost_arg = nullptr;
        int64_t rows_arg = 0;
        if (host_boost_to) {
            const size_t bytes = per_event_boost ? (size_t)n * 32 : 32;
            PS_CUDA(cudaMemcpyAsync(d_boost, host_boost_to + (per_event_boost ? 4 * done : 0), bytes, cudaMemcpyHostToDevice, streams[b]));
            boost_arg = d_boost;
            rows_arg = per_event_boost ? n : 1;
        }
        rc = ps_generate(tree, seed, first_event + (uint64_t)done, n, boost_arg, rows_arg, nullptr, nullptr, d_w,
                         want_wmax ? d_wmax : nullptr, d_p, 4 * chunk_events, nullptr, arena->d_err, flags, streams[b], 1.0);
        if (rc) break;
        PS_CUDA(cudaMemcpyAsync(host_weights + done, d_w, (size_t)n * 8, cudaMemcpyDeviceToHost, streams[b]));
        if (want_wmax) PS_CUDA(cudaMemcpyAsync(host_weights_max + done, d_wmax, (size_t)n * 8, cudaMemcpyDeviceToHost, streams[b]));
        for (int s = 0; s < P; ++s)
            PS_CUDA(cudaMemcpyAsync(host_particles + (size_t)s * 4 * n_events + 4 * done, d_p + (size_t)s * 4 * chunk_events,
                                    (size_t)n * 32, cudaMemcpyDeviceToHost, streams[b]));
        PS_CUDA(cudaEventRecord(arena->done[b], streams[b]));
        done += n;
    }
    int32_t h_err = 0;
    cudaError_t e0 = cudaStreamSynchronize(streams[0]), e1 = cudaStreamSynchronize(streams[1]);
    if (rc) return rc;
    if (e0 != cudaSuccess || e1 != cudaSuccess) return fail(PS_E_CUDA, "ps_generate_host: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
    PS_CUDA(cudaMemcpy(&h_err, arena->d_err, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (h_err) return fail(PS_E_FORBIDDEN, "Forbidden decay");
    return 0;
}

// ---------------------------------------------------------------------------------- unweighting
namespace {

// Exclusive scan of the accepted-event counts per tile of PS_BLOCK events into int64 offsets.  A tile's count is the
// population count of its eight mask words (tile_count), so the mask pass needs no CTA-wide reduction.  Three small
// kernels over segments of PS_SCAN_SEG tiles:
//   1. every CTA sums its segment (coalesced loads) and parks the sum in offsets[first tile of the segment];
//   2. one CTA scans those parked sums in place (exclusive) and writes the grand total;
//   3. every CTA reads its segment's offset, scans its counts in shared memory and writes the final offsets
//      (the first tile of a segment gets exactly the value parked there, so no scratch buffer is needed).
#define PS_SCAN_SEG 2048
#define PS_SCAN_THREADS 256
static_assert(PS_BLOCK == 256, "tile_count reads the eight mask words of a tile as two 16-byte vectors");

// accepted events in tile `tile`: mask holds n_words words (16-byte aligned), one bit per event
__device__ __forceinline__ int tile_count(const unsigned int* mask, long long tile, long long n_words) {
    const long long w0 = tile * (PS_BLOCK / 32);
    if (w0 + 8 <= n_words) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(mask + w0));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(mask + w0 + 4));
        return __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
    }
    int c = 0;
    for (long long w = w0; w < n_words; ++w) c += __popc(mask[w]);
    return c;
}

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* warp_sums, long long* block_total) {
    // exclusive scan of one value per thread over a CTA of PS_SCAN_THREADS threads (or 1024 in phase 2)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        long long w = lane < n_warps ? warp_sums[lane] : 0;
        long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += up;
        }
        if (lane < n_warps) warp_sums[lane] = wi - w;  // exclusive prefix of the warp totals
        if (lane == 31) *block_total = wi;
    }
    __syncthreads();
    const long long out = warp_sums[wid] + incl - v;
    __syncthreads();  // warp_sums / block_total may be reused by the caller's next round
    return out;
}

__global__ void __launch_bounds__(PS_SCAN_THREADS) scan_phase1_kernel(const unsigned int* mask, long long n_words, long long n,
                                                                      long long* offsets) {
    __shared__ long long warp_sums[32];
    __shared__ long long total;
    const long long base = (long long)blockIdx.x * PS_SCAN_SEG;
    long long s = 0;
#pragma unroll
    for (int k = 0; k < PS_SCAN_SEG / PS_SCAN_THREADS; ++k) {
        const long long i = base + k * PS_SCAN_THREADS + threadIdx.x;
        if (i < n) s += tile_count(mask, i, n_words);
    }
    (void)block_exclusive_scan(s, warp_sums, &total);
    if (threadIdx.x == 0) offsets[base] = total;
}

__global__ void __launch_bounds__(1024) scan_phase2_kernel(long long* offsets, long long n_segments, long long* total_out) {
    __shared__ long long warp_sums[32];
    __shared__ long long total;
    long long carry = 0;
    for (long long first = 0; first < n_segments; first += blockDim.x) {
        const long long seg = first + threadIdx.x;
        const long long v = seg < n_segments ? offsets[seg * PS_SCAN_SEG] : 0;
        const long long excl = block_exclusive_scan(v, warp_sums, &total);
        if (seg < n_segments) offsets[seg * PS_SCAN_SEG] = carry + excl;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(PS_SCAN_THREADS) scan_phase3_kernel(const unsigned int* mask, long long n_words, long long n,
                                                                      long long* offsets) {
    __shared__ int seg_counts[PS_SCAN_SEG];
    __shared__ long long warp_sums[32];
    __shared__ long long total;
    constexpr int PER = PS_SCAN_SEG / PS_SCAN_THREADS;
    const long long base = (long long)blockIdx.x * PS_SCAN_SEG;
    const long long seg_offset = offsets[base];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const long long i = base + k * PS_SCAN_THREADS + threadIdx.x;
        seg_counts[k * PS_SCAN_THREADS + threadIdx.x] = i < n ? tile_count(mask, i, n_words) : 0;
    }
    __syncthreads();  // also orders the read of offsets[base] before thread 0 overwrites it below
    long long mine = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) mine += seg_counts[threadIdx.x * PER + k];
    long long run = seg_offset + block_exclusive_scan(mine, warp_sums, &total);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const long long i = base + threadIdx.x * PER + k;
        if (i < n) offsets[i] = run;
        run += seg_counts[threadIdx.x * PER + k];
    }
}

// One thread per mask word (32 events): its first output position is the tile's offset plus the accepted events of the
// preceding words of the tile (prefix sum over the eight lanes that share a tile); it then writes the global index of
// each set bit.  One word per thread keeps a few hundred thousand independent loads in flight -- a CTA per tile with one
// dependent load chain per iteration was latency-bound (0.69 ms per 1.3e8 candidates).
__global__ void __launch_bounds__(PS_BLOCK) index_kernel(const unsigned int* mask, const long long* tile_offsets,
                                                         unsigned long long first_event, long long n_words, long long* out) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * PS_BLOCK;
    // the loop bound is rounded up to whole warps: the shuffles below need every lane
    for (long long w0 = (long long)blockIdx.x * PS_BLOCK + (threadIdx.x - lane); w0 < n_words; w0 += stride) {
        const long long w = w0 + lane;
        unsigned int word = w < n_words ? __ldg(mask + w) : 0u;
        const int c = __popc(word);
        int incl = c;
#pragma unroll
        for (int o = 1; o < PS_BLOCK / 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o, PS_BLOCK / 32);
            if ((lane & (PS_BLOCK / 32 - 1)) >= o) incl += up;
        }
        if (word) {
            long long pos = __ldg(tile_offsets + w / (PS_BLOCK / 32)) + (incl - c);
            const unsigned long long ev0 = first_event + (unsigned long long)w * 32ull;
            while (word) {
                const int bit = __ffs(word) - 1;
                word &= word - 1u;
                out[pos++] = (long long)(ev0 + (unsigned long long)bit);
            }
        }
    }
}

}  // namespace

extern "C" int ps_unweight_mask(const ps_tree* tree, uint64_t seed, uint64_t first_event, int64_t n_events, double w_bound,
                                uint32_t* mask, uint64_t* overflow, int32_t* err_flag, uint32_t flags, void* stream) {
    if (!tree || !mask) return fail(PS_E_INVALID, "null argument");
    if (tree->n_user > 0) return fail(PS_E_INVALID, "unweighting does not take user-mass trees");
    if (!(w_bound > 0.0)) return fail(PS_E_INVALID, "w_bound must be > 0");
    PsParams p;
    int rc = fill_params(tree, p, seed, first_event, nullptr, n_events, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                         nullptr, 0, nullptr, err_flag, flags | PS_GEN_NORMALIZE);
    if (rc) return rc;
    if (n_events == 0) return 0;
    const void* fn = nullptr;
    rc = get_kernel(tree, PS_VARIANT_MASK, 0, 0, (flags & PS_GEN_EXACT) ? 1 : 0, &fn);
    if (rc) return rc;
    int sms = 0;
    rc = sm_count(&sms);
    if (rc) return rc;
    const long long tiles = (n_events + PS_BLOCK - 1) / PS_BLOCK;
    int grid = (int)((long long)sms * 8 < tiles ? (long long)sms * 8 : tiles);
    void* args[] = {&p, &w_bound, &mask, &overflow};
    PS_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(PS_BLOCK), args, 0, (cudaStream_t)stream));
    g_launches.fetch_add(1);
    return 0;
}

extern "C" int ps_unweight_scan(const uint32_t* mask, int64_t n_events, int64_t* tile_offsets, int64_t* total, void* stream) {
    if (!mask || !tile_offsets || n_events < 0) return fail(PS_E_INVALID, "null argument");
    if (reinterpret_cast<uintptr_t>(mask) & 15u) return fail(PS_E_INVALID, "mask must be 16-byte aligned");
    const long long n_words = (n_events + 31) / 32, n_tiles = (n_events + PS_BLOCK - 1) / PS_BLOCK;
    const long long n_segments = (n_tiles + PS_SCAN_SEG - 1) / PS_SCAN_SEG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_segments > 0) scan_phase1_kernel<<<(unsigned)n_segments, PS_SCAN_THREADS, 0, st>>>(mask, n_words, n_tiles, (long long*)tile_offsets);
    scan_phase2_kernel<<<1, 1024, 0, st>>>((long long*)tile_offsets, n_segments, (long long*)total);
    if (n_segments > 0) scan_phase3_kernel<<<(unsigned)n_segments, PS_SCAN_THREADS, 0, st>>>(mask, n_words, n_tiles, (long long*)tile_offsets);
    PS_CUDA(cudaGetLastError());
    g_launches.fetch_add(n_segments > 0 ? 3 : 1);
    return 0;
}

extern "C" int ps_unweight_index(const uint32_t* mask, const int64_t* tile_offsets, uint64_t first_event, int64_t n_events,
                                 int64_t* event_index, void* stream) {
    if (!mask || !tile_offsets || !event_index) return fail(PS_E_INVALID, "null argument");
    if (n_events == 0) return 0;
    int sms = 0;
    if (int rc = sm_count(&sms)) return rc;
    const long long n_words = (n_events + 31) / 32;
    const long long ctas = (n_words + PS_BLOCK - 1) / PS_BLOCK;
    const long long grid = (long long)sms * 8 < ctas ? (long long)sms * 8 : ctas;
    index_kernel<<<(int)grid, PS_BLOCK, 0, (cudaStream_t)stream>>>(mask, (const long long*)tile_offsets, first_event, n_words, (long long*)event_index);
    PS_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
}

// ---------------------------------------------------------------------------------- fused histograms
namespace {
struct PsHistParamsHost {  // must match psk::PsHistParams (ps_kernel.cuh)
    int n_bins;
    int pad_;
    double p_lo, p_inv_width, e_lo, e_inv_width, w_inv_width;
    double w_cap, fixed_scale;
    double* hist;
};
}  // namespace

extern "C" int ps_histogram(const ps_tree* tree, uint64_t seed, uint64_t first_event, int64_t n_events,
                            const double* boost_to, int64_t boost_rows, int32_t n_bins, double p_lo, double p_hi,
                            double e_lo, double e_hi, double w_cap, double* hist, int32_t* err_flag, uint32_t flags,
                            void* stream) {
    if (!tree || !hist) return fail(PS_E_INVALID, "null argument");
    if (tree->n_user > 0) return fail(PS_E_INVALID, "ps_histogram does not take user-mass trees");
    if (n_bins < 1 || !(p_hi > p_lo) || !(e_hi > e_lo)) return fail(PS_E_INVALID, "bad histogram binning");
    if (!(w_cap > 0.0) || !std::isfinite(w_cap)) return fail(PS_E_INVALID, "w_cap must be a positive finite number");
    // per cell: one fp64 accumulator and PS_HIST_LIMBS 32-bit limbs (ps_kernel.cuh: genbod_hist_body)
    const size_t smem = (size_t)(4 * tree->n_particles + 1) * (size_t)n_bins * (sizeof(double) + PS_HIST_LIMBS * sizeof(uint32_t));
    if (smem > 200 * 1024) return fail(PS_E_INVALID, "histograms need %zu bytes of shared memory per CTA (limit 200 KiB): use fewer bins", smem);
    const double cap2 = std::exp2(std::ceil(std::log2(w_cap)));  // power of two >= w_cap: the scale below is exact
    PsParams p;
    int rc = fill_params(tree, p, seed, first_event, nullptr, n_events, boost_to, boost_rows, nullptr, nullptr, nullptr,
                         nullptr, nullptr, 0, nullptr, err_flag, flags | PS_GEN_NORMALIZE);
    if (rc) return rc;
    if (n_events == 0) return 0;
    const void* fn = nullptr;
    rc = get_kernel(tree, PS_VARIANT_HIST, boost_to ? 1 : 0, 0, (flags & PS_GEN_EXACT) ? 1 : 0, &fn);
    if (rc) return rc;
    if (smem > 48 * 1024) PS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0, per_sm = 0;
    rc = sm_count(&sms);
    if (rc) return rc;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PS_BLOCK, smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
    }
    const long long tiles = (n_events + PS_BLOCK - 1) / PS_BLOCK;
    const int grid = (int)((long long)sms * per_sm < tiles ? (long long)sms * per_sm : tiles);
    PsHistParamsHost h{n_bins, 0, p_lo, n_bins / (p_hi - p_lo), e_lo, n_bins / (e_hi - e_lo), n_bins / (1.0 + 1e-8),
                       cap2, std::exp2((double)(PS_HIST_LIMBS * PS_HIST_LIMB_BITS)) / cap2, hist};
    void* args[] = {&p, &h};
    PS_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(PS_BLOCK), args, smem, (cudaStream_t)stream));
    g_launches.fetch_add(1);
    return 0;
}

// ---------------------------------------------------------------------------------- small kernels
namespace {

__global__ void __launch_bounds__(PS_BLOCK) sample_mass_kernel(int kind, const __grid_constant__ psk::SamplerConst k,
                                                               const double* lo, const double* hi, long long n,
                                                               unsigned long long seed, unsigned long long first_event,
                                                               double* out) {
    for (long long i = (long long)blockIdx.x * PS_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * PS_BLOCK) {
        const double u = psk::philox_uniform0(seed, first_event + (unsigned long long)i, 2u);
        out[i] = psk::sample_mass<false>(kind, k, nullptr, lo[i], hi[i], u);
    }
}

// kinematics.py:122-149 with per-event zero-boost handling
__global__ void __launch_bounds__(PS_BLOCK) boost_kernel(const double* v, const double* b, long long b_stride, int b_cols,
                                                         long long n, double* out) {
    for (long long i = (long long)blockIdx.x * PS_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * PS_BLOCK) {
        double x = v[4 * i], y = v[4 * i + 1], z = v[4 * i + 2], e = v[4 * i + 3];
        const double bx = b[i * b_stride], by = b[i * b_stride + 1], bz = b[i * b_stride + 2];
        (void)b_cols;
        const double b2 = __dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz));
        if (b2 != 0.0) {
            const double gamma = 1.0 / sqrt(__dadd_rn(1.0, -b2));
            const double gamma2 = __dadd_rn(gamma, -1.0) / b2;
            const double bp = __dadd_rn(__dadd_rn(__dmul_rn(x, bx), __dmul_rn(y, by)), __dmul_rn(z, bz));
            const double t = __dadd_rn(__dmul_rn(gamma2, bp), __dmul_rn(gamma, e));
            x = __dadd_rn(x, __dmul_rn(t, bx));
            y = __dadd_rn(y, __dmul_rn(t, by));
            z = __dadd_rn(z, __dmul_rn(t, bz));
            e = __dmul_rn(gamma, __dadd_rn(e, bp));
        }
        out[4 * i] = x;
        out[4 * i + 1] = y;
        out[4 * i + 2] = z;
        out[4 * i + 3] = e;
    }
}

// kinematics.py:93-104
__global__ void __launch_bounds__(PS_BLOCK) mass_kernel(const double* v, long long n, double* out) {
    for (long long i = (long long)blockIdx.x * PS_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * PS_BLOCK) {
        const double x = v[4 * i], y = v[4 * i + 1], z = v[4 * i + 2], e = v[4 * i + 3];
        double s = -__dmul_rn(x, x);
        s = __dadd_rn(s, -__dmul_rn(y, y));
        s = __dadd_rn(s, -__dmul_rn(z, z));
        s = __dadd_rn(s, __dmul_rn(e, e));
        out[i] = sqrt(s);
    }
}

// numpy.random.choice(centres, p=probabilities): cdf.searchsorted(u, side="right")
__device__ __forceinline__ int choice_index(const double* cdf, int n, double u) {
    int lo = 0, hi = n;  // first index with cdf[index] > u
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    return lo < n ? lo : n - 1;
}

// tests/helpers/rapidsim.py:23-51
__global__ void __launch_bounds__(PS_BLOCK) fonll_boost_kernel(const double* pt_centers, const double* pt_cdf, int n_pt,
                                                               const double* eta_centers, const double* eta_cdf, int n_eta,
                                                               double mass, double pt_scale, unsigned long long seed,
                                                               unsigned long long first_event, long long n, double* out) {
    for (long long i = (long long)blockIdx.x * PS_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * PS_BLOCK) {
        const unsigned long long ev = first_event + (unsigned long long)i;
        const double u_pt = psk::philox_uniform_col<0>(seed, ev, 4u), u_eta = psk::philox_uniform_col<1>(seed, ev, 4u);
        const double u_phi = psk::philox_uniform_col<2>(seed, ev, 4u);
        const double pt = __dmul_rn(pt_scale, fabs(pt_centers[choice_index(pt_cdf, n_pt, u_pt)]));
        const double eta = eta_centers[choice_index(eta_cdf, n_eta, u_eta)];
        double s, c;
        sincos(__dmul_rn(u_phi, 2.0 * 3.141592653589793), &s, &c);
        const double px = __dmul_rn(pt, c), py = __dmul_rn(pt, s), pz = __dmul_rn(pt, sinh(eta));
        double e2 = __dmul_rn(px, px);
        e2 = __dadd_rn(e2, __dmul_rn(py, py));
        e2 = __dadd_rn(e2, __dmul_rn(pz, pz));
        e2 = __dadd_rn(e2, __dmul_rn(mass, mass));
        reinterpret_cast<double2*>(out + 4 * i)[0] = make_double2(px, py);
        reinterpret_cast<double2*>(out + 4 * i)[1] = make_double2(pz, sqrt(e2));
    }
}

int small_grid(long long n) {
    long long g = (n + PS_BLOCK - 1) / PS_BLOCK;
    if (g > 148 * 16) g = 148 * 16;
    return g < 1 ? 1 : (int)g;
}

}  // namespace

extern "C" int ps_sample_mass(int32_t kind, double mass, double width, const double* min_mass, const double* max_mass,
                              int64_t n_events, uint64_t seed, uint64_t first_event, double* out, void* stream) {
    if (!is_sampled(kind)) return fail(PS_E_INVALID, "ps_sample_mass: kind must be GAUSS, BW or RELBW");
    if (!(width > 0.0)) return fail(PS_E_INVALID, "ps_sample_mass: width must be > 0");
    if (!min_mass || !max_mass || !out || n_events < 0) return fail(PS_E_INVALID, "null argument");
    if (n_events == 0) return 0;
    double sc[5];
    host_sampler_constants(kind, mass, width, sc);
    const psk::SamplerConst k{mass, width, sc[0], sc[1], sc[2], sc[3], sc[4]};
    sample_mass_kernel<<<small_grid(n_events), PS_BLOCK, 0, (cudaStream_t)stream>>>(kind, k, min_mass, max_mass, n_events, seed, first_event, out);
    PS_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
}

extern "C" int ps_fonll_boost(const double* pt_centers, const double* pt_cdf, int32_t n_pt, const double* eta_centers,
                              const double* eta_cdf, int32_t n_eta, double mass, double pt_scale, uint64_t seed,
                              uint64_t first_event, int64_t n_events, double* out, void* stream) {
    if (!pt_centers || !pt_cdf || !eta_centers || !eta_cdf || !out || n_events < 0) return fail(PS_E_INVALID, "null argument");
    if (n_pt < 1 || n_eta < 1) return fail(PS_E_INVALID, "ps_fonll_boost: empty spectrum");
    if (!aligned32(out)) return fail(PS_E_INVALID, "ps_fonll_boost: out must be 32-byte aligned");
    if (n_events == 0) return 0;
    fonll_boost_kernel<<<small_grid(n_events), PS_BLOCK, 0, (cudaStream_t)stream>>>(pt_centers, pt_cdf, n_pt, eta_centers, eta_cdf, n_eta, mass, pt_scale, seed, first_event, n_events, out);
    PS_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
}

extern "C" int ps_lorentz_boost(const double* vectors, const double* boost, int64_t boost_rows, int32_t boost_cols,
                                int64_t n, double* out, void* stream) {
    if (!vectors || !boost || !out || n < 0) return fail(PS_E_INVALID, "null argument");
    if (boost_cols != 3 && boost_cols != 4) return fail(PS_E_INVALID, "boost vector must have 3 or 4 columns");
    if (boost_rows != 1 && boost_rows != n) return fail(PS_E_INVALID, "boost rows must be 1 or n");
    if (n == 0) return 0;
    boost_kernel<<<small_grid(n), PS_BLOCK, 0, (cudaStream_t)stream>>>(vectors, boost, (boost_rows == n && n != 1) ? boost_cols : 0, boost_cols, n, out);
    PS_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
}

extern "C" int ps_mass(const double* vectors, int64_t n, double* out, void* stream) {
    if (!vectors || !out || n < 0) return fail(PS_E_INVALID, "null argument");
    if (n == 0) return 0;
    mass_kernel<<<small_grid(n), PS_BLOCK, 0, (cudaStream_t)stream>>>(vectors, n, out);
    PS_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
}
