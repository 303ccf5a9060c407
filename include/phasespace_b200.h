/* phasespace_b200 -- C ABI of the B200-native GENBOD event generator.
 *
 * The reference (zfit/phasespace) is pure Python on TensorFlow and has no FFI of its own; its
 * boundary for this path is the Python API.  This header is the C-ABI layer
 * underneath the Python mirror of that API: plain pointers and sizes, no torch types.  Every entry
 * cites the reference interface it replaces (paths relative to /root/reference/src/phasespace).
 *
 * Ownership: every device buffer is allocated and owned by the caller; the library keeps no global
 * state beyond immutable ps_tree objects, a cache of compiled kernels and (ps_generate_host only) a
 * per-device staging arena of two chunk buffers.  Calls are asynchronous on
 * the stream passed in (a CUstream / cudaStream_t handle; 0 = default stream) and thread-safe.
 * All functions return 0 on success or a negative PS_E_* code; ps_last_error() (thread local)
 * describes the failure.  4-vectors are [px, py, pz, E] (kinematics.py:32,45), fp64 throughout.
 */
#ifndef PHASESPACE_B200_H_
#define PHASESPACE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_ABI_VERSION 4

/* mass model of a particle (GenParticle(name, mass): phasespace.py:100-119) */
#define PS_MASS_FIXED 0 /* float mass: has_fixed_mass (phasespace.py:186-189) */
#define PS_MASS_GAUSS 1 /* truncated normal: mass_functions.py:16-40, tests/helpers/decays.py:26-37 */
#define PS_MASS_BW 2    /* Cauchy with gamma = width: mass_functions.py:43-67 */
#define PS_MASS_RELBW 3 /* relativistic Breit-Wigner: mass_functions.py:70-102 */
#define PS_MASS_USER 4  /* arbitrary callable, pre-evaluated by the caller into user_masses */

#define PS_E_INVALID -1   /* bad argument / malformed tree (ValueError in the reference) */
#define PS_E_CUDA -2      /* CUDA runtime failure */
#define PS_E_COMPILE -3   /* NVRTC failed to specialise the kernel for this tree */
#define PS_E_FORBIDDEN -4 /* kinematically forbidden decay (phasespace.py:357-363) */

/* generate() flags */
#define PS_GEN_NORMALIZE 1u /* weights = w / w_max (normalize_weights=True, phasespace.py:713-717) */
#define PS_GEN_EXACT 2u     /* reference operation order, no FMA (parity / debug variant) */

/* One particle of a decay tree.  The top particle has parent = -1; the children of a particle are
 * the nodes naming it as parent, in array order (GenParticle.set_children, phasespace.py:191-221). */
typedef struct ps_node_desc {
    int32_t parent; /* index into the node array, -1 for the top particle */
    int32_t kind;   /* PS_MASS_* */
    double mass;    /* fixed mass, or pole mass of a resonance */
    double width;   /* resonance width (ignored for PS_MASS_FIXED / PS_MASS_USER) */
} ps_node_desc;

typedef struct ps_tree ps_tree; /* immutable, opaque */

/* Replaces the tree bookkeeping done at trace time by GenParticle._recursive_generate
 * (phasespace.py:497-626): validates the tree (>= 2 children per decaying particle,
 * phasespace.py:213-216; fixed-mass top, phasespace.py:543), assigns output slots in the reference's
 * dict order (phasespace.py:547-570), the uniform draw order (phasespace.py:344-353,367,444-451),
 * recurse_stable() minimum masses (phasespace.py:326-333) and, when it is event independent, the
 * analytic maximum weight (phasespace.py:288-298, 571-625). */
int ps_tree_create(const ps_node_desc* nodes, int32_t n_nodes, ps_tree** out);
void ps_tree_destroy(ps_tree* tree);

/* Number of particles in the output (every node except the top) and, for node i, its plane in the
 * particles buffer (-1 for the top).  Slot order == key order of the reference's output dict. */
int32_t ps_tree_n_particles(const ps_tree* tree);
int32_t ps_tree_slot(const ps_tree* tree, int32_t node);
/* Uniform draws consumed per event (3n-4 per decaying node + one per sampled resonance). */
int32_t ps_tree_n_draws(const ps_tree* tree);
/* Number of PS_MASS_USER particles and the user_masses column of node i (-1 if not USER). */
int32_t ps_tree_n_user(const ps_tree* tree);
int32_t ps_tree_user_column(const ps_tree* tree, int32_t node);
/* recurse_stable() of node i: the min_mass handed to its mass function (phasespace.py:349). */
double ps_tree_min_mass(const ps_tree* tree, int32_t node);
/* Event-independent w_max, or NaN when it depends on the event (boost_to / resonant leaves). */
double ps_tree_w_max(const ps_tree* tree);
/* Canonical C++ type of the tree (the kernel template argument) and whether a kernel for it was
 * compiled ahead of time (otherwise the first ps_generate JIT-compiles it with NVRTC). */
const char* ps_tree_signature(const ps_tree* tree);
/* Nodes of the inverse-CDF table built for a relativistic Breit-Wigner resonance whose mass limits are the same for every
 * event (the first resonance below a top particle at rest), 0 if the tree has none: with a table that resonance is sampled
 * without transcendental functions (ps_samplers.cuh sample_relbw_table; mass_functions.py:70-102). */
int32_t ps_tree_relbw_nodes(const ps_tree* tree);
/* Size in doubles of the inverse-CDF node table of the tree, whichever sampler it serves -- the relativistic Breit-Wigner
 * above (one double per node) or a truncated normal (gauss_factory, mass_functions.py:16-40, and the tfp TruncatedNormal of
 * tests/helpers/decays.py:26-37: two doubles per node, ps_samplers.cuh sample_gauss_table) -- 0 if it has none. */
int32_t ps_tree_node_table_size(const ps_tree* tree);
/* Diagnostic copy of that table (host memory; tests/test_cabi.py checks the nodes against a multi-precision inverse CDF):
 * returns its size in doubles and copies min(size, capacity) of them to `out` (may be NULL); meta = {mass kind it serves
 * or -1, log2 of the nodes per level M, levels - 1 of the lower half, of the upper half} (layout: ps_params.h), *w_min =
 * distance from an end of the uniform's range below which the kernel leaves a draw to the library inverse. */
int32_t ps_tree_node_table(const ps_tree* tree, double* out, int32_t capacity, int32_t* meta, double* w_min);
int32_t ps_tree_is_aot(const ps_tree* tree, int32_t with_boost, int32_t with_debug_uniforms, uint32_t flags);
/* Resolves the generator kernel of the tree ahead of its first ps_generate -- a registry lookup, or for a tree without an
 * ahead-of-time kernel the NVRTC compilation (one to three seconds; cached in the process and on disk) -- on `device`
 * (>= 0: made current in the calling thread first; < 0: the current device).  Thread-safe and concurrent: the modes of a
 * GenMultiDecay (genmultidecay.py:155-197: one GenParticle.generate, i.e. one tf.function trace, per mode) are prepared
 * by several host threads at once instead of one after the other inside the first generate(). */
int ps_tree_prepare(const ps_tree* tree, int32_t with_boost, int32_t with_debug_uniforms, uint32_t flags, int32_t device);

/* Replaces GenParticle.generate / _recursive_generate / _generate / _generate_part2
 * (phasespace.py:628-717, 497-626, 300-397, 399-495) and the rng.uniform calls of random.py:
 * generates events [first_event, first_event + n_events) of the stream `seed` into device buffers.
 *   boost_to       NULL (rest frame, phasespace.py:537-541) or device (n_events,4) with boost_rows =
 *                  n_events, or (1,4) with boost_rows = 1 (phasespace.py:534-535, 697-703); 32-byte aligned
 *   debug_uniforms NULL (in-kernel Philox4x32-10 keyed on (seed, event index)) or device
 *                  (n_events, ps_tree_n_draws) uniforms in the reference's draw order (parity mode)
 *   user_masses    device (n_events, ps_tree_n_user) for PS_MASS_USER particles, else NULL
 *   weights        device (n_events)
 *   weights_max    device (n_events) or NULL; written only without PS_GEN_NORMALIZE
 *   particles      device; plane s (slot s) starts at particles + s*particle_stride doubles and holds
 *                  (n_events,4) rows; particle_stride >= 4*n_events and a multiple of 4; 32-byte aligned
 *   stats          NULL or device double[4], caller-zeroed, accumulated atomically:
 *                  [max w_out, sum w_out, sum w_out^2, max w_max]
 *   err_flag       device int, caller-zeroed; set to 1 on a forbidden decay
 *   weight_scale   1.0, or (with PS_GEN_NORMALIZE and an event-independent w_max only) a factor multiplied into the
 *                  normalised weights at no cost: GenMultiDecay.generate divides the weights of every decay mode by the
 *                  largest maximum weight of all modes (genmultidecay.py:192-195) and passes w_max_mode / w_max_all here
 *                  instead of generating unnormalised weights and making two more passes over them */
int ps_generate(const ps_tree* tree, uint64_t seed, uint64_t first_event, int64_t n_events, const double* boost_to,
                int64_t boost_rows, const double* debug_uniforms, const double* user_masses, double* weights,
                double* weights_max, double* particles, int64_t particle_stride, double* stats, int32_t* err_flag,
                uint32_t flags, void* stream, double weight_scale);

/* As ps_generate but local event i is global event event_index[i] (device int64 array): the second
 * pass of unweighting regenerates exactly the accepted events. */
int ps_generate_indexed(const ps_tree* tree, uint64_t seed, const int64_t* event_index, int64_t n_events,
                        const double* boost_to, int64_t boost_rows, const double* user_masses, double* weights,
                        double* weights_max, double* particles, int64_t particle_stride, double* stats,
                        int32_t* err_flag, uint32_t flags, void* stream);

/* Host-buffer end-to-end call (what a numpy user of the reference gets back): generates in chunks on
 * the device and streams results into HOST memory (pinned memory overlaps copy and compute).
 * Synchronous.  host_particles plane s starts at host_particles + s*4*n_events. */
int ps_generate_host(const ps_tree* tree, uint64_t seed, uint64_t first_event, int64_t n_events,
                     const double* host_boost_to, int64_t boost_rows, double* host_weights, double* host_weights_max,
                     double* host_particles, int64_t chunk_events, uint32_t flags);

/* Accept-reject unweighting, deterministic and ordered by event index:
 *   ps_unweight_mask   recomputes the weights of [first_event, first_event+n_events) (no 4-vectors are
 *                      formed or stored), draws one extra uniform per event (Philox stream 1) and sets
 *                      bit (i%32) of mask[i/32] when u * w_bound < w / w_max; mask holds (n_events+31)/32
 *                      words and every one of them is written (bits past n_events are 0).
 *                      overflow (device uint64, caller-zeroed, may be NULL) is incremented once per candidate
 *                      with w / w_max > w_bound: those are accepted with probability 1 instead of > 1, so a
 *                      non-zero count means w_bound was too small and the accepted sample is biased.
 *   ps_unweight_scan   exclusive prefix sum of the accepted events per tile of 256 events (the population
 *                      counts of the tile's eight mask words): tile_offsets[t] (device, (n_events+255)/256
 *                      entries) = accepted events before tile t, *total (device) = all accepted events.
 *                      mask must be 16-byte aligned.
 *   ps_unweight_index  turns mask + tile_offsets into the sorted list of accepted global event indices
 *                      (feed to ps_generate_indexed). */
int ps_unweight_mask(const ps_tree* tree, uint64_t seed, uint64_t first_event, int64_t n_events, double w_bound,
                     uint32_t* mask, uint64_t* overflow, int32_t* err_flag, uint32_t flags, void* stream);
int ps_unweight_scan(const uint32_t* mask, int64_t n_events, int64_t* tile_offsets, int64_t* total, void* stream);
int ps_unweight_index(const uint32_t* mask, const int64_t* tile_offsets, uint64_t first_event, int64_t n_events,
                      int64_t* event_index, void* stream);

/* Fused weighted histograms -- the consumer of the reference's physics tests (tests/test_physics.py:96-107,
 * tests/helpers/plotting.py:12-15): generates events [first_event, first_event + n_events) exactly as ps_generate
 * would, but instead of storing them accumulates, per particle slot s, weighted histograms (weight = w / w_max) of
 * px, py, pz on [p_lo, p_hi) and of E on [e_lo, e_hi), plus an unweighted histogram of w / w_max on [0, 1+1e-8).
 *   w_cap expected upper bound of the normalised weights (1 is always safe; a tighter bound, e.g. a few times the
 *         largest weight of a pilot sample, keeps more of a many-body decay's tiny weights on the fast fixed-point
 *         path); weights outside [w_cap 2^-26, w_cap) are still accumulated correctly, only more slowly.
 *   hist  device double[(4 P + 1) * n_bins], caller-zeroed, ADDED to: row 4 s + k = coordinate k of slot s,
 *         last row = weights.  Accumulated per CTA in exact fixed point (ps_kernel.cuh); only the final per-CTA
 *         fp64 adds into hist are order dependent. */
int ps_histogram(const ps_tree* tree, uint64_t seed, uint64_t first_event, int64_t n_events, const double* boost_to,
                 int64_t boost_rows, int32_t n_bins, double p_lo, double p_hi, double e_lo, double e_hi, double w_cap,
                 double* hist, int32_t* err_flag, uint32_t flags, void* stream);

/* Stand-alone resonance sampler, the device implementation behind the callables of
 * fromdecay/mass_functions.py:31-38, 58-65, 89-100: out[i] ~ kind(mass, width) within
 * [min_mass[i], max_mass[i]], one Philox draw per event (stream 2). */
int ps_sample_mass(int32_t kind, double mass, double width, const double* min_mass, const double* max_mass,
                   int64_t n_events, uint64_t seed, uint64_t first_event, double* out, void* stream);

/* kinematics.py helpers on device (N,4) arrays: lorentz_boost (122-149), mass (93-104). */
int ps_lorentz_boost(const double* vectors, const double* boost, int64_t boost_rows, int32_t boost_cols,
                     int64_t n, double* out, void* stream);
int ps_mass(const double* vectors, int64_t n, double* out, void* stream);

/* FONLL-style boost_to generator, the input-side caller of the path in the reference's tests
 * (tests/helpers/rapidsim.py:23-51, generate_fonll): per event pT and eta are drawn from binned spectra (bin centres
 * with probabilities, numpy.random.choice semantics = inverse CDF with side="right"), phi uniformly in [0, 2 pi);
 * out row i = [pT cos phi, pT sin phi, pT sinh eta, sqrt(p^2 + mass^2)] with pT = pt_scale * |centre| -- directly usable
 * as boost_to.  Three Philox draws per event (stream 4, columns 0-2).
 *   pt_centers, pt_cdf   device double[n_pt]: bin centres and the normalised cumulative sum of the bin contents
 *   eta_centers, eta_cdf device double[n_eta]
 *   out                  device double[n_events * 4], 32-byte aligned */
int ps_fonll_boost(const double* pt_centers, const double* pt_cdf, int32_t n_pt, const double* eta_centers,
                   const double* eta_cdf, int32_t n_eta, double mass, double pt_scale, uint64_t seed, uint64_t first_event,
                   int64_t n_events, double* out, void* stream);

const char* ps_last_error(void);
int32_t ps_abi_version(void);
/* Kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t ps_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
