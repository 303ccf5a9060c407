/* Kernel parameter block shared by the AOT kernels, the NVRTC-compiled kernels and the host
 * launcher.  Plain C so that NVRTC can include it without any system header. */
#ifndef PS_PARAMS_H_
#define PS_PARAMS_H_

/* trees with at most this many particles are memory-bound and use the covering (chunked) grid */
#ifndef PS_CHUNKED_MAX_P
#define PS_CHUNKED_MAX_P 2
#endif
#define PS_MAX_SLOTS 40 /* particles in one decay tree, excluding the top */

/* mass-model kinds (mirrors include/phasespace_b200.h) */
#define PS_KIND_FIXED 0
#define PS_KIND_GAUSS 1 /* truncated normal: zfit Gauss / tfp TruncatedNormal (mass_functions.py:16-40) */
#define PS_KIND_BW 2    /* Cauchy, gamma = width (mass_functions.py:43-67) */
#define PS_KIND_RELBW 3 /* relativistic Breit-Wigner density in the mass (mass_functions.py:70-102) */
#define PS_KIND_USER 4  /* per-event masses supplied by the caller (arbitrary Python callables) */

/* fused histograms (ps_kernel.cuh: genbod_hist_body): fixed-point weight = PS_HIST_LIMBS limbs of PS_HIST_LIMB_BITS bits */
#ifndef PS_HIST_LIMBS
#define PS_HIST_LIMBS 3
#endif
#ifndef PS_HIST_LIMB_BITS
#define PS_HIST_LIMB_BITS 20
#endif

#define PS_ERR_FORBIDDEN 1 /* available_mass < 0 or NaN: "Forbidden decay" (phasespace.py:357-363) */

typedef struct PsParams {
    /* index = slot + 1; entry 0 is the top particle */
    double mass[PS_MAX_SLOTS + 1];     /* fixed mass, or pole mass of a resonance */
    double width[PS_MAX_SLOTS + 1];    /* resonance width (0 for fixed) */
    double min_mass[PS_MAX_SLOTS + 1]; /* recurse_stable(): sum of fixed masses below (phasespace.py:326-333) */
    int user_col[PS_MAX_SLOTS + 1];    /* column in user_masses for PS_KIND_USER */
    /* sampler constants derived on the host from (mass, width) (ps_api.cu: fill_sampler_constants).
     *   GAUSS, BW: sc[0] = 1/width
     *   RELBW:     sc[0] = p, sc[1] = q (roots of the quartic, ps_samplers.cuh), sc[2] = kappa = 2p/q,
     *              sc[3] = 1/kappa, sc[4] = 1/(kappa^2 + 4) */
    double sc[5][PS_MAX_SLOTS + 1];
    /* The first non-fixed child of the top particle, when it is a sampled resonance and the top decays at rest,
     * has event-independent mass limits [min_mass, M - fixed sum]: everything of its sampler that depends only
     * on the limits is evaluated once on the host.  lim_slot1 = its slot + 1, or 0 if there is none.
     *   BW:    lim[0] = atan((lo-m)/w), lim[1] = atan((hi-m)/w) - lim[0]
     *   GAUSS: lim[0] = Phi(a), lim[1] = Phi(b) - Phi(a), lim[2] = 1 if mirrored (ps_samplers.cuh) else 0
     *   RELBW: lim[0] = t_lo, lim[1] = t_hi - t_lo, lim[2] = G(lo), lim[3] = G(hi) - G(lo) */
    int lim_slot1;
    int lim_pad_;
    double lim[4];

    /* The same resonance when it is a relativistic Breit-Wigner or a truncated normal: nodes of its inverse CDF
     * (ps_samplers.cuh sample_relbw_table / sample_gauss_table; built by ps_tree_create, one device copy per GPU).
     * node_tab = NULL: the iterative solver / library inverse.
     * Layout: lower half of the uniform (u < 1/2: w = u) then upper half (w = 1 - u); per half J+1 dyadic levels in w
     * (level j < J: w in [2^-(j+2), 2^-(j+1)), level J: w in [0, 2^-(J+1))), per level M+1 nodes uniform in w.
     * A node is one double (RELBW: z) or two (GAUSS: z, 1/phi(z)).  tab_wmin (GAUSS): draws with min(u, 1-u) below it
     * take the library path (0 unless an end of the range lies where Phi is 0 or 1 in fp64). */
    const double* node_tab;
    int tab_mlog, tab_J_lo, tab_J_hi, tab_pad_; /* M = 2^tab_mlog */
    double tab_wmin;

    double wmax_const;     /* event-independent maximum weight (rest frame, fixed leaf masses) */
    double inv_wmax_const; /* its reciprocal */

    unsigned long long seed;
    unsigned int philox_keys[20];   /* round keys of Philox4x32-10 for this seed: (lo, hi) + r (0x9E3779B9, 0xBB67AE85) */
    unsigned long long first_event; /* global index of local event 0: the Philox counter */
    long long n_events;

    const long long* event_index; /* optional: local event i generates global event event_index[i] */
    const double* boost_to;       /* optional (n,4) or (1,4) [px,py,pz,E] */
    long long boost_stride;       /* 4 for per-event rows, 0 to broadcast one row */
    const double* dbg_u;          /* debug mode: (n_events, dbg_ncols) uniforms in reference draw order */
    int dbg_ncols;
    int n_user;
    const double* user_masses; /* (n_events, n_user) */

    double* weights;           /* (n) */
    double* weights_max;       /* (n) or NULL */
    double* particles;         /* (P, particle_stride/4, 4) */
    long long particle_stride; /* doubles between two particle planes */
    double* stats;             /* NULL or [max w_out, sum w_out, sum w_out^2, max w_max] */
    int* err_flag;
    int normalize; /* 1: weights = w / w_max ; 0: weights = w and weights_max = w_max */
    int tiles_per_cta; /* chunked grid only: CTA b takes tiles [b*K, (b+1)*K) */
} PsParams;

#endif
