/* Scalar C restatement of the GENBOD / Raubold-Lynch n-body generator.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path never does.
 *
 * Role: (1) regenerates the Kolmogorov-Smirnov reference samples that the reference's
 * tests/test_physics.py:44-80 reads from data/bto{2,3,4}pi.root (files stripped from the tree;
 * they were produced by ROOT's TGenPhaseSpace through scripts/prepare_test_samples.cxx:13-112:
 * B0 5279.58 -> n x pi 139.57018, weight = TGenPhaseSpace::Generate() = prod(pd)/wtmax);
 * (2) is the compiled all-cores CPU baseline, the analogue of
 * benchmark/bench_tgenphasespace.cxx:9-21.
 *
 * PARITY STATUS: parity unpinned against ROOT (ROOT is absent).  The algorithm is the published
 * one (F. James, "Monte Carlo Phase Space", CERN 68-15, GENBOD W515) which both ROOT's
 * TGenPhaseSpace and the reference follow (phasespace.py:7-12, 305-306): sorted uniforms ->
 * invariant masses -> two-body momenta pdk (eq. 9.17) -> weight -> per-step random rotation
 * about z then y and a boost along y.  It is written independently of oracle/genbod_numpy.py
 * (scalar, event at a time, its own xoshiro256** stream) so that the two can cross-check each
 * other statistically (tests/test_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define GENBOD_MAX_N 18

typedef struct { uint64_t s[4]; } rng_t;

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline void rng_seed(rng_t *r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed);
}

static inline double rng_uniform(rng_t *r) { /* xoshiro256** -> 53-bit U[0,1) */
    uint64_t *s = r->s;
    const uint64_t result = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return (double)(result >> 11) * 0x1.0p-53;
}

/* CERN 68-15 eq. 9.17 (phasespace.py:56-71). */
static inline double pdk(double a, double b, double c) {
    double x = (a - b - c) * (a + b + c) * (a - b + c) * (a + b - c);
    return sqrt(x) / (2.0 * a);
}

/* Maximum weight of the decay (phasespace.py:288-298). */
double genbod_wtmax(double m_top, int n, const double *m) {
    double tecm = m_top;
    for (int i = 0; i < n; ++i) tecm -= m[i];
    double emmax = tecm + m[0], emmin = 0.0, wtmax = 1.0;
    for (int i = 1; i < n; ++i) {
        emmin += m[i - 1];
        emmax += m[i];
        wtmax *= pdk(emmax, emmin, m[i]);
    }
    return wtmax;
}

/* One event in the rest frame of the parent.  p is [n][4] = (px,py,pz,E). Returns prod(pd). */
static inline double genbod_event(rng_t *rng, double m_top, int n, const double *m, double tecm, double (*p)[4]) {
    double rno[GENBOD_MAX_N], inv[GENBOD_MAX_N], pd[GENBOD_MAX_N];
    (void)m_top;
    rno[0] = 0.0;
    for (int i = 1; i < n - 1; ++i) { /* insertion sort of the n-2 uniforms */
        double u = rng_uniform(rng);
        int j = i;
        while (j > 1 && rno[j - 1] > u) { rno[j] = rno[j - 1]; --j; }
        rno[j] = u;
    }
    rno[n - 1] = 1.0;
    double sum = 0.0;
    for (int i = 0; i < n; ++i) { sum += m[i]; inv[i] = rno[i] * tecm + sum; }
    double wt = 1.0;
    for (int i = 0; i < n - 1; ++i) { pd[i] = pdk(inv[i + 1], inv[i], m[i + 1]); wt *= pd[i]; }
    p[0][0] = 0.0; p[0][1] = pd[0]; p[0][2] = 0.0; p[0][3] = sqrt(pd[0] * pd[0] + m[0] * m[0]);
    for (int k = 1;; ++k) {
        p[k][0] = 0.0; p[k][1] = -pd[k - 1]; p[k][2] = 0.0; p[k][3] = sqrt(pd[k - 1] * pd[k - 1] + m[k] * m[k]);
        double cz = 2.0 * rng_uniform(rng) - 1.0, sz = sqrt(1.0 - cz * cz);
        double ang = 2.0 * M_PI * rng_uniform(rng), cy = cos(ang), sy = sin(ang);
        for (int j = 0; j <= k; ++j) {
            double x = p[j][0], y = p[j][1];
            p[j][0] = cz * x - sz * y; p[j][1] = sz * x + cz * y;
            x = p[j][0]; double z = p[j][2];
            p[j][0] = cy * x - sy * z; p[j][2] = sy * x + cy * z;
        }
        if (k == n - 1) break;
        double beta = pd[k] / sqrt(pd[k] * pd[k] + inv[k] * inv[k]);
        double gamma = 1.0 / sqrt(1.0 - beta * beta);
        for (int j = 0; j <= k; ++j) { /* boost along +y */
            double y = p[j][1], e = p[j][3];
            p[j][1] = gamma * (y + beta * e);
            p[j][3] = gamma * (e + beta * y);
        }
    }
    return wt;
}

typedef struct {
    double m_top, tecm, inv_wtmax;
    int n, tid, nth;
    const double *m;
    int64_t n_events;
    uint64_t seed;
    double *weights, *particles;
} genbod_job;

static void *genbod_worker(void *arg) {
    const genbod_job *jb = (const genbod_job *)arg;
    rng_t rng;
    rng_seed(&rng, jb->seed * 0x9E3779B97F4A7C15ull + (uint64_t)jb->tid + 1);
    const int64_t lo = jb->n_events * jb->tid / jb->nth, hi = jb->n_events * (jb->tid + 1) / jb->nth;
    double p[GENBOD_MAX_N][4];
    double sink = 0.0;
    for (int64_t e = lo; e < hi; ++e) {
        double wt = genbod_event(&rng, jb->m_top, jb->n, jb->m, jb->tecm, p) * jb->inv_wtmax;
        if (jb->weights) jb->weights[e] = wt;
        if (jb->particles) {
            for (int j = 0; j < jb->n; ++j)
                memcpy(jb->particles + ((int64_t)j * jb->n_events + e) * 4, p[j], 4 * sizeof(double));
        } else {
            sink += p[jb->n - 1][3];
        }
    }
    if (sink == -1.0 && jb->weights && hi > lo) jb->weights[lo] = sink; /* keep the work alive when not storing */
    return NULL;
}

/* Fill weights[n_events] (normalised: prod(pd)/wtmax, like TGenPhaseSpace::Generate) and
 * particles[n][n_events][4] for a rest-frame decay.  particles may be NULL (timing only).
 * n_threads <= 0 means all online cores.  Returns the number of threads used. */
int genbod_generate(double m_top, int n, const double *m, int64_t n_events, uint64_t seed,
                    double *weights, double *particles, int n_threads) {
    if (n < 2 || n > GENBOD_MAX_N) return -1;
    double tecm = m_top;
    for (int i = 0; i < n; ++i) tecm -= m[i];
    if (!(tecm >= 0.0)) return -2;
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    genbod_job jobs[256];
    pthread_t th[256];
    for (int t = 0; t < n_threads; ++t) {
        genbod_job jb = {m_top, tecm, 1.0 / genbod_wtmax(m_top, n, m), n, t, n_threads, m, n_events, seed, weights, particles};
        jobs[t] = jb;
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], NULL, genbod_worker, &jobs[t]);
    genbod_worker(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], NULL);
    return n_threads;
}
