/* oracle_c.c -- plain-C restatement of the reference's hot path, for TIMING the CPU arm.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (see oracle/qfall_oracle.py header).  Built by
 * oracle/Makefile into oracle/liboracle_c.so; used by bench.py's cpu_baseline leg and
 * `bench.py --impl reference`, and checked against qfall_oracle.py in tests/test_oracle_c.py.
 *
 * Same loop structure as the reference (file:line cited per function, relative to
 * /root/reference), with float64 in place of qfall-math's exact rationals and
 * int64/__int128 in place of FLINT integers -- i.e. this baseline is FASTER than the real
 * crate.  Parity status: as qfall_oracle.py (deterministic parts pinned through it; sampler
 * outputs unpinned, the reference has no seed API).
 *
 * Threading: the reference is single-threaded; `threads` > 1 runs one independent
 * instance per pthread over disjoint targets (the image has no OpenMP runtime).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef unsigned __int128 u128;
typedef __int128 i128;

/* ---- RNG: xoshiro256** (the reference uses rand's ThreadRng; any good PRNG will do) ---- */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {
    uint64_t* s = r->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static void rng_seed(rng_t* r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) {
        seed += 0x9E3779B97F4A7C15ull;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        r->s[i] = z ^ (z >> 31);
    }
}
static inline double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * 0x1.0p-53; }
static inline double rng_normal(rng_t* r) {
    double u1 = 1.0 - rng_unit(r), u2 = rng_unit(r);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

/* GPV08 SampleZ as qfall-math implements Z::sample_discrete_gauss (CONTRIBUTING.md:35-45):
 * uniform proposal on [c - ceil(6s), c + floor(6s)], accept with rho_{s,c}(x). */
static int64_t sample_z(rng_t* r, double s, double c) {
    const double lo = ceil(c - ceil(6.0 * s)), hi = floor(c + floor(6.0 * s));
    const double width = hi - lo + 1.0;
    for (;;) {
        double x = lo + floor(rng_unit(r) * width);
        double d = x - c;
        if (rng_unit(r) < exp(-3.141592653589793 * d * d / (s * s))) return (int64_t)x;
    }
}

int orc_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

/* ---- pthread work sharing (no OpenMP runtime in this image): dynamic chunks of `grain` ---- */
typedef void (*body_fn)(long lo, long hi, void* arg);
typedef struct { body_fn fn; void* arg; long n, grain; long next; pthread_mutex_t mu; } par_t;
static void* par_worker(void* vp) {
    par_t* p = (par_t*)vp;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        long lo = p->next;
        p->next += p->grain;
        pthread_mutex_unlock(&p->mu);
        if (lo >= p->n) break;
        long hi = lo + p->grain < p->n ? lo + p->grain : p->n;
        p->fn(lo, hi, p->arg);
    }
    return NULL;
}
static void par_for(long n, long grain, int threads, body_fn fn, void* arg) {
    if (threads <= 1 || n <= grain) { fn(0, n, arg); return; }
    par_t p = { fn, arg, n, grain, 0 };
    pthread_mutex_init(&p.mu, NULL);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, par_worker, &p);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&p.mu);
}

void orc_sample_z(double s, double c, uint64_t seed, long count, int64_t* out) {
    rng_t r;
    rng_seed(&r, seed);
    for (long i = 0; i < count; ++i) out[i] = sample_z(&r, s, c);
}

/* ---- f_a = A * sigma mod q  (gpv.rs:190-193) ------------------------------------------- */
typedef struct { const int64_t* A; const int32_t* sigma; int64_t* u; long n, m; uint64_t q; } fa_t;
static void fa_body(long lo, long hi, void* vp) {
    fa_t* a = (fa_t*)vp;
    for (long b = lo; b < hi; ++b) {
        const int32_t* s = a->sigma + b * a->m;
        for (long i = 0; i < a->n; ++i) {
            const int64_t* row = a->A + i * a->m;
            i128 acc = 0;
            for (long j = 0; j < a->m; ++j) acc += (i128)row[j] * s[j];
            i128 r = acc % (i128)a->q;
            if (r < 0) r += a->q;
            a->u[b * a->n + i] = (int64_t)r;
        }
    }
}
void orc_f_a(const int64_t* A, const int32_t* sigma, int64_t* u, long B, long n, long m, uint64_t q, int threads) {
    fa_t a = { A, sigma, u, n, m, q };
    par_for(B, 1, threads, fa_body, &a);
}

/* ---- FIPS 203 compress / decompress (lossy_compression_fips203.rs:101-111, 159-169) ------ */
typedef struct { const uint16_t* in; uint16_t* out; uint32_t q, d; int dec; } cp_t;
static void cp_body(long lo, long hi, void* vp) {
    cp_t* a = (cp_t*)vp;
    const uint64_t q = a->q, two_d = 1ull << a->d, half_q = a->q / 2, round = 1ull << (a->d - 1);
    for (long i = lo; i < hi; ++i) {
        uint64_t x = a->in[i];
        a->out[i] = a->dec ? (uint16_t)((x * q + round) / two_d) : (uint16_t)(((x * two_d + half_q) / q) % two_d);
    }
}
void orc_compress_u16(const uint16_t* in, uint16_t* out, size_t count, uint32_t q, uint32_t d, int dec, int threads) {
    cp_t a = { in, out, q, d, dec };
    par_for((long)count, 1 << 16, threads, cp_body, &a);
}

/* ---- GPV08 SampleD (MatZ::sample_d_precomputed_gso as used at gpv.rs:160) ----------------
 * Bt : dim x dim, row i = basis vector b_i;  Gt : row i = b~_i;  inv_n2[i] = 1/||b~_i||^2.
 * c (length dim) is consumed; v accumulates sum z_i b_i. */
static void sample_d(rng_t* rng, const double* Bt, const double* Gt, const double* n2, long dim, double* c,
                     double* v, double s) {
    for (long i = dim - 1; i >= 0; --i) {
        const double* g = Gt + i * dim;
        const double* b = Bt + i * dim;
        double dot = 0;
        for (long t = 0; t < dim; ++t) dot += c[t] * g[t];
        double z = (double)sample_z(rng, s / sqrt(n2[i]), dot / n2[i]);
        if (z != 0.0)
            for (long t = 0; t < dim; ++t) {
                c[t] -= z * b[t];
                v[t] += z * b[t];
            }
    }
}

/* PSFGPV::samp_p (gpv.rs:152-161).  The per-call Gaussian elimination of the reference
 * (:153-156) is hoisted: piv/Ainv give sol[piv[k]] = sum_j Ainv[k][j] u_j mod q
 * (favours the CPU arm).  Bt/Gt as above (dim = m). */
typedef struct {
    const double *Bt, *Gt, *n2; const int32_t* piv; const int64_t* Ainv; long npiv; const int64_t* u; int32_t* e;
    long n, dim; uint64_t q; double s; uint64_t seed;
} gpv_t;
static void gpv_body(long lo, long hi, void* vp) {
    gpv_t* a = (gpv_t*)vp;
    const long dim = a->dim, n = a->n;
    double* c = (double*)malloc(sizeof(double) * dim);
    double* v = (double*)malloc(sizeof(double) * dim);
    double* sol = (double*)malloc(sizeof(double) * dim);
    for (long b = lo; b < hi; ++b) {
        rng_t rng;
        rng_seed(&rng, a->seed + 0x1000003ull * (uint64_t)b);
        memset(sol, 0, sizeof(double) * dim);
        for (long kk = 0; kk < a->npiv; ++kk) {
            u128 acc = 0;
            for (long j = 0; j < n; ++j)
                acc = (acc + (u128)(uint64_t)a->Ainv[kk * a->npiv + j] * (uint64_t)a->u[b * n + j]) % a->q;
            sol[a->piv[kk]] = (double)(uint64_t)acc;
        }
        for (long t = 0; t < dim; ++t) { c[t] = -sol[t]; v[t] = 0; }
        sample_d(&rng, a->Bt, a->Gt, a->n2, dim, c, v, a->s);
        for (long t = 0; t < dim; ++t) a->e[b * dim + t] = (int32_t)llrint(sol[t] + v[t]);
    }
    free(c); free(v); free(sol);
}
void orc_samp_p_gpv(const double* Bt, const double* Gt, const int32_t* piv, const int64_t* Ainv, long npiv,
                    const int64_t* u, int32_t* e, long B, long n, long dim, uint64_t q, double s, uint64_t seed,
                    int threads) {
    double* n2 = (double*)malloc(sizeof(double) * dim);
    for (long i = 0; i < dim; ++i) {
        double acc = 0;
        for (long t = 0; t < dim; ++t) acc += Gt[i * dim + t] * Gt[i * dim + t];
        n2[i] = acc;
    }
    gpv_t a = { Bt, Gt, n2, piv, Ainv, npiv, u, e, n, dim, q, s, seed };
    par_for(B, 1, threads, gpv_body, &a);
    free(n2);
}

/* PSFPerturbation::samp_p (mp_perturbation.rs:304-336), dense like the reference:
 *  L   : m x m lower-triangular sqrt(Sigma_2) (row-major)
 *  A   : n x m residues, R : m_bar x nk (int8)
 *  SGt : nk x nk, row i = b_i of the gadget short basis I_n (x) S_k; SGgt: its GSO rows.
 * The gadget nearest plane runs over the full nk x nk matrices as the reference does
 * (randomized_nearest_plane_gadget, :173-191). */
typedef struct {
    const double* L; const int64_t* A; const int8_t* R; const double *SGt, *SGgt, *n2; const int64_t* u; int32_t* e;
    long n, k, m_bar, base; uint64_t q; double r; uint64_t seed;
} pert_t;
static void pert_body(long lo, long hi, void* vp) {
    pert_t* a = (pert_t*)vp;
    const long n = a->n, k = a->k, m_bar = a->m_bar, nk = n * k, m = m_bar + nk, base = a->base;
    const double r = a->r, s_g = r * sqrt((double)(base * base + 1));
    double* g = (double*)malloc(sizeof(double) * m);
    int64_t* p = (int64_t*)malloc(sizeof(int64_t) * m);
    double* c = (double*)malloc(sizeof(double) * nk);
    double* v = (double*)malloc(sizeof(double) * nk);
    int64_t* x0 = (int64_t*)malloc(sizeof(int64_t) * nk);
    for (long b = lo; b < hi; ++b) {
        rng_t rng;
        rng_seed(&rng, a->seed + 0x1000003ull * (uint64_t)b);
        /* p <- sample_d_common_non_spherical(sqrt(Sigma_2), r)   (:315) */
        for (long i = 0; i < m; ++i) g[i] = rng_normal(&rng);
        for (long i = 0; i < m; ++i) {
            const double* row = a->L + i * m;
            double x2 = 0;
            for (long j = 0; j <= i; ++j) x2 += row[j] * g[j];
            p[i] = sample_z(&rng, r, x2);
        }
        /* v = u - A p   (:318), then digits x0 (gadget_classical.rs:219-229) */
        for (long i = 0; i < n; ++i) {
            const int64_t* row = a->A + i * m;
            i128 acc = 0;
            for (long j = 0; j < m; ++j) acc += (i128)row[j] * p[j];
            i128 vv = ((i128)a->u[b * n + i] - acc) % (i128)a->q;
            if (vv < 0) vv += a->q;
            uint64_t val = (uint64_t)vv;
            for (long t = 0; t < k; ++t) {
                x0[i * k + t] = (int64_t)(val % (uint64_t)base);
                val /= (uint64_t)base;
            }
        }
        /* z = x0 + SampleD(S, S~, -x0, s_G)   (:185-190) */
        for (long t = 0; t < nk; ++t) { c[t] = -(double)x0[t]; v[t] = 0; }
        sample_d(&rng, a->SGt, a->SGgt, a->n2, nk, c, v, s_g);
        /* e = p + [R; I] z   (:328-335) */
        for (long i = 0; i < m_bar; ++i) {
            const int8_t* row = a->R + i * nk;
            double acc = 0;
            for (long t = 0; t < nk; ++t) acc += (double)row[t] * ((double)x0[t] + v[t]);
            a->e[b * m + i] = (int32_t)(p[i] + llrint(acc));
        }
        for (long t = 0; t < nk; ++t) a->e[b * m + m_bar + t] = (int32_t)(p[m_bar + t] + x0[t] + llrint(v[t]));
    }
    free(g); free(p); free(c); free(v); free(x0);
}
void orc_samp_p_pert(const double* L, const int64_t* A, const int8_t* R, const double* SGt, const double* SGgt,
                     const int64_t* u, int32_t* e, long B, long n, long k, long m_bar, long base, uint64_t q,
                     double r, uint64_t seed, int threads) {
    const long nk = n * k;
    double* n2 = (double*)malloc(sizeof(double) * nk);
    for (long i = 0; i < nk; ++i) {
        double acc = 0;
        for (long t = 0; t < nk; ++t) acc += SGgt[i * nk + t] * SGgt[i * nk + t];
        n2[i] = acc;
    }
    pert_t a = { L, A, R, SGt, SGgt, n2, u, e, n, k, m_bar, base, q, r, seed };
    par_for(B, 1, threads, pert_body, &a);
    free(n2);
}
