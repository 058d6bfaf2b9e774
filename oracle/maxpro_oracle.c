/*
 * maxpro_oracle.c -- CPU restatement of the mrhheffernan/maxpro hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under maxpro_b200/ may include, link or
 * call this file.  Allowed callers: tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status
 *   - Criteria (maxpro / maximin / l2), argbest tie rule, annealer control
 *     flow and greedy ordering follow the reference line by line and are
 *     PINNED by the reference's own vectors (tests/test_oracle.py):
 *       README.md:90 / python/comparison_r.py:14-115  (R 100x2 design,
 *       95.98505099515626), src/maxpro_utils.rs:117-133,
 *       src/maximin_utils.rs:85-101, python/pymaxpro.py:7-62 (author's NumPy
 *       restatement, imported to generate tests/golden/pymaxpro_vectors.json).
 *   - RANDOM STREAM: parity unpinned.  The reference draws from rand 0.9.2
 *     StdRng (ChaCha12; crate not vendored under /root/reference, no
 *     Cargo.lock) and none of its tests pins a stream.  This file instead
 *     defines the two design streams of the engine -- "LHD-Philox v1"
 *     (Philox4x32-10, Random123 KATs) and "LHD-mix v2" (one Philox call per
 *     column + a 32-bit mixer per element; the default) -- which the CUDA
 *     engine implements independently; generated designs are checked
 *     bit-exactly between the two and for LHD validity, not against rand.
 *
 * Compile with -ffp-contract=off: Rust never contracts a*b+c into an FMA.
 * All citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_METRIC_MAXPRO 0
#define ORACLE_METRIC_MAXIMIN 1

/* design stream version used by oracle_generate_lhd and everything built on it: 1 = LHD-Philox v1,
 * 2 = LHD-mix v2 (the engine's default) */
static int g_stream_version = 2;
void oracle_set_stream_version(int v) { g_stream_version = v == 1 ? 1 : 2; }
int oracle_get_stream_version(void) { return g_stream_version; }
int oracle_generate_lhd_v1(uint64_t n, uint64_t d, uint64_t stream, double *out);
int oracle_generate_lhd_v2(uint64_t n, uint64_t d, uint64_t stream, double *out);
int oracle_generate_lhd(uint64_t n, uint64_t d, uint64_t stream, double *out)
{
    return g_stream_version == 1 ? oracle_generate_lhd_v1(n, d, stream, out) : oracle_generate_lhd_v2(n, d, stream, out);
}

/* ------------------------------------------------------------------ */
/* Criteria                                                            */
/* ------------------------------------------------------------------ */

/* src/maxpro_utils.rs:28-60 -- rows i ascending, j = i+1..n, product seeded
 * with 1.0, diff*diff then *=, eps added AFTER the product, 1.0/p summed.
 * rayon's .sum() order is unspecified; the sequential order used here
 * reproduces README.md:90 to the last digit. */
double oracle_maxpro_sum(const double *x, uint64_t n, uint64_t d)
{
    if (n < 2) return 0.0; /* :31-33 */
    const double epsilon = 1e-12; /* :35 */
    double total = 0.0;
    for (uint64_t i = 0; i < n; ++i) {
        const double *ri = x + i * d;
        double inverse_product = 0.0; /* :42 */
        for (uint64_t j = i + 1; j < n; ++j) {
            const double *rj = x + j * d;
            double p = 1.0; /* :45 */
            for (uint64_t k = 0; k < d; ++k) {
                double diff = ri[k] - rj[k]; /* :48 */
                double diff_sq = diff * diff; /* :49 */
                p *= diff_sq; /* :50 */
            }
            p += epsilon; /* :53 */
            inverse_product += 1.0 / p; /* :54 */
        }
        total += inverse_product; /* :58 */
    }
    return total;
}

/* src/maxpro_utils.rs:72-83.  n == 0 and d == 0 panic in the reference; the
 * callers here reject them before reaching this function. */
double oracle_maxpro_criterion(const double *x, uint64_t n, uint64_t d)
{
    if (n < 2) return 0.0; /* :75-77 */
    double s = oracle_maxpro_sum(x, n, d);
    double n_pairs = (double)n * ((double)n - 1.0) / 2.0; /* :81 */
    return pow(s / n_pairs, 1.0 / (double)d); /* :82 */
}

/* src/maximin_utils.rs:26-45 -- sum left to right from 0, powi(2) == a*a. */
double oracle_l2_distance(const double *a, const double *b, uint64_t d)
{
    double acc = 0.0;
    for (uint64_t k = 0; k < d; ++k) {
        double diff = a[k] - b[k];
        acc += diff * diff;
    }
    return sqrt(acc);
}

/* src/maximin_utils.rs:54-67 -- strict <, n == 1 gives +INF. */
double oracle_maximin_criterion(const double *x, uint64_t n, uint64_t d)
{
    double min_distance = INFINITY;
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = i + 1; j < n; ++j) {
            double dist = oracle_l2_distance(x + i * d, x + j * d, d);
            if (dist < min_distance) min_distance = dist;
        }
    return min_distance;
}

static double criterion(const double *x, uint64_t n, uint64_t d, int metric)
{
    return metric == ORACLE_METRIC_MAXPRO ? oracle_maxpro_criterion(x, n, d)
                                          : oracle_maximin_criterion(x, n, d);
}

/* B designs, each n*d row-major.  argbest follows src/lib.rs:76-98: the fold
 * keeps the left element only on a strict win, so ties go to the LATER index;
 * identity is +INF (minimise, maxpro) / -INF (maximise, maximin). */
void oracle_score_batch(const double *designs, uint64_t b, uint64_t n, uint64_t d,
                        int metric, double *scores, uint64_t *argbest)
{
    int minimize = metric == ORACLE_METRIC_MAXPRO;
    double best = minimize ? INFINITY : -INFINITY;
    uint64_t best_i = UINT64_MAX;
    for (uint64_t c = 0; c < b; ++c) {
        double m = criterion(designs + c * n * d, n, d, metric);
        scores[c] = m;
        int keep_left = minimize ? (best < m) : (best > m);
        if (!keep_left) { best = m; best_i = c; }
    }
    if (argbest) *argbest = best_i;
}

/* ------------------------------------------------------------------ */
/* LHD-Philox v1 stream                                                */
/* ------------------------------------------------------------------ */

/* Philox4x32-10, Salmon et al. SC'11 (Random123 philox.h constants). */
void oracle_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define KEY_LHD0     0x4C484431u /* "LHD1": fixed Philox key of the design stream */
#define KEY_LHD1     0x243F6A88u
#define DOMAIN_SWAP  0x53574150u /* "SWAP" */
#define DOMAIN_JACC  0x4A414343u /* "JACC" */
#define DOMAIN_JIT   0x4A495454u /* "JITT" */

/* Structure of src/lhd.rs:97-132: per column, a uniform random permutation of
 * the n strata and one uniform draw inside each stratum.  Stream definition
 * (ours, see header): fixed key {KEY_LHD0, KEY_LHD1}; for stream id s, column k
 * and element pair b the block Philox(ctr = {b, k, s_lo, s_hi}) yields
 * {jit, idx} for element 2b and {jit, idx} for element 2b+1.  Inside-out Fisher-Yates:
 *     j = (idx * (i+1)) >> 32;  a[i] = a[j];  a[j] = v_i
 *     v_i = (i + (jit + 0.5) * 2^-32) / n   evaluated as ((i<<32|jit) + 0.5) * (2^-32 * (1/n))
 * so that floor(v*n) == i for every draw (no edge rounding into i+1). */
int oracle_generate_lhd_v1(uint64_t n, uint64_t d, uint64_t stream, double *out)
{
    if (n == 0) return 0;              /* :98-101 empty design */
    if (d == 0) return -1;             /* :106 panics */
    if (n > (1ull << 18)) return -2;   /* stratum-exactness bound of v1 */
    const double c1 = ldexp(1.0 / (double)n, -32);
    const uint32_t key[2] = { KEY_LHD0, KEY_LHD1 };
    double *col = (double *)malloc(n * sizeof(double));
    if (!col) return -3;
    for (uint64_t k = 0; k < d; ++k) {
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t ctr[4] = { (uint32_t)(i >> 1), (uint32_t)k, (uint32_t)stream, (uint32_t)(stream >> 32) };
            uint32_t w[4];
            oracle_philox4x32_10(ctr, key, w);
            uint32_t jit = (i & 1) ? w[2] : w[0];
            uint32_t idx = (i & 1) ? w[3] : w[1];
            uint64_t j = ((uint64_t)idx * (uint64_t)(i + 1)) >> 32;
            uint64_t m = (i << 32) | jit;
            double v = ((double)m + 0.5) * c1;
            col[i] = col[j];
            col[j] = v;
        }
        for (uint64_t i = 0; i < n; ++i) out[i * d + k] = col[i]; /* lhd[i][k] :128 */
    }
    free(col);
    return 0;
}

/* LHD-mix v2 (the engine's default design stream; maxpro_b200/csrc/philox.cuh has the rationale).
 * One Philox4x32-10 call per column gives the key, a counter-based 32-bit mixer gives the elements:
 *     {A, B} = Philox(ctr = {k, "LHD2", s_lo, s_hi}, key = {KEY_LHD0, KEY_LHD1}).{x, y | 1}
 *     w_i = mix32(A + i*B);  P = w_i * (i+1);  j_i = P >> 32;  jit_i = (uint32)P
 *     v_i = fma(2^52 + (i*2^32 + jit_i), c1, c0'),  c1 = 2^-32/n,  c0' = c1/2 - 2^52*c1
 * Same inside-out Fisher-Yates as v1. */
static uint32_t lhd2_mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x735a2d97u;   /* one xorshift-multiply round on a Weyl sequence with random start and step */
    return x;
}

int oracle_generate_lhd_v2(uint64_t n, uint64_t d, uint64_t stream, double *out)
{
    if (n == 0) return 0;
    if (d == 0) return -1;
    if (n > (1ull << 18)) return -2;
    const double c1 = ldexp(1.0 / (double)n, -32);
    const double c0p = 0.5 * c1 - 0x1p52 * c1;
    const uint32_t key[2] = { KEY_LHD0, KEY_LHD1 };
    double *col = (double *)malloc(n * sizeof(double));
    if (!col) return -3;
    for (uint64_t k = 0; k < d; ++k) {
        uint32_t ctr[4] = { (uint32_t)k, 0x4C484432u, (uint32_t)stream, (uint32_t)(stream >> 32) };
        uint32_t w[4];
        oracle_philox4x32_10(ctr, key, w);
        const uint32_t a = w[0], b = w[1] | 1u;
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t wi = lhd2_mix32(a + (uint32_t)i * b);
            uint64_t p = (uint64_t)wi * (uint64_t)(i + 1);
            uint64_t j = p >> 32;
            uint32_t jit = (uint32_t)p;
            uint64_t bits = 0x4330000000000000ull | (i << 32) | jit;   /* the double 2^52 + (i*2^32 + jit) */
            double m;
            memcpy(&m, &bits, 8);
            double v = fma(m, c1, c0p);
            col[i] = col[j];
            col[j] = v;
        }
        for (uint64_t i = 0; i < n; ++i) out[i * d + k] = col[i];
    }
    free(col);
    return 0;
}

/* `count` consecutive designs of the search seeded `seed`, starting at candidate `lo` (design c is
 * generate_lhd of stream seed + lo + c, src/lib.rs:70).  Used by the statistical tests of the streams. */
int oracle_generate_batch(uint64_t n, uint64_t d, uint64_t seed, uint64_t lo, uint64_t count, double *out)
{
    int status = 0;
#pragma omp parallel for schedule(static)
    for (uint64_t c = 0; c < count; ++c)
        if (oracle_generate_lhd(n, d, seed + lo + c, out + c * n * d) != 0) status = -1;
    return status;
}

/* ------------------------------------------------------------------ */
/* Random search (src/lib.rs:30-100)                                   */
/* ------------------------------------------------------------------ */

/* Candidates lo..hi-1 of the search seeded `seed`: candidate i is the design
 * of stream seed+i (lib.rs:70, wrapping like a release build).  Returns the
 * best score and its GLOBAL index under the lib.rs:85-97 tie rule (equal
 * score -> later index).  OpenMP over contiguous chunks, folded in order. */
int oracle_search_range(uint64_t n, uint64_t d, int metric, uint64_t seed,
                        uint64_t lo, uint64_t hi, double *best_score, uint64_t *best_idx)
{
    if (n == 0 || d == 0) return -1;
    int minimize = metric == ORACLE_METRIC_MAXPRO;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    double *tb = (double *)malloc(sizeof(double) * nt);
    uint64_t *ti = (uint64_t *)malloc(sizeof(uint64_t) * nt);
    int status = 0;
#pragma omp parallel num_threads(nt)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        uint64_t total = hi - lo;
        uint64_t a = lo + total * (uint64_t)t / (uint64_t)nt;
        uint64_t b = lo + total * (uint64_t)(t + 1) / (uint64_t)nt;
        double best = minimize ? INFINITY : -INFINITY; /* lib.rs:77-83 */
        uint64_t bi = UINT64_MAX;
        double *x = (double *)malloc(n * d * sizeof(double));
        for (uint64_t i = a; i < b; ++i) {
            if (oracle_generate_lhd(n, d, seed + i, x) != 0) { status = -2; break; }
            double m = criterion(x, n, d, metric);
            int keep_left = minimize ? (best < m) : (best > m);
            if (!keep_left) { best = m; bi = i; }
        }
        free(x);
        tb[t] = best; ti[t] = bi;
    }
    double best = minimize ? INFINITY : -INFINITY;
    uint64_t bi = UINT64_MAX;
    for (int t = 0; t < nt; ++t) {
        if (ti[t] == UINT64_MAX) continue; /* empty chunk == identity */
        int keep_left = minimize ? (best < tb[t]) : (best > tb[t]);
        if (!keep_left) { best = tb[t]; bi = ti[t]; }
    }
    free(tb); free(ti);
    *best_score = best; *best_idx = bi;
    return status;
}

/* src/lib.rs:30-100.  metric < 0 means None (:44-50): one design of stream
 * `seed`.  Writes the winning design, its score and its index. */
int oracle_build_lhd(uint64_t n, uint64_t d, uint64_t iters, int metric, uint64_t seed,
                     double *out, double *out_score, uint64_t *out_idx)
{
    if (n == 0 || d == 0 || iters == 0) return -1; /* :37-42 */
    if (metric < 0) {
        if (out_score) *out_score = NAN;
        if (out_idx) *out_idx = 0;
        return oracle_generate_lhd(n, d, seed, out);
    }
    double s; uint64_t idx;
    int rc = oracle_search_range(n, d, metric, seed, 0, iters, &s, &idx);
    if (rc) return rc;
    if (out_score) *out_score = s;
    if (out_idx) *out_idx = idx;
    return oracle_generate_lhd(n, d, seed + idx, out);
}

/* ------------------------------------------------------------------ */
/* Annealer (src/anneal.rs:23-143)                                     */
/* ------------------------------------------------------------------ */

/* One chain.  Control flow is the reference's: clone, move, FULL re-score,
 * Metropolis on the criterion difference, strict-improvement global best,
 * geometric cooling.  Draws come from stream `seed` (Philox key = {seed_lo, seed_hi}):
 *   swap   (anneal.rs:23-40): Philox({t_lo,t_hi,0,DOMAIN_SWAP}) -> r1,r2,c,u
 *          r = (w*n)>>32, c = (w*d)>>32, u = w*2^-32
 *   jitter (anneal.rs:102-107): element e=i*d+k uses word e%4 of
 *          Philox({t_lo,t_hi,e/4,DOMAIN_JIT}); delta = u*(2*step) - step
 *          (reference: random_range(-step..step) = low + u*(high-low));
 *          accept draw = word 0 of Philox({t_lo,t_hi,0,DOMAIN_JACC}).
 * best_metric_out (optional) receives the global best metric. */
int oracle_anneal_lhd(const double *design, uint64_t n, uint64_t d, uint64_t iters,
                      double initial_temp, double cooling_rate, int metric, int minimize,
                      uint64_t seed, int swap, double *out, double *best_metric_out)
{
    if (n == 0) return 0;                           /* :69-71 */
    if (iters == 0 || !(initial_temp > 0.0)) return -1; /* :72-79 */
    uint64_t nd = n * d;
    double temp = initial_temp;                     /* :81 */
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    double step = 0.01 / (double)n;                 /* :86-87 */
    double *cur = (double *)malloc(nd * sizeof(double));
    double *cand = (double *)malloc(nd * sizeof(double));
    memcpy(cur, design, nd * sizeof(double));       /* :89 */
    double cur_metric = criterion(design, n, d, metric); /* :90 */
    memcpy(out, design, nd * sizeof(double));       /* :93 */
    double global_best = cur_metric;                /* :94 */
    for (uint64_t t = 0; t < iters; ++t) {
        memcpy(cand, cur, nd * sizeof(double));     /* :97 */
        uint32_t w[4];
        double u;
        if (swap) {
            uint32_t ctr[4] = { (uint32_t)t, (uint32_t)(t >> 32), 0u, DOMAIN_SWAP };
            oracle_philox4x32_10(ctr, key, w);
            uint64_t r1 = ((uint64_t)w[0] * n) >> 32; /* :29 */
            uint64_t r2 = ((uint64_t)w[1] * n) >> 32; /* :30 */
            uint64_t c = ((uint64_t)w[2] * d) >> 32;  /* :31 */
            double v1 = cand[r1 * d + c], v2 = cand[r2 * d + c];
            cand[r1 * d + c] = v2;                    /* :38 */
            cand[r2 * d + c] = v1;                    /* :39 */
            u = (double)w[3] * 0x1p-32;
        } else {
            double two_step = step + step;
            for (uint64_t e = 0; e < nd; ++e) {
                if ((e & 3) == 0) {
                    uint32_t ctr[4] = { (uint32_t)t, (uint32_t)(t >> 32), (uint32_t)(e >> 2), DOMAIN_JIT };
                    oracle_philox4x32_10(ctr, key, w);
                }
                double uu = (double)w[e & 3] * 0x1p-32;
                double delta = uu * two_step - step;
                double v = cand[e] + delta;           /* :105 */
                v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); /* clamp(0,1) */
                cand[e] = v;
            }
            uint32_t ctr[4] = { (uint32_t)t, (uint32_t)(t >> 32), 0u, DOMAIN_JACC };
            oracle_philox4x32_10(ctr, key, w);
            u = (double)w[0] * 0x1p-32;
        }
        double new_metric = criterion(cand, n, d, metric); /* :110 */
        double diff = new_metric - cur_metric;             /* :111 */
        if (!minimize) diff *= -1.0;                       /* :112-115 */
        double p = diff < 0.0 ? 1.0 : exp(-diff / temp);   /* :118-122 */
        if (u < p) {                                       /* :125 */
            double *tmp = cur; cur = cand; cand = tmp;
            cur_metric = new_metric;
        }
        if ((minimize && cur_metric < global_best) || (!minimize && cur_metric > global_best)) {
            memcpy(out, cur, nd * sizeof(double));         /* :131-136 */
            global_best = cur_metric;
        }
        temp *= cooling_rate;                              /* :139 */
    }
    if (best_metric_out) *best_metric_out = global_best;
    free(cur); free(cand);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Greedy ordering (src/order.rs:34-105)                               */
/* ------------------------------------------------------------------ */

/* Reference-faithful O(n^4 d): every trial re-scores the whole prefix.
 * `perm[m]` = original row index placed at position m.  Position order of
 * the unordered pool follows Vec::swap_remove (last element fills the hole). */
int oracle_order_design(const double *design, uint64_t n, uint64_t d, int metric,
                        int minimize, double *out, uint64_t *perm)
{
    if (n == 0 || d == 0) return 0; /* :38-46 returned unchanged */
    uint64_t *pool = (uint64_t *)malloc(n * sizeof(uint64_t));
    double *ord = (double *)malloc(n * d * sizeof(double));
    double *center = (double *)malloc(d * sizeof(double));
    for (uint64_t k = 0; k < d; ++k) center[k] = 0.5; /* :49 */
    for (uint64_t i = 0; i < n; ++i) pool[i] = i;
    double dc = INFINITY; uint64_t ci = 0;
    for (uint64_t i = 0; i < n; ++i) {               /* :53-59 */
        double pd = oracle_l2_distance(center, design + i * d, d);
        if (pd < dc) { dc = pd; ci = i; }
    }
    uint64_t pool_n = n, m = 0;
    uint64_t first = pool[ci]; pool[ci] = pool[--pool_n]; /* swap_remove :63 */
    memcpy(ord, design + first * d, d * sizeof(double)); perm[m++] = first;
    while (pool_n > 0) {                              /* :72 */
        double best = minimize ? INFINITY : -INFINITY;
        uint64_t bi = 0;
        for (uint64_t p = 0; p < pool_n; ++p) {       /* :78 */
            memcpy(ord + m * d, design + pool[p] * d, d * sizeof(double));
            double mv = criterion(ord, m + 1, d, metric); /* :84 */
            if (minimize ? (mv < best) : (mv > best)) { best = mv; bi = p; } /* :89-99 */
        }
        uint64_t pick = pool[bi]; pool[bi] = pool[--pool_n]; /* :101 */
        memcpy(ord + m * d, design + pick * d, d * sizeof(double)); perm[m++] = pick;
    }
    memcpy(out, ord, n * d * sizeof(double));
    free(pool); free(ord); free(center);
    return 0;
}

/* Same selection rule with O(n d) work per step.  Adding candidate c to the
 * ordered prefix changes the criterion only through
 *   maxpro : delta_c = sum_{i in prefix} 1/(prod_k (x_ik-x_ck)^2 + eps)
 *            psi = ((S + delta_c)/pairs)^(1/d) is monotone in delta_c
 *   maximin: m_c = min_{i in prefix} dist(i,c); value = min(cur_min, m_c) (exact)
 * Validated against oracle_order_design in tests at sizes where that finishes. */
int oracle_order_design_incremental(const double *design, uint64_t n, uint64_t d, int metric,
                                    int minimize, double *out, uint64_t *perm)
{
    if (n == 0 || d == 0) return 0;
    uint64_t *pool = (uint64_t *)malloc(n * sizeof(uint64_t));
    double *acc = (double *)malloc(n * sizeof(double)); /* by original index */
    double *center = (double *)malloc(d * sizeof(double));
    for (uint64_t k = 0; k < d; ++k) center[k] = 0.5;
    for (uint64_t i = 0; i < n; ++i) { pool[i] = i; acc[i] = metric == ORACLE_METRIC_MAXPRO ? 0.0 : INFINITY; }
    double dc = INFINITY; uint64_t ci = 0;
    for (uint64_t i = 0; i < n; ++i) {
        double pd = oracle_l2_distance(center, design + i * d, d);
        if (pd < dc) { dc = pd; ci = i; }
    }
    uint64_t pool_n = n, m = 0;
    uint64_t last = pool[ci]; pool[ci] = pool[--pool_n];
    perm[m++] = last;
    double s_prefix = 0.0;      /* maxpro sum over the ordered prefix */
    double cur_min = INFINITY;  /* maximin of the ordered prefix */
    while (pool_n > 0) {
        const double *xl = design + last * d;
        double best = minimize ? INFINITY : -INFINITY;
        uint64_t bi = 0;
        double n_pairs = (double)(m + 1) * ((double)(m + 1) - 1.0) / 2.0;
        for (uint64_t p = 0; p < pool_n; ++p) {
            uint64_t c = pool[p];
            const double *xc = design + c * d;
            double mv;
            if (metric == ORACLE_METRIC_MAXPRO) {
                double prod = 1.0;
                for (uint64_t k = 0; k < d; ++k) { double df = xl[k] - xc[k]; prod *= df * df; }
                acc[c] += 1.0 / (prod + 1e-12);
                mv = pow((s_prefix + acc[c]) / n_pairs, 1.0 / (double)d);
            } else {
                double dist = oracle_l2_distance(xl, xc, d);
                if (dist < acc[c]) acc[c] = dist;
                mv = acc[c] < cur_min ? acc[c] : cur_min;
            }
            if (minimize ? (mv < best) : (mv > best)) { best = mv; bi = p; }
        }
        uint64_t pick = pool[bi]; pool[bi] = pool[--pool_n];
        if (metric == ORACLE_METRIC_MAXPRO) s_prefix += acc[pick];
        else if (acc[pick] < cur_min) cur_min = acc[pick];
        perm[m++] = pick; last = pick;
    }
    for (uint64_t i = 0; i < n; ++i) memcpy(out + i * d, design + perm[i] * d, d * sizeof(double));
    free(pool); free(acc); free(center);
    return 0;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use every host core. */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
