/* TEST INFRASTRUCTURE ONLY -- see bp_oracle.h.  Plain-C restatement of the sbm-bp BP path.
 *
 * Arithmetic follows the reference operation by operation (same association order, no FMA
 * contraction: built with -ffp-contract=off) so that, on the same seed, it tracks the reference's
 * own compiled code to the last bit where libm agrees.  State is flat: the message that the
 * reference keeps in mmap_[i][l][q] lives at msg[(row_ptr[i]+l)*Q+q].
 */
#include "bp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LARGE_DEGREE 50u /* belief_propagation.h:68 */
#define EPS 1.0e-50      /* belief_propagation.h:69 */

/* ---- MT19937 (Matsumoto & Nishimura 1998), what std::mt19937 is ---- */
typedef struct {
    uint32_t s[624];
    int idx;
} mt_t;

static void mt_seed(mt_t *m, uint32_t seed) {
    m->s[0] = seed;
    for (int i = 1; i < 624; ++i) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}

static uint32_t mt_next(mt_t *m) {
    if (m->idx >= 624) {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (m->s[k] & 0x80000000u) | (m->s[(k + 1) % 624] & 0x7fffffffu);
            m->s[k] = m->s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->s[m->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* libstdc++ uniform_real_distribution<double>(0,1) == generate_canonical<double,53>: two 32-bit
 * draws, low word first, summed in double and divided by 2^64, clamped below 1. */
static double mt_uniform(mt_t *m) {
    double sum = 0.0, tmp = 1.0;
    sum += (double)mt_next(m) * tmp;
    tmp *= 4294967296.0;
    sum += (double)mt_next(m) * tmp;
    tmp *= 4294967296.0;
    double r = sum / tmp;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

struct orc {
    uint32_t N, Q, dc, E, max_deg;
    uint64_t M;
    uint64_t *row_ptr; /* N+1 */
    uint32_t *col;     /* graph_neis_ flattened */
    uint32_t *inv;     /* graph_neis_inv_ flattened */
    uint32_t *deg;
    uint32_t *conf_true;
    int32_t *conf_planted; /* belief_propagation.h:39: -1 = unknown; set by init_messages flags 1-3 */
    int conditional;       /* 1: bp_conditional (main.cpp:322, -m infer) -- planted nodes of degree < 50 are not updated */
    double beta;
    uint32_t *na;
    double *cab, *pab, *logcab, *eta, *logeta; /* Q or Q*Q, row-major [a][b] */
    double *msg, *marg, *h, *exph;
    double *na_expect, *nna_expect, *cab_expect;
    /* per-node scratch sized by max degree (belief_propagation.cpp:267-282) */
    double *field_iter, *mmap_total, *maxpom, *q_nb /* Q*max_deg */, *psi_q;
    mt_t rng;
};

/* ------------------------------------------------------------------ graph */

/* operator>>(unsigned) on a std::stringstream, as far as load_edge_list exercises it: returns 1 on
 * success; on a parse failure the target is zeroed (C++11) and the stream stays failed; at end of
 * input the sentry fails and the target is left alone. */
static int extract_uint(const char **p, int *failed, uint32_t *out) {
    if (*failed) return 0;
    const char *s = *p;
    while (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\v' || *s == '\f' || *s == '\n') ++s;
    if (*s == 0) {
        *failed = 1;
        *p = s;
        return 0;
    }
    const char *d = s;
    if (*d == '+') ++d;
    if (*d < '0' || *d > '9') {
        *out = 0;
        *failed = 1;
        *p = s;
        return 0;
    }
    unsigned long long v = 0;
    int over = 0;
    while (*d >= '0' && *d <= '9') {
        v = v * 10 + (unsigned)(*d - '0');
        if (v > 0xffffffffull) over = 1, v = 0xffffffffull;
        ++d;
    }
    *out = over ? 0xffffffffu : (uint32_t)v;
    if (over) *failed = 1;
    *p = d;
    return 1;
}

uint64_t orc_load_edge_list(const char *path, uint32_t *u, uint32_t *v, uint64_t cap) {
    FILE *f = fopen(path, "r");
    if (!f) return 0;
    uint64_t n = 0;
    uint32_t a = 0, b = 0; /* declared once outside the loop, graph_utilities.cpp:48 */
    size_t lcap = 1 << 16;
    char *line = (char *)malloc(lcap);
    while (fgets(line, (int)lcap, f)) {
        size_t len = strlen(line);
        while (len + 1 == lcap && line[len - 1] != '\n') { /* long line: grow and continue reading */
            lcap *= 2;
            line = (char *)realloc(line, lcap);
            if (!fgets(line + len, (int)(lcap - len), f)) break;
            len = strlen(line);
        }
        if (len && line[len - 1] == '\n') line[len - 1] = 0;
        const char *p = line;
        int failed = 0;
        extract_uint(&p, &failed, &a);
        extract_uint(&p, &failed, &b);
        if (n < cap) {
            u[n] = a;
            v[n] = b;
        }
        ++n;
    }
    free(line);
    fclose(f);
    return n;
}

static int cmp_u64(const void *x, const void *y) {
    uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
    return a < b ? -1 : a > b;
}

orc_t *orc_create(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, const uint32_t *block_sizes, uint32_t Q,
                  uint32_t dc_flag) {
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    uint32_t N = 0;
    for (uint32_t r = 0; r < Q; ++r) N += block_sizes[r];
    o->N = N;
    o->Q = Q;
    o->dc = dc_flag;
    o->beta = 1.0;
    /* edge_to_adj: a std::set per vertex == sorted unique (dst, src) pairs; self-loops stay once */
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (2 * n_pairs + 1));
    uint64_t nk = 0;
    for (uint64_t k = 0; k < n_pairs; ++k) {
        if (u[k] >= N || v[k] >= N) continue; /* ids must be < sum(n): the reference would index out of bounds */
        keys[nk++] = ((uint64_t)u[k] << 32) | v[k];
        keys[nk++] = ((uint64_t)v[k] << 32) | u[k];
    }
    qsort(keys, nk, sizeof(uint64_t), cmp_u64);
    uint64_t M = 0;
    for (uint64_t k = 0; k < nk; ++k)
        if (k == 0 || keys[k] != keys[k - 1]) keys[M++] = keys[k];
    o->M = M;
    o->row_ptr = (uint64_t *)calloc((size_t)N + 1, sizeof(uint64_t));
    o->col = (uint32_t *)malloc(sizeof(uint32_t) * (M + 1));
    o->inv = (uint32_t *)malloc(sizeof(uint32_t) * (M + 1));
    o->deg = (uint32_t *)calloc((size_t)N + 1, sizeof(uint32_t));
    for (uint64_t e = 0; e < M; ++e) {
        uint32_t i = (uint32_t)(keys[e] >> 32);
        o->col[e] = (uint32_t)keys[e];
        o->deg[i]++;
    }
    for (uint32_t i = 0; i < N; ++i) {
        o->row_ptr[i + 1] = o->row_ptr[i] + o->deg[i];
        if (o->deg[i] >= o->max_deg) o->max_deg = o->deg[i]; /* blockmodel.cpp:35-37 */
    }
    o->E = (uint32_t)(M / 2); /* blockmodel.cpp:32,46: directed count halved (a self-loop counts 1/2 -> truncated) */
    /* graph_neis_inv_: rank of i among j's neighbours (belief_propagation.cpp:260-262) */
    for (uint32_t i = 0; i < N; ++i) {
        for (uint64_t e = o->row_ptr[i]; e < o->row_ptr[i + 1]; ++e) {
            uint32_t j = o->col[e];
            uint64_t lo = o->row_ptr[j], hi = o->row_ptr[j + 1];
            while (lo < hi) {
                uint64_t mid = (lo + hi) / 2;
                if (o->col[mid] < i) lo = mid + 1;
                else hi = mid;
            }
            o->inv[e] = (uint32_t)(lo - o->row_ptr[j]);
        }
    }
    free(keys);
    /* main.cpp:239-252, :284-286: true conf defaults to the -n block ordering */
    o->conf_true = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)N + 1));
    o->conf_planted = (int32_t *)malloc(sizeof(int32_t) * ((size_t)N + 1));
    for (uint32_t i = 0; i < N; ++i) o->conf_planted[i] = -1; /* belief_propagation.cpp:284 */
    o->conditional = 0;
    uint32_t shift = 0;
    for (uint32_t r = 0; r < Q; ++r) {
        for (uint32_t i = 0; i < block_sizes[r]; ++i) o->conf_true[shift + i] = r;
        shift += block_sizes[r];
    }
    size_t QQ = (size_t)Q * Q, md = o->max_deg ? o->max_deg : 1;
    o->na = (uint32_t *)calloc(Q, sizeof(uint32_t));
    o->cab = (double *)calloc(QQ, sizeof(double));
    o->pab = (double *)calloc(QQ, sizeof(double));
    o->logcab = (double *)calloc(QQ, sizeof(double));
    o->eta = (double *)calloc(Q, sizeof(double));
    o->logeta = (double *)calloc(Q, sizeof(double));
    o->msg = (double *)calloc((size_t)M * Q + 1, sizeof(double));
    o->marg = (double *)calloc((size_t)N * Q + 1, sizeof(double));
    o->h = (double *)calloc(Q, sizeof(double));
    o->exph = (double *)calloc(Q, sizeof(double));
    o->na_expect = (double *)calloc(Q, sizeof(double));
    o->nna_expect = (double *)calloc(Q, sizeof(double));
    o->cab_expect = (double *)calloc(QQ, sizeof(double));
    o->field_iter = (double *)calloc(md, sizeof(double));
    o->mmap_total = (double *)calloc(md, sizeof(double));
    o->maxpom = (double *)calloc(md, sizeof(double));
    o->q_nb = (double *)calloc(md * Q, sizeof(double));
    o->psi_q = (double *)calloc(Q, sizeof(double));
    mt_seed(&o->rng, 5489u);
    return o;
}

void orc_destroy(orc_t *o) {
    if (!o) return;
    free(o->row_ptr); free(o->col); free(o->inv); free(o->deg); free(o->conf_true); free(o->conf_planted);
    free(o->na); free(o->cab); free(o->pab); free(o->logcab); free(o->eta); free(o->logeta);
    free(o->msg); free(o->marg); free(o->h); free(o->exph);
    free(o->na_expect); free(o->nna_expect); free(o->cab_expect);
    free(o->field_iter); free(o->mmap_total); free(o->maxpom); free(o->q_nb); free(o->psi_q);
    free(o);
}

uint32_t orc_N(const orc_t *o) { return o->N; }
uint32_t orc_Q(const orc_t *o) { return o->Q; }
uint64_t orc_M(const orc_t *o) { return o->M; }
uint32_t orc_E(const orc_t *o) { return o->E; }
uint32_t orc_max_degree(const orc_t *o) { return o->max_deg; }

void orc_get_csr(const orc_t *o, uint64_t *row_ptr, uint32_t *col, uint32_t *rev_local, uint64_t *rev_global) {
    memcpy(row_ptr, o->row_ptr, sizeof(uint64_t) * ((size_t)o->N + 1));
    for (uint64_t e = 0; e < o->M; ++e) {
        col[e] = o->col[e];
        if (rev_local) rev_local[e] = o->inv[e];
        if (rev_global) rev_global[e] = o->row_ptr[o->col[e]] + o->inv[e];
    }
}

/* ------------------------------------------------------------------ parameters */

/* belief_propagation.cpp:290-317 */
static void expand_params(orc_t *o) {
    uint32_t Q = o->Q;
    for (uint32_t q = 0; q < Q; ++q) {
        o->eta[q] = 1.0 * o->na[q] / o->N;
        o->logeta[q] = log(o->eta[q]);
        for (uint32_t j = 0; j < Q; ++j) {
            o->pab[q * Q + j] = o->cab[q * Q + j] / o->N;
            o->logcab[q * Q + j] = log(o->cab[q * Q + j]);
        }
    }
}

/* blockmodel.cpp:274-302.  The na[Q-1] = N - tot_size assignment (:282-284) is overwritten at :286,
 * so eta need not sum to one; --cab is the upper triangle in row-major order (:295-297). */
void orc_set_params_direct(orc_t *o, const double *pa, const double *cab_upper) {
    uint32_t Q = o->Q;
    for (uint32_t q = 0; q < Q; ++q) o->na[q] = (uint32_t)(int)(pa[q] * o->N);
    for (uint32_t q = 0; q < Q; ++q) {
        uint32_t base = q * Q - q * (q - 1) / 2;
        o->cab[q * Q + q] = cab_upper[base];
        for (uint32_t t = q + 1; t < Q; ++t) {
            o->cab[q * Q + t] = cab_upper[base + t - q];
            o->cab[t * Q + q] = o->cab[q * Q + t];
        }
    }
    expand_params(o);
}

/* blockmodel.cpp:229-272 */
void orc_set_params_epsilon_c(orc_t *o, double eps, double c) {
    uint32_t Q = o->Q;
    double cin, co;
    for (uint32_t q = 0; q < Q; ++q) {
        double pa = 1.0 / Q;
        o->na[q] = (uint32_t)(int)(pa * o->N); /* :248 overwrites the remainder fix-up of :243-245 */
    }
    if (eps < 0) {
        cin = 0;
        co = c * Q / (Q - 1);
    } else {
        cin = c * Q / ((Q - 1) * eps + 1);
        co = eps * cin;
    }
    for (uint32_t q = 0; q < Q; ++q) {
        o->cab[q * Q + q] = cin;
        for (uint32_t t = q + 1; t < Q; ++t) {
            o->cab[q * Q + t] = co;
            o->cab[t * Q + q] = co;
        }
    }
    expand_params(o);
}

void orc_set_params_raw(orc_t *o, const uint32_t *na, const double *cab) {
    memcpy(o->na, na, sizeof(uint32_t) * o->Q);
    memcpy(o->cab, cab, sizeof(double) * o->Q * o->Q);
    expand_params(o);
}

void orc_get_params(const orc_t *o, uint32_t *na, double *cab, double *eta) {
    if (na) memcpy(na, o->na, sizeof(uint32_t) * o->Q);
    if (cab) memcpy(cab, o->cab, sizeof(double) * o->Q * o->Q);
    if (eta) memcpy(eta, o->eta, sizeof(double) * o->Q);
}

void orc_set_beta(orc_t *o, double beta) { o->beta = beta; }

/* ------------------------------------------------------------------ state */

void orc_seed(orc_t *o, uint32_t seed) { mt_seed(&o->rng, seed); }
double orc_uniform(orc_t *o) { return mt_uniform(&o->rng); }

/* ---- std::shuffle over n elements as libstdc++ (GCC >= 11, bits/stl_algo.h + bits/uniform_int_dist.h) runs it with a
 * std::mt19937: what blockmodel_t::shuffle (blockmodel.cpp:103-106) costs the run's generator under --mb_rand
 * (main.cpp:299-301).  Third-party algorithm, restated from its published source:
 *   - a bounded integer in [0, range) from a 32-bit generator is Lemire's nearly-divisionless method (_S_nd): the high
 *     word of draw * range, redrawing while the low word falls under (-range) % range;
 *   - while n * n fits the generator's range, two swap positions come from ONE bounded draw over b0 * b1
 *     (__gen_two_uniform_ints: x / b1, x % b1); an even n takes a single {0, 1} draw first;
 *   - otherwise element i swaps with a bounded draw from [0, i].
 * perm (n entries, or NULL) receives the permutation applied to 0..n-1. */
static uint32_t mt_bounded(mt_t *m, uint32_t range) {
    uint64_t product = (uint64_t)mt_next(m) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        const uint32_t threshold = (uint32_t)(-range) % range;
        while (low < threshold) {
            product = (uint64_t)mt_next(m) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return (uint32_t)(product >> 32);
}

static void swap_u32(uint32_t *p, uint64_t a, uint64_t b) {
    if (!p) return;
    uint32_t t = p[a];
    p[a] = p[b];
    p[b] = t;
}

static void mt_shuffle(mt_t *m, uint32_t *perm, uint64_t n) {
    if (n == 0) return;
    const uint64_t urngrange = 0xffffffffull;
    if (urngrange / n >= n) {
        uint64_t i = 1;
        if ((n % 2) == 0) swap_u32(perm, i++, mt_bounded(m, 2u));
        while (i != n) {
            const uint64_t swap_range = i + 1;
            const uint32_t x = mt_bounded(m, (uint32_t)(swap_range * (swap_range + 1)));
            swap_u32(perm, i, x / (swap_range + 1));
            ++i;
            swap_u32(perm, i, x % (swap_range + 1));
            ++i;
        }
        return;
    }
    for (uint64_t i = 1; i != n; ++i) swap_u32(perm, i, mt_bounded(m, (uint32_t)(i + 1)));
}

/* --mb_rand: seed, shuffle the N memberships (their order is not read by the BP path), then init_messages flag 0 */
void orc_init_messages_mb_rand(orc_t *o, uint32_t seed) {
    mt_seed(&o->rng, seed);
    mt_shuffle(&o->rng, NULL, o->N);
    orc_init_messages_continue(o);
}

void orc_shuffle(orc_t *o, uint32_t *perm, uint64_t n) { mt_shuffle(&o->rng, perm, n); }

/* belief_propagation.cpp:110-131: per node, Q uniforms -> normalised marginal; then per neighbour in
 * ascending order Q uniforms -> normalised OUTGOING message stored in the neighbour's in-slot. */
void orc_init_messages(orc_t *o, uint32_t seed) {
    mt_seed(&o->rng, seed);
    orc_init_messages_continue(o);
}

/* the same from the generator where it stands */
void orc_init_messages_continue(orc_t *o) {
    uint32_t Q = o->Q;
    for (uint32_t i = 0; i < o->N; ++i) {
        double norm = 0.0;
        double *mp = o->marg + (size_t)i * Q;
        for (uint32_t q = 0; q < Q; ++q) {
            mp[q] = mt_uniform(&o->rng);
            norm += mp[q];
        }
        for (uint32_t q = 0; q < Q; ++q) mp[q] /= norm;
        for (uint64_t e = o->row_ptr[i]; e < o->row_ptr[i + 1]; ++e) {
            double *slot = o->msg + (o->row_ptr[o->col[e]] + o->inv[e]) * Q;
            norm = 0.0;
            for (uint32_t q = 0; q < Q; ++q) {
                slot[q] = mt_uniform(&o->rng);
                norm += slot[q];
            }
            for (uint32_t q = 0; q < Q; ++q) slot[q] /= norm;
        }
    }
}

/* belief_propagation.cpp:101-215, all four flags, draw for draw.  conf has N entries (-1 = unknown).
 * Quirks kept: flag 1 plants one-hot marginals / outgoing messages and draws the rest; flag 2 loops q OUTSIDE the
 * neighbours, writes the node's own INCOMING slots, un-normalised, with a float noise constant (:177); flag 3's loop
 * advances its index twice per turn (:205), so only even-ranked neighbours receive the one-hot message and the other
 * slots keep the zeros of bp_allocate.  Flags 2 and 3 carry assert(conf_planted_[i] != 1) (:179,:197): the reference
 * aborts when a planted label equals 1; this function then returns -1 and leaves the state alone. */
int orc_init_messages_flag(orc_t *o, uint32_t flag, const int32_t *conf, uint32_t seed) {
    const uint32_t Q = o->Q;
    if (flag == 0) {
        orc_init_messages(o, seed);
        return 0;
    }
    if (flag > 3 || !conf) return -2;
    if (flag >= 2)
        for (uint32_t i = 0; i < o->N; ++i)
            if (conf[i] == 1) return -1;
    mt_seed(&o->rng, seed);
    memset(o->msg, 0, sizeof(double) * (size_t)o->M * Q);       /* bp_allocate: fresh zero vectors */
    memset(o->marg, 0, sizeof(double) * (size_t)o->N * Q);
    memcpy(o->conf_planted, conf, sizeof(int32_t) * o->N);
    const float planted_noise = 0.1f; /* :177 */
    for (uint32_t i = 0; i < o->N; ++i) {
        double *mp = o->marg + (size_t)i * Q;
        const uint64_t r0 = o->row_ptr[i];
        const uint32_t d = o->deg[i];
        if (flag == 1) {
            double norm = 0.0;
            if (conf[i] != -1) {
                for (uint32_t q = 0; q < Q; ++q) mp[q] = (q == (uint32_t)conf[i]) ? 1.0 : 0.0;
            } else {
                for (uint32_t q = 0; q < Q; ++q) {
                    mp[q] = mt_uniform(&o->rng);
                    norm += mp[q];
                }
                for (uint32_t q = 0; q < Q; ++q) mp[q] /= norm;
            }
            for (uint32_t l = 0; l < d; ++l) {
                double *slot = o->msg + (o->row_ptr[o->col[r0 + l]] + o->inv[r0 + l]) * Q;
                if (conf[i] != -1) {
                    for (uint32_t q = 0; q < Q; ++q) slot[q] = (q == (uint32_t)conf[i]) ? 1.0 : 0.0;
                } else {
                    norm = 0.0;
                    for (uint32_t q = 0; q < Q; ++q) {
                        slot[q] = mt_uniform(&o->rng);
                        norm += slot[q];
                    }
                    for (uint32_t q = 0; q < Q; ++q) slot[q] /= norm;
                }
            }
        } else if (flag == 2) {
            for (uint32_t q = 0; q < Q; ++q) {
                if (q == (uint32_t)conf[i]) mp[q] = planted_noise + (1.0 - planted_noise) * mt_uniform(&o->rng);
                else mp[q] = mt_uniform(&o->rng) * (1.0 - planted_noise);
                for (uint32_t l = 0; l < d; ++l) {
                    double *slot = o->msg + (r0 + l) * Q; /* mmap_[i][idxij]: the node's own in-slots */
                    if (q == (uint32_t)conf[i]) slot[q] = planted_noise + (1.0 - planted_noise) * mt_uniform(&o->rng);
                    else slot[q] = mt_uniform(&o->rng) * (1.0 - planted_noise);
                }
            }
        } else {
            for (uint32_t q = 0; q < Q; ++q) mp[q] = (q == (uint32_t)conf[i]) ? 1.0 : 0.000;
            for (uint32_t l = 0; l < d; l += 2) { /* :203-205: idxij++ inside the body and in the loop header */
                double *slot = o->msg + (o->row_ptr[o->col[r0 + l]] + o->inv[r0 + l]) * Q;
                for (uint32_t q = 0; q < Q; ++q) slot[q] = (q == (uint32_t)conf[i]) ? 1.0 : 0.0;
            }
        }
    }
    return 0;
}

/* main.cpp:318-323: -m infer runs bp_conditional, -m learn bp_basic */
void orc_set_conditional(orc_t *o, int on) { o->conditional = on ? 1 : 0; }
void orc_get_conf_planted(const orc_t *o, int32_t *conf) { memcpy(conf, o->conf_planted, sizeof(int32_t) * o->N); }

void orc_get_state(const orc_t *o, double *msg, double *marg, double *h) {
    if (msg) memcpy(msg, o->msg, sizeof(double) * o->M * o->Q);
    if (marg) memcpy(marg, o->marg, sizeof(double) * (size_t)o->N * o->Q);
    if (h) memcpy(h, o->h, sizeof(double) * o->Q);
}

/* belief_propagation.cpp:334-360 */
static void update_h(orc_t *o, uint32_t i, int mode) {
    uint32_t Q = o->Q;
    double di = o->deg[i];
    const double *psi = o->marg + (size_t)i * Q;
    for (uint32_t q1 = 0; q1 < Q; ++q1) {
        for (uint32_t q2 = 0; q2 < Q; ++q2) {
            double term = (o->dc == 0) ? o->cab[q2 * Q + q1] * psi[q2] : di * o->cab[q2 * Q + q1] * psi[q2];
            if (mode < 0) o->h[q1] -= term;
            else o->h[q1] += term;
        }
    }
}

/* belief_propagation.cpp:363-368 */
static void update_exph(orc_t *o) {
    for (uint32_t q = 0; q < o->Q; ++q) o->exph[q] = exp(-o->beta * o->h[q] / o->N);
}

/* belief_propagation.cpp:320-332 */
void orc_init_h(orc_t *o) {
    for (uint32_t q = 0; q < o->Q; ++q) o->h[q] = 0.0;
    for (uint32_t i = 0; i < o->N; ++i) update_h(o, i, +1);
    update_exph(o);
}

void orc_set_state(orc_t *o, const double *msg, const double *marg) {
    if (msg) memcpy(o->msg, msg, sizeof(double) * o->M * o->Q);
    if (marg) memcpy(o->marg, marg, sizeof(double) * (size_t)o->N * o->Q);
    orc_init_h(o);
}

/* ------------------------------------------------------------------ node update */

/* the edge kernel K(t,q;i,l) times the message, as each dc branch associates it
 * (belief_propagation.cpp:1003-1011); with_beta=0 is the large-degree path (:834-842) */
static inline double kernel_term(const orc_t *o, int with_beta, double di, double dn, uint32_t t, uint32_t q,
                                 double m) {
    uint32_t Q = o->Q;
    if (o->dc == 0) {
        double c = o->cab[t * Q + q];
        return (with_beta ? pow(c, o->beta) : c) * m;
    } else if (o->dc == 1) {
        return di * dn * o->cab[t * Q + q] * m;
    } else {
        double tmp = di * dn * o->pab[t * Q + q];
        return tmp / (1.0 + tmp) * m;
    }
}

/* belief_propagation.cpp:1079-1098 (bp_basic; bp_conditional is identical while conf_planted_ == -1):
 * clean (:422-426), sum_all_messages_to_i (:991-1049), h -= (:1088), norm_m_at_i (:1051-1071), h += , exph */
static double update_small(orc_t *o, uint32_t i, double damping) {
    const uint32_t Q = o->Q;
    const uint64_t r0 = o->row_ptr[i];
    const uint32_t d = o->deg[i];
    const double di = d;
    for (uint32_t l = 0; l < d; ++l) o->mmap_total[l] = 0.;
    double psi_total = 0.0;
    for (uint32_t q = 0; q < Q; ++q) {
        double a = 1.0;
        for (uint32_t l = 0; l < d; ++l) {
            double b = 0.0;
            double dn = o->deg[o->col[r0 + l]];
            for (uint32_t t = 0; t < Q; ++t) b += kernel_term(o, 1, di, dn, t, q, o->msg[(r0 + l) * Q + t]);
            if (b == 0.) continue; /* :1013-1016: field_iter_[l] keeps its stale value */
            a *= b;
            o->field_iter[l] = b;
        }
        if (o->dc == 0) o->psi_q[q] = a * o->eta[q] * o->exph[q];
        else o->psi_q[q] = a * o->eta[q] * exp(-1.0 * di * o->h[q] / o->N);
        psi_total += o->psi_q[q];
        for (uint32_t l = 0; l < d; ++l) {
            if (o->field_iter[l] < EPS) { /* :1029-1042: the fallback drops eta and the field */
                double tmprob = 1.0;
                for (uint32_t lx = 0; lx < d; ++lx) {
                    if (lx == l) continue;
                    if (o->field_iter[lx] != 0) tmprob *= o->field_iter[lx];
                }
                o->q_nb[(size_t)q * o->max_deg + l] = tmprob;
            } else {
                o->q_nb[(size_t)q * o->max_deg + l] = o->psi_q[q] / o->field_iter[l];
            }
            o->mmap_total[l] += o->q_nb[(size_t)q * o->max_deg + l];
        }
    }
    update_h(o, i, -1);
    double mymaxdiff = -100.0;
    for (uint32_t q = 0; q < Q; ++q) {
        o->marg[(size_t)i * Q + q] = o->psi_q[q] / psi_total;
        for (uint32_t l = 0; l < d; ++l) {
            double *slot = o->msg + (o->row_ptr[o->col[r0 + l]] + o->inv[r0 + l]) * Q;
            double nb = o->q_nb[(size_t)q * o->max_deg + l];
            double mydiff = fabs(slot[q] - nb / o->mmap_total[l]);
            if (mydiff > mymaxdiff) mymaxdiff = mydiff;
            slot[q] = (damping)*nb / o->mmap_total[l] + (1.0 - damping) * slot[q];
        }
    }
    update_h(o, i, +1);
    update_exph(o);
    return mymaxdiff;
}

/* belief_propagation.cpp:813-890: log domain, beta ignored (:835,:850) */
static double update_large(orc_t *o, uint32_t i, double damping) {
    const uint32_t Q = o->Q;
    const uint64_t r0 = o->row_ptr[i];
    const uint32_t d = o->deg[i];
    const double di = d;
    double psi_total = 0.0;
    for (uint32_t l = 0; l < d; ++l) o->mmap_total[l] = 0.;
    double maxpom_psi = -100000000.0;
    for (uint32_t l = 0; l < d; ++l) o->maxpom[l] = -100000000.0;
    for (uint32_t q = 0; q < Q; ++q) {
        double a = 0.0;
        for (uint32_t l = 0; l < d; ++l) {
            double b = 0.0;
            double dn = o->deg[o->col[r0 + l]];
            for (uint32_t t = 0; t < Q; ++t) b += kernel_term(o, 0, di, dn, t, q, o->msg[(r0 + l) * Q + t]);
            double tmp = log(b);
            a += tmp;
            o->field_iter[l] = tmp;
        }
        if (o->dc == 0) o->psi_q[q] = a + o->logeta[q] - o->h[q] / o->N;
        else o->psi_q[q] = a + o->logeta[q] - 1.0 * di * o->h[q] / o->N;
        if (o->psi_q[q] > maxpom_psi) maxpom_psi = o->psi_q[q];
        for (uint32_t l = 0; l < d; ++l) {
            double v = o->psi_q[q] - o->field_iter[l];
            o->q_nb[(size_t)q * o->max_deg + l] = v;
            if (v > o->maxpom[l]) o->maxpom[l] = v;
        }
    }
    for (uint32_t q = 0; q < Q; ++q) {
        psi_total += exp(o->psi_q[q] - maxpom_psi);
        for (uint32_t l = 0; l < d; ++l)
            o->mmap_total[l] += exp(o->q_nb[(size_t)q * o->max_deg + l] - o->maxpom[l]);
    }
    update_h(o, i, -1);
    double mymaxdiff = -100.0;
    for (uint32_t q = 0; q < Q; ++q) {
        o->marg[(size_t)i * Q + q] = exp(o->psi_q[q] - maxpom_psi) / psi_total;
        for (uint32_t l = 0; l < d; ++l) {
            double *slot = o->msg + (o->row_ptr[o->col[r0 + l]] + o->inv[r0 + l]) * Q;
            double thisvalue = exp(o->q_nb[(size_t)q * o->max_deg + l] - o->maxpom[l]) / o->mmap_total[l];
            double mydiff = fabs(slot[q] - thisvalue);
            if (mydiff > mymaxdiff) mymaxdiff = mydiff;
            slot[q] = (damping)*thisvalue + (1 - damping) * slot[q];
        }
    }
    update_h(o, i, +1);
    update_exph(o);
    return mymaxdiff;
}

/* dispatch of belief_propagation.cpp:397-401 */
double orc_update_node(orc_t *o, uint32_t i, double damping) {
    if (o->deg[i] >= LARGE_DEGREE) return update_large(o, i, damping); /* ignores conf_planted_ */
    /* bp_conditional::bp_iter_update_psi (:1100-1126): a planted node emits constant messages */
    if (o->conditional && o->conf_planted[i] != -1) return 0;
    return update_small(o, i, damping);
}

double orc_jacobi_sweep(orc_t *o, double damping, double *new_msg, double *new_marg, double *node_diff) {
    const uint32_t Q = o->Q;
    orc_init_h(o);
    double *h0 = (double *)malloc(sizeof(double) * 2 * Q);
    memcpy(h0, o->h, sizeof(double) * Q);
    memcpy(h0 + Q, o->exph, sizeof(double) * Q);
    double *saved = (double *)malloc(sizeof(double) * ((size_t)o->max_deg + 1) * Q);
    double maxdiff = -100.0;
    for (uint32_t i = 0; i < o->N; ++i) {
        const uint64_t r0 = o->row_ptr[i];
        const uint32_t d = o->deg[i];
        for (uint32_t l = 0; l < d; ++l)
            memcpy(saved + (size_t)l * Q, o->msg + (o->row_ptr[o->col[r0 + l]] + o->inv[r0 + l]) * Q,
                   sizeof(double) * Q);
        memcpy(saved + (size_t)d * Q, o->marg + (size_t)i * Q, sizeof(double) * Q);
        double diff = orc_update_node(o, i, damping);
        if (diff > maxdiff) maxdiff = diff;
        if (node_diff) node_diff[i] = diff;
        for (uint32_t l = 0; l < d; ++l) {
            uint64_t g = o->row_ptr[o->col[r0 + l]] + o->inv[r0 + l];
            if (new_msg) memcpy(new_msg + g * Q, o->msg + g * Q, sizeof(double) * Q);
            memcpy(o->msg + g * Q, saved + (size_t)l * Q, sizeof(double) * Q);
        }
        if (new_marg) memcpy(new_marg + (size_t)i * Q, o->marg + (size_t)i * Q, sizeof(double) * Q);
        memcpy(o->marg + (size_t)i * Q, saved + (size_t)d * Q, sizeof(double) * Q);
        memcpy(o->h, h0, sizeof(double) * Q);
        memcpy(o->exph, h0 + Q, sizeof(double) * Q);
    }
    free(saved);
    free(h0);
    return maxdiff;
}

/* ------------------------------------------------------------------ extended-precision referee
 * The SAME mathematical update as update_small / update_large (belief_propagation.cpp:991-1071, :813-890) -- kernel,
 * field term and beta handling of whichever routine the reference dispatches to (:397-401) -- evaluated in long
 * double (x87, 64-bit mantissa, eps 1.1e-19) in the log domain from the frozen state, like orc_jacobi_sweep.  It
 * restates no reference code path: it is the yardstick that says which of two FP64 evaluations (the reference's
 * sequential sum of d logarithms, the engine's tree reductions) is closer to the exact value at hub nodes.
 * h is taken as the reference holds it (double, init_h) unless h_in is given: the exponent d_i h_q / N of a hub is
 * ill-conditioned in h (a last-bit difference in h moves a degree-400 dc update by ~1e-12), so an implementation is
 * measured against the exact update for ITS OWN h, and its h against the reference's separately.  The b < EPS fallback (:1029-1042) is not a function of the
 * inputs alone (stale scratch) and is not modelled: nodes that hit it are flagged in skipped[] and left unchanged. */
int orc_referee_sweep(orc_t *o, double damping, const double *h_in, double *new_msg, double *new_marg, uint8_t *skipped) {
    const uint32_t Q = o->Q;
    int nskipped = 0;
    orc_init_h(o);
    const double *hh = h_in ? h_in : o->h; /* the field the update is evaluated with (default: init_h of this state) */
    long double *L = (long double *)malloc(sizeof(long double) * Q);
    long double *lb = (long double *)malloc(sizeof(long double) * ((size_t)o->max_deg + 1) * Q);
    long double *cav = (long double *)malloc(sizeof(long double) * Q);
    if (new_msg) memcpy(new_msg, o->msg, sizeof(double) * o->M * Q);
    for (uint32_t i = 0; i < o->N; ++i) {
        const uint64_t r0 = o->row_ptr[i];
        const uint32_t d = o->deg[i];
        const int large = d >= LARGE_DEGREE;
        if (skipped) skipped[i] = 0;
        if (new_marg) memcpy(new_marg + (size_t)i * Q, o->marg + (size_t)i * Q, sizeof(double) * Q);
        if (!large && o->conditional && o->conf_planted[i] != -1) continue; /* frozen (:1100-1126) */
        const long double di = d;
        int tiny = 0;
        for (uint32_t q = 0; q < Q; ++q) {
            long double a = 0.0L;
            for (uint32_t l = 0; l < d; ++l) {
                const long double dn = o->deg[o->col[r0 + l]];
                long double b = 0.0L;
                for (uint32_t t = 0; t < Q; ++t) {
                    const long double m = o->msg[(r0 + l) * Q + t];
                    long double k;
                    if (o->dc == 0) {
                        const long double c = o->cab[t * Q + q];
                        k = large ? c : powl(c, (long double)o->beta);
                    } else if (o->dc == 1) {
                        k = di * dn * (long double)o->cab[t * Q + q];
                    } else {
                        const long double tau = di * dn * (long double)o->pab[t * Q + q];
                        k = tau / (1.0L + tau);
                    }
                    b += k * m;
                }
                if (!large && !(b >= (long double)EPS)) tiny = 1;
                lb[(size_t)l * Q + q] = logl(b);
                a += lb[(size_t)l * Q + q];
            }
            long double fexp;
            if (o->dc == 0) fexp = (large ? 1.0L : (long double)o->beta) * (long double)hh[q] / (long double)o->N;
            else fexp = di * (long double)hh[q] / (long double)o->N;
            L[q] = a + logl((long double)o->eta[q]) - fexp;
        }
        if (tiny) {
            if (skipped) skipped[i] = 1;
            ++nskipped;
            continue;
        }
        long double mx = L[0], tot = 0.0L;
        for (uint32_t q = 1; q < Q; ++q) mx = L[q] > mx ? L[q] : mx;
        for (uint32_t q = 0; q < Q; ++q) tot += expl(L[q] - mx);
        if (new_marg)
            for (uint32_t q = 0; q < Q; ++q) new_marg[(size_t)i * Q + q] = (double)(expl(L[q] - mx) / tot);
        for (uint32_t l = 0; l < d; ++l) {
            long double cm = L[0] - lb[(size_t)l * Q], ct = 0.0L;
            for (uint32_t q = 0; q < Q; ++q) {
                cav[q] = L[q] - lb[(size_t)l * Q + q];
                cm = cav[q] > cm ? cav[q] : cm;
            }
            for (uint32_t q = 0; q < Q; ++q) ct += expl(cav[q] - cm);
            const uint64_t g = o->row_ptr[o->col[r0 + l]] + o->inv[r0 + l];
            if (new_msg)
                for (uint32_t q = 0; q < Q; ++q) {
                    const long double v = expl(cav[q] - cm) / ct;
                    new_msg[g * Q + q] = (double)((long double)damping * v + (1.0L - (long double)damping) * (long double)o->msg[g * Q + q]);
                }
        }
    }
    free(L);
    free(lb);
    free(cav);
    return nskipped;
}

double orc_sync_sweep(orc_t *o, double damping) {
    size_t nm = (size_t)o->M * o->Q, nn = (size_t)o->N * o->Q;
    double *nmsg = (double *)malloc(sizeof(double) * (nm + 1));
    double *nmarg = (double *)malloc(sizeof(double) * (nn + 1));
    double md = orc_jacobi_sweep(o, damping, nmsg, nmarg, NULL);
    memcpy(o->msg, nmsg, sizeof(double) * nm);
    memcpy(o->marg, nmarg, sizeof(double) * nn);
    free(nmsg);
    free(nmarg);
    orc_init_h(o);
    return md;
}

int orc_sync_converge(orc_t *o, float crit, uint32_t max_iter, float damping) {
    for (int it = 0; it < (int)max_iter; ++it) {
        double md = orc_sync_sweep(o, damping);
        if (md < crit) return it;
    }
    return -1;
}

/* belief_propagation.cpp:386-415 */
int orc_converge(orc_t *o, float bp_err, uint32_t max_iter_time, float dumping_rate) {
    orc_init_h(o);
    for (int iter_time = 0; iter_time < (int)max_iter_time; ++iter_time) {
        double maxdiffm = -100.0;
        for (uint32_t k = 0; k < o->N; ++k) {
            uint32_t i = (uint32_t)(int)(mt_uniform(&o->rng) * o->N);
            double diffm = orc_update_node(o, i, dumping_rate);
            if (diffm > maxdiffm) maxdiffm = diffm;
        }
        if (maxdiffm < bp_err) return iter_time;
    }
    return -1;
}

/* ------------------------------------------------------------------ free energy */

/* belief_propagation.cpp:442-504 (the unused _diff_ twins are dropped) */
double orc_f_site(orc_t *o) {
    const uint32_t Q = o->Q;
    double f_site = 0;
    for (uint32_t i = 0; i < o->N; ++i) {
        double di = o->deg[i];
        double rescale = -100000.;
        for (uint32_t q = 0; q < Q; ++q) {
            double a = 0.0;
            for (uint64_t e = o->row_ptr[i]; e < o->row_ptr[i + 1]; ++e) {
                double b = 0;
                double dn = o->deg[o->col[e]];
                for (uint32_t t = 0; t < Q; ++t) b += kernel_term(o, 1, di, dn, t, q, o->msg[e * Q + t]);
                a += log(b);
            }
            if (o->dc == 0) o->psi_q[q] = a + o->logeta[q] - o->beta * o->h[q] / o->N;
            else o->psi_q[q] = a + o->logeta[q] - di * o->h[q] / o->N;
            if (o->psi_q[q] > rescale) rescale = o->psi_q[q];
        }
        double norm = 0.;
        for (uint32_t q = 0; q < Q; ++q) norm += exp(o->psi_q[q] - rescale);
        f_site += rescale + log(norm);
    }
    return f_site / o->N;
}

/* symmetric two-point weight of a directed edge (:576-605 / :907-931): sum over q1<=q2 of
 * K(q1,q2) * (psi_in[q1]*psi_out[q2] [+ psi_in[q2]*psi_out[q1]]) */
static double edge_norm(const orc_t *o, int with_beta, double di, double dl, const double *min, const double *mout) {
    const uint32_t Q = o->Q;
    double norm_L = 0;
    for (uint32_t q1 = 0; q1 < Q; ++q1) {
        for (uint32_t q2 = q1; q2 < Q; ++q2) {
            double pair = (q1 == q2) ? (min[q1] * mout[q2]) : (min[q1] * mout[q2] + min[q2] * mout[q1]);
            if (o->dc == 0) {
                double c = o->cab[q1 * Q + q2];
                norm_L += (with_beta ? pow(c, o->beta) : c) * pair;
            } else if (o->dc == 1) {
                norm_L += di * dl * o->cab[q1 * Q + q2] * pair;
            } else {
                double tmp = di * dl * o->pab[q1 * Q + q2];
                norm_L += tmp / (1.0 + tmp) * pair;
            }
        }
    }
    return norm_L;
}

/* belief_propagation.cpp:562-612 */
double orc_f_edge(orc_t *o) {
    const uint32_t Q = o->Q;
    double f_link = 0;
    for (uint32_t i = 0; i < o->N; ++i) {
        double di = o->deg[i];
        for (uint64_t e = o->row_ptr[i]; e < o->row_ptr[i + 1]; ++e) {
            uint32_t i2 = o->col[e];
            uint64_t e2 = o->row_ptr[i2] + o->inv[e];
            f_link += log(edge_norm(o, 1, di, (double)o->deg[i2], o->msg + e * Q, o->msg + e2 * Q));
        }
    }
    return f_link / (2. * o->N);
}

/* belief_propagation.cpp:675-709: O(N^2); dc != 0 adds nothing (:692-697) */
double orc_f_non_edge(orc_t *o) {
    const uint32_t Q = o->Q;
    double acc = 0;
    for (uint32_t i = 0; i < o->N; ++i) {
        for (uint32_t l = 0; l < o->N; ++l) {
            int is_nb = 0;
            for (uint64_t e = o->row_ptr[i]; e < o->row_ptr[i + 1]; ++e)
                if (o->col[e] == l) { is_nb = 1; break; }
            if (is_nb) continue;
            double f = 0;
            if (o->dc == 0)
                for (uint32_t q1 = 0; q1 < Q; ++q1)
                    for (uint32_t q2 = 0; q2 < Q; ++q2)
                        f += pow((1 - o->cab[q1 * Q + q2] / o->N), o->beta) * o->marg[(size_t)i * Q + q1] *
                             o->marg[(size_t)l * Q + q2];
            if (f != 0.) acc += log(f);
        }
    }
    return acc / (2. * o->N);
}

/* belief_propagation.cpp:744-750 */
double orc_free_energy(orc_t *o) {
    double f_site = -orc_f_site(o);
    double f_link = orc_f_edge(o);
    double f_non_edge = orc_f_non_edge(o);
    return f_site + f_link + f_non_edge;
}

/* ------------------------------------------------------------------ entropy (first stdout field) */

/* belief_propagation.cpp:506-560 (dc == 0 only; otherwise 0/0 = NaN as the reference yields) */
double orc_entropy_site(orc_t *o) {
    const uint32_t Q = o->Q;
    double e_site = 0;
    for (uint32_t i = 0; i < o->N; ++i) {
        const uint64_t r0 = o->row_ptr[i];
        const uint32_t d = o->deg[i];
        double numerator = 0., denominator = 0.;
        for (uint32_t q = 0; q < Q; ++q) {
            double a = 0.0;
            for (uint32_t l = 0; l < d; ++l) {
                double b = 0;
                if (o->dc == 0)
                    for (uint32_t t = 0; t < Q; ++t) b += o->cab[t * Q + q] * o->msg[(r0 + l) * Q + t];
                a += log(b);
            }
            double a2 = 0.;
            for (uint32_t l = 0; l < d; ++l) {
                double b2 = 0.;
                for (uint32_t t = 0; t < Q; ++t)
                    b2 += log(o->cab[t * Q + q]) * o->cab[t * Q + q] * o->msg[(r0 + l) * Q + t];
                double sum_logs = 0;
                for (uint32_t l2 = 0; l2 < d; ++l2) {
                    double ex = 0.;
                    for (uint32_t t = 0; t < Q; ++t)
                        if (l2 != l) ex += o->cab[t * Q + q] * o->msg[(r0 + l2) * Q + t];
                    sum_logs += log(ex); /* l2 == l contributes log(0) = -inf, as in the reference (:538-545) */
                }
                a2 += b2 * exp(sum_logs);
            }
            if (o->dc == 0) {
                denominator += exp(a + o->logeta[q] - o->h[q] / o->N);
                numerator += exp(a + o->logeta[q] - o->h[q] / o->N) * (-o->h[q] / o->N);
                numerator += a2 * exp(o->logeta[q]) / exp(o->h[q] / o->N);
            }
        }
        e_site += numerator / denominator;
    }
    return e_site / o->N;
}

/* belief_propagation.cpp:614-672 */
double orc_entropy_edge(orc_t *o) {
    const uint32_t Q = o->Q;
    double s_link = 0;
    for (uint32_t i = 0; i < o->N; ++i) {
        double di = o->deg[i];
        for (uint64_t e = o->row_ptr[i]; e < o->row_ptr[i + 1]; ++e) {
            uint32_t i2 = o->col[e];
            const double *min = o->msg + e * Q, *mout = o->msg + (o->row_ptr[i2] + o->inv[e]) * Q;
            double dl = o->deg[i2];
            double numerator = 0., denominator = 0.;
            for (uint32_t q1 = 0; q1 < Q; ++q1) {
                for (uint32_t q2 = q1; q2 < Q; ++q2) {
                    double pair = (q1 == q2) ? (min[q1] * mout[q2]) : (min[q1] * mout[q2] + min[q2] * mout[q1]);
                    double c = o->cab[q1 * Q + q2];
                    if (o->dc == 0) {
                        denominator += c * pair;
                        numerator += c * log(c) * pair;
                    } else if (o->dc == 1) {
                        denominator += di * dl * c * pair;
                        numerator += di * dl * c * log(c) * pair;
                    } else {
                        double tmp = di * dl * o->pab[q1 * Q + q2];
                        denominator += tmp / (1.0 + tmp) * pair;
                        numerator += tmp / (1.0 + tmp) * log(c) * pair;
                    }
                }
            }
            s_link += numerator / denominator;
        }
    }
    return s_link / (2. * o->N);
}

/* belief_propagation.cpp:711-741 */
double orc_entropy_non_edge(orc_t *o) {
    const uint32_t Q = o->Q;
    double acc = 0;
    for (uint32_t i = 0; i < o->N; ++i) {
        for (uint32_t l = 0; l < o->N; ++l) {
            int is_nb = 0;
            for (uint64_t e = o->row_ptr[i]; e < o->row_ptr[i + 1]; ++e)
                if (o->col[e] == l) { is_nb = 1; break; }
            if (is_nb) continue;
            double numerator = 0., denominator = 0.;
            if (o->dc == 0)
                for (uint32_t q1 = 0; q1 < Q; ++q1)
                    for (uint32_t q2 = 0; q2 < Q; ++q2) {
                        double c = o->cab[q1 * Q + q2];
                        double pp = o->marg[(size_t)i * Q + q1] * o->marg[(size_t)l * Q + q2];
                        denominator += (1 - c / o->N) * o->marg[(size_t)i * Q + q1] * o->marg[(size_t)l * Q + q2];
                        numerator += (c / o->N) * log(c) * o->marg[(size_t)i * Q + q1] * o->marg[(size_t)l * Q + q2];
                        (void)pp;
                    }
            if (numerator * denominator != 0) acc += numerator / denominator;
        }
    }
    return acc / (2. * o->N);
}

/* belief_propagation.cpp:752-758 */
double orc_entropy(orc_t *o) {
    double e_site = -orc_entropy_site(o);
    double e_link = +orc_entropy_edge(o);
    double e_non_edge = -orc_entropy_non_edge(o);
    return e_site + e_link + e_non_edge;
}

/* ------------------------------------------------------------------ overlap, EM */

/* belief_propagation.cpp:428-440 */
static void na_expect(orc_t *o) {
    const uint32_t Q = o->Q;
    for (uint32_t q = 0; q < Q; ++q) o->na_expect[q] = o->nna_expect[q] = 0.0;
    for (uint32_t i = 0; i < o->N; ++i)
        for (uint32_t q = 0; q < Q; ++q) {
            o->na_expect[q] += o->marg[(size_t)i * Q + q];
            o->nna_expect[q] += o->deg[i] * o->marg[(size_t)i * Q + q];
        }
}

static int next_perm(uint32_t *p, uint32_t n) { /* std::next_permutation */
    if (n < 2) return 0;
    int i = (int)n - 2;
    while (i >= 0 && p[i] >= p[i + 1]) --i;
    if (i < 0) return 0;
    int j = (int)n - 1;
    while (p[j] <= p[i]) --j;
    uint32_t t = p[i]; p[i] = p[j]; p[j] = t;
    for (int a = i + 1, b = (int)n - 1; a < b; ++a, --b) { t = p[a]; p[a] = p[b]; p[b] = t; }
    return 1;
}

void orc_set_true_conf(orc_t *o, const uint32_t *conf) { memcpy(o->conf_true, conf, sizeof(uint32_t) * o->N); }

/* belief_propagation.cpp:775-811: all Q! relabellings for Q <= 8, identity only above; not normalised */
double orc_overlap(orc_t *o) {
    const uint32_t Q = o->Q;
    uint32_t *perm = (uint32_t *)malloc(sizeof(uint32_t) * Q);
    for (uint32_t q = 0; q < Q; ++q) perm[q] = q;
    na_expect(o); /* side effect kept (:794) */
    double max_ov = -1.0;
    do {
        double ov = 0.0;
        for (uint32_t i = 0; i < o->N; ++i) ov += o->marg[(size_t)i * Q + perm[o->conf_true[i]]];
        ov /= o->N;
        if (ov > max_ov) max_ov = ov;
    } while (Q <= 8 && next_perm(perm, Q));
    free(perm);
    return max_ov;
}

/* belief_propagation.cpp:892-989 */
static void cab_expect(orc_t *o) {
    const uint32_t Q = o->Q;
    double *ce = o->cab_expect;
    for (uint32_t k = 0; k < Q * Q; ++k) ce[k] = 0.;
    for (uint32_t i = 0; i < o->N; ++i) {
        double di = o->deg[i];
        for (uint64_t e = o->row_ptr[i]; e < o->row_ptr[i + 1]; ++e) {
            uint32_t i2 = o->col[e];
            const double *min = o->msg + e * Q, *mout = o->msg + (o->row_ptr[i2] + o->inv[e]) * Q;
            double dl = o->deg[i2];
            double norm_L = edge_norm(o, 0, di, dl, min, mout);
            for (uint32_t q1 = 0; q1 < Q; ++q1) {
                for (uint32_t q2 = q1; q2 < Q; ++q2) {
                    double pair = (q1 == q2) ? (min[q1] * mout[q2]) : (min[q1] * mout[q2] + min[q2] * mout[q1]);
                    if (o->dc == 0) {
                        ce[q1 * Q + q2] += 0.5 * o->cab[q1 * Q + q2] * pair / norm_L;
                    } else if (o->dc == 1) {
                        ce[q1 * Q + q2] += 0.5 * di * dl * o->cab[q1 * Q + q2] * pair / norm_L;
                    } else {
                        double tmp = di * dl * o->pab[q1 * Q + q2];
                        ce[q1 * Q + q2] += 0.5 * tmp / (1.0 + tmp) * pair / norm_L;
                    }
                    if (q1 != q2) ce[q2 * Q + q1] = ce[q1 * Q + q2];
                }
            }
        }
    }
    for (uint32_t q1 = 0; q1 < Q; ++q1) {
        for (uint32_t q2 = q1; q2 < Q; ++q2) {
            if ((o->na_expect[q1] > EPS) && (o->na_expect[q2] > EPS)) {
                const double *w = (o->dc == 0) ? o->na_expect : o->nna_expect;
                if (q1 != q2) {
                    ce[q1 * Q + q2] *= o->N / (w[q1] * w[q2]);
                    ce[q2 * Q + q1] = ce[q1 * Q + q2];
                } else {
                    ce[q1 * Q + q2] *= 2. * o->N / (w[q1] * w[q2]);
                }
            }
        }
    }
}

void orc_em_stats(orc_t *o, double *na_exp, double *nna_exp, double *cab_exp) {
    na_expect(o);
    cab_expect(o);
    if (na_exp) memcpy(na_exp, o->na_expect, sizeof(double) * o->Q);
    if (nna_exp) memcpy(nna_exp, o->nna_expect, sizeof(double) * o->Q);
    if (cab_exp) memcpy(cab_exp, o->cab_expect, sizeof(double) * o->Q * o->Q);
}

/* belief_propagation.cpp:53-75: n_a is an unsigned int truncated every step */
void orc_learning_step(orc_t *o, float learning_rate) {
    const uint32_t Q = o->Q;
    uint32_t rest = o->N;
    for (uint32_t i = 0; i + 1 < Q; ++i) {
        o->na[i] = (uint32_t)(int)(learning_rate * o->na_expect[i] + (1.0 - learning_rate) * o->na[i]);
        rest -= o->na[i];
    }
    o->na[Q - 1] = rest;
    for (uint32_t i = 0; i < Q; ++i) {
        o->eta[i] = (double)o->na[i] / o->N;
        o->logeta[i] = log(o->eta[i]);
        for (uint32_t j = 0; j < Q; ++j) {
            o->cab[i * Q + j] = learning_rate * o->cab_expect[i * Q + j] + (1.0 - learning_rate) * o->cab[i * Q + j];
            o->logcab[i * Q + j] = log(o->cab[i * Q + j]);
            o->pab[i * Q + j] = o->cab[i * Q + j] / o->N;
        }
    }
}

/* belief_propagation.cpp:14-51; returns the number of EM iterations entered */
int orc_learning(orc_t *o, float learning_conv_crit, uint32_t learning_max_time, float learning_rate,
                 float dumping_rate, int sync) {
    double fold = 0.0, fdiff = 1.0;
    int learning_time;
    for (learning_time = 0; learning_time < (int)learning_max_time; learning_time++) {
        if (fdiff < learning_conv_crit) learning_conv_crit *= 0.1;
        if (sync) orc_sync_converge(o, learning_conv_crit, learning_max_time, dumping_rate);
        else orc_converge(o, learning_conv_crit, learning_max_time, dumping_rate);
        na_expect(o);
        cab_expect(o);
        double fnew = orc_free_energy(o);
        fdiff = fabs(fnew - fold);
        fold = fnew;
        if (isnan(fold) || isinf(fold)) break;
        if (fdiff < learning_conv_crit) break;
        orc_learning_step(o, learning_rate);
    }
    return learning_time;
}
