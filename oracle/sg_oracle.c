/*
 * sg_oracle.c -- CPU restatement of the reference's Monte Carlo sweep path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * as the timed CPU baseline.  The product path (spin_glass_anneal_rl_b200)
 * never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file
 * against traces recorded from the reference itself (device='cpu', dense
 * couplings) by tests/golden/make_golden.py and, for UpdateRule.WOLFF, by
 * tests/golden/make_wolff_golden.py: same seed => same energy history, best
 * energy, best configuration, final spins and RNG consumption.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/spin_glass_rl/).
 *
 * RNG model.  The reference draws from torch's global CPU generator
 * (mt19937).  Measured in this container (torch 2.11): every
 * torch.randint(0, n, (1,)) consumes exactly one raw 32-bit output x and
 * returns x % n; every torch.rand(1) consumes one raw output and returns
 * (x & 0xFFFFFF) * 2^-24.  The oracle therefore takes the RAW 32-bit stream
 * (produced by the test with numpy's MT19937 seeded the legacy way, which is
 * bit-identical to torch.manual_seed) and a cursor, and consumes it exactly
 * where the reference would.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SGO_RULE_METROPOLIS 0
#define SGO_RULE_GLAUBER 1
#define SGO_RULE_HEAT_BATH 2
/* UpdateRule.WOLFF has its own entry point (sgo_wolff_sweeps) */

typedef struct {
    const uint32_t *raw; /* raw mt19937 outputs                           */
    int64_t len;         /* number of outputs available                   */
    int64_t pos;         /* cursor: next output to consume                */
    int overrun;         /* set if the stream ran dry                     */
} sgo_stream;

static inline uint32_t sgo_next(sgo_stream *s) {
    if (s->pos >= s->len) {
        s->overrun = 1;
        return 0u;
    }
    return s->raw[s->pos++];
}

/* torch.randint(0, n, (1,)).item()  -- core/spin_dynamics.py:69 */
static inline int sgo_randint(sgo_stream *s, int n) { return (int)(sgo_next(s) % (uint32_t)n); }

/* torch.rand(1).item() (float32)    -- core/spin_dynamics.py:146,162,181 */
static inline float sgo_rand(sgo_stream *s) {
    return (float)(sgo_next(s) & 0xFFFFFFu) * (1.0f / 16777216.0f);
}

/* fp32 dot as torch.dot does it (fp32 accumulate; summation order is
 * ATen's own, so float couplings agree to rounding, integer couplings
 * exactly). */
static inline float sgo_dotf(const float *a, const float *b, int n) {
    float acc = 0.0f;
#pragma omp simd reduction(+ : acc)
    for (int j = 0; j < n; ++j) acc += a[j] * b[j];
    return acc;
}

/* IsingModel.get_local_field(i) -- core/ising_model.py:176-185.
 * coupling_field = torch.dot(J[i], spins).item()  (fp32, diagonal included)
 * return coupling_field + h[i].item()              (python double add)      */
double sgo_local_field(const float *J, int64_t ld, const float *h, const float *spins, int n,
                       int i) {
    float cf = sgo_dotf(J + (int64_t)i * ld, spins, n);
    return (double)cf + (double)h[i];
}

/* IsingModel.compute_energy() -- core/ising_model.py:149-174.
 * interaction = -0.5 * torch.dot(s, torch.mv(J, s)).item()
 * field       = -torch.dot(h, s).item()                                     */
double sgo_energy(const float *J, int64_t ld, const float *h, const float *spins, int n,
                  float *scratch) {
    for (int i = 0; i < n; ++i) scratch[i] = sgo_dotf(J + (int64_t)i * ld, spins, n);
    double inter = -0.5 * (double)sgo_dotf(spins, scratch, n);
    double field = -(double)sgo_dotf(h, spins, n);
    return inter + field;
}

/* One SpinDynamics.single_spin_update(site) for the three local rules.
 * Returns 1 if the spin changed ("accepted"), 0 otherwise ("rejected").
 *   _metropolis_update  core/spin_dynamics.py:131-152
 *   _glauber_update     core/spin_dynamics.py:154-170
 *   _heat_bath_update   core/spin_dynamics.py:172-191
 * flip_spin (core/ising_model.py:125-147) recomputes the same dot and
 * negates the spin; the recomputation has no observable effect, so only the
 * baseline timing variant (sgo_sweep with redo_flip_dot=1) performs it.     */
static inline int sgo_attempt(const float *J, int64_t ld, const float *h, float *spins, int n,
                              int site, double T, int rule, sgo_stream *st, int redo_flip_dot,
                              float *u_out, int *drew) {
    double lf = sgo_local_field(J, ld, h, spins, n, site);
    *drew = 0;
    if (rule == SGO_RULE_METROPOLIS) {
        double dE = 2.0 * (double)spins[site] * lf; /* :135 */
        int accept;
        if (dE <= 0.0) { /* :138 */
            accept = 1;
        } else {
            /* :145  torch.exp(torch.tensor(-dE / T)): the double quotient is
             * rounded to float32, exp is evaluated in float32.              */
            float p = expf((float)(-dE / T));
            float u = sgo_rand(st); /* :146, drawn ONLY on this branch */
            *u_out = u;
            *drew = 1;
            accept = ((double)u < (double)p);
        }
        if (accept) {
            if (redo_flip_dot) {
                volatile double sink = sgo_local_field(J, ld, h, spins, n, site);
                (void)sink;
            }
            spins[site] = -spins[site];
            return 1;
        }
        return 0;
    }
    /* Glauber / heat bath: u is ALWAYS drawn (:162, :181). */
    float arg;
    if (rule == SGO_RULE_GLAUBER) {
        arg = (float)(-2.0 * lf / T); /* :159 */
    } else {
        double beta = 1.0 / T; /* :177 */
        arg = (float)(-2.0 * beta * lf); /* :178 */
    }
    float prob_up = 1.0f / (1.0f + expf(arg)); /* float32 tensor arithmetic */
    float u = sgo_rand(st);
    *u_out = u;
    *drew = 1;
    float new_spin = (u < prob_up) ? 1.0f : -1.0f;
    if (new_spin != spins[site]) {
        if (redo_flip_dot && rule == SGO_RULE_GLAUBER) {
            volatile double sink = sgo_local_field(J, ld, h, spins, n, site);
            (void)sink;
        }
        spins[site] = new_spin;
        return 1;
    }
    return 0;
}

/*
 * SpinDynamics.sweep() x n_sweeps for ONE replica -- core/spin_dynamics.py:73-94:
 * n attempts at sites drawn with replacement, then compute_energy().
 *
 *   temps[n_sweeps]        temperature used for each sweep (already clamped
 *                          by set_temperature, core/spin_dynamics.py:57-59)
 *   energies[n_sweeps]     out: energy after each sweep
 *   accepted[n_sweeps]     out: accepted attempts in each sweep
 *   trace_site/trace_u     optional out, n_sweeps*n entries: the site of every
 *                          attempt and the uniform it consumed (NaN if none)
 * Returns 0, or -1 if the raw stream ran dry.
 */
int sgo_sweeps(const float *J, int64_t ld, const float *h, float *spins, int n, int rule,
               const double *temps, int n_sweeps, const uint32_t *raw, int64_t raw_len,
               int64_t *raw_pos, double *energies, int64_t *accepted, int32_t *trace_site,
               float *trace_u, int redo_flip_dot) {
    sgo_stream st = {raw, raw_len, *raw_pos, 0};
    float *scratch = (float *)malloc(sizeof(float) * (size_t)n);
    for (int sw = 0; sw < n_sweeps; ++sw) {
        double T = temps[sw];
        int64_t acc = 0;
        for (int k = 0; k < n; ++k) {
            int site = sgo_randint(&st, n);
            float u = NAN;
            int drew = 0;
            acc += sgo_attempt(J, ld, h, spins, n, site, T, rule, &st, redo_flip_dot, &u, &drew);
            if (trace_site) trace_site[(int64_t)sw * n + k] = site;
            if (trace_u) trace_u[(int64_t)sw * n + k] = drew ? u : NAN;
        }
        if (energies) energies[sw] = sgo_energy(J, ld, h, spins, n, scratch);
        if (accepted) accepted[sw] = acc;
    }
    free(scratch);
    *raw_pos = st.pos;
    return st.overrun ? -1 : 0;
}

/*
 * The same sweep driven by an explicit per-attempt schedule instead of the raw
 * stream: sites[k] is visited at attempt k and uniforms[k] is the uniform that
 * attempt may use.  This is the form the CUDA sweep kernel consumes in its
 * injected mode (sites are shared by the replicas of a block; one uniform per
 * attempt per replica), so oracle and kernel can be compared attempt for
 * attempt.  Metropolis ignores uniforms[k] when dE <= 0, exactly like the
 * reference (core/spin_dynamics.py:138-142).
 */
int sgo_sweeps_scheduled(const float *J, int64_t ld, const float *h, float *spins, int n, int rule,
                         const double *temps, int n_sweeps, const int32_t *sites,
                         const float *uniforms, double *energies, int64_t *accepted) {
    float *scratch = (float *)malloc(sizeof(float) * (size_t)n);
    for (int sw = 0; sw < n_sweeps; ++sw) {
        double T = temps[sw];
        int64_t acc = 0;
        for (int k = 0; k < n; ++k) {
            int64_t idx = (int64_t)sw * n + k;
            int site = sites[idx];
            float u = uniforms[idx];
            double lf = sgo_local_field(J, ld, h, spins, n, site);
            int flip = 0;
            if (rule == SGO_RULE_METROPOLIS) {
                double dE = 2.0 * (double)spins[site] * lf;
                if (dE <= 0.0) {
                    flip = 1;
                } else {
                    float p = expf((float)(-dE / T));
                    flip = ((double)u < (double)p);
                }
            } else {
                float arg = (rule == SGO_RULE_GLAUBER) ? (float)(-2.0 * lf / T)
                                                       : (float)(-2.0 * (1.0 / T) * lf);
                float prob_up = 1.0f / (1.0f + expf(arg));
                float ns = (u < prob_up) ? 1.0f : -1.0f;
                flip = (ns != spins[site]);
            }
            if (flip) {
                spins[site] = -spins[site];
                ++acc;
            }
        }
        if (energies) energies[sw] = sgo_energy(J, ld, h, spins, n, scratch);
        if (accepted) accepted[sw] = acc;
    }
    free(scratch);
    return 0;
}

/*
 * Wolff cluster updates -- core/spin_dynamics.py:193-262 (_wolff_update -> _wolff_cluster_dense;
 * the dense branch is the one a dense model takes, :206-209).
 *
 * One update: breadth-first growth from `start` (:218-243).  The queue is FIFO (queue.pop(0)), the
 * neighbours of the dequeued site are visited in index order, a neighbour already in the cluster
 * is skipped WITHOUT a draw, and a uniform is drawn only for a neighbour with coupling < 0 whose
 * spin equals the dequeued site's (:234-239):
 *     prob_add = 1.0 - torch.exp(torch.tensor(2.0 * coupling / T))      float32 tensor arithmetic
 *     torch.rand(1).item() < prob_add
 * coupling = couplings[current, neighbour] (row of the dequeued site; J need not be symmetric).
 * The spins tested are the clone taken before the cluster is flipped (:216).  Then every cluster
 * site is flipped (:246-247) and n_accepted grows by the cluster size (:254); the update never
 * counts as rejected.  External fields play no role in the move.  The two compute_energy() calls
 * around it (:213, :250) only feed the returned delta, which sweep() ignores.
 *
 * The uniforms come either from the raw mt19937 stream (src->st) or, for replays, from an explicit
 * list consumed in order (src->u_in); every uniform consumed can be logged (src->u_out).
 */
typedef struct {
    sgo_stream *st;
    const float *u_in;
    int64_t u_len, u_pos;
    float *u_out;
    int64_t u_cap, u_cnt;
    int overrun;
} sgo_usource;

static inline float sgo_usource_next(sgo_usource *s) {
    float u;
    if (s->st) {
        u = sgo_rand(s->st);
    } else if (s->u_pos < s->u_len) {
        u = s->u_in[s->u_pos++];
    } else {
        s->overrun = 1;
        u = 1.0f;
    }
    if (s->u_out) {
        if (s->u_cnt < s->u_cap) s->u_out[s->u_cnt] = u;
        else s->overrun = 1;
    }
    s->u_cnt++;
    return u;
}

static int sgo_wolff_cluster(const float *J, int64_t ld, float *spins, int n, int start, double T,
                             sgo_usource *src, int32_t *queue, uint8_t *in_cluster) {
    memset(in_cluster, 0, (size_t)n);
    int head = 0, tail = 0;
    queue[tail++] = start;
    in_cluster[start] = 1;
    while (head < tail) {
        const int cur = queue[head++];
        const float cs = spins[cur];
        const float *row = J + (int64_t)cur * ld;
        for (int nb = 0; nb < n; ++nb) {
            if (nb == cur || in_cluster[nb]) continue;
            const float c = row[nb];
            if (c < 0.0f && cs == spins[nb]) {
                const float p = 1.0f - expf((float)(2.0 * (double)c / T));
                const float u = sgo_usource_next(src);
                if (u < p) {
                    in_cluster[nb] = 1;
                    queue[tail++] = nb;
                }
            }
        }
    }
    for (int k = 0; k < tail; ++k) spins[queue[k]] = -spins[queue[k]];
    return tail;
}

/*
 * SpinDynamics.sweep() x n_sweeps with UpdateRule.WOLFF for ONE replica (core/spin_dynamics.py:73-94:
 * n updates per sweep, each from a start site drawn with torch.randint, then compute_energy()).
 *   raw != NULL : start sites and uniforms from the raw mt19937 stream (the reference's own order:
 *                 one randint, then the cluster's uniforms); trace_site / trace_u / draws_per_update
 *                 optionally record them for a replay
 *   raw == NULL : start sites from sites[n_sweeps * n], uniforms from uniforms[0 .. n_uniforms) in
 *                 consumption order (what the CUDA kernel takes in injected mode)
 * accepted[s] = sum of the cluster sizes of sweep s; *uniforms_used = uniforms consumed.
 * Returns 0, or -1 if a stream ran dry / a trace buffer was too small.
 */
int sgo_wolff_sweeps(const float *J, int64_t ld, const float *h, float *spins, int n,
                     const double *temps, int n_sweeps, const uint32_t *raw, int64_t raw_len,
                     int64_t *raw_pos, const int32_t *sites, const float *uniforms,
                     int64_t n_uniforms, double *energies, int64_t *accepted, int32_t *trace_site,
                     float *trace_u, int64_t trace_u_cap, int32_t *draws_per_update,
                     int64_t *uniforms_used) {
    sgo_stream st = {raw, raw_len, raw_pos ? *raw_pos : 0, 0};
    sgo_usource src = {raw ? &st : NULL, uniforms, n_uniforms, 0, trace_u, trace_u_cap, 0, 0};
    float *scratch = (float *)malloc(sizeof(float) * (size_t)n);
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    uint8_t *in_cluster = (uint8_t *)malloc((size_t)n);
    for (int sw = 0; sw < n_sweeps; ++sw) {
        const double T = temps[sw];
        int64_t acc = 0;
        for (int k = 0; k < n; ++k) {
            const int64_t idx = (int64_t)sw * n + k;
            const int start = raw ? sgo_randint(&st, n) : sites[idx];
            const int64_t before = src.u_cnt;
            acc += sgo_wolff_cluster(J, ld, spins, n, start, T, &src, queue, in_cluster);
            if (trace_site) trace_site[idx] = start;
            if (draws_per_update) draws_per_update[idx] = (int32_t)(src.u_cnt - before);
        }
        if (energies) energies[sw] = sgo_energy(J, ld, h, spins, n, scratch);
        if (accepted) accepted[sw] = acc;
    }
    free(scratch);
    free(queue);
    free(in_cluster);
    if (raw_pos) *raw_pos = st.pos;
    if (uniforms_used) *uniforms_used = src.u_cnt;
    return (st.overrun || src.overrun) ? -1 : 0;
}

/* Batched energies / local fields, one row per configuration:
 *   BatchProcessor.process_batch_energies      optimization/high_performance_computing.py:98-165
 *   VectorizedOperations.vectorized_local_fields   same file :357-372
 * fields[b][j] = sum_i S[b][i] J[i][j] + h[j];  E[b] = -0.5 sum_j fields'[b][j] S[b][j] - sum_j h[j] S[b][j]
 * (fields' = fields without h).  Accumulated in double: this is the checker. */
void sgo_batch_fields_energies(const float *J, int64_t ld, const float *h, const float *S, int n,
                               int batch, double *fields, double *energies) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < batch; ++b) {
        const float *s = S + (int64_t)b * n;
        double e_int = 0.0, e_h = 0.0;
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int i = 0; i < n; ++i) acc += (double)s[i] * (double)J[(int64_t)i * ld + j];
            if (fields) fields[(int64_t)b * n + j] = acc + (double)h[j];
            e_int += acc * (double)s[j];
            e_h += (double)h[j] * (double)s[j];
        }
        if (energies) energies[b] = -0.5 * e_int - e_h;
    }
}

/*
 * CPU baseline: the reference path for many independent replicas, one replica
 * per OpenMP thread at a time (the reference itself has no replica batching:
 * R replicas cost R x).  Same arithmetic as sgo_sweeps with the redundant
 * flip_spin dot included, sites/uniforms from a per-replica xorshift stream
 * (timing only; not used for parity).  Returns attempts performed.
 */
int64_t sgo_baseline_run(const float *J, int64_t ld, const float *h, float *spins, int n,
                         int n_replicas, int n_sweeps, double T, uint64_t seed, int n_threads,
                         double *energies_out) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int r = 0; r < n_replicas; ++r) {
        float *s = spins + (int64_t)r * n;
        float *scratch = (float *)malloc(sizeof(float) * (size_t)n);
        uint64_t x = seed * 0x9E3779B97F4A7C15ull + (uint64_t)(r + 1) * 0xBF58476D1CE4E5B9ull;
        uint32_t buf[256];
        for (int sw = 0; sw < n_sweeps; ++sw) {
            for (int k = 0; k < n; ++k) {
                /* two raw words per attempt are enough (site, maybe uniform) */
                for (int q = 0; q < 2; ++q) {
                    x ^= x << 13;
                    x ^= x >> 7;
                    x ^= x << 17;
                    buf[q] = (uint32_t)(x >> 16);
                }
                sgo_stream st = {buf, 2, 0, 0};
                int site = sgo_randint(&st, n);
                float u;
                int drew;
                (void)sgo_attempt(J, ld, h, s, n, site, T, SGO_RULE_METROPOLIS, &st, 1, &u, &drew);
            }
            double e = sgo_energy(J, ld, h, s, n, scratch);
            if (energies_out) energies_out[r] = e;
            total += n;
        }
        free(scratch);
    }
    return total;
}

/* The same baseline for models the reference's callers hold as sparse COO and that cannot be
 * densified at full size (cfg2: 65 536 spins, cfg5: 50 000 spins -- SURVEY 8c: "cfg2/cfg5 are timed
 * with the restatement and labelled as such"): the reference's algorithm with every dense dot
 * product replaced by the sum over the row's non-zeros, which is what its sparse branch computes
 * (core/ising_model.py:160-166, 180-183) without materialising the dense matrix per call.  CSR
 * rows of J.  Timing arm only; parity is always checked on dense (down-scaled) instances. */
int64_t sgo_baseline_run_csr(const int64_t *rowptr, const int32_t *colidx, const float *val, const float *h,
                             float *spins, int n, int n_replicas, int n_sweeps, double T, uint64_t seed,
                             int n_threads, double *energies_out) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int r = 0; r < n_replicas; ++r) {
        float *s = spins + (int64_t)r * n;
        uint64_t x = seed * 0x9E3779B97F4A7C15ull + (uint64_t)(r + 1) * 0xBF58476D1CE4E5B9ull;
        for (int sw = 0; sw < n_sweeps; ++sw) {
            for (int k = 0; k < n; ++k) {
                uint32_t w[2];
                for (int q = 0; q < 2; ++q) {
                    x ^= x << 13;
                    x ^= x >> 7;
                    x ^= x << 17;
                    w[q] = (uint32_t)(x >> 16);
                }
                const int site = (int)(w[0] % (uint32_t)n);
                float cf = 0.0f;
                for (int64_t e = rowptr[site]; e < rowptr[site + 1]; ++e) cf += val[e] * s[colidx[e]];
                const double lf = (double)cf + (double)h[site];
                const double dE = 2.0 * (double)s[site] * lf;
                int accept = dE <= 0.0;
                if (!accept) {
                    const float p = expf((float)(-dE / T));
                    const float u = (float)(w[1] & 0xFFFFFFu) * (1.0f / 16777216.0f);
                    accept = ((double)u < (double)p);
                }
                if (accept) s[site] = -s[site];
            }
            /* compute_energy() after the sweep */
            double inter = 0.0, field = 0.0;
            for (int i = 0; i < n; ++i) {
                float cf = 0.0f;
                for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) cf += val[e] * s[colidx[e]];
                inter += (double)s[i] * (double)cf;
                field += (double)h[i] * (double)s[i];
            }
            if (energies_out) energies_out[r] = -0.5 * inter - field;
            total += n;
        }
    }
    return total;
}

int sgo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
