/*
 * CPU oracle for the UMNN Clenshaw-Curtis integration hot path -- plain C restatement.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (cross-check against the numpy oracle and the
 * golden vectors) and by bench.py's cpu_baseline / --impl reference legs as the CPU arm that is
 * timed beside the GPU path.  Nothing in the product (umnn_b200/, models/) links or loads it.
 *
 * Pinned against the reference through tests/golden/ (tests/test_oracle_c.py).
 *
 * What it restates (AWehenkel/UMNN @ 59118c14):
 *   - integrate(...) forward branch, models/UMNN/ParallelNeuralIntegral.py:37-65
 *       xT = x0 + Q*((x-x0)/Q); X_i = x0 + ((xT-x0)*(t_i+1))/2; z = sum_i w_i f(X_i,h);
 *       result = (z*(xT-x0))/2
 *   - IntegrandNetwork.forward slot layout, models/UMNN/UMNNMAF.py:263-284
 *       slot (n,d) input = [x[n,d], h[n,0*D+d], ..., h[n,(E-1)*D+d]]       (layout 0, "strided")
 *   - IntegrandNN.forward, models/UMNN/MonotonicNN.py:26-27                 (layout 1, "contig")
 *   - the nn.Sequential MLP: Linear, LeakyReLU(0.01) | ReLU, ..., Linear, ELU(.)+1 | Sigmoid
 *       UMNNMAF.py:11-19,245-254; MonotonicNN.py:15-24
 *
 * Threading: POSIX threads over contiguous ranges of slots (every (sample, dimension) slot is
 * independent; libgomp is not in the image, so no OpenMP).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORC_SIMD 1
#endif

#define ORC_MAX_LAYERS 16

typedef struct {
    int n_layers;                 /* number of Linear layers (hidden + output) */
    int widths[ORC_MAX_LAYERS + 1]; /* n_in, H1, ..., HL, 1 */
    int hidden_act;               /* 0 = ReLU, 1 = LeakyReLU(0.01) */
    int out_act;                  /* 0 = ELU+1, 1 = Sigmoid */
} orc_mlp;

static inline float orc_hidden(float v, int kind) {
    if (kind == 1) return v > 0.f ? v : v * 0.01f;
    return v > 0.f ? v : 0.f;
}

static inline float orc_out(float v, int kind) {
    if (kind == 0) return (v > 0.f ? v : expm1f(v)) + 1.0f;
    return 1.0f / (1.0f + expf(-v));
}

/* One dense layer on a block of rows: out[r][j] = act(b[j] + sum_k in[r][k] * W[j][k]).
 * The weights arrive packed in panels of ORC_CB output columns, Wp[panel][k][ORC_CB] (zero padded),
 * so the k loop streams one L1-resident panel; the register block is ORC_RB rows x ORC_CB outputs.
 * Per output the k-sum runs in increasing k (a fixed, documented order). */
#define ORC_RB 6
#define ORC_CB 16
static void orc_layer(const float *in, int R, int nin, const float *Wp, const float *bias, int nout,
                      float *out, int last, int hidden_act, int out_act) {
    for (int j0 = 0; j0 < nout; j0 += ORC_CB) {
        const int cb = (nout - j0 < ORC_CB) ? (nout - j0) : ORC_CB;
        const float *panel = Wp + (size_t)(j0 / ORC_CB) * nin * ORC_CB;
        for (int r0 = 0; r0 < R; r0 += ORC_RB) {
            const int rb = (R - r0 < ORC_RB) ? (R - r0) : ORC_RB;
            float acc[ORC_RB][ORC_CB];
#ifdef ORC_SIMD
            if (rb == ORC_RB) {
                __m256 c[ORC_RB][2];
                for (int r = 0; r < ORC_RB; ++r) c[r][0] = c[r][1] = _mm256_setzero_ps();
                const float *ip = in + (size_t)r0 * nin;
                for (int k = 0; k < nin; ++k) {
                    const __m256 w0 = _mm256_loadu_ps(panel + (size_t)k * ORC_CB);
                    const __m256 w1 = _mm256_loadu_ps(panel + (size_t)k * ORC_CB + 8);
#pragma GCC unroll 6
                    for (int r = 0; r < ORC_RB; ++r) {
                        const __m256 av = _mm256_broadcast_ss(ip + (size_t)r * nin + k);
                        c[r][0] = _mm256_fmadd_ps(av, w0, c[r][0]);
                        c[r][1] = _mm256_fmadd_ps(av, w1, c[r][1]);
                    }
                }
                for (int r = 0; r < ORC_RB; ++r) {
                    _mm256_storeu_ps(&acc[r][0], c[r][0]);
                    _mm256_storeu_ps(&acc[r][8], c[r][1]);
                }
            } else
#endif
            {
                for (int r = 0; r < ORC_RB; ++r)
                    for (int j = 0; j < ORC_CB; ++j) acc[r][j] = 0.f;
                for (int k = 0; k < nin; ++k) {
                    const float *w = panel + (size_t)k * ORC_CB;
                    for (int r = 0; r < rb; ++r) {
                        const float ak = in[(size_t)(r0 + r) * nin + k];
                        for (int j = 0; j < ORC_CB; ++j) acc[r][j] += ak * w[j];
                    }
                }
            }
            for (int r = 0; r < rb; ++r)
                for (int j = 0; j < cb; ++j) {
                    const float v = acc[r][j] + bias[j0 + j];
                    out[(size_t)(r0 + r) * nout + j0 + j] = last ? orc_out(v, out_act) : orc_hidden(v, hidden_act);
                }
        }
    }
}

/* rows[R][n_in] -> f[R]; scratch a/b hold R*maxw floats each; Wt[l] is W_l packed in column panels */
static void orc_mlp_rows(const orc_mlp *m, float *const *Wt, float *const *bias, int R,
                         const float *rows, float *a, float *b, float *f) {
    const float *cur = rows;
    float *nxt = a;
    for (int l = 0; l < m->n_layers; ++l) {
        orc_layer(cur, R, m->widths[l], Wt[l], bias[l], m->widths[l + 1], nxt, l == m->n_layers - 1,
                  m->hidden_act, m->out_act);
        cur = nxt;
        nxt = (nxt == a) ? b : a;
    }
    for (int r = 0; r < R; ++r) f[r] = cur[r];
}

typedef struct {
    const orc_mlp *m;
    float *const *Wt;
    float *const *bias;
    int D, E, layout, Q, maxw;
    const float *t, *w, *x0, *x, *h;
    float *out_integral, *out_fx, *out_fx0;
    long s_begin, s_end;
    int fail;
} orc_job;

static void *orc_worker(void *arg) {
    orc_job *jb = (orc_job *)arg;
    const int Q = jb->Q, E = jb->E, D = jb->D;
    const int R = Q + 3; /* Q+1 nodes, then x, then x0 */
    float *rows = (float *)malloc(sizeof(float) * (size_t)R * (1 + E));
    float *a = (float *)malloc(sizeof(float) * (size_t)R * jb->maxw);
    float *b = (float *)malloc(sizeof(float) * (size_t)R * jb->maxw);
    float *f = (float *)malloc(sizeof(float) * (size_t)R);
    if (!rows || !a || !b || !f) {
        jb->fail = 1;
    } else {
        for (long s = jb->s_begin; s < jb->s_end; ++s) {
            const long n = s / D;
            const int d = (int)(s % D);
            const float lo = jb->x0[s], hi = jb->x[s];
            const float step = (hi - lo) / (float)Q;
            const float xT = lo + (float)Q * step;
            const float span = xT - lo;
            for (int i = 0; i < R; ++i) {
                float xi;
                if (i <= Q) xi = lo + (span * (jb->t[i] + 1.0f)) / 2.0f;
                else xi = (i == Q + 1) ? hi : lo;
                float *row = rows + (size_t)i * (1 + E);
                row[0] = xi;
                if (jb->layout == 0)
                    for (int e = 0; e < E; ++e) row[1 + e] = jb->h[(size_t)n * E * D + (size_t)e * D + d];
                else
                    for (int e = 0; e < E; ++e) row[1 + e] = jb->h[(size_t)n * E + e];
            }
            orc_mlp_rows(jb->m, jb->Wt, jb->bias, R, rows, a, b, f);
            float z = 0.f;
            for (int i = 0; i <= Q; ++i) z += f[i] * jb->w[i];
            jb->out_integral[s] = (z * span) / 2.0f;
            if (jb->out_fx) jb->out_fx[s] = f[Q + 1];
            if (jb->out_fx0) jb->out_fx0[s] = f[Q + 2];
        }
    }
    free(rows); free(a); free(b); free(f);
    return NULL;
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/*
 * Forward integral for every slot.
 *   layout 0: x0,x [B][D], h [B][E*D] (h[b][e*D+d]);  layout 1: D must be 1, h [B][E].
 *   nodes t[Q+1], weights w[Q+1] (float32 tables of compute_cc_weights).
 *   out_integral [B][D]; out_fx, out_fx0 [B][D] or NULL (f at x and at x0, exact extra rows).
 *   n_threads <= 0: one thread per online core.
 * Returns 0, or -1 on bad arguments / allocation failure.
 */
int orc_cc_forward(const orc_mlp *m, const float *flat_params, int B, int D, int E, int layout, int Q,
                   const float *t, const float *w, const float *x0, const float *x, const float *h,
                   float *out_integral, float *out_fx, float *out_fx0, int n_threads) {
    if (!m || m->n_layers < 1 || m->n_layers > ORC_MAX_LAYERS || m->widths[0] != 1 + E ||
        m->widths[m->n_layers] != 1 || (layout == 1 && D != 1))
        return -1;
    const int L = m->n_layers;
    float *Wt[ORC_MAX_LAYERS], *bias[ORC_MAX_LAYERS];
    int maxw = 1 + E;
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
        const int nin = m->widths[l], nout = m->widths[l + 1];
        if (nout > maxw) maxw = nout;
        const int npanel = (nout + ORC_CB - 1) / ORC_CB;
        Wt[l] = (float *)calloc((size_t)npanel * nin * ORC_CB, sizeof(float));
        bias[l] = (float *)malloc(sizeof(float) * (size_t)nout);
        if (!Wt[l] || !bias[l]) return -1;
        for (int j = 0; j < nout; ++j)
            for (int k = 0; k < nin; ++k)
                Wt[l][((size_t)(j / ORC_CB) * nin + k) * ORC_CB + (j % ORC_CB)] = flat_params[off + (size_t)j * nin + k];
        off += (size_t)nin * nout;
        memcpy(bias[l], flat_params + off, sizeof(float) * nout);
        off += nout;
    }
    const long S = (long)B * D;
    int nt = n_threads > 0 ? n_threads : orc_max_threads();
    if (nt > 256) nt = 256;
    if ((long)nt > S) nt = S > 0 ? (int)S : 1;
    orc_job jobs[256];
    pthread_t tids[256];
    int fail = 0;
    for (int i = 0; i < nt; ++i) {
        orc_job *jb = &jobs[i];
        jb->m = m; jb->Wt = Wt; jb->bias = bias; jb->D = D; jb->E = E; jb->layout = layout; jb->Q = Q;
        jb->maxw = maxw; jb->t = t; jb->w = w; jb->x0 = x0; jb->x = x; jb->h = h;
        jb->out_integral = out_integral; jb->out_fx = out_fx; jb->out_fx0 = out_fx0;
        jb->s_begin = S * i / nt; jb->s_end = S * (i + 1) / nt; jb->fail = 0;
    }
    for (int i = 1; i < nt; ++i)
        if (pthread_create(&tids[i], NULL, orc_worker, &jobs[i]) != 0) { jobs[i].fail = 2; }
    orc_worker(&jobs[0]);
    for (int i = 1; i < nt; ++i) {
        if (jobs[i].fail == 2) { jobs[i].fail = 0; orc_worker(&jobs[i]); }
        else pthread_join(tids[i], NULL);
    }
    for (int i = 0; i < nt; ++i) fail |= jobs[i].fail;
    for (int l = 0; l < L; ++l) { free(Wt[l]); free(bias[l]); }
    return fail ? -1 : 0;
}
