/*
 * vq_oracle.c -- CPU restatement of MedTok's vector-quantizer arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under medtok_b200/ may link, load or call
 * this file; it exists so tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg can check the CUDA path against an independent statement of
 * the reference algorithm.
 *
 * The reference is pure PyTorch (no C sources to compile), so this file restates
 * the *arithmetic the reference's CPU path performs*, operation order included.
 * Every order below was pinned by running the unmodified reference modules in the
 * authoring container (torch 2.11 / MKL 2024.2, see oracle/make_golden.py and
 * tests/test_oracle_golden.py) and comparing bit patterns:
 *
 *   F.normalize / vector_norm over a contiguous last dim  -> vqo_l2norm_rows
 *   x.pow(2).sum(dim=1)  (ATen cascade_sum, 4 x 8-lane)    -> vqo_sumsq_rows
 *   einsum('bd,nd->bn') == MKL sgemm, K<=512               -> sequential fmaf chain
 *   ema_inplace / norm_ema_inplace (mul_, add_(alpha))     -> fmaf(new, alpha, fl(x*decay))
 *   l2norm over the strided (.t()) per-code mean           -> sequential, no FMA
 *
 * Reference lines (relative to /root/reference/MedTok/):
 *   norm_ema_quantizer.py:8-12      l2norm, ema_inplace
 *   norm_ema_quantizer.py:136-138   norm_ema_inplace
 *   norm_ema_quantizer.py:166-218   NormEMAVectorQuantizer.forward
 *   vector_quantization_soft_one_new.py:120-125,187-217   soft top-k quantizer
 *   norm_ema_quantizer.py:24-57     kmeans
 *
 * Build: see oracle/build_oracle.py  (gcc -O2 -ffp-contract=off -mfma -fopenmp).
 * -ffp-contract=off is REQUIRED: every FMA below is an explicit fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VQO_API __attribute__((visibility("default")))

/* ---- row reductions -------------------------------------------------------- */

/* torch.linalg.vector_norm(x, 2, dim=-1) on a contiguous row: one 8-lane
 * accumulator, acc += fl(x*x) (no FMA), lanes summed 0..7 in order, scalar tail. */
static float norm2_row(const float *x, int d)
{
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int main_ = (d / 8) * 8;
    for (int c = 0; c < main_; c += 8)
        for (int v = 0; v < 8; ++v) {
            float sq = x[c + v] * x[c + v];
            acc[v] = acc[v] + sq;
        }
    float s = acc[0];
    for (int v = 1; v < 8; ++v) s = s + acc[v];
    for (int i = main_; i < d; ++i) {
        float sq = x[i] * x[i];
        s = s + sq;
    }
    return sqrtf(s);
}

/* x.pow(2).sum(dim=1): ATen cascade_sum inner kernel -- 4 accumulators of 8 lanes,
 * leftover vectors into accumulator 0, accumulators folded 1,2,3 into 0, then
 * scalar tail first and the 8 lanes added in order. */
static float sumsq_row(const float *x, int d)
{
    float acc[4][8];
    memset(acc, 0, sizeof(acc));
    int nvec = d / 8, n4 = nvec / 4;
    for (int c = 0; c < n4; ++c)
        for (int u = 0; u < 4; ++u)
            for (int v = 0; v < 8; ++v) {
                float e = x[(c * 4 + u) * 8 + v];
                float sq = e * e;
                acc[u][v] = acc[u][v] + sq;
            }
    for (int i = n4 * 4; i < nvec; ++i)
        for (int v = 0; v < 8; ++v) {
            float e = x[i * 8 + v];
            float sq = e * e;
            acc[0][v] = acc[0][v] + sq;
        }
    for (int u = 1; u < 4; ++u)
        for (int v = 0; v < 8; ++v) acc[0][v] = acc[0][v] + acc[u][v];
    float fin = 0.0f;
    for (int i = nvec * 8; i < d; ++i) {
        float sq = x[i] * x[i];
        fin = fin + sq;
    }
    for (int v = 0; v < 8; ++v) fin = fin + acc[0][v];
    return fin;
}

/* vector_norm when the reduced dim is NOT unit-stride (the 'b c h w -> b h w c' view with
 * h*w > 1, or the .t() view of the per-code mean): plain sequential sum, no FMA. */
static float norm2_row_seq(const float *x, int d)
{
    float s = 0.0f;
    for (int i = 0; i < d; ++i) {
        float sq = x[i] * x[i];
        s = s + sq;
    }
    return sqrtf(s);
}

/* norm_ema_quantizer.py:8-9  l2norm == F.normalize(p=2, dim=-1, eps=1e-12)
 * seq_mode = 0: rows are unit-stride in the reference ([N,D,1,1] inputs, codebooks)
 * seq_mode = 1: the reference normalises a channel-strided view (b c h w with h*w > 1) */
VQO_API void vqo_l2norm_rows(const float *x, int64_t n, int d, float *out, float *denom_out,
                             int seq_mode)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float *r = x + i * d;
        float den = seq_mode ? norm2_row_seq(r, d) : norm2_row(r, d);
        if (den < 1e-12f) den = 1e-12f; /* clamp_min(eps); NaN propagates like torch */
        for (int j = 0; j < d; ++j) out[i * d + j] = r[j] / den;
        if (denom_out) denom_out[i] = den;
    }
}

VQO_API void vqo_sumsq_rows(const float *x, int64_t n, int d, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = sumsq_row(x + i * d, d);
}

/* ---- distance + selection -------------------------------------------------- */

/* transposed copy wt[d][k] so the k loop vectorises while each (n,k) dot keeps the
 * sequential-over-d fmaf chain (lane-wise vfmadd is the same correctly-rounded op) */
static float *transpose_kd(const float *w, int k, int d)
{
    float *wt = (float *)malloc((size_t)k * d * sizeof(float));
    for (int j = 0; j < k; ++j)
        for (int i = 0; i < d; ++i) wt[(size_t)i * k + j] = w[(size_t)j * d + i];
    return wt;
}

/* norm_ema_quantizer.py:175-177 (and soft :120-125):
 *   d[n,k] = fl( fl(zz[n] + cc[k]) - fl(2 * dot[n,k]) ),  dot = fmaf chain over d */
static void dist_row(const float *zrow, float zz, const float *wt, const float *cc,
                     int k, int d, float *acc /* [k] scratch, returns d */)
{
    for (int j = 0; j < k; ++j) acc[j] = 0.0f;
    for (int i = 0; i < d; ++i) {
        float zi = zrow[i];
        const float *wr = wt + (size_t)i * k;
        for (int j = 0; j < k; ++j) acc[j] = fmaf(zi, wr[j], acc[j]);
    }
    for (int j = 0; j < k; ++j) {
        float t = zz + cc[j];
        float two = 2.0f * acc[j];
        acc[j] = t - two;
    }
}

/* norm_ema_quantizer.py:175-179: distances + first-index argmin.
 * zf: [n,d] (already l2-normalised by the caller), w: [k,d].
 * idx_out int64 [n]; dmin_out (optional) the winning distance. */
VQO_API void vqo_hard_assign(const float *zf, const float *w, int64_t n, int d, int k,
                             int64_t *idx_out, float *dmin_out)
{
    float *wt = transpose_kd(w, k, d);
    float *cc = (float *)malloc((size_t)k * sizeof(float));
    for (int j = 0; j < k; ++j) cc[j] = sumsq_row(w + (size_t)j * d, d);
#pragma omp parallel
    {
        float *acc = (float *)malloc((size_t)k * sizeof(float));
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < n; ++i) {
            const float *zr = zf + i * d;
            dist_row(zr, sumsq_row(zr, d), wt, cc, k, d, acc);
            int best = 0;
            float bv = acc[0];
            for (int j = 1; j < k; ++j)
                if (acc[j] < bv) { bv = acc[j]; best = j; } /* strict <  => first index wins */
            idx_out[i] = best;
            if (dmin_out) dmin_out[i] = bv;
        }
        free(acc);
    }
    free(cc);
    free(wt);
}

/* full distance matrix for small cases (debug / fixtures) */
VQO_API void vqo_dist_matrix(const float *zf, const float *w, int64_t n, int d, int k, float *dout)
{
    float *wt = transpose_kd(w, k, d);
    float *cc = (float *)malloc((size_t)k * sizeof(float));
    for (int j = 0; j < k; ++j) cc[j] = sumsq_row(w + (size_t)j * d, d);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float *zr = zf + i * d;
        dist_row(zr, sumsq_row(zr, d), wt, cc, k, d, dout + i * k);
    }
    free(cc);
    free(wt);
}

/* soft :157-160 / :203-204  topk(d, kk, largest=False) (ascending, lowest index first on
 * exact ties -- the reference's tie order is unspecified) and softmax(-vals).
 * xn: [n,d] normalised inputs, cn: [k,d] normalised codebook slice. */
VQO_API void vqo_soft_topk(const float *xn, const float *cn, int64_t n, int d, int k, int kk,
                           int64_t *idx_out /*[n,kk]*/, float *val_out /*[n,kk]*/,
                           float *wgt_out /*[n,kk]*/, float *zq_out /*[n,d]*/)
{
    float *wt = transpose_kd(cn, k, d);
    float *cc = (float *)malloc((size_t)k * sizeof(float));
    for (int j = 0; j < k; ++j) cc[j] = sumsq_row(cn + (size_t)j * d, d);
#pragma omp parallel
    {
        float *acc = (float *)malloc((size_t)k * sizeof(float));
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < n; ++i) {
            const float *xr = xn + i * d;
            dist_row(xr, sumsq_row(xr, d), wt, cc, k, d, acc);
            int64_t *ti = idx_out + i * kk;
            float *tv = val_out + i * kk;
            int cnt = 0;
            for (int j = 0; j < k; ++j) { /* insertion into an ascending list of kk */
                float v = acc[j];
                if (cnt < kk || v < tv[cnt - 1]) {
                    int p = (cnt < kk) ? cnt : kk - 1;
                    while (p > 0 && v < tv[p - 1]) { tv[p] = tv[p - 1]; ti[p] = ti[p - 1]; --p; }
                    tv[p] = v; ti[p] = j;
                    if (cnt < kk) ++cnt;
                }
            }
            /* softmax(-vals): torch CPU softmax = exp(x - max) / sum, fp32 */
            float mx = -tv[0];
            float e[64], s = 0.0f;
            for (int j = 0; j < kk; ++j) { e[j] = expf(-tv[j] - mx); s += e[j]; }
            float *wg = wgt_out + i * kk;
            for (int j = 0; j < kk; ++j) wg[j] = e[j] / s;
            if (zq_out) { /* soft :205  sum_j w_j * cn[idx_j], products then sequential sum over j */
                for (int c = 0; c < d; ++c) {
                    float a = 0.0f;
                    for (int j = 0; j < kk; ++j) {
                        float p = wg[j] * cn[(size_t)ti[j] * d + c];
                        a = a + p;
                    }
                    zq_out[i * d + c] = a;
                }
            }
        }
        free(acc);
    }
    free(cc);
    free(wt);
}

/* ---- statistics + EMA ------------------------------------------------------ */

/* norm_ema_quantizer.py:183,187,194,202: bins = one_hot(idx).sum(0) (exact integers),
 * embed_sum = zf.t() @ one_hot (here [k,d], accumulated in double: the reference's
 * fp32 GEMM order over n is opaque, see DESIGN.md "stats tolerance"). */
VQO_API void vqo_stats(const float *zf, const int64_t *idx, int64_t n, int d, int k,
                       double *bins /*[k]*/, double *embed_sum /*[k,d]*/)
{
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = idx[i];
        bins[c] += 1.0;
        const float *zr = zf + i * d;
        double *er = embed_sum + c * d;
        for (int j = 0; j < d; ++j) er[j] += (double)zr[j];
    }
}

/* norm_ema_quantizer.py:11-12,189,197  cluster_size <- decay*cs + (1-decay)*bins */
VQO_API void vqo_ema_cluster_size(float *cluster_size, const double *bins, int k, double decay_d)
{
    /* python: mul_(decay).add_(new, alpha=1-decay) -- both scalars are doubles cast to fp32 */
    float decay = (float)decay_d;
    float alpha = (float)(1.0 - decay_d);
    for (int j = 0; j < k; ++j) {
        float t = cluster_size[j] * decay;
        cluster_size[j] = fmaf((float)bins[j], alpha, t);
    }
}

/* norm_ema_quantizer.py:199-210 + :136-138
 *   e_k = l2norm(embed_sum_k / max(bins_k,1))  (sequential, no-FMA norm: strided .t() view)
 *   e_k = w_k where bins_k == 0
 *   w_k <- l2norm( fmaf(e_k, 1-decay, fl(w_k*decay)) )           (contiguous l2norm) */
VQO_API void vqo_ema_codebook(float *weight /*[k,d] in place*/, const double *bins,
                              const double *embed_sum, int k, int d, double decay_d)
{
    float decay = (float)decay_d;
    float alpha = (float)(1.0 - decay_d);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < k; ++j) {
        float e[1024], t[1024];
        float *wr = weight + (size_t)j * d;
        if (bins[j] == 0.0) {
            for (int c = 0; c < d; ++c) e[c] = wr[c];
        } else {
            float b = (float)bins[j];
            for (int c = 0; c < d; ++c) e[c] = (float)embed_sum[(size_t)j * d + c] / b;
            float den = norm2_row_seq(e, d);
            if (den < 1e-12f) den = 1e-12f;
            for (int c = 0; c < d; ++c) e[c] = e[c] / den;
        }
        for (int c = 0; c < d; ++c) {
            float m = wr[c] * decay;
            t[c] = fmaf(e[c], alpha, m);
        }
        float den = norm2_row(t, d);
        if (den < 1e-12f) den = 1e-12f;
        for (int c = 0; c < d; ++c) wr[c] = t[c] / den;
    }
}

/* norm_ema_quantizer.py:181,212-214: z_q = w[idx]; sse = sum (z_q - zf)^2 (double);
 * out = zf + (z_q - zf)  (straight-through value, fp32, this exact order) */
VQO_API double vqo_gather_ste(const float *zf, const float *w, const int64_t *idx, int64_t n,
                              int d, float *out)
{
    double sse = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const float *zr = zf + i * d;
        const float *wr = w + idx[i] * d;
        for (int j = 0; j < d; ++j) {
            float diff = wr[j] - zr[j];
            float sq = diff * diff;
            sse += (double)sq;
            out[i * d + j] = zr[j] + diff;
        }
    }
    return sse;
}

VQO_API int vqo_abi_version(void) { return 1; }
