// K1 and its scalar companion: the unary factors of a batch of sentences.
//
// Reference semantics (do not "fix"): a unary factor's message is normalize(copy(table column)) (LBP.py:492-498),
// its belief ignores the variable's other messages (LBP.py:540), and var->unary-factor messages do not exist
// (LBP.py:211-212, :237).  So everything a variable's unary factors contribute to inference is ONE constant
// vector -- the product of their normalised messages -- and their gradient is a closed form of per-theta
// statistics (column sums of the table planes, per-German-word sums of psi) plus a few sparse corrections.
#include "common.cuh"

namespace mlbp {

struct ThetaED { double t[6]; };

__device__ __forceinline__ double psi_base(const float *edT, const float *pedT, int ldf, int d, int e,
                                           const ThetaED &th) {
    return exp(th.t[0] * (double)edT[(size_t)d * ldf + e] + th.t[1] * (double)pedT[(size_t)d * ldf + e] + th.t[5]);
}

// total sparse exponent shift at english index e for the entries [s0, s1)
__device__ __forceinline__ double sparse_delta(const int32_t *sp_en, const int32_t *sp_feat, const float *sp_val,
                                               int s0, int s1, int e, const ThetaED &th) {
    double dl = 0.0;
    for (int s = s0; s < s1; ++s)
        if (sp_en[s] == e) dl += th.t[sp_feat[s]] * (double)sp_val[s];
    return dl;
}

// one thread per variable
__global__ void unary_stats_kernel(int nv, const int32_t *__restrict__ var_de, const int32_t *__restrict__ var_label,
                                   const int32_t *__restrict__ sp_off, const int32_t *__restrict__ sp_en,
                                   const int32_t *__restrict__ sp_feat, const float *__restrict__ sp_val,
                                   const int32_t *__restrict__ giv_off, const int32_t *__restrict__ giv_label,
                                   const int32_t *__restrict__ giv_gap1, const float *__restrict__ pmi,
                                   const float *__restrict__ w1, const float *__restrict__ edT,
                                   const float *__restrict__ pedT, int V, int ldf, ThetaED th,
                                   const double *__restrict__ edstats, const double *__restrict__ colsums,
                                   double *__restrict__ inv_sigma, double *__restrict__ g_unary) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int d = var_de[v], y = var_label[v];
    const int s0 = sp_off[v], s1 = sp_off[v + 1];
    double g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double S0 = 1.0, S1 = 0.0, S2 = 0.0;
    if (d >= 0) {                                          // d < 0: the variable has no en_de factor
    S0 = edstats[3 * d]; S1 = edstats[3 * d + 1]; S2 = edstats[3 * d + 2];
    for (int s = s0; s < s1; ++s) {  // every DISTINCT touched english index once
        const int e = sp_en[s];
        bool first = true;
        for (int q = s0; q < s; ++q) first = first && (sp_en[q] != e);
        if (!first) continue;
        const double b = psi_base(edT, pedT, ldf, d, e, th);
        const double f = b * exp(sparse_delta(sp_en, sp_feat, sp_val, s0, s1, e, th));
        S0 += f - b;
        S1 += (f - b) * (double)edT[(size_t)d * ldf + e];
        S2 += (f - b) * (double)pedT[(size_t)d * ldf + e];
    }
    g[3] = (double)edT[(size_t)d * ldf + y] - S1 / S0;    // ed
    g[4] = (double)pedT[(size_t)d * ldf + y] - S2 / S0;   // ped
    for (int s = s0; s < s1; ++s) {                        // correct / full_history / hit_history
        const int e = sp_en[s];
        const double f = psi_base(edT, pedT, ldf, d, e, th) * exp(sparse_delta(sp_en, sp_feat, sp_val, s0, s1, e, th));
        g[3 + sp_feat[s]] += (double)sp_val[s] * ((e == y ? 1.0 : 0.0) - f / S0);
    }
    }
    // bias features: observed 1 - expected 1 (the reference yields ~1e-16 noise here)
    const double *csT = colsums, *csT1 = colsums + V, *csG = colsums + 2 * (size_t)V, *csG1 = colsums + 3 * (size_t)V,
                 *csG1w = colsums + 4 * (size_t)V;
    for (int j = giv_off[v]; j < giv_off[v + 1]; ++j) {   // unary en_en factors: column giv_label of T / T1
        const int o = giv_label[j];
        if (giv_gap1[j]) {
            g[0] += (double)pmi[(size_t)y * ldf + o] - csG1[o] / csT1[o];
            g[1] += (double)w1[(size_t)y * ldf + o] - csG1w[o] / csT1[o];
        } else {
            g[0] += (double)pmi[(size_t)y * ldf + o] - csG[o] / csT[o];  // phi_en_en has a zero pmi_w1 plane (train.py:595)
        }
    }
    inv_sigma[v] = 1.0 / S0;
#pragma unroll
    for (int i = 0; i < 9; ++i) g_unary[(size_t)v * 9 + i] = g[i];
}

// grid (nv, chunks), 256 threads, 4 consecutive elements per thread.
__global__ void __launch_bounds__(256)
unary_products_kernel(const int32_t *__restrict__ var_de, const int32_t *__restrict__ sp_off,
                      const int32_t *__restrict__ sp_en, const int32_t *__restrict__ sp_feat,
                      const float *__restrict__ sp_val, const int32_t *__restrict__ giv_off,
                      const int32_t *__restrict__ giv_label, const int32_t *__restrict__ giv_gap1,
                      const float *__restrict__ edT, const float *__restrict__ pedT, int V, int ldf, ThetaED th,
                      const double *__restrict__ inv_sigma, const __half *__restrict__ planes, int64_t ps, int ldv,
                      double unscale, const double *__restrict__ colsums, float *__restrict__ U) {
    const int v = blockIdx.x;
    const int d = var_de[v];
    const int e0 = (blockIdx.y * 256 + threadIdx.x) * 4;
    const int g0 = giv_off[v], g1 = giv_off[v + 1];
    const double scale0 = (double)V * inv_sigma[v];
    float *urow = U + (size_t)v * ldv;
    if (e0 < V) {
        // fp32 exp / products here (1 ulp expf, 1e-7 relative): the fp64 pipe of B200 issues ~3 lanes/clk/SM.  The
        // normaliser inv_sigma and the sparse fix-ups below stay in float64 (a handful of values per variable).
        float u[4] = {1.f, 1.f, 1.f, 1.f};                  // d < 0: no en_de factor on this variable
        if (d >= 0) {
            const float4 ed4 = *reinterpret_cast<const float4 *>(edT + (size_t)d * ldf + e0);
            const float4 pd4 = *reinterpret_cast<const float4 *>(pedT + (size_t)d * ldf + e0);
            const float ed[4] = {ed4.x, ed4.y, ed4.z, ed4.w}, pd[4] = {pd4.x, pd4.y, pd4.z, pd4.w};
            const float t0 = (float)th.t[0], t1 = (float)th.t[1], lsc = (float)(th.t[5] + log(scale0));
#pragma unroll
            for (int i = 0; i < 4; ++i) u[i] = expf(fmaf(t0, ed[i], fmaf(t1, pd[i], lsc)));
        }
        for (int j = g0; j < g1; ++j) {
            const int o = giv_label[j];
            const int tp = giv_gap1[j] ? 6 : 2;  // T1t / Tt plane pair: row o = column o of T1 / T
            const double cs = colsums[(size_t)(giv_gap1[j] ? 1 : 0) * V + o];
            const float f = (float)(unscale * (double)V / cs);
            const __half2 *hi = reinterpret_cast<const __half2 *>(planes + (size_t)tp * ps + (size_t)o * ldv + e0);
            const __half2 *lo = reinterpret_cast<const __half2 *>(planes + (size_t)(tp + 1) * ps + (size_t)o * ldv + e0);
            const float2 h0 = __half22float2(hi[0]), h1 = __half22float2(hi[1]);
            const float2 l0 = __half22float2(lo[0]), l1 = __half22float2(lo[1]);
            u[0] *= (h0.x + l0.x) * f;
            u[1] *= (h0.y + l0.y) * f;
            u[2] *= (h1.x + l1.x) * f;
            u[3] *= (h1.y + l1.y) * f;
        }
        // columns >= V of the padded row stay zero (ldf, ldv are multiples of 4; reads stay inside the row)
        float4 o4;
        o4.x = (e0 + 0 < V) ? u[0] : 0.f;
        o4.y = (e0 + 1 < V) ? u[1] : 0.f;
        o4.z = (e0 + 2 < V) ? u[2] : 0.f;
        o4.w = (e0 + 3 < V) ? u[3] : 0.f;
        *reinterpret_cast<float4 *>(urow + e0) = o4;
    }
    // sparse per-sentence features (train.py:176-215): re-evaluate the touched entries of this chunk
    const int s0 = sp_off[v], s1 = sp_off[v + 1];
    if (s1 > s0 && d >= 0) {
        __syncthreads();
        const int c0 = blockIdx.y * 1024, c1 = min(c0 + 1024, V);
        for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
            const int e = sp_en[s];
            if (e < c0 || e >= c1) continue;
            bool first = true;
            for (int q = s0; q < s; ++q) first = first && (sp_en[q] != e);
            if (!first) continue;
            double u = psi_base(edT, pedT, ldf, d, e, th) * exp(sparse_delta(sp_en, sp_feat, sp_val, s0, s1, e, th)) * scale0;
            for (int j = g0; j < g1; ++j) {
                const int o = giv_label[j];
                const int tp = giv_gap1[j] ? 6 : 2;
                const double cs = colsums[(size_t)(giv_gap1[j] ? 1 : 0) * V + o];
                const double t = (double)__half2float(planes[(size_t)tp * ps + (size_t)o * ldv + e]) +
                                 (double)__half2float(planes[(size_t)(tp + 1) * ps + (size_t)o * ldv + e]);
                u *= t * unscale * (double)V / cs;
            }
            urow[e] = (float)u;
        }
    }
}

}  // namespace mlbp

using namespace mlbp;

extern "C" int mlbp_unary_stats(int nv, const int32_t *var_de, const int32_t *var_label, const int32_t *sp_off,
                                const int32_t *sp_en, const int32_t *sp_feat, const float *sp_val,
                                const int32_t *giv_off, const int32_t *giv_label, const int32_t *giv_gap1,
                                const float *pmi, const float *pmi_w1, const float *edT, const float *pedT, int V,
                                int ldf, const double *h_theta_ed, const double *edstats, const double *colsums,
                                double *inv_sigma, double *g_unary, void *stream) {
    if (nv == 0) return MLBP_OK;
    MLBP_CHECK_ARG(nv > 0 && var_de && var_label && sp_off && giv_off && pmi && pmi_w1 && edT && pedT && edstats &&
                   colsums && inv_sigma && g_unary && h_theta_ed, "unary_stats: null pointer");
    ThetaED th;
    for (int i = 0; i < 6; ++i) th.t[i] = h_theta_ed[i];
    unary_stats_kernel<<<(nv + 127) / 128, 128, 0, as_stream(stream)>>>(
        nv, var_de, var_label, sp_off, sp_en, sp_feat, sp_val, giv_off, giv_label, giv_gap1, pmi, pmi_w1, edT, pedT, V,
        ldf, th, edstats, colsums, inv_sigma, g_unary);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_unary_products(int nv, const int32_t *var_de, const int32_t *sp_off, const int32_t *sp_en,
                                   const int32_t *sp_feat, const float *sp_val, const int32_t *giv_off,
                                   const int32_t *giv_label, const int32_t *giv_gap1, const float *edT,
                                   const float *pedT, int V, int ldf, const double *h_theta_ed,
                                   const double *inv_sigma, const void *planes, int64_t plane_stride, int ldv,
                                   int scale_exp, const double *colsums, float *U, void *stream) {
    if (nv == 0) return MLBP_OK;
    MLBP_CHECK_ARG(nv > 0 && var_de && sp_off && giv_off && edT && pedT && inv_sigma && planes && colsums && U &&
                   h_theta_ed, "unary_products: null pointer");
    MLBP_CHECK_ARG((ldf % 4) == 0 && (ldv % 64) == 0 && ldv >= V && ldf >= V, "unary_products: bad ld");
    ThetaED th;
    for (int i = 0; i < 6; ++i) th.t[i] = h_theta_ed[i];
    dim3 grid(nv, (V + 1023) / 1024);
    unary_products_kernel<<<grid, 256, 0, as_stream(stream)>>>(var_de, sp_off, sp_en, sp_feat, sp_val, giv_off,
                                                               giv_label, giv_gap1, edT, pedT, V, ldf, th, inv_sigma,
                                                               (const __half *)planes, plane_stride, ldv,
                                                               ldexp(1.0, -scale_exp), colsums, U);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}
