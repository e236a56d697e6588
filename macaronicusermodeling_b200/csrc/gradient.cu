// K6: observed-minus-expected feature counts of the pairwise factors and the per-sentence reduction.
//
// The reference materialises the dense V x V belief  normalize((c r') o T)  per pairwise factor and contracts it
// with the (V,V,3) feature tensor (LBP.py:544-569, :610): ~7 passes over V^2 doubles per factor.  In closed form
//  only three inner products per factor are needed,
//     Z = c . (T r) = r . (T'c),   N1 = c . ((T o PMI) r),   N2 = c . ((T1 o PMI_w1) r)
// where the matrix-vector products are rows of the batched GEMM (K4).  This file does the inner products (K6a)
// and the segmented sum over each sentence's factors and variables (K6b).
#include "common.cuh"

namespace mlbp {

// one CTA per pairwise factor
__global__ void __launch_bounds__(256)
pair_expectations_kernel(const int32_t *__restrict__ c_row, const int32_t *__restrict__ z_row,
                         const int32_t *__restrict__ u0_row,
                         const int32_t *__restrict__ u1_row, const int32_t *__restrict__ u2_row,
                         const __half *__restrict__ A_hi, const __half *__restrict__ A_lo, const float *__restrict__ D,
                         int ldv, int V, double *__restrict__ stats, const int32_t *__restrict__ r_row,
                         const int32_t *__restrict__ gap1, const int32_t *__restrict__ spike_words,
                         const int32_t *__restrict__ spike_cnt, const int2 *__restrict__ spike_entries,
                         const __half *__restrict__ planes, int64_t ps, float alpha) {
    __shared__ double red[32];
    const int f = blockIdx.x;
    const __half *ch = A_hi + (size_t)c_row[f] * ldv, *cl = A_lo + (size_t)c_row[f] * ldv;
    // Z = z . u0: (z, u0) = (c, T r) in general, or (r, T'c) when the plan reuses the D row of a message update
    const bool zc = z_row[f] == c_row[f];
    const __half *zh = A_hi + (size_t)z_row[f] * ldv, *zl = A_lo + (size_t)z_row[f] * ldv;
    const float *u0 = D + (size_t)u0_row[f] * ldv, *u1 = D + (size_t)u1_row[f] * ldv;
    const float *u2 = u2_row[f] >= 0 ? D + (size_t)u2_row[f] * ldv : nullptr;
    // fp32 products, short per-thread fp32 partial sums (V / 256 terms), float64 across the block: the fp64 pipe of
    // B200 issues ~3 lanes/clk/SM, and the ratio N/Z only needs ~1e-6
    // spike counts of the two messages: loaded with the row indices above (same dependency depth as the first loads of the loop)
    int ncs = 0, nrs = 0;
    if (spike_words && spike_words[0] == 0) {
        ncs = min(spike_cnt[c_row[f]], MLBP_SPIKE_SLOTS);
        nrs = min(spike_cnt[r_row[f]], MLBP_SPIKE_SLOTS);
    }
    // 8 consecutive elements per thread and step: 16-byte loads of the fp16 rows, two float4 per D row; rows are padded to a
    // multiple of 64 elements, so whole chunks (the tail chunk is masked: the padding of the A rows is never written).  All
    // loads of a step are independent -- the kernel is a stream of ~124 KB per factor and lives on memory-level parallelism.
    float zf = 0.f, n1f = 0.f, n2f = 0.f;
    const int n8 = (V + 7) >> 3;
    for (int k8 = threadIdx.x; k8 < n8; k8 += blockDim.x) {
        const uint4 ch8 = __ldg(reinterpret_cast<const uint4 *>(ch) + k8), cl8 = __ldg(reinterpret_cast<const uint4 *>(cl) + k8);
        const float4 a0 = __ldg(reinterpret_cast<const float4 *>(u0) + 2 * k8), a1 = __ldg(reinterpret_cast<const float4 *>(u0) + 2 * k8 + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(u1) + 2 * k8), b1 = __ldg(reinterpret_cast<const float4 *>(u1) + 2 * k8 + 1);
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        if (u2) { c0 = __ldg(reinterpret_cast<const float4 *>(u2) + 2 * k8); c1 = __ldg(reinterpret_cast<const float4 *>(u2) + 2 * k8 + 1); }
        uint4 zh8 = ch8, zl8 = cl8;
        if (!zc) { zh8 = __ldg(reinterpret_cast<const uint4 *>(zh) + k8); zl8 = __ldg(reinterpret_cast<const uint4 *>(zl) + k8); }
        float c[8], zz[8];
        {
            const __half2 *hh = reinterpret_cast<const __half2 *>(&ch8), *ll = reinterpret_cast<const __half2 *>(&cl8);
            const __half2 *zhh = reinterpret_cast<const __half2 *>(&zh8), *zll = reinterpret_cast<const __half2 *>(&zl8);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 h = __half22float2(hh[q]), l = __half22float2(ll[q]);
                const float2 zh2 = __half22float2(zhh[q]), zl2 = __half22float2(zll[q]);
                c[2 * q] = h.x + l.x; c[2 * q + 1] = h.y + l.y;
                zz[2 * q] = zh2.x + zl2.x; zz[2 * q + 1] = zh2.y + zl2.y;
            }
        }
        if (8 * k8 + 8 > V) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (8 * k8 + i >= V) { c[i] = 0.f; zz[i] = 0.f; }
        }
        const float ua[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float ub[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float uc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            zf = fmaf(zz[i], ua[i], zf);
            n1f = fmaf(c[i], ub[i], n1f);
            n2f = fmaf(c[i], uc[i], n2f);
        }
    }
    // One-pass gradient rows (u = alpha * r_hi . B_hi) drop the lo half of the table planes.  Its rounding averages away
    // over the cells a belief spreads over -- except where BOTH messages have a spike: those few cells are restored here,
    //   sum over spikes a* of c, b* of r:  alpha * c[a*] * r_hi[b*] * B_lo[a*, b*]
    // one cell per thread (at most MLBP_SPIKE_SLOTS^2), folded into the block sums below (fixed order: deterministic).  Only
    // factors whose two messages BOTH have spikes pay for the chain of dependent loads (entries -> cells).
    // (spike lists: mlbp_var_to_factor; skipped when a row had more spikes than slots -- then the rows ran two passes).
    double cz = 0.0, c1 = 0.0, c2 = 0.0;
    if (ncs > 0 && nrs > 0 && threadIdx.x < MLBP_SPIKE_SLOTS * MLBP_SPIKE_SLOTS) {
        const int rr = r_row[f], cr = c_row[f];
        const int i = threadIdx.x / MLBP_SPIKE_SLOTS, j = threadIdx.x % MLBP_SPIKE_SLOTS;
        if (i < ncs && j < nrs) {
            const int g1 = gap1[f];
            const int a = spike_entries[(size_t)cr * MLBP_SPIKE_SLOTS + i].x, b = spike_entries[(size_t)rr * MLBP_SPIKE_SLOTS + j].x;
            const double w = (double)alpha * (double)(__half2float(ch[a]) + __half2float(cl[a])) *
                             (double)__half2float(A_hi[(size_t)rr * ldv + b]);
            const size_t cell = (size_t)a * ldv + b;
            if (zc) cz = w * (double)__half2float(planes[(size_t)(2 * (g1 ? MLBP_TABLE_T1 : MLBP_TABLE_T) + 1) * ps + cell]);
            c1 = w * (double)__half2float(planes[(size_t)(2 * (g1 ? MLBP_TABLE_G1 : MLBP_TABLE_G) + 1) * ps + cell]);
            if (u2) c2 = w * (double)__half2float(planes[(size_t)(2 * MLBP_TABLE_G1W + 1) * ps + cell]);
        }
    }
    double z = block_sum((double)zf + cz, red), n1 = block_sum((double)n1f + c1, red), n2 = block_sum((double)n2f + c2, red);
    if (threadIdx.x == 0) { stats[3 * (size_t)f] = z; stats[3 * (size_t)f + 1] = n1; stats[3 * (size_t)f + 2] = n2; }
}

// one warp per sentence; deterministic (no atomics).  The observed feature values phi[l0, l1, :] are gathered here from the
// supervised labels of the factor's two variables (FactorNode.get_observed_factor, LBP.py:584-589).
__global__ void gradient_reduce_kernel(int n_sent, const int32_t *__restrict__ sent_var_off,
                                       const int32_t *__restrict__ sent_fac_off, const double *__restrict__ g_unary,
                                       const double *__restrict__ pair_stats, const int32_t *__restrict__ pair_v0,
                                       const int32_t *__restrict__ pair_v1, const int32_t *__restrict__ var_label,
                                       const int32_t *__restrict__ gap1,
                                       const float *__restrict__ pmi, const float *__restrict__ w1, int ldf,
                                       const double *__restrict__ logp_var, double *__restrict__ grad,
                                       double *__restrict__ logp_sent) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n_sent) return;
    double g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double lp = 0.0;
    for (int v = sent_var_off[s] + lane; v < sent_var_off[s + 1]; v += 32) {
#pragma unroll
        for (int i = 0; i < 9; ++i) g[i] += g_unary[(size_t)v * 9 + i];
        if (logp_var) lp += logp_var[v];
    }
    const int f0 = sent_fac_off ? sent_fac_off[s] : 0, f1 = sent_fac_off ? sent_fac_off[s + 1] : 0;
    for (int f = f0 + lane; f < f1; f += 32) {
        const double z = pair_stats[3 * (size_t)f];
        const size_t cell = (size_t)var_label[pair_v0[f]] * ldf + var_label[pair_v1[f]];
        // Z <= 0: the reference's normalize zero-fills the belief (pyx:39-40) -> expected counts are 0
        const double e1 = z > 0.0 ? pair_stats[3 * (size_t)f + 1] / z : 0.0;
        g[0] += (double)pmi[cell] - e1;
        if (gap1[f]) {
            const double e2 = z > 0.0 ? pair_stats[3 * (size_t)f + 2] / z : 0.0;
            g[1] += (double)w1[cell] - e2;
        }
        g[2] += z > 0.0 ? 0.0 : 1.0;  // bias: 1 - sum(belief)
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] = warp_sum(g[i]);
    lp = warp_sum(lp);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) grad[(size_t)s * 9 + i] = g[i];
        if (logp_sent) logp_sent[s] = lp;
    }
}

// Batch level of the reduction (train_mp.py:405-424: the parent sums what the workers return): ONE CTA adds this micro-batch's
// per-sentence gradients, log-posteriors and precision counts (LBP.py:80-106: P@0 = rank 0, P@25 = rank < 26, P@50 = rank < 50
// as counted by Result.precision_counts) into the 16-double vector the NCCL all-reduce ships,
//   [g_ee(3), g_ed(6), sum logp, p@0, p@25, p@50, n_vars, n_sent, peaked].  Fixed summation order: deterministic.
__global__ void __launch_bounds__(1024)
batch_reduce_kernel(int n_sent, const double *__restrict__ grad, const double *__restrict__ logp_sent, int n_vars,
                    const int32_t *__restrict__ rank, const int32_t *__restrict__ peak_flag, double *__restrict__ out16) {
    __shared__ double red[32];
    double acc[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) acc[i] = 0.0;
    for (int s = threadIdx.x; s < n_sent; s += blockDim.x) {
        if (grad) {
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[i] += grad[(size_t)s * 9 + i];
        }
        if (logp_sent) acc[9] += logp_sent[s];
    }
    if (rank)
        for (int v = threadIdx.x; v < n_vars; v += blockDim.x) {
            const int r = rank[v];
            acc[10] += r == 0 ? 1.0 : 0.0; acc[11] += r < 26 ? 1.0 : 0.0; acc[12] += r < 50 ? 1.0 : 0.0;
        }
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const double t = block_sum(acc[i], red);
        if (threadIdx.x == 0) out16[i] += t;
    }
    if (threadIdx.x == 0) {
        out16[13] += rank ? (double)n_vars : 0.0;
        out16[14] += (double)n_sent;
        if (peak_flag) {                                           // 1 = PEAK (message rows ran three passes), 2 = SPIKE word
            const double code = (peak_flag[0] ? 1.0 : 0.0) + (peak_flag[3] ? 2.0 : 0.0);
            if (code > out16[15]) out16[15] = code;
        }
    }
}

// D rows 0..4 of a theta: row 0 = the constant-one row (messages still uniform), rows 1..4 = the factor->variable message
// of a pairwise factor that is fed the uniform initial message (LBP.py:211-216, :509, :518): row / column sums of T, T1,
// mean-one scaled (messages are scale-free).  Table order MLBP_TABLE_T, TT, T1, T1T <- colsums rows 5, 0, 6, 1.
__global__ void __launch_bounds__(256)
const_rows_kernel(const double *__restrict__ colsums, int V, int ldv, float *__restrict__ out) {
    __shared__ double red[32];
    const int t = blockIdx.x;                                      // 0: ones, 1..4: tables
    float *row = out + (size_t)t * ldv;
    if (t == 0) {
        for (int e = threadIdx.x; e < ldv; e += blockDim.x) row[e] = 1.0f;
        return;
    }
    const int src[4] = {5, 0, 6, 1};
    const double *cs = colsums + (size_t)src[t - 1] * V;
    double s = 0.0;
    for (int e = threadIdx.x; e < V; e += blockDim.x) s += cs[e];
    s = block_sum(s, red);
    const double inv_mean = s > 0.0 ? (double)V / s : 1.0;
    for (int e = threadIdx.x; e < ldv; e += blockDim.x) row[e] = e < V ? (float)(cs[e] * inv_mean) : 0.f;
}

}  // namespace mlbp

using namespace mlbp;

extern "C" int mlbp_pair_expectations(int n_factors, const int32_t *c_row, const int32_t *z_row,
                                      const int32_t *u0_row, const int32_t *u1_row, const int32_t *u2_row, const void *A_hi,
                                      const void *A_lo, const float *D, int ldv, int V, double *stats,
                                      const int32_t *r_row, const int32_t *pair_gap1, const int32_t *spike_words,
                                      const int32_t *spike_cnt, const int32_t *spike_entries, const void *planes,
                                      int64_t plane_stride, float alpha, void *stream) {
    if (n_factors == 0) return MLBP_OK;
    MLBP_CHECK_ARG(n_factors > 0 && c_row && z_row && u0_row && u1_row && u2_row && A_hi && A_lo && D && stats,
                   "pair_expectations: null pointer");
    MLBP_CHECK_ARG((ldv % 8) == 0 && ldv >= V && ((reinterpret_cast<uintptr_t>(A_hi) | reinterpret_cast<uintptr_t>(A_lo) |
                                                  reinterpret_cast<uintptr_t>(D)) % 16) == 0,
                   "pair_expectations: rows must be 16-byte aligned and padded to a multiple of 8 elements");
    MLBP_CHECK_ARG(!spike_words || (r_row && pair_gap1 && spike_cnt && spike_entries && planes),
                   "pair_expectations: the spike-cell correction needs r_row, pair_gap1, the spike lists and the table planes");
    pair_expectations_kernel<<<n_factors, 256, 0, as_stream(stream)>>>(c_row, z_row, u0_row, u1_row, u2_row,
                                                                       (const __half *)A_hi, (const __half *)A_lo, D,
                                                                       ldv, V, stats, r_row, pair_gap1, spike_words, spike_cnt,
                                                                       reinterpret_cast<const int2 *>(spike_entries),
                                                                       (const __half *)planes, plane_stride, alpha);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_gradient_reduce(int n_sent, const int32_t *sent_var_off, const int32_t *sent_fac_off,
                                    const double *g_unary, const double *pair_stats, const int32_t *pair_v0,
                                    const int32_t *pair_v1, const int32_t *var_label, const int32_t *pair_gap1,
                                    const float *pmi, const float *pmi_w1, int ldf, const double *logp_var, double *grad,
                                    double *logp_sent, void *stream) {
    if (n_sent == 0) return MLBP_OK;
    MLBP_CHECK_ARG(n_sent > 0 && sent_var_off && g_unary && pmi && pmi_w1 && grad, "gradient_reduce: null pointer");
    MLBP_CHECK_ARG(!sent_fac_off || (pair_stats && pair_v0 && pair_v1 && var_label && pair_gap1),
                   "gradient_reduce: pairwise factors need pair_stats, pair_v0, pair_v1, var_label, pair_gap1");
    const int threads = 128, warps_per_block = threads / 32;
    gradient_reduce_kernel<<<(n_sent + warps_per_block - 1) / warps_per_block, threads, 0, as_stream(stream)>>>(
        n_sent, sent_var_off, sent_fac_off, g_unary, pair_stats, pair_v0, pair_v1, var_label, pair_gap1, pmi, pmi_w1, ldf,
        logp_var, grad, logp_sent);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_batch_reduce(int n_sent, const double *grad, const double *logp_sent, int n_vars, const int32_t *rank,
                                 const int32_t *peak_flag, double *out16, void *stream) {
    MLBP_CHECK_ARG(n_sent >= 0 && n_vars >= 0 && out16, "batch_reduce: bad argument");
    if (n_sent == 0 && n_vars == 0) return MLBP_OK;
    batch_reduce_kernel<<<1, 1024, 0, as_stream(stream)>>>(n_sent, grad, logp_sent, n_vars, rank, peak_flag, out16);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_const_rows(const double *colsums, int V, int ldv, float *rows, void *stream) {
    MLBP_CHECK_ARG(colsums && rows && V > 0 && ldv >= V, "const_rows: bad argument");
    const_rows_kernel<<<MLBP_D_CONST_ROWS, 256, 0, as_stream(stream)>>>(colsums, V, ldv, rows);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}
