// K3 (variable -> factor messages) and K5 (marginals / log-posterior / argmax / rank).
//
// Both multiply the messages coming into one variable.  The reference does it one np.multiply per incoming
// message and per OUTGOING edge (LBP.py:377-389, :717-730): O(deg^2) vector passes per variable and sweep.
// Here one CTA owns one (variable, schedule level) group, reads each incoming row once per phase and emits
// every needed leave-one-out product from a prefix / suffix product in registers.  Products are carried in
// float64: a variable can have ~40 incoming messages and the raw product of fp32 values would leave the fp32
// range (the reference multiplies fp64 values of size 1/V and never rescales, LBP.py:381-385).
//
// Phase 1 sums every outgoing product (the renormalisation of Message.renormalize, LBP.py:649-657);
// phase 2 recomputes it, scales to 2^14 / sum and splits into the fp16 hi / lo operand rows the pairwise GEMM
// (K4) consumes through TMA.
//
// T = float when the caller's bound (range_log2) proves that no product can leave the fp32 range, else double.
// Measured on B200 the fp64 pipe issues only ~3 lanes/clk/SM (profiles/README.md), so the double variant is
// compute-bound at ~16 % of HBM bandwidth; it is the always-safe fallback.  A bulk-async (cp.async.bulk + mbarrier)
// shared-memory staging of the inputs, and a cluster/DSMEM single-read variant (8 CTAs per group, slices kept in shared
// memory between the phases), were both tried in round 1 and were slower than these plain coalesced loads.
#include <stdlib.h>

#include "common.cuh"

namespace mlbp {

constexpr int K3_THREADS = 256;

__global__ void fill_uniform_rows_kernel(__half *__restrict__ A_hi, __half *__restrict__ A_lo, int ldv, int V,
                                         const int32_t *__restrict__ rows, const uint8_t *__restrict__ keep) {
    const size_t r = (size_t)rows[blockIdx.x];
    __half hi, lo;
    split_f16(ldexpf(1.0f, MLBP_A_SCALE_LOG2) / (float)V, hi, lo);
    const __half z = __float2half_rn(0.f);
    for (int e = threadIdx.x; e < ldv; e += blockDim.x) {
        const bool on = e < V && (!keep || keep[e]);
        A_hi[r * ldv + e] = on ? hi : z;
        A_lo[r * ldv + e] = on ? lo : z;
    }
}

template <int NMAX, typename T>
__global__ void __launch_bounds__(K3_THREADS)
var_to_factor_kernel(const int32_t *__restrict__ grp_u, const int32_t *__restrict__ grp_off,
                     const int32_t *__restrict__ in_row, const int32_t *__restrict__ dest_off,
                     const int32_t *__restrict__ dest, const float *__restrict__ U, const float *__restrict__ D,
                     int ldv, int V, __half *__restrict__ A_hi, __half *__restrict__ A_lo) {
    __shared__ const float *s_src[NMAX];
    __shared__ int s_d0[NMAX], s_d1[NMAX];
    __shared__ double s_scale[NMAX];
    __shared__ double red[32];
    const int g = blockIdx.x;
    const int i0 = grp_off[g], n = grp_off[g + 1] - i0;
    const float *urow = U + (size_t)grp_u[g] * ldv;
    if (threadIdx.x < NMAX) {
        const int j = threadIdx.x;
        if (j < n) {
            const int r = in_row[i0 + j];
            s_src[j] = r >= 0 ? D + (size_t)r * ldv : nullptr;   // nullptr: uniform message (scale-free -> 1)
            s_d0[j] = dest_off[i0 + j];
            s_d1[j] = dest_off[i0 + j + 1];
        } else {
            s_src[j] = nullptr; s_d0[j] = 0; s_d1[j] = 0;
        }
    }
    __syncthreads();

    T acc[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) acc[j] = (T)0;

    // ---- phase 1: sums of the leave-one-out products
    for (int e = threadIdx.x; e < V; e += K3_THREADS) {
        float d[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) d[j] = (j < n && s_src[j]) ? __ldg(s_src[j] + e) : 1.0f;
        T pre[NMAX];
        T p = (T)__ldg(urow + e);
#pragma unroll
        for (int j = 0; j < NMAX; ++j) { pre[j] = p; p *= (T)d[j]; }
        T suf = (T)1;
#pragma unroll
        for (int j = NMAX - 1; j >= 0; --j) {
            acc[j] += pre[j] * suf;
            suf *= (T)d[j];
        }
    }
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
        if (j < n && s_d1[j] > s_d0[j]) {                         // block-uniform condition
            const double s = block_sum((double)acc[j], red);
            if (threadIdx.x == 0)
                s_scale[j] = (s > 0.0 && isfinite(s)) ? ldexp(1.0, MLBP_A_SCALE_LOG2) / s : -1.0;  // -1: uniform fallback
        }
    }
    __syncthreads();

    // ---- phase 2: recompute, normalise, split, scatter to the consuming GEMM blocks
    const float uni = ldexpf(1.0f, MLBP_A_SCALE_LOG2) / (float)V;
    for (int e = threadIdx.x; e < V; e += K3_THREADS) {
        float d[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) d[j] = (j < n && s_src[j]) ? __ldg(s_src[j] + e) : 1.0f;
        T pre[NMAX];
        T p = (T)__ldg(urow + e);
#pragma unroll
        for (int j = 0; j < NMAX; ++j) { pre[j] = p; p *= (T)d[j]; }
        T suf = (T)1;
#pragma unroll
        for (int j = NMAX - 1; j >= 0; --j) {
            if (j < n && s_d1[j] > s_d0[j]) {
                const double sc = s_scale[j];
                const float x = sc > 0.0 ? (float)(pre[j] * suf * (T)sc) : uni;
                __half hi, lo;
                split_f16(x, hi, lo);
                for (int t = s_d0[j]; t < s_d1[j]; ++t) {
                    const size_t o = (size_t)dest[t] * ldv + e;
                    A_hi[o] = hi;
                    A_lo[o] = lo;
                }
            }
            suf *= (T)d[j];
        }
    }
}

// Top-K masking of message rows: the reference's approximate paths (use_approx_inference / use_approx_beliefs,
// LBP.py:506-507, :515-516, :554-563) contract only the K = 100 largest entries of a message
// (au.sparse_vec_mat_dot pyx:193-205, au.sparse_dot pyx:117-129).  Zeroing every other entry of the operand row and
// running the same dense GEMM gives the same sums.  One CTA per row: 4-pass radix select (8 bits per pass) on the bit
// patterns of hi + lo (non-negative floats order like their bits), then one masking pass; ties at the threshold are
// kept in arbitrary order up to K, like np.argpartition.
__global__ void __launch_bounds__(256)
topk_mask_rows_kernel(__half *__restrict__ A_hi, __half *__restrict__ A_lo, int ldv, int V, int64_t row0, int K) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_need, s_taken;
    __half *hi = A_hi + (size_t)(row0 + blockIdx.x) * ldv, *lo = A_lo + (size_t)(row0 + blockIdx.x) * ldv;
    auto bits_of = [&](int e) { return __float_as_uint(__half2float(hi[e]) + __half2float(lo[e])); };
    if (threadIdx.x == 0) { s_prefix = 0u; s_need = (unsigned)K; s_taken = 0u; }
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int e = threadIdx.x; e < V; e += 256) {
            const unsigned b = bits_of(e);
            if (shift == 24 || (b >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(b >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned cum = 0, need = s_need;
            for (int b = 255; b >= 0; --b) {
                if (cum + hist[b] >= need) { s_need = need - cum; s_prefix = prefix | ((unsigned)b << shift); break; }
                cum += hist[b];
            }
        }
        __syncthreads();
    }
    const unsigned thr = s_prefix, need = s_need;     // K-th largest bit pattern; keep `need` of the entries equal to it
    const __half z = __float2half_rn(0.f);
    for (int e = threadIdx.x; e < V; e += 256) {
        const unsigned b = bits_of(e);
        bool keep = b > thr;
        if (b == thr) keep = atomicAdd(&s_taken, 1u) < need;
        if (!keep) { hi[e] = z; lo[e] = z; }
    }
}

// one CTA per variable: total product of all incoming messages (T = float under the same range bound as K3)
template <typename T>
__global__ void __launch_bounds__(256)
marginals_kernel(const int32_t *__restrict__ grp_u, const int32_t *__restrict__ grp_off,
                 const int32_t *__restrict__ in_row, const int32_t *__restrict__ label, const float *__restrict__ U,
                 const float *__restrict__ D, int ldv, int V, double *__restrict__ logp, int32_t *__restrict__ top1,
                 int32_t *__restrict__ rank, float *__restrict__ beliefs) {
    __shared__ double red[32];
    __shared__ T s_best[8];
    __shared__ int s_besti[8];
    const int g = blockIdx.x;
    const int i0 = grp_off[g], n = grp_off[g + 1] - i0;
    const float *urow = U + (size_t)grp_u[g] * ldv;
    const int lab = label[g];
    __shared__ const float *s_rows[64];
    const int nn = min(n, 64);
    if (threadIdx.x < 64) {
        const int r = threadIdx.x < nn ? in_row[i0 + threadIdx.x] : -1;
        s_rows[threadIdx.x] = r >= 0 ? D + (size_t)r * ldv : nullptr;
    }
    __syncthreads();
    auto prod = [&](int e) {
        T p = (T)__ldg(urow + e);
        for (int j = 0; j < nn; ++j)
            if (s_rows[j]) p *= (T)__ldg(s_rows[j] + e);
        return p;
    };
    const T plab = prod(lab);
    T sp = (T)0, best = (T)-1;
    int besti = 0x7fffffff, cnt = 0;
    for (int e = threadIdx.x; e < V; e += blockDim.x) {
        const T p = prod(e);
        sp += p;
        cnt += (p > plab) ? 1 : 0;
        if (p > best) { best = p; besti = e; }                    // strided ascending e: first index wins per thread
    }
    double s = block_sum((double)sp, red);
    const double c = block_sum((double)cnt, red);
    // argmax with first-index tie break (np.argmax)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_best[threadIdx.x >> 5] = best; s_besti[threadIdx.x >> 5] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (s_best[w] > best || (s_best[w] == best && s_besti[w] < besti)) { best = s_best[w]; besti = s_besti[w]; }
        const bool ok = s > 0.0 && isfinite(s);
        // renormalize falls back to uniform when the sum is not positive (LBP.py:650-657)
        const double b = ok ? (double)plab / s : 1.0 / (double)V;
        logp[g] = b > 0.0 ? log(b) : -99.99;                      // LBP.py:252-258
        top1[g] = ok ? besti : 0;
        rank[g] = ok ? (int)(c + 0.5) : 0;
    }
    if (beliefs) {
        const bool ok = s > 0.0 && isfinite(s);
        float *brow = beliefs + (size_t)g * ldv;
        for (int e = threadIdx.x; e < V; e += blockDim.x) brow[e] = ok ? (float)((double)prod(e) / s) : 1.0f / (float)V;
    }
}

}  // namespace mlbp

using namespace mlbp;

extern "C" int mlbp_fill_uniform_rows(void *A_hi, void *A_lo, int ldv, int V, const int32_t *rows, int n_rows,
                                      const uint8_t *keep, void *stream) {
    if (n_rows == 0) return MLBP_OK;
    MLBP_CHECK_ARG(A_hi && A_lo && rows && n_rows > 0 && V > 0 && ldv >= V, "fill_uniform_rows: bad argument");
    fill_uniform_rows_kernel<<<n_rows, 256, 0, as_stream(stream)>>>((__half *)A_hi, (__half *)A_lo, ldv, V, rows, keep);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_var_to_factor(int n_groups, const int32_t *grp_u, const int32_t *grp_off, const int32_t *in_row,
                                  const int32_t *dest_off, const int32_t *dest, const float *U, const float *D,
                                  int ldv, int V, void *A_hi, void *A_lo, int max_in, float range_log2, void *stream) {
    if (n_groups == 0) return MLBP_OK;
    MLBP_CHECK_ARG(n_groups > 0 && grp_u && grp_off && in_row && dest_off && dest && U && D && A_hi && A_lo,
                   "var_to_factor: null pointer");
    const bool fp32_ok = range_log2 >= 0.f && range_log2 < 100.f;   // products provably stay inside 2^+-100
    MLBP_CHECK_ARG(V > 0 && ldv >= V && (ldv % 4) == 0, "var_to_factor: bad V/ldv");
    cudaStream_t st = as_stream(stream);
#define MLBP_K3_LAUNCH(N)                                                                                        \
    do {                                                                                                         \
        if (fp32_ok)                                                                                             \
            var_to_factor_kernel<N, float><<<n_groups, K3_THREADS, 0, st>>>(                                     \
                grp_u, grp_off, in_row, dest_off, dest, U, D, ldv, V, (__half *)A_hi, (__half *)A_lo);           \
        else                                                                                                     \
            var_to_factor_kernel<N, double><<<n_groups, K3_THREADS, 0, st>>>(                                    \
                grp_u, grp_off, in_row, dest_off, dest, U, D, ldv, V, (__half *)A_hi, (__half *)A_lo);           \
    } while (0)
    if (max_in <= 4) MLBP_K3_LAUNCH(4);
    else if (max_in <= 8) MLBP_K3_LAUNCH(8);
    else if (max_in <= 16) MLBP_K3_LAUNCH(16);
    else if (max_in <= 24) MLBP_K3_LAUNCH(24);
    else if (max_in <= 32) MLBP_K3_LAUNCH(32);
    else if (max_in <= 48) MLBP_K3_LAUNCH(48);
    else {
        set_error("var_to_factor: a variable with %d pairwise factors exceeds the supported 48", max_in);
        return MLBP_ERR_UNSUPPORTED;
    }
#undef MLBP_K3_LAUNCH
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_topk_mask_rows(void *A_hi, void *A_lo, int ldv, int V, int64_t row0, int n_rows, int K, void *stream) {
    if (n_rows == 0 || K >= V) return MLBP_OK;
    MLBP_CHECK_ARG(A_hi && A_lo && n_rows > 0 && K > 0 && row0 >= 0 && ldv >= V, "topk_mask_rows: bad argument");
    topk_mask_rows_kernel<<<n_rows, 256, 0, as_stream(stream)>>>((__half *)A_hi, (__half *)A_lo, ldv, V, row0, K);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_marginals(int n_groups, const int32_t *grp_u, const int32_t *grp_off, const int32_t *in_row,
                              const int32_t *label, const float *U, const float *D, int ldv, int V, double *logp,
                              int32_t *top1, int32_t *rank, float *beliefs, float range_log2, void *stream) {
    if (n_groups == 0) return MLBP_OK;
    MLBP_CHECK_ARG(n_groups > 0 && grp_u && grp_off && in_row && label && U && D && logp && top1 && rank,
                   "marginals: null pointer");
    if (range_log2 >= 0.f && range_log2 < 100.f)
        marginals_kernel<float><<<n_groups, 256, 0, as_stream(stream)>>>(grp_u, grp_off, in_row, label, U, D, ldv, V,
                                                                         logp, top1, rank, beliefs);
    else
        marginals_kernel<double><<<n_groups, 256, 0, as_stream(stream)>>>(grp_u, grp_off, in_row, label, U, D, ldv, V,
                                                                          logp, top1, rank, beliefs);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}
