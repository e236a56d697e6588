// K4a / K4b: spikes of the var->factor messages and their compensation in reduced-pass GEMM rows.
//
// A two-pass message GEMM (MLBP_GEMM_A_HI_ONLY) computes  D[r, :] = alpha * A_hi[r, :] . B'  and drops  A_lo[r, :] . B'.
// For the bulk of a message that rounding averages away in the contraction; for an element that carries a visible share of
// the mass (a "spike": the word a history feature points at, say) it does not.
//
// K4a, spike_scan_kernel: one CTA per A row of a GEMM block, run right before the block's GEMM.  It streams the row's hi
// half (16-byte loads), collects the elements above the limit, and files up to MLBP_SPIKE_SLOTS of them as
// (column, A_lo value) in ascending column order -- the lo half is exactly what the dropped pass would have multiplied; rows
// with spikes are appended to the block's list; a row with more spikes than slots raises the PEAK word (the gated GEMMs then
// keep the lo half).  (Round 2 first recorded spikes inside the var->factor kernel; the extra code in its hot loop cost
// that kernel 25 % -- more than this separate pass over the hi halves, 8.7 GB per micro-batch of C3, costs.)
//
// K4b, spike_correct_kernel: the dropped term of the recorded spikes is restored exactly:
//     D[r, n] += alpha * sum_s lo_s * B[n, col_s]
// B[n, col] for all n is row `col` of the TRANSPOSED table's plane pair (K2 stores both orientations K-major), so a spike costs
// one contiguous row read (hi + lo) and the row's spikes share one read-modify-write of D[r, :].  Reference semantics are
// untouched: this is arithmetic the float64 dgemv of LBP.py:509 / :518 does implicitly.
//
// One launch per (level, table) GEMM block, after the block's gated launches.  The var->factor kernel keeps one list of spiky
// rows per GEMM block, so a launch touches only its own rows; persistent CTAs walk that list (its length is only known on the
// device); the spikes of a row are applied in ascending column order (deterministic).  The kernel returns at once when the
// gate word is set: then the block ran with all its passes and nothing was dropped.
#include "common.cuh"

namespace mlbp {

constexpr int SP_THREADS = 256;
constexpr int SCAN_THREADS = 128;
constexpr int SCAN_MAX = 64;          // spikes of one row collected before the overflow is certain anyway

__global__ void __launch_bounds__(SCAN_THREADS)
spike_scan_kernel(const __half *__restrict__ A_hi, const __half *__restrict__ A_lo, int ldv, int V, int a0, int n_rows,
                  float limit, int32_t *__restrict__ words, int32_t *__restrict__ cnt, int2 *__restrict__ entries,
                  int32_t *__restrict__ block_rows, int32_t *__restrict__ block_n) {
    __shared__ int s_n;
    __shared__ int s_col[SCAN_MAX];
    const __half2 lim2 = __float2half2_rn(limit);
    for (int row = a0 + blockIdx.x; row < a0 + n_rows; row += gridDim.x) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const uint4 *h8 = reinterpret_cast<const uint4 *>(A_hi + (size_t)row * ldv);
        // the padding of an A row is never written (uninitialised): only whole chunks below V, then the tail element-wise
        const int n8 = V >> 3;
        for (int k8 = threadIdx.x; k8 < n8; k8 += SCAN_THREADS) {
            const uint4 v = __ldg(h8 + k8);
            const __half2 *p = reinterpret_cast<const __half2 *>(&v);
            const __half2 m = __hmax2(__hmax2(p[0], p[1]), __hmax2(p[2], p[3]));
            const bool any = __hgt(__hmax(__low2half(m), __high2half(m)), __low2half(lim2));
            if (any) {                                             // rare
                const __half *e = reinterpret_cast<const __half *>(&v);
                for (int i = 0; i < 8; ++i)
                    if (__half2float(e[i]) > limit) {
                        const int k = atomicAdd(&s_n, 1);
                        if (k < SCAN_MAX) s_col[k] = 8 * k8 + i;
                    }
            }
        }
        for (int c = 8 * n8 + threadIdx.x; c < V; c += SCAN_THREADS)
            if (__half2float(A_hi[(size_t)row * ldv + c]) > limit) {
                const int k = atomicAdd(&s_n, 1);
                if (k < SCAN_MAX) s_col[k] = c;
            }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int n = s_n;
            if (n > 0) {
                words[3] = 1;                                      // SPIKE (diagnostics)
                atomicAdd(&words[4], 1);
                if (n > MLBP_SPIKE_SLOTS) {
                    words[0] = 1;                                  // PEAK: reduced-pass rows of this theta keep the lo half
                } else {
                    for (int i = 1; i < n; ++i) {                  // ascending columns: deterministic order downstream
                        const int c = s_col[i];
                        int q = i;
                        while (q > 0 && s_col[q - 1] > c) { s_col[q] = s_col[q - 1]; --q; }
                        s_col[q] = c;
                    }
                    float mx = 0.f;
                    for (int i = 0; i < n; ++i) {
                        const size_t o = (size_t)row * ldv + s_col[i];
                        entries[(size_t)row * MLBP_SPIKE_SLOTS + i] = make_int2(s_col[i], __float_as_int(__half2float(A_lo[o])));
                        mx = fmaxf(mx, __half2float(A_hi[o]));
                    }
                    atomicMax(&words[2], __float_as_int(mx));     // largest spike seen (2^14 * probability; diagnostics)
                    if (block_rows) block_rows[atomicAdd(block_n, 1)] = row;
                }
            }
            cnt[row] = min(n, MLBP_SPIKE_SLOTS + 1);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SP_THREADS)
spike_correct_kernel(const int32_t *__restrict__ words, const int32_t *__restrict__ cnt, const int2 *__restrict__ entries,
                     const int32_t *__restrict__ rows, const int32_t *__restrict__ n_list, int a0, int n_rows, const __half *__restrict__ Bt_hi,
                     const __half *__restrict__ Bt_lo, int V, int ldv, float *__restrict__ D, int64_t d_row0, int ldd,
                     float alpha, const __half *__restrict__ A_hi, const __half *__restrict__ Rt_hi, float tbar) {
    if (words[0] != 0) return;                                     // three passes ran: nothing to restore
    __shared__ int s_col[MLBP_SPIKE_SLOTS];
    __shared__ float s_lo[MLBP_SPIKE_SLOTS], s_hi[MLBP_SPIKE_SLOTS];
    const int total = min(*n_list, n_rows);
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
        const int row = rows[i];
        if (row < a0 || row >= a0 + n_rows) continue;              // (cannot happen: the list belongs to this block)
        const int n = min(cnt[row], MLBP_SPIKE_SLOTS);
        __syncthreads();
        if (threadIdx.x < n) {                                     // entries are filed in ascending column order (spike_scan_kernel)
            const int2 e = entries[(size_t)row * MLBP_SPIKE_SLOTS + threadIdx.x];
            s_col[threadIdx.x] = e.x; s_lo[threadIdx.x] = __int_as_float(e.y);
            // one-pass rows (A_hi . B_hi) also dropped  A_hi . B_lo : restored at the spikes with the hi value as weight of B_lo
            s_hi[threadIdx.x] = A_hi ? __half2float(A_hi[(size_t)row * ldv + e.x]) : 0.f;
        }
        __syncthreads();
        float *drow = D + (d_row0 + (int64_t)(row - a0)) * (int64_t)ldd;
        // rows are padded to a multiple of 64 elements: whole chunks of 8 (the padding of the planes is zero)
        // two 8-element chunks per thread and iteration: all loads of both are issued before the first store
        for (int c8 = threadIdx.x; c8 < (ldv >> 3); c8 += 2 * SP_THREADS) {
            const int cc[2] = {c8, c8 + SP_THREADS};
            const bool on1 = cc[1] < (ldv >> 3);
            float acc[2][8];
            float4 x[2], y[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[u][q] = 0.f;
                if (u == 0 || on1) {
                    const float4 *d4 = reinterpret_cast<const float4 *>(drow + 8 * (size_t)cc[u]);
                    x[u] = d4[0]; y[u] = d4[1];
                }
            }
            for (int s = 0; s < n; ++s) {
                const float w = s_lo[s], wl = w + s_hi[s];          // weights of B_hi and of B_lo
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u == 1 && !on1) continue;
                    const size_t o = (size_t)s_col[s] * ldv + 8 * (size_t)cc[u];
                    const uint4 h = __ldg(reinterpret_cast<const uint4 *>(Bt_hi + o)), l = __ldg(reinterpret_cast<const uint4 *>(Bt_lo + o));
                    const __half2 *hh = reinterpret_cast<const __half2 *>(&h), *ll = reinterpret_cast<const __half2 *>(&l);
                    if (Rt_hi) {
                        // residual-plane product: the GEMM held  hi_s * fp16(T - tbar)  of this element (+ the constant tbar * sum);
                        // exact is (hi_s + lo_s) * (T - tbar):  add  (hi_s + lo_s) * (T - tbar) - hi_s * R_hi
                        const uint4 rr = __ldg(reinterpret_cast<const uint4 *>(Rt_hi + o));
                        const __half2 *rh = reinterpret_cast<const __half2 *>(&rr);
                        const float hs = s_hi[s];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 a = __half22float2(hh[q]), b = __half22float2(ll[q]), r2 = __half22float2(rh[q]);
                            acc[u][2 * q] = fmaf(wl, (a.x - tbar) + b.x, fmaf(-hs, r2.x, acc[u][2 * q]));
                            acc[u][2 * q + 1] = fmaf(wl, (a.y - tbar) + b.y, fmaf(-hs, r2.y, acc[u][2 * q + 1]));
                        }
                        continue;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 a = __half22float2(hh[q]), b = __half22float2(ll[q]);
                        acc[u][2 * q] = fmaf(w, a.x, fmaf(wl, b.x, acc[u][2 * q]));
                        acc[u][2 * q + 1] = fmaf(w, a.y, fmaf(wl, b.y, acc[u][2 * q + 1]));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !on1) continue;
                const int e = 8 * cc[u];
                x[u].x += e + 0 < V ? alpha * acc[u][0] : 0.f; x[u].y += e + 1 < V ? alpha * acc[u][1] : 0.f;
                x[u].z += e + 2 < V ? alpha * acc[u][2] : 0.f; x[u].w += e + 3 < V ? alpha * acc[u][3] : 0.f;
                y[u].x += e + 4 < V ? alpha * acc[u][4] : 0.f; y[u].y += e + 5 < V ? alpha * acc[u][5] : 0.f;
                y[u].z += e + 6 < V ? alpha * acc[u][6] : 0.f; y[u].w += e + 7 < V ? alpha * acc[u][7] : 0.f;
                float4 *d4 = reinterpret_cast<float4 *>(drow + 8 * (size_t)cc[u]);
                d4[0] = x[u]; d4[1] = y[u];
            }
        }
    }
}

}  // namespace mlbp

using namespace mlbp;

extern "C" int mlbp_spike_scan(const void *A_hi, const void *A_lo, int ldv, int V, int a_row0, int n_rows, float spike_prob,
                               int32_t *spike_words, int32_t *spike_cnt, int32_t *spike_entries, int32_t *block_rows,
                               int32_t *block_n, void *stream) {
    if (n_rows == 0) return MLBP_OK;
    MLBP_CHECK_ARG(A_hi && A_lo && spike_words && spike_cnt && spike_entries && n_rows > 0 && a_row0 >= 0 && spike_prob > 0.f &&
                   (!block_rows || block_n), "spike_scan: bad argument");
    MLBP_CHECK_ARG((ldv % 64) == 0 && ldv >= V && (reinterpret_cast<uintptr_t>(A_hi) % 16) == 0,
                   "spike_scan: rows must be 16-byte aligned and padded to a multiple of 64");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device());
    const int grid = n_rows < 64 * sms ? n_rows : 64 * sms;
    spike_scan_kernel<<<grid, SCAN_THREADS, 0, as_stream(stream)>>>((const __half *)A_hi, (const __half *)A_lo, ldv, V, a_row0, n_rows,
                                                                    ldexpf(spike_prob, MLBP_A_SCALE_LOG2), spike_words, spike_cnt,
                                                                    reinterpret_cast<int2 *>(spike_entries), block_rows, block_n);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_spike_correct(const int32_t *spike_words, const int32_t *spike_cnt, const int32_t *spike_entries,
                                  const int32_t *block_rows, const int32_t *block_n, int a_row0, int n_rows, const void *Bt_hi, const void *Bt_lo,
                                  int V, int ldv, float *D, int64_t d_row0, int ldd, float alpha, const void *A_hi_one_pass,
                                  const void *Rt_hi, float tbar, void *stream) {
    if (n_rows == 0) return MLBP_OK;
    MLBP_CHECK_ARG(spike_words && spike_cnt && spike_entries && block_rows && block_n && Bt_hi && Bt_lo && D && n_rows > 0 && a_row0 >= 0,
                   "spike_correct: bad argument");
    MLBP_CHECK_ARG((ldv % 64) == 0 && ldv >= V && (ldd % 64) == 0 && ldd >= V &&
                   ((reinterpret_cast<uintptr_t>(Bt_hi) | reinterpret_cast<uintptr_t>(Bt_lo) | reinterpret_cast<uintptr_t>(D)) % 16) == 0,
                   "spike_correct: rows must be 16-byte aligned and padded to a multiple of 64");
    MLBP_CHECK_ARG(!Rt_hi || (A_hi_one_pass && (reinterpret_cast<uintptr_t>(Rt_hi) % 16) == 0),
                   "spike_correct: a residual plane goes with one-pass rows (A_hi) and must be 16-byte aligned");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device());
    const int grid = n_rows < 8 * sms ? n_rows : 8 * sms;
    spike_correct_kernel<<<grid, SP_THREADS, 0, as_stream(stream)>>>(spike_words, spike_cnt, reinterpret_cast<const int2 *>(spike_entries),
                                                                    block_rows, block_n, a_row0, n_rows, (const __half *)Bt_hi,
                                                                    (const __half *)Bt_lo, V, ldv, D, d_row0, ldd, alpha,
                                                                    (const __half *)A_hi_one_pass, (const __half *)Rt_hi, tbar);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}
