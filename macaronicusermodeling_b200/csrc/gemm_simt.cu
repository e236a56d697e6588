// CUDA-core cross-check for the tcgen05 message GEMM (impl = 1 of mlbp_factor_to_var_gemm).
// TEST-ONLY: tests compare K4's tensor-core result against this kernel at sizes the CPU oracle cannot reach.
// Same operands (fp16 hi/lo planes), exact products, float64 accumulation.
#include "common.cuh"

namespace mlbp {

constexpr int ST = 64, SK = 16;

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const __half *__restrict__ A_hi, const __half *__restrict__ A_lo, int64_t a_rows_total, int a_row0,
                 int n_rows, const __half *__restrict__ B_hi, const __half *__restrict__ B_lo, int V, int ldv,
                 float *__restrict__ D, int64_t d_row0, int ldd, float alpha, int a_terms, int b_terms) {
    __shared__ float sA[SK][ST + 1], sB[SK][ST + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < V; k0 += SK) {
        for (int i = threadIdx.x; i < ST * SK; i += 256) {
            const int r = i / SK, kk = i % SK, k = k0 + kk;
            const int64_t ar = (int64_t)a_row0 + m0 + r;
            float a = 0.f, b = 0.f;
            if (m0 + r < n_rows && ar < a_rows_total && k < V)
                a = __half2float(A_hi[ar * ldv + k]) + (a_terms == 2 ? __half2float(A_lo[ar * ldv + k]) : 0.f);
            if (n0 + r < V && k < V)
                b = __half2float(B_hi[(size_t)(n0 + r) * ldv + k]) + (b_terms == 2 ? __half2float(B_lo[(size_t)(n0 + r) * ldv + k]) : 0.f);
            sA[kk][r] = a;
            sB[kk][r] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < n_rows && n < V) D[(d_row0 + m) * ldd + n] = (float)((double)alpha * acc[i][j]);
        }
}

int launch_gemm_simt(const void *A_hi, const void *A_lo, int64_t a_rows_total, int a_row0, int n_rows,
                     const void *B_hi, const void *B_lo, int V, int ldv, float *D, int64_t d_row0, int ldd, float alpha,
                     int a_terms, int b_terms, cudaStream_t st) {
    dim3 grid((V + ST - 1) / ST, (n_rows + ST - 1) / ST);
    gemm_simt_kernel<<<grid, 256, 0, st>>>((const __half *)A_hi, (const __half *)A_lo, a_rows_total, a_row0, n_rows,
                                           (const __half *)B_hi, (const __half *)B_lo, V, ldv, D, d_row0, ldd, alpha, a_terms, b_terms);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

}  // namespace mlbp
