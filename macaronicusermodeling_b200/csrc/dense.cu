// Device versions of the dense array_utils/c_array_utils.pyx helpers, float64 like the reference.
// They back the drop-in `array_utils.c_array_utils` module; the batched engine never calls them (it fuses
// the same arithmetic into K1..K6), so these are written for fidelity to the reference contract, one
// launch per call, not for roofline.
#include "common.cuh"

namespace mlbp {

__global__ void pointwise_multiply_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                          double *__restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a[i] * b[i];
}

__global__ void __launch_bounds__(256) sum_kernel(const double *__restrict__ a, int64_t n, double *__restrict__ sum) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += a[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(sum, s);
}

// pyx:29-40: s > 0 ? m1 / s : zero-fill
__global__ void normalize_apply_kernel(const double *__restrict__ a, double *__restrict__ out, int64_t n,
                                       const double *__restrict__ sum) {
    const double s = *sum;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = s > 0.0 ? a[i] / s : 0.0;
}

// out[m,n] = A[m,k] . B[k,n], row-major.  One warp per output row block; covers GEMV (n == 1), vector-matrix
// (m == 1) and the outer product (k == 1) the reference uses (LBP.py:509, :518, :566).
__global__ void __launch_bounds__(256)
dense_dot_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C, int m, int k,
                 int n) {
    if (n == 1) {  // GEMV: one warp per row, coalesced along k
        const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        if (row >= m) return;
        double s = 0.0;
        for (int j = lane; j < k; j += 32) s += A[(size_t)row * k + j] * B[j];
        s = warp_sum(s);
        if (lane == 0) C[row] = s;
        return;
    }
    // general: thread per output element, coalesced along n
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)m * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    double s = 0.0;
    for (int t = 0; t < k; ++t) s += A[(size_t)i * k + t] * B[(size_t)t * n + j];
    C[idx] = s;
}

static int grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace mlbp

using namespace mlbp;

extern "C" int mlbp_pointwise_multiply_f64(const double *m1, const double *m2, double *out, int64_t n, void *stream) {
    if (n == 0) return MLBP_OK;
    MLBP_CHECK_ARG(m1 && m2 && out && n > 0, "pointwise_multiply: bad argument");
    pointwise_multiply_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(m1, m2, out, n);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_dense_pointwise_multiply_f64(const double *m1, const double *m2, double *out, int64_t n,
                                                 void *stream) {
    return mlbp_pointwise_multiply_f64(m1, m2, out, n, stream);
}

extern "C" int mlbp_normalize_f64(const double *m1, double *out, int64_t n, double *d_sum, void *stream) {
    if (n == 0) return MLBP_OK;
    MLBP_CHECK_ARG(m1 && out && d_sum && n > 0, "normalize: bad argument");
    cudaStream_t st = as_stream(stream);
    MLBP_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), st));
    sum_kernel<<<grid_for(n), 256, 0, st>>>(m1, n, d_sum);
    MLBP_LAUNCH_CHECK();
    normalize_apply_kernel<<<grid_for(n), 256, 0, st>>>(m1, out, n, d_sum);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_dense_dot_f64(const double *m1, const double *m2, double *out, int m, int k, int n,
                                  void *stream) {
    MLBP_CHECK_ARG(m1 && m2 && out && m > 0 && k > 0 && n > 0, "dense_dot: bad argument");
    const int64_t work = (n == 1) ? (int64_t)m * 32 : (int64_t)m * n;
    const int64_t blocks = (work + 255) / 256;
    MLBP_CHECK_ARG(blocks < (1ll << 31), "dense_dot: too large");
    dense_dot_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(m1, m2, out, m, k, n);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}
