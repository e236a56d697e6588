// K5c: the K most probable words of a belief row, best first, ON THE DEVICE.
//
// The reference lists them per variable with np.argpartition + np.argsort on the host marginal (VariableNode.get_max_vocab,
// LBP.py:402-411; K = 50 for to_string / get_precision_counts, LBP.py:87, :115).  Reading a whole belief row back costs
// 4 V bytes per variable; this kernel leaves K (index, probability) pairs instead.
//
// One CTA per row.  Non-negative floats order like their bit patterns, so the K-th largest value is found by a 4-pass radix
// select (8 bits per pass, shared-memory histogram); then every entry above the threshold is collected, and of the entries
// EQUAL to it the ones with the smallest indices (chunks are walked in ascending index order with a block scan, so the choice
// is deterministic), and the K keys (value bits, ~index) are sorted in shared memory by a bitonic network: descending value,
// ascending index among equal values.  NumPy's order among exactly equal values is an implementation detail of its
// introselect; callers that must reproduce it bit for bit (LBP.py's dumps) fall back to the host calls when the device list
// reports a tie (n_ties > 0).
#include "common.cuh"

namespace mlbp {

constexpr int TK_THREADS = 256;
constexpr int TK_MAX = 1024;          // largest K (keys are sorted in shared memory)

__global__ void __launch_bounds__(TK_THREADS)
topk_rows_kernel(const float *__restrict__ X, int ldx, int V, int K, int32_t *__restrict__ idx, float *__restrict__ val,
                 int32_t *__restrict__ n_ties) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_need, s_n, s_eq_base;
    __shared__ unsigned s_wsum[TK_THREADS / 32];
    __shared__ unsigned long long keys[TK_MAX];
    const float *x = X + (size_t)blockIdx.x * ldx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // negative or NaN entries cannot occur in a belief; map them to zero so that the bit order stays a value order
    auto bits_of = [&](int e) { const float v = __ldg(x + e); return v > 0.f ? __float_as_uint(v) : 0u; };
    if (threadIdx.x == 0) { s_prefix = 0u; s_need = (unsigned)K; s_n = 0u; s_eq_base = 0u; }
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int e = threadIdx.x; e < V; e += TK_THREADS) {
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
    const unsigned thr = s_prefix, need = s_need;     // K-th largest bit pattern; `need` of the entries equal to it are listed
    for (int i = threadIdx.x; i < TK_MAX; i += TK_THREADS) keys[i] = 0ull;
    __syncthreads();
    // entries above the threshold: any order (sorted below)
    for (int e = threadIdx.x; e < V; e += TK_THREADS) {
        const unsigned b = bits_of(e);
        if (b > thr) keys[atomicAdd(&s_n, 1u)] = ((unsigned long long)b << 32) | (unsigned)(0x7fffffff - e);
    }
    __syncthreads();
    const unsigned n_gt = s_n;
    // entries equal to the threshold: the `need` smallest indices (ascending chunks, ballot + block scan inside a chunk)
    unsigned n_eq_total = 0;
    for (int e0 = 0; e0 < V; e0 += TK_THREADS) {
        const int e = e0 + threadIdx.x;
        const bool eq = e < V && bits_of(e) == thr;
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_wsum[warp] = __popc(m);
        __syncthreads();
        unsigned before = s_eq_base, total = 0;
        for (int w = 0; w < TK_THREADS / 32; ++w) {
            if (w < warp) before += s_wsum[w];
            total += s_wsum[w];
        }
        const unsigned r = before + __popc(m & ((1u << lane) - 1u));
        if (eq && r < need) keys[n_gt + r] = ((unsigned long long)thr << 32) | (unsigned)(0x7fffffff - e);
        n_eq_total += total;
        __syncthreads();
        if (threadIdx.x == 0) s_eq_base += total;
        __syncthreads();
    }
    // bitonic sort of the (power-of-two padded) key list, descending; padding keys are zero and sink to the end
    int n2 = 1;
    while (n2 < K) n2 <<= 1;
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n2; i += TK_THREADS) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const unsigned long long a = keys[i], c = keys[j];
                    if (desc ? a < c : a > c) { keys[i] = c; keys[j] = a; }
                }
            }
            __syncthreads();
        }
    int ties = 0;
    for (int i = threadIdx.x; i < K; i += TK_THREADS) {
        const unsigned long long kv = keys[i];
        idx[(size_t)blockIdx.x * K + i] = 0x7fffffff - (int)(unsigned)(kv & 0xffffffffull);
        val[(size_t)blockIdx.x * K + i] = __uint_as_float((unsigned)(kv >> 32));
        if (i + 1 < K && (unsigned)(keys[i + 1] >> 32) == (unsigned)(kv >> 32)) ++ties;   // equal neighbours inside the list
    }
    if (n_ties) {
        __shared__ double red[32];
        const double t = block_sum((double)ties, red);
        // ... and a tie ACROSS the boundary: more entries equal the K-th value than were listed
        if (threadIdx.x == 0) n_ties[blockIdx.x] = (int)(t + 0.5) + (n_eq_total > need ? 1 : 0);
    }
}

}  // namespace mlbp

using namespace mlbp;

extern "C" int mlbp_topk_rows(const float *X, int ldx, int V, int n_rows, int K, int32_t *idx, float *val, int32_t *n_ties,
                              void *stream) {
    if (n_rows == 0) return MLBP_OK;
    MLBP_CHECK_ARG(X && idx && val && n_rows > 0 && V > 0 && ldx >= V, "topk_rows: bad argument");
    MLBP_CHECK_ARG(K > 0 && K <= V && K <= TK_MAX, "topk_rows: K must lie in [1, min(V, 1024)]");
    topk_rows_kernel<<<n_rows, TK_THREADS, 0, as_stream(stream)>>>(X, ldx, V, K, idx, val, n_ties);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}
