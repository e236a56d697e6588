// K2: everything that depends on theta only.  The reference rebuilds exp(phi . theta) over the full
// (V,V,3) and (V,Vd,6) tensors for EVERY sentence (train.py:218-253); here it is one pass per SGD step that
// emits the tensor-core operand planes (fp16 hi/lo, both orientations) plus the small per-column /
// per-German-word statistics that make every unary factor O(1) in the gradient.
#include "common.cuh"

namespace mlbp {

constexpr int TS = 64;  // tile edge

struct ThetaEE { double pmi, w1, bias; };
struct ThetaED { double t[6]; };

// One 64x64 tile of the (a, b) plane per CTA, 256 threads: thread (tx = tid % 64, ty = tid / 64) walks rows
// ty, ty+4, ... so global reads and the direct-orientation writes are coalesced along b; the transposed planes
// go through shared memory and are written coalesced along a.
__global__ void __launch_bounds__(256)
build_pairwise_tables_kernel(const float *__restrict__ pmi, const float *__restrict__ w1, int V, int ldf, ThetaEE th,
                             int scale_exp, __half *__restrict__ planes, int64_t ps, int ldv,
                             double *__restrict__ colsums, int with_grad) {
    __shared__ __half sT[4][TS][TS + 2];  // T.hi T.lo T1.hi T1.lo, indexed [a][b]
    __shared__ double sSum[5][4][TS];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int a0 = blockIdx.y * TS, b0 = blockIdx.x * TS;
    const int b = b0 + tx;
    double cs[5] = {0, 0, 0, 0, 0};
    for (int r = ty; r < TS; r += 4) {
        const int a = a0 + r;
        double rs_t = 0.0, rs_t1 = 0.0;
        __half h[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) h[i] = __float2half_rn(0.f);
        if (a < V && b < V) {
            const double p = (double)pmi[(size_t)a * ldf + b];
            const double w = (double)w1[(size_t)a * ldf + b];
            const double z = th.pmi * p + th.bias;
            const double t = exp(z), t1 = exp(z + th.w1 * w);
            const double g = t * p, g1 = t1 * p, g1w = t1 * w;
            cs[0] += t; cs[1] += t1; cs[2] += g; cs[3] += g1; cs[4] += g1w;
            rs_t = t; rs_t1 = t1;
            const double v[5] = {t, t1, g, g1, g1w};
#pragma unroll
            for (int i = 0; i < 5; ++i) split_f16((float)ldexp(v[i], scale_exp), h[2 * i], h[2 * i + 1]);
            const size_t o = (size_t)a * ldv + b;
            planes[0 * ps + o] = h[0];  planes[1 * ps + o] = h[1];      // T
            planes[4 * ps + o] = h[2];  planes[5 * ps + o] = h[3];      // T1
            if (with_grad) {
                planes[8 * ps + o] = h[4];   planes[9 * ps + o] = h[5];   // G   = T  o PMI
                planes[10 * ps + o] = h[6];  planes[11 * ps + o] = h[7];  // G1  = T1 o PMI
                planes[12 * ps + o] = h[8];  planes[13 * ps + o] = h[9];  // G1w = T1 o PMI_w1
            }
        }
        sT[0][r][tx] = h[0]; sT[1][r][tx] = h[1]; sT[2][r][tx] = h[2]; sT[3][r][tx] = h[3];
        // row sums of T and T1 (message of a pairwise factor whose input is still the uniform initial message)
        rs_t = warp_sum(rs_t); rs_t1 = warp_sum(rs_t1);
        if ((threadIdx.x & 31) == 0 && a < V) {
            atomicAdd(&colsums[(size_t)5 * V + a], rs_t);
            atomicAdd(&colsums[(size_t)6 * V + a], rs_t1);
        }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) sSum[i][ty][tx] = cs[i];
    __syncthreads();
    // transposed planes: Tt[b][a] = T[a][b]; threads now run along a
    for (int r = ty; r < TS; r += 4) {
        const int bb = b0 + r, aa = a0 + tx;
        if (bb < V && aa < V) {
            const size_t o = (size_t)bb * ldv + aa;
            planes[2 * ps + o] = sT[0][tx][r];  planes[3 * ps + o] = sT[1][tx][r];
            planes[6 * ps + o] = sT[2][tx][r];  planes[7 * ps + o] = sT[3][tx][r];
        }
    }
    if (ty == 0 && b < V) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double s = sSum[i][0][tx] + sSum[i][1][tx] + sSum[i][2][tx] + sSum[i][3][tx];
            atomicAdd(&colsums[(size_t)i * V + b], s);
        }
    }
}

// one CTA per German word d: sum_e psi, sum_e psi*ed, sum_e psi*ped over the de-major rows
__global__ void __launch_bounds__(256)
build_unary_tables_kernel(const float *__restrict__ edT, const float *__restrict__ pedT, int V, int ldf, ThetaED th,
                          double *__restrict__ edstats) {
    __shared__ double red[32];
    const int d = blockIdx.x;
    const float *e = edT + (size_t)d * ldf, *p = pedT + (size_t)d * ldf;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const double x = e[i], y = p[i];
        const double psi = exp(th.t[0] * x + th.t[1] * y + th.t[5]);
        s0 += psi; s1 += psi * x; s2 += psi * y;
    }
    s0 = block_sum(s0, red); s1 = block_sum(s1, red); s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        edstats[3 * d + 0] = s0; edstats[3 * d + 1] = s1; edstats[3 * d + 2] = s2;
    }
}

}  // namespace mlbp

using namespace mlbp;

extern "C" int mlbp_build_pairwise_tables(const float *pmi, const float *pmi_w1, int V, int ldf,
                                          const double *h_theta_ee, int scale_exp, void *planes,
                                          int64_t plane_stride, int ldv, double *colsums, int with_grad_planes,
                                          void *stream) {
    MLBP_CHECK_ARG(pmi && pmi_w1 && planes && colsums && h_theta_ee, "build_pairwise_tables: null pointer");
    MLBP_CHECK_ARG(V > 0 && ldf >= V && ldv >= V && (ldv % 64) == 0, "build_pairwise_tables: bad V/ld (%d,%d,%d)", V, ldf, ldv);
    MLBP_CHECK_ARG(plane_stride >= (int64_t)V * ldv, "build_pairwise_tables: plane_stride too small");
    cudaStream_t st = as_stream(stream);
    MLBP_CUDA(cudaMemsetAsync(colsums, 0, sizeof(double) * MLBP_N_SUMS * (size_t)V, st));
    ThetaEE th{h_theta_ee[0], h_theta_ee[1], h_theta_ee[2]};
    dim3 grid((V + TS - 1) / TS, (V + TS - 1) / TS);
    build_pairwise_tables_kernel<<<grid, 256, 0, st>>>(pmi, pmi_w1, V, ldf, th, scale_exp, (__half *)planes,
                                                       plane_stride, ldv, colsums, with_grad_planes);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

extern "C" int mlbp_build_unary_tables(const float *edT, const float *pedT, int V, int Vd, int ldf,
                                       const double *h_theta_ed, double *edstats, void *stream) {
    MLBP_CHECK_ARG(edT && pedT && edstats && h_theta_ed, "build_unary_tables: null pointer");
    MLBP_CHECK_ARG(V > 0 && Vd > 0 && ldf >= V, "build_unary_tables: bad shape");
    ThetaED th;
    for (int i = 0; i < 6; ++i) th.t[i] = h_theta_ed[i];
    build_unary_tables_kernel<<<Vd, 256, 0, as_stream(stream)>>>(edT, pedT, V, ldf, th, edstats);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}
