// K4: pairwise factor -> variable messages as a split-precision tcgen05 GEMM (sm_100a).
//
// Reference: FactorNode.update_message_to, LBP.py:499-526 -- one float64 dgemv  T.m  or  m'.T  per pairwise factor
// and direction (au.dense_dot, c_array_utils.pyx:90-91).  All messages of one schedule level that share a
// potential table are stacked into the rows of A, so the level becomes  D = alpha * A . B^T  with B one of the
// table planes of K2 (K-major for both orientations because K2 also stores the transposed tables).
//
// Precision: the reference is float64 and parity needs bit-exact top-1 and 1e-4 beliefs, so a single fp16/bf16
// rounding of either operand is not enough.  Both operands are stored as fp16 hi + lo
// (22 significant bits) and the product is issued as three MMAs  hi*hi + hi*lo + lo*hi  into one fp32 TMEM
// accumulator (the dropped lo*lo term is 2^-22 relative).  Roofline flops are counted once (2*M*N*K).
//
// Accumulation: measured on B200, the tensor core aligns every product to the accumulator's exponent and
// truncates, so a K-long accumulation in TMEM is biased by about (K/4) * 2^-24 relative (1.8e-4 at K = 10^4) with
// an element-dependent part of the same order -- enough to flip near-tied arg-maxes.  The accumulator is therefore
// kept SHORT: every CHUNK_KB k-blocks (CHUNK_KB * 64 products) the MMA issuer switches to the other of two TMEM
// accumulators and the epilogue warps drain the finished one into fp32 registers (round-to-nearest adds), which
// brings the error back to ~1e-6 relative.
//
// Two kernels share this structure: gemm_split_f16_pair_kernel (further down; a CTA PAIR per 256 x 256 tile with
// tcgen05.mma.cta_group::2 -- the product path for V > 2048) and the one-CTA kernel described here (small V, probes).
//
// Structure (one 128 x BN output tile per CTA, K = V streamed in 64-wide slabs):
//   warp 0 lane 0 : TMA producer  -- 4 cp.async.bulk.tensor loads per stage (A.hi A.lo B.hi B.lo, swizzle-128B)
//   warp 1 lane 0 : MMA issuer    -- 3 x (BK/16) tcgen05.mma.cta_group::1.kind::f16 per stage, tcgen05.commit
//   warp 2        : TMEM allocate / free (2 x BN fp32 columns)
//   warps 4..     : epilogue      -- one warp per (32 TMEM lanes, 128 columns): tcgen05.ld 32x32b.x32 per chunk into
//                                    128 register accumulators, finally scale by alpha and 16-byte stores
// full/empty mbarriers ring the shared-memory stages; tmem_full/tmem_empty mbarriers ring the two accumulators.
// CTAs are rasterised 16 M-tiles deep so that co-resident CTAs share A and B slabs in L2.
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace mlbp {

int launch_gemm_simt(const void *, const void *, int64_t, int, int, const void *, const void *, int, int, float *,
                     int64_t, int, float, int, int, cudaStream_t);

constexpr int BM = 128;
constexpr int UMMA_K = 16;
constexpr int GROUP_M = 16;
constexpr unsigned long long WAIT_LIMIT_CYCLES = 8000000000ull;  // ~4 s: a stuck barrier traps instead of hanging the GPU

// BK = 64 fp16 = 128-byte rows (swizzle-128B) or BK = 32 = 64-byte rows (swizzle-64B: twice the stages per KB)
template <int BN, int STAGES, int BK = 64>
struct Cfg {
    static_assert(BK == 64 || BK == 32, "BK");
    static constexpr int EPI_WARPS = 4 * (BN / 128);          // one warp per (lane quarter, 128-column half)
    static constexpr int THREADS = 128 + 32 * EPI_WARPS;
    static constexpr int TMEM_COLS = 2 * BN;                   // two accumulators
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + alignment slack
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(BN == 128 || BN == 256, "UMMA N");
};

// ------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int *err_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > WAIT_LIMIT_CYCLES) {
            if (err_flag) atomicExch(err_flag, code);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, swizzle-128B shared-memory matrix descriptor: rows are 128 bytes, 8-row core-matrix groups are
// 1024 bytes apart (SBO), LBO unused for a single swizzle atom along K, descriptor version 1 (sm_100).
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
    constexpr uint64_t row_bytes = BK * 2;                // 128 (SWIZZLE_128B = 2) or 64 (SWIZZLE_64B = 4)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                               // leading byte offset (ignored)
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;          // stride byte offset: 8-row core-matrix group, bits [32,46)
    d |= (uint64_t)1 << 46;                               // version = 1
    d |= (uint64_t)(BK == 64 ? 2 : 4) << 61;              // layout type
    return d;
}
// kind::f16 instruction descriptor: fp16 A/B (format 0), fp32 accumulator, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------- kernel
template <int BN, int STAGES, int CHUNK_KB, int BK>
__global__ void __launch_bounds__(Cfg<BN, STAGES, BK>::THREADS, 1)
gemm_split_f16_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                      float *__restrict__ D, int64_t d_row0, int ldd, int a_row0, int n_rows, int V, float alpha, float add_const,
                      int m_tiles, int n_tiles, int a_terms, int b_terms, const int32_t *__restrict__ gate, int run_if_set,
                      int *err_flag) {
    using C = Cfg<BN, STAGES, BK>;
    // device-side launch decision (mlbp_factor_to_var_gemm_gated): the whole grid returns unless the gate matches
    if (gate != nullptr && ((*reinterpret_cast<const volatile int32_t *>(gate) != 0) != (run_if_set != 0))) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * C::STAGE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * STAGES + 2 + b); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // rasterisation: GROUP_M m-tiles deep, then along n
    const int per_group = GROUP_M * n_tiles;
    const int grp = blockIdx.x / per_group, in_grp = blockIdx.x % per_group;
    const int first_m = grp * GROUP_M;
    const int gsize = min(GROUP_M, m_tiles - first_m);
    const int m_blk = first_m + in_grp % gsize, n_blk = in_grp / gsize;
    const int num_kb = (V + BK - 1) / BK;
    const int num_chunks = (num_kb + CHUNK_KB - 1) / CHUNK_KB;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
        tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), C::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            const int row_a = a_row0 + m_blk * BM, row_b = n_blk * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u, err_flag, 1);
                // a_terms == 1 (gradient rows): A.lo is neither loaded nor multiplied
                mbar_expect_tx(full_bar(s), (uint32_t)(C::STAGE_BYTES - (a_terms == 2 ? 0 : C::A_BYTES) - (b_terms == 2 ? 0 : C::B_BYTES)));
                const uint32_t st = smem_base + (uint32_t)s * C::STAGE_BYTES;
                tma_load_2d(&tm_a_hi, full_bar(s), st, kb * BK, row_a);
                if (a_terms == 2) tma_load_2d(&tm_a_lo, full_bar(s), st + C::A_BYTES, kb * BK, row_a);
                tma_load_2d(&tm_b_hi, full_bar(s), st + 2 * C::A_BYTES, kb * BK, row_b);
                if (b_terms == 2) tma_load_2d(&tm_b_lo, full_bar(s), st + 2 * C::A_BYTES + C::B_BYTES, kb * BK, row_b);
            }
        } else if (warp == 1 && lane == 0) {
            // ===================== MMA issuer =====================
            // the last N tile is usually narrow (V = 10 000: 16 of 256 columns): issue it with the matching UMMA N so that
            // it costs 1/16 of a full tile instead of multiplying 240 zero-filled rows of B
            const int n_valid = min(BN, V - n_blk * BN);
            const uint32_t idesc = umma_idesc_f16(BM, (n_valid + 15) & ~15);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                const int chunk = kb / CHUNK_KB, in_chunk = kb % CHUNK_KB;
                const int buf = chunk & 1;
                if (in_chunk == 0) {                              // accumulator must have been drained
                    mbar_wait(tempty_bar(buf), ((uint32_t)(chunk >> 1) & 1u) ^ 1u, err_flag, 4);
                    tc_fence_after();
                }
                mbar_wait(full_bar(s), ph, err_flag, 2);
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                const uint32_t st = smem_base + (uint32_t)s * C::STAGE_BYTES;
                const uint64_t a_hi = umma_desc_kmajor<BK>(st), a_lo = umma_desc_kmajor<BK>(st + C::A_BYTES);
                const uint64_t b_hi = umma_desc_kmajor<BK>(st + 2 * C::A_BYTES);
                const uint64_t b_lo = umma_desc_kmajor<BK>(st + 2 * C::A_BYTES + C::B_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);   // 32 bytes per K step inside the atom
                    tc_mma_f16(acc, a_hi + adv, b_hi + adv, idesc, (in_chunk | k) != 0 ? 1u : 0u);
                    if (b_terms == 2) tc_mma_f16(acc, a_hi + adv, b_lo + adv, idesc, 1u);
                    if (a_terms == 2) tc_mma_f16(acc, a_lo + adv, b_hi + adv, idesc, 1u);
                }
                tc_commit(empty_bar(s));                          // frees the stage when the MMAs above have read it
                if (in_chunk == CHUNK_KB - 1 || kb == num_kb - 1) tc_commit(tfull_bar(buf));   // chunk accumulator complete
            }
        }
    } else {
        // ===================== epilogue: drain every chunk into fp32 registers =====================
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;                         // which 128-column half of the tile
        float acc[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) acc[i] = 0.f;
        for (int c = 0; c < num_chunks; ++c) {
            const int buf = c & 1;
            mbar_wait(tfull_bar(buf), (uint32_t)(c >> 1) & 1u, err_flag, 3);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * 128);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t r[32];
                tmem_ld_32x32(t0 + (uint32_t)(j * 32), r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[j * 32 + i] += __uint_as_float(r[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        const int m = m_blk * BM + q * 32 + lane;
        if (m < n_rows) {
            float *drow = D + (d_row0 + (int64_t)m) * (int64_t)ldd;
            const float c0 = add_const;
            const int n0 = n_blk * BN + half * 128;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int nc = n0 + 4 * j;
                if (nc < ldd) {                                   // ldd is a multiple of 64: a float4 is all-in or all-out
                    float4 o;                                     // columns >= V (row padding) are written as zeros
                    o.x = nc + 0 < V ? fmaf(acc[4 * j + 0], alpha, c0) : 0.f; o.y = nc + 1 < V ? fmaf(acc[4 * j + 1], alpha, c0) : 0.f;
                    o.z = nc + 2 < V ? fmaf(acc[4 * j + 2], alpha, c0) : 0.f; o.w = nc + 3 < V ? fmaf(acc[4 * j + 3], alpha, c0) : 0.f;
                    *reinterpret_cast<float4 *>(drow + nc) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------- CTA-pair kernel
// The one-CTA kernel above moves (BM + BN) * BK * 4 bytes from L2 into shared memory per 3 * 2 * BM * BN * BK flops:
// 128 flop/B, i.e. ~9 300 B/clk over 148 SMs at full tensor rate -- above the ~6 300 B/clk the L2 slices deliver, so it
// is L2-bandwidth-bound (75 % tensor-pipe active measured; the two-pass gradient rows gained only 10 %).
// A CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2, M = 256) halves the B traffic: each CTA loads its own 128 rows
// of A and only HALF of the 256-row B tile, and the tensor cores of both SMs read both halves (192 flop/B).  Per CTA a
// stage is 64 KB (A.hi A.lo B.hi/2 B.lo/2), so three stages fit.  Rank 0 (leader) issues the MMAs for the pair; both
// CTAs' TMA loads signal the leader's full barrier (mbarrier address with the peer bit cleared); tcgen05.commit
// multicasts to the empty / tmem-full barriers of both CTAs; both CTAs' epilogue warps drain their own TMEM (their 128
// rows) and arrive on the leader's tmem-empty barrier.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address of the even CTA of a pair

// A_T / B_T: how many fp16 planes of A / B a stage holds (2 = hi + lo, 1 = hi only).  Fewer planes -> smaller stages ->
// more of them: the one-pass rows need ~6 stages, their k-block lasts only ~512 clk against a ~1000 clk TMA round trip.
template <int STAGES, int A_T = 2, int B_T = 2>
struct PairCfg {
    static constexpr int BN = 256, BK = 64;
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 128 + 32 * EPI_WARPS;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int A_BYTES = BM * BK * 2;                // 128 rows of A (this CTA's half of M = 256)
    static constexpr int B_BYTES = (BN / 2) * BK * 2;          // this CTA's half of the B tile
    static constexpr int STAGE_BYTES = A_T * A_BYTES + B_T * B_BYTES;
    static constexpr int B_OFF = A_T * A_BYTES;                 // A.hi [A.lo] B.hi [B.lo]
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(2 * STAGES + 4 <= 24, "barrier block");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *map, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {            // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same A tile feeds two consecutive MMAs of a k-step (hi*hi, hi*lo): the first keeps it in the tensor core's A
// collector buffer (SASS UTCHMMA ... .A_KEEP), the second reads it from there (.A_REUSE) instead of shared memory.
__device__ __forceinline__ void tc_mma_f16_pair_a_fill(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                       uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair_a_lastuse(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar, uint32_t rank) {   // arrive on the leader CTA's barrier
    if (rank == 0) {
        mbar_arrive(bar);
    } else {
        asm volatile(
            "{\n\t.reg .b32 rem;\n\t"
            "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
            "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}"
            ::"r"(bar), "r"(0) : "memory");
    }
}

template <int STAGES, int CHUNK_KB, int A_T, int B_T, bool A_REUSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg<STAGES, A_T, B_T>::THREADS, 1)
gemm_split_f16_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                           const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                           float *__restrict__ D, int64_t d_row0, int ldd, int a_row0, int n_rows, int V, float alpha, float add_const,
                           int m_pairs, int n_tiles, const int32_t *__restrict__ gate, int run_if_set, int kb0, int kb_n,
                           int *err_flag) {
    using C = PairCfg<STAGES, A_T, B_T>;
    // device-side launch decision (mlbp_factor_to_var_gemm_gated): every CTA of every pair reads the same word, written by an
    // earlier kernel of the stream, and returns before any barrier or tensor-memory allocation unless the gate matches
    if (gate != nullptr && ((*reinterpret_cast<const volatile int32_t *>(gate) != 0) != (run_if_set != 0))) return;
    constexpr int a_terms = A_T, b_terms = B_T;
    constexpr int BN = C::BN, BK = C::BK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * C::STAGE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * STAGES + 2 + b); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                      // 0 = leader of the pair
    const int pair = blockIdx.x >> 1;

    // rasterisation: GROUP_M / 2 pairs deep (the same 16 M-tiles as the one-CTA kernel), then along n
    constexpr int GROUP_P = GROUP_M / 2;
    // (Measured and removed in round 2: an L2 prefetch (cp.async.bulk.prefetch.tensor ... L2) of the A / B tiles 4-32 k-blocks
    // ahead of the shared-memory ring, on the theory that the one-pass rows wait for DRAM misses of slabs their wave streams
    // for the first time -- 3 % SLOWER at every distance, alone and in the bench: profiles/r2f_gemm_probe_l2_prefetch.txt.)
    // (Measured and removed in round 2: issuing the narrow last N tiles of all M pairs at the END of the launch instead of
    // inside the raster -- the idea was that equal-length tiles keep co-resident pairs in step -- was +0.8 % in the bench,
    // inside the run-to-run noise: profiles/r2a_bench_c3_ab_narrow_last.txt.)
    const int per_group = GROUP_P * n_tiles;
    const int grp = pair / per_group, in_grp = pair % per_group;
    const int first_p = grp * GROUP_P;
    const int gsize = min(GROUP_P, m_pairs - first_p);
    const int p_blk = first_p + in_grp % gsize;
    const int n_blk = in_grp / gsize;
    const int m_blk = 2 * p_blk + (int)rank;                      // this CTA's 128-row tile
    // K range of this launch: k-blocks [kb0, kb0 + kb_n).  A long K (V = 50 000: 782 k-blocks) is issued as several launches
    // that ADD into D (kb0 > 0): the CTA pairs of a launch drift apart in K and stop sharing A / B slabs in L2; a kernel
    // boundary re-aligns them (same reason as the row slices of a level, see Engine._gemm_slice_rows).
    const int num_kb = kb_n;
    const int num_chunks = (num_kb + CHUNK_KB - 1) / CHUNK_KB;
    const int n_valid = min(BN, V - n_blk * BN);
    const int n_cur = (n_valid + 15) & ~15;                       // UMMA N of this tile (cta_group::2: multiples of 16)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
        tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * C::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    cluster_sync_all();                                           // the peer's barriers exist before anything targets them
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer (both CTAs) =====================
            const int row_a = a_row0 + m_blk * BM;
            const int row_b = n_blk * BN + (int)rank * (n_cur >> 1);      // this CTA's half of the N extent in use
            const uint32_t bytes_cta = (uint32_t)C::STAGE_BYTES;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u, err_flag, 1);
                if (rank == 0) mbar_expect_tx(full_bar(s), 2u * bytes_cta);      // bytes of BOTH CTAs land on the leader
                const uint32_t lb = full_bar(s) & PEER_BIT_MASK;
                const uint32_t st = smem_base + (uint32_t)s * C::STAGE_BYTES;
                const int kc = (kb0 + kb) * BK;
                tma_load_2d_pair(&tm_a_hi, lb, st, kc, row_a);
                if (a_terms == 2) tma_load_2d_pair(&tm_a_lo, lb, st + C::A_BYTES, kc, row_a);
                tma_load_2d_pair(&tm_b_hi, lb, st + C::B_OFF, kc, row_b);
                if (b_terms == 2) tma_load_2d_pair(&tm_b_lo, lb, st + C::B_OFF + C::B_BYTES, kc, row_b);
            }
        } else if (warp == 1 && lane == 0 && rank == 0) {
            // ===================== MMA issuer (leader only, for both SMs) =====================
            const uint32_t idesc = umma_idesc_f16(2 * BM, n_cur);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                const int chunk = kb / CHUNK_KB, in_chunk = kb % CHUNK_KB;
                const int buf = chunk & 1;
                if (in_chunk == 0) {                              // both CTAs have drained this accumulator
                    mbar_wait(tempty_bar(buf), ((uint32_t)(chunk >> 1) & 1u) ^ 1u, err_flag, 4);
                    tc_fence_after();
                }
                mbar_wait(full_bar(s), ph, err_flag, 2);
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                const uint32_t st = smem_base + (uint32_t)s * C::STAGE_BYTES;
                const uint64_t a_hi = umma_desc_kmajor<BK>(st), a_lo = umma_desc_kmajor<BK>(st + C::A_BYTES);
                const uint64_t b_hi = umma_desc_kmajor<BK>(st + C::B_OFF);
                const uint64_t b_lo = umma_desc_kmajor<BK>(st + C::B_OFF + C::B_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                    if (A_REUSE && b_terms == 2) {
                        tc_mma_f16_pair_a_fill(acc, a_hi + adv, b_hi + adv, idesc, (in_chunk | k) != 0 ? 1u : 0u);
                        tc_mma_f16_pair_a_lastuse(acc, a_hi + adv, b_lo + adv, idesc, 1u);
                    } else {
                        tc_mma_f16_pair(acc, a_hi + adv, b_hi + adv, idesc, (in_chunk | k) != 0 ? 1u : 0u);
                        if (b_terms == 2) tc_mma_f16_pair(acc, a_hi + adv, b_lo + adv, idesc, 1u);
                    }
                    if (a_terms == 2) tc_mma_f16_pair(acc, a_lo + adv, b_hi + adv, idesc, 1u);
                }
                tc_commit_pair(empty_bar(s));                     // frees the stage in both CTAs
                if (in_chunk == CHUNK_KB - 1 || kb == num_kb - 1) tc_commit_pair(tfull_bar(buf));
            }
        }
    } else {
        // ===================== epilogue (both CTAs): drain every chunk into fp32 registers =====================
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        float acc[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) acc[i] = 0.f;
        for (int c = 0; c < num_chunks; ++c) {
            const int buf = c & 1;
            mbar_wait(tfull_bar(buf), (uint32_t)(c >> 1) & 1u, err_flag, 3);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * 128);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t r[32];
                tmem_ld_32x32(t0 + (uint32_t)(j * 32), r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[j * 32 + i] += __uint_as_float(r[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty_bar(buf), rank);
        }
        const int m = m_blk * BM + q * 32 + lane;
        if (m < n_rows) {
            float *drow = D + (d_row0 + (int64_t)m) * (int64_t)ldd;
            const float c0 = kb0 > 0 ? 0.f : add_const;       // the constant of a residual-plane product: once per K
            const int n0 = n_blk * BN + half * 128;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int nc = n0 + 4 * j;
                if (nc < ldd) {
                    float4 o;                                     // columns >= V (row padding) are written as zeros
                    o.x = nc + 0 < V ? fmaf(acc[4 * j + 0], alpha, c0) : 0.f; o.y = nc + 1 < V ? fmaf(acc[4 * j + 1], alpha, c0) : 0.f;
                    o.z = nc + 2 < V ? fmaf(acc[4 * j + 2], alpha, c0) : 0.f; o.w = nc + 3 < V ? fmaf(acc[4 * j + 3], alpha, c0) : 0.f;
                    if (kb0 > 0) {                                // a later K range of a split launch: add to what is there
                        const float4 p = *reinterpret_cast<const float4 *>(drow + nc);
                        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                    }
                    *reinterpret_cast<float4 *>(drow + nc) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // neither CTA frees tensor memory the pair still uses
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 2-D fp16 [rows, cols] tensor with row pitch ld elements; box = [box_rows, 64 cols], swizzle-128B, zero OOB fill
static int make_map(CUtensorMap *m, const void *ptr, int64_t rows, int cols, int ld, int box_rows, int bk) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return MLBP_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return MLBP_ERR_CUDA; }
    return MLBP_OK;
}

struct MapKey {
    const void *ptr; int64_t rows; int cols, ld, box, bk;
    bool operator<(const MapKey &o) const {
        return std::tie(ptr, rows, cols, ld, box, bk) < std::tie(o.ptr, o.rows, o.cols, o.ld, o.box, o.bk);
    }
};
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

static int cached_map(CUtensorMap *out, const void *ptr, int64_t rows, int cols, int ld, int box, int bk) {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    MapKey k{ptr, rows, cols, ld, box, bk};
    auto it = g_maps.find(k);
    if (it == g_maps.end()) {
        CUtensorMap m;
        int rc = make_map(&m, ptr, rows, cols, ld, box, bk);
        if (rc != MLBP_OK) return rc;
        if (g_maps.size() > 4096) g_maps.clear();
        it = g_maps.emplace(k, m).first;
    }
    *out = it->second;
    return MLBP_OK;
}

static int *g_err_flag = nullptr;   // pinned, mapped: the kernel records which barrier timed out before trapping

template <int BN, int STAGES, int CHUNK_KB, int BK = 64>
static int launch_tc(const void *A_hi, const void *A_lo, int64_t a_rows_total, int a_row0, int n_rows, const void *B_hi,
                     const void *B_lo, int V, int ldv, float *D, int64_t d_row0, int ldd, float alpha, float add_const, int a_terms,
                     int b_terms, const int32_t *gate, int run_if_set, cudaStream_t st) {
    using C = Cfg<BN, STAGES, BK>;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = cached_map(&ma_hi, A_hi, a_rows_total, V, ldv, BM, BK)) != MLBP_OK) return rc;
    if ((rc = cached_map(&ma_lo, A_lo, a_rows_total, V, ldv, BM, BK)) != MLBP_OK) return rc;
    if ((rc = cached_map(&mb_hi, B_hi, V, V, ldv, BN, BK)) != MLBP_OK) return rc;
    if ((rc = cached_map(&mb_lo, B_lo, V, V, ldv, BN, BK)) != MLBP_OK) return rc;
    static bool attr_set_dev[MLBP_MAX_DEVICES] = {};
    bool &attr_set = attr_set_dev[current_device()];
    if (!attr_set) {
        MLBP_CUDA(cudaFuncSetAttribute(gemm_split_f16_kernel<BN, STAGES, CHUNK_KB, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_BYTES));
        attr_set = true;
    }
    if (!g_err_flag) {
        int *h = nullptr;
        if (cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped) == cudaSuccess) { *h = 0; g_err_flag = h; }
    }
    int *d_flag = nullptr;
    if (g_err_flag) cudaHostGetDevicePointer(&d_flag, g_err_flag, 0);
    const int m_tiles = (n_rows + BM - 1) / BM, n_tiles = (V + BN - 1) / BN;
    gemm_split_f16_kernel<BN, STAGES, CHUNK_KB, BK><<<m_tiles * n_tiles, C::THREADS, C::SMEM_BYTES, st>>>(
        ma_hi, ma_lo, mb_hi, mb_lo, D, d_row0, ldd, a_row0, n_rows, V, alpha, add_const, m_tiles, n_tiles, a_terms, b_terms, gate, run_if_set, d_flag);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

template <int STAGES, int CHUNK_KB, int A_T, int B_T, bool A_REUSE = false>
static int launch_pair_t(const void *A_hi, const void *A_lo, int64_t a_rows_total, int a_row0, int n_rows, const void *B_hi,
                         const void *B_lo, int V, int ldv, float *D, int64_t d_row0, int ldd, float alpha, float add_const,
                         const int32_t *gate, int run_if_set, int kb0, int kb_n, cudaStream_t st) {
    using C = PairCfg<STAGES, A_T, B_T>;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = cached_map(&ma_hi, A_hi, a_rows_total, V, ldv, BM, C::BK)) != MLBP_OK) return rc;
    if ((rc = cached_map(&ma_lo, A_lo, a_rows_total, V, ldv, BM, C::BK)) != MLBP_OK) return rc;
    if ((rc = cached_map(&mb_hi, B_hi, V, V, ldv, C::BN / 2, C::BK)) != MLBP_OK) return rc;
    if ((rc = cached_map(&mb_lo, B_lo, V, V, ldv, C::BN / 2, C::BK)) != MLBP_OK) return rc;
    static bool attr_set_dev[MLBP_MAX_DEVICES] = {};
    bool &attr_set = attr_set_dev[current_device()];
    if (!attr_set) {
        MLBP_CUDA(cudaFuncSetAttribute(gemm_split_f16_pair_kernel<STAGES, CHUNK_KB, A_T, B_T, A_REUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_BYTES));
        attr_set = true;
    }
    if (!g_err_flag) {
        int *h = nullptr;
        if (cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped) == cudaSuccess) { *h = 0; g_err_flag = h; }
    }
    int *d_flag = nullptr;
    if (g_err_flag) cudaHostGetDevicePointer(&d_flag, g_err_flag, 0);
    const int m_pairs = (n_rows + 2 * BM - 1) / (2 * BM), n_tiles = (V + C::BN - 1) / C::BN;
    gemm_split_f16_pair_kernel<STAGES, CHUNK_KB, A_T, B_T, A_REUSE><<<2 * m_pairs * n_tiles, C::THREADS, C::SMEM_BYTES, st>>>(
        ma_hi, ma_lo, mb_hi, mb_lo, D, d_row0, ldd, a_row0, n_rows, V, alpha, add_const, m_pairs, n_tiles, gate, run_if_set, kb0, kb_n, d_flag);
    MLBP_LAUNCH_CHECK();
    return MLBP_OK;
}

// stages per operand-plane count: 3 x 64 KB (hi+lo, hi+lo), 4 x 48 KB, 6 x 32 KB (hi, hi); `stages_scale`: probe variants.
// CHUNK_11: k-blocks per accumulator drain of the ONE-pass rows.  Their k-block lasts a third of a three-pass one, so at
// CHUNK_KB = 2 the epilogue warps had to drain 128 x 256 accumulators every ~1 000 clk and held the MMA issuer back:
// 4 k-blocks per drain is +13 % on these rows in the bench (0.93 -> 1.05 PFLOP/s).  The accumulation bias grows with the
// chunk (-2.2e-6 -> -4.4e-6, uniform) but these rows only feed the ratio N/Z, where it cancels (<= 3e-6 even between a
// dense and a 90 % sparse plane, scaled from profiles/r1f_gemm_probe_bias_sparse.txt).
template <int CHUNK_KB, int S22 = 3, int S12 = 4, int S11 = 6, bool A_REUSE = false, int CHUNK_11 = CHUNK_KB>
static int launch_pair(const void *A_hi, const void *A_lo, int64_t a_rows_total, int a_row0, int n_rows, const void *B_hi,
                       const void *B_lo, int V, int ldv, float *D, int64_t d_row0, int ldd, float alpha, float add_const, int a_terms,
                       int b_terms, const int32_t *gate, int run_if_set, int kb0, int kb_n, cudaStream_t st) {
#define MLBP_PAIR(S, AT, BT, CH) \
    return launch_pair_t<S, CH, AT, BT, A_REUSE>(A_hi, A_lo, a_rows_total, a_row0, n_rows, B_hi, B_lo, V, ldv, D, d_row0, ldd, alpha, add_const, gate, run_if_set, kb0, kb_n, st)
    if (a_terms == 2 && b_terms == 2) MLBP_PAIR(S22, 2, 2, CHUNK_KB);
    if (a_terms == 1 && b_terms == 2) MLBP_PAIR(S12, 1, 2, CHUNK_KB);
    if (a_terms == 2 && b_terms == 1) MLBP_PAIR(S12, 2, 1, CHUNK_KB);
    MLBP_PAIR(S11, 1, 1, CHUNK_11);
#undef MLBP_PAIR
}

}  // namespace mlbp

using namespace mlbp;

extern "C" int mlbp_gemm_barrier_timeout_code(void) { return g_err_flag ? *g_err_flag : 0; }

static int gemm_dispatch(const void *A_hi, const void *A_lo, int64_t a_rows_total, int a_row0, int n_rows, const void *B_hi,
                         const void *B_lo, int V, int ldv, float *D, int64_t d_row0, int ldd, float alpha, float add_const, int impl,
                         const int32_t *gate, int run_if_set, int k0, int k_len, void *stream) {
    if (n_rows == 0) return MLBP_OK;
    MLBP_CHECK_ARG(A_hi && A_lo && B_hi && B_lo && D, "factor_to_var_gemm: null pointer");
    MLBP_CHECK_ARG(n_rows > 0 && V > 0 && a_row0 >= 0 && a_row0 + (int64_t)n_rows <= a_rows_total,
                   "factor_to_var_gemm: row range [%d, %d) outside the A buffer (%lld rows)", a_row0, a_row0 + n_rows,
                   (long long)a_rows_total);
    MLBP_CHECK_ARG((ldv % 64) == 0 && ldv >= V && (ldd % 64) == 0 && ldd >= V, "factor_to_var_gemm: ld must be a multiple of 64 and >= V");
    MLBP_CHECK_ARG((reinterpret_cast<uintptr_t>(A_hi) % 128) == 0 && (reinterpret_cast<uintptr_t>(A_lo) % 128) == 0 &&
                   (reinterpret_cast<uintptr_t>(B_hi) % 128) == 0 && (reinterpret_cast<uintptr_t>(B_lo) % 128) == 0 &&
                   (reinterpret_cast<uintptr_t>(D) % 16) == 0, "factor_to_var_gemm: misaligned buffer");
    cudaStream_t st = as_stream(stream);
    // K range [k0, k0 + k_len) in elements (k_len == 0: all of K); a range that does not start at 0 ADDS to D
    const int kb_all = (V + 63) / 64;
    MLBP_CHECK_ARG(k0 >= 0 && k_len >= 0 && (k0 % 64) == 0 && (k_len % 64 == 0 || k0 + k_len >= V) && k0 < V + 64,
                   "factor_to_var_gemm: K range [%d, +%d) must be aligned to 64 elements", k0, k_len);
    const int kb0 = k0 / 64, kb_n = k_len == 0 ? kb_all - kb0 : ((k0 + k_len >= V ? kb_all : (k0 + k_len) / 64) - kb0);
    const bool whole_k = kb0 == 0 && kb_n == kb_all;
    const int a_terms = (impl & MLBP_GEMM_A_HI_ONLY) ? 1 : 2, b_terms = (impl & MLBP_GEMM_B_HI_ONLY) ? 1 : 2;
    impl &= ~(MLBP_GEMM_A_HI_ONLY | MLBP_GEMM_B_HI_ONLY);
    MLBP_CHECK_ARG(whole_k || ((impl == 0 && V > 2048) || impl == 2), "factor_to_var_gemm: only the CTA-pair kernel takes a partial K range");
    MLBP_CHECK_ARG(kb_n > 0, "factor_to_var_gemm: empty K range");
    if (impl == 1) {
        MLBP_CHECK_ARG(gate == nullptr && add_const == 0.f, "factor_to_var_gemm_gated: the SIMT cross-check kernel has no device-side gate and no constant term");
        return launch_gemm_simt(A_hi, A_lo, a_rows_total, a_row0, n_rows, B_hi, B_lo, V, ldv, D, d_row0, ldd, alpha, a_terms, b_terms, st);
    }
#define MLBP_ARGS A_hi, A_lo, a_rows_total, a_row0, n_rows, B_hi, B_lo, V, ldv, D, d_row0, ldd, alpha, add_const, a_terms, b_terms
#define MLBP_TC(BN_, ST_, CH_) return launch_tc<BN_, ST_, CH_>(MLBP_ARGS, gate, run_if_set, st)
    switch (impl) {
        case 0:                                       // product configuration: CTA-pair kernel for large V
            if (V > 2048) return launch_pair<2, 3, 4, 6, true, 4>(MLBP_ARGS, gate, run_if_set, kb0, kb_n, st);
            MLBP_TC(128, 3, 2);
        case 2: return launch_pair<2, 3, 4, 6, true, 4>(MLBP_ARGS, gate, run_if_set, kb0, kb_n, st);   // the CTA-pair kernel at any V (tests)
        case 3: MLBP_TC(128, 3, 2);                                                                        // the one-CTA kernel at any V (tests)
#ifdef MLBP_PROBES                                    // variants for scripts/gemm_probe.py only (build with -DMLBP_PROBES)
        case 10: MLBP_TC(256, 2, 1);
        case 11: MLBP_TC(256, 2, 2);
        case 12: MLBP_TC(256, 2, 4);
        case 13: MLBP_TC(256, 2, 1 << 20);      // never flush: one TMEM accumulation over all of K
        case 14: MLBP_TC(128, 3, 1);
        case 15: MLBP_TC(128, 3, 2);
        case 16: MLBP_TC(128, 3, 4);
        case 17: MLBP_TC(128, 3, 1 << 20);
        case 30: return launch_pair<2>(MLBP_ARGS, gate, run_if_set, kb0, kb_n, st);
        case 31: return launch_pair<1>(MLBP_ARGS, gate, run_if_set, kb0, kb_n, st);
        case 33: return launch_pair<2, 3, 4, 6, true>(MLBP_ARGS, gate, run_if_set, kb0, kb_n, st);       // one-pass rows drained every 2 too
        case 34: return launch_pair<4, 3, 4, 6, true>(MLBP_ARGS, gate, run_if_set, kb0, kb_n, st);       // all rows drained every 4 k-blocks
        case 35: return launch_pair<2, 3, 4, 6, true, 8>(MLBP_ARGS, gate, run_if_set, kb0, kb_n, st);    // one-pass rows drained every 8
        case 32: return launch_pair<2, 2, 3, 3>(MLBP_ARGS, gate, run_if_set, kb0, kb_n, st);
        case 20: return launch_tc<256, 4, 4, 32>(MLBP_ARGS, gate, run_if_set, st);
        case 21: return launch_tc<256, 4, 2, 32>(MLBP_ARGS, gate, run_if_set, st);
        case 22: return launch_tc<128, 6, 4, 32>(MLBP_ARGS, gate, run_if_set, st);
#endif
        default: break;
    }
#undef MLBP_TC
#undef MLBP_ARGS
    set_error("factor_to_var_gemm: unknown impl %d (probe variants need a -DMLBP_PROBES build)", impl);
    return MLBP_ERR_INVALID;
}

extern "C" int mlbp_factor_to_var_gemm(const void *A_hi, const void *A_lo, int64_t a_rows_total, int a_row0,
                                       int n_rows, const void *B_hi, const void *B_lo, int V, int ldv, float *D,
                                       int64_t d_row0, int ldd, float alpha, int impl, void *stream) {
    return gemm_dispatch(A_hi, A_lo, a_rows_total, a_row0, n_rows, B_hi, B_lo, V, ldv, D, d_row0, ldd, alpha, 0.f, impl, nullptr, 0, 0, 0, stream);
}

extern "C" int mlbp_factor_to_var_gemm_gated(const void *A_hi, const void *A_lo, int64_t a_rows_total, int a_row0,
                                             int n_rows, const void *B_hi, const void *B_lo, int V, int ldv, float *D,
                                             int64_t d_row0, int ldd, float alpha, int impl, const int32_t *gate,
                                             int run_if_set, int k0, int k_len, float add_const, void *stream) {
    return gemm_dispatch(A_hi, A_lo, a_rows_total, a_row0, n_rows, B_hi, B_lo, V, ldv, D, d_row0, ldd, alpha, add_const, impl, gate, run_if_set, k0, k_len, stream);
}
