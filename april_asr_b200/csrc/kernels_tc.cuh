// Tensor-core tier for sm_100a: tcgen05.mma (UMMA) fed by TMA into 128B-swizzled shared memory,
// fp32 accumulators in TMEM, warp-specialised persistent CTAs. All PTX is written by hand here.
//
// Numerics: the reference path is fp32 (ONNXRuntime CPU). To keep emitted tokens identical while
// using the fp16 tensor pipe, every operand x is carried as an fp16 pair (hi, lo) with
// hi = fp16(x), lo = fp16(x - hi) (22-23 significant bits), weights pre-scaled by a power of two so
// that lo never underflows, and each product is issued as three MMAs
//     A_hi*B_lo + A_lo*B_hi + A_hi*B_hi            (the dropped lo*lo term is < 2^-22 relative)
// accumulated in fp32 in TMEM ("fp16x3"). NPASS = 1 issues only A_hi*B_hi (plain fp16 inputs).
// The four sub-tiles of a k-block are loaded ONCE per stage and used by all three MMA groups, so
// the split costs 3x tensor work but only 2x shared-memory fill (64 MAC/B at 128x256x64 tiles).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace april {
namespace tc {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t_start = clock64();
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if ((spin & 0xFFFF) == 0xFFFF && clock64() - t_start > 4000000000LL) { printf("libapril: mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// 2-CTA (cta_group::2) forms: the load may signal an mbarrier of the pair's leader CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t mapa_tc(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory"); }
__device__ __forceinline__ uint32_t cluster_rank_tc() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_tc() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
// true in exactly one lane of a fully converged warp. The warp stays convergent around the single-thread tcgen05
// instructions, so their descriptor operands are computed warp-uniformly (uniform registers, no per-lane waterfall loop).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp receives lane (quarter base + i)
// D[tmem of both CTAs] (+)= A[smem, each CTA its M half] * B[smem, each CTA its N half]^T, issued by the pair's leader
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// K-major operand tile in shared memory, 128-byte swizzle (what TMA SWIZZLE_128B writes for a
// [rows x 64 fp16] box): 8-row groups of 1024 B (SBO), LBO unused, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                              // version = 1
    d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dswish_f(float x) { return x * sigmoid_f(x - 1.0f); }
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

// ------------------------------------------------------------------------------------------------
// GEMM  C[M,N] = (sum_p A_p[M,K] * B_p[N,K]^T) * inv_scale + bias  (+ epilogue)
// ------------------------------------------------------------------------------------------------
enum GemmEpi { GE_F32 = 0,            // out_f32 = v
               GE_F32_RESIDUAL = 1,   // out_f32 = v + R
               GE_DSWISH_SPLIT = 2,   // (out_hi, out_lo) = split(double_swish(v))
               GE_F32_RESIDUAL_SPLIT = 3 }; // out_f32 = v + R, and (out_hi, out_lo) = split(out_f32)

struct GemmArgs {
    int M, N, K;
    float inv_scale;
    const float* bias;        // [N] or null
    const float* residual;    // [M,N] fp32 (GE_F32_RESIDUAL)
    float* out_f32;           // [M,N]
    __half* out_hi;           // [M,N]
    __half* out_lo;
    int ksplit;               // > 1: K is cut into `ksplit` ranges of k-blocks, one work item each; the raw fp32 partial tiles go to
    float* partial;           //      partial[ks][M][N] and gemm_splitk_finish_kernel sums them IN ORDER and applies the epilogue
};

constexpr int GEMM_BM = 128, GEMM_BN = 256, GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;     // two per TMEM lane quarter, alternating 32-column chunks (the DoubleSwish + split epilogue of FF1 is
                                      // ~40 instructions per element: with 4 warps it, not the MMA, paced that GEMM)
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;     // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue

template <int NPASS, int CL = 1> struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;                    // 16 KB
    static constexpr int B_BYTES = (GEMM_BN / CL) * GEMM_BK * 2;             // 32 KB; a CTA of a pair holds half of the N rows
    static constexpr int STAGE_BYTES = (NPASS == 3 ? 2 : 1) * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (NPASS == 3 ? 2 : 4) + (CL == 2 ? 1 : 0);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// CL = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x 256 tile: CTA `rank` holds rows
// [rank * 128, +128) of A and N-rows [rank * 128, +128) of B in its own shared memory, the leader (rank 0) issues one
// M = 256 MMA that reads both halves, and each CTA's TMEM receives its 128 rows x 256 columns. Per k-block a CTA then
// ingests 64 KB instead of 96 KB for the same 128 x 256 outputs; the kernel is bound by what an SM can ingest
// (~25 B/clk measured), not by the tensor pipe (12 MMAs of 128 clk per k-block).
template <int NPASS, int EPI, int CL = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, GemmArgs g) {
    using Cfg = GemmCfg<NPASS, CL>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + S * Cfg::STAGE_BYTES;            // full[S], empty[S], tfull[2], tempty[2], tmem slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (g.M + GEMM_BM - 1) / GEMM_BM, tiles_n = (g.N + GEMM_BN - 1) / GEMM_BN;
    const int nkb = (g.K + GEMM_BK - 1) / GEMM_BK;
    // work item w = (group of CL vertically adjacent m-tiles, n-tile); CTA `rank` of the pair takes m-tile group * CL + rank
    // (past the last tile for odd tiles_m: its A loads are zero-filled by TMA and its epilogue stores nothing)
    const int rank = CL == 2 ? (int)cluster_rank_tc() : 0;
    const int cid = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
    // split-K (small M: a handful of tiles cannot occupy the GPU, and one CTA streaming all of K is paced by what an SM ingests):
    // work item w = (tile, k-range ks); the ranges of one tile are consecutive work items, so neighbouring CTAs share the A rows in L2
    const int ksplit = g.ksplit > 1 ? g.ksplit : 1;
    const int num_work = ((tiles_m + CL - 1) / CL) * tiles_n * ksplit;
    auto work_m0 = [&](int w) { return (((w / ksplit) / tiles_n) * CL + rank) * GEMM_BM; };
    auto work_n0 = [&](int w) { return ((w / ksplit) % tiles_n) * GEMM_BN; };
    auto work_kb0 = [&](int w) { return (int)(((long long)(w % ksplit) * nkb) / ksplit); };
    auto work_kb1 = [&](int w) { return (int)(((long long)(w % ksplit + 1) * nkb) / ksplit); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), GEMM_EPI_WARPS * CL); }   // leader: both CTAs' epilogue warps
        fence_barrier_init();
    }
    if (warp == 1) { if (CL == 2) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_tc();                        // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer: convergent warp, one elected lane issues =====================
        {
            if (lane == 0) {
                tma_prefetch_desc(&mapAhi); tma_prefetch_desc(&mapBhi);
                if (NPASS == 3) { tma_prefetch_desc(&mapAlo); tma_prefetch_desc(&mapBlo); }
            }
            __syncwarp();
            uint32_t it = 0;
            for (int w = cid; w < num_work; w += ncl) {
                const int m0 = work_m0(w), n0 = work_n0(w) + rank * (GEMM_BN / CL);   // pair: my half of the B rows
                for (int kb = work_kb0(w), kbe = work_kb1(w); kb < kbe; kb++, it++) {
                    const int s = it % S; const uint32_t ph = (it / S) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);               // pair: the leader's commit is multicast to both CTAs
                    const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
                    if (elect_one()) {
                        if (CL == 2) {
                            const uint32_t fb = mapa_tc(full_bar(s), 0);                    // everything lands on the LEADER's barrier
                            if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);   // both CTAs' bytes
                            tma_load_2d_pair(st, &mapAhi, fb, kb * GEMM_BK, m0);
                            tma_load_2d_pair(st + Cfg::A_BYTES, &mapBhi, fb, kb * GEMM_BK, n0);
                            if (NPASS == 3) {
                                tma_load_2d_pair(st + Cfg::A_BYTES + Cfg::B_BYTES, &mapAlo, fb, kb * GEMM_BK, m0);
                                tma_load_2d_pair(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapBlo, fb, kb * GEMM_BK, n0);
                            }
                        } else {
                            mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
                            tma_load_2d(st, &mapAhi, full_bar(s), kb * GEMM_BK, m0);
                            tma_load_2d(st + Cfg::A_BYTES, &mapBhi, full_bar(s), kb * GEMM_BK, n0);
                            if (NPASS == 3) {
                                tma_load_2d(st + Cfg::A_BYTES + Cfg::B_BYTES, &mapAlo, full_bar(s), kb * GEMM_BK, m0);
                                tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapBlo, full_bar(s), kb * GEMM_BK, n0);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair: the leader CTA only): convergent warp, one elected lane issues ==========
        if (CL == 1 || rank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(GEMM_BM * CL, GEMM_BN);
            uint32_t it = 0, tcount = 0;
            for (int w = cid; w < num_work; w += ncl, tcount++) {
                const int acc = tcount & 1; const uint32_t aph = (tcount >> 1) & 1;
                mbar_wait(tempty_bar(acc), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
                const int kbs = work_kb0(w);
                for (int kb = kbs, kbe = work_kb1(w); kb < kbe; kb++, it++) {
                    const int s = it % S; const uint32_t ph = (it / S) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
                    const uint64_t ahi = make_sw128_desc(st), bhi = make_sw128_desc(st + Cfg::A_BYTES);
                    const uint64_t alo = make_sw128_desc(st + Cfg::A_BYTES + Cfg::B_BYTES), blo = make_sw128_desc(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
                    const uint32_t first = kb != kbs;
                    if (elect_one()) {
                        auto mma = [&](uint64_t ad, uint64_t bd, uint32_t accum) { if (CL == 2) umma_f16_pair(d_tmem, ad, bd, idesc, accum); else umma_f16(d_tmem, ad, bd, idesc, accum); };
                        if (NPASS == 3) {
#pragma unroll
                            for (int k = 0; k < GEMM_BK / 16; k++) mma(ahi + 2 * k, blo + 2 * k, first || k != 0);
#pragma unroll
                            for (int k = 0; k < GEMM_BK / 16; k++) mma(alo + 2 * k, bhi + 2 * k, 1);
#pragma unroll
                            for (int k = 0; k < GEMM_BK / 16; k++) mma(ahi + 2 * k, bhi + 2 * k, 1);
                        } else {
#pragma unroll
                            for (int k = 0; k < GEMM_BK / 16; k++) mma(ahi + 2 * k, bhi + 2 * k, first || k != 0);
                        }
                        if (CL == 2) umma_commit_pair(empty_bar(s), 0x3); else umma_commit(empty_bar(s));   // stage free once these MMAs have read it
                    }
                    __syncwarp();
                }
                if (elect_one()) { if (CL == 2) umma_commit_pair(tfull_bar(acc), 0x3); else umma_commit(tfull_bar(acc)); }   // accumulator complete (in both CTAs)
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        uint32_t tcount = 0;
        for (int w = cid; w < num_work; w += ncl, tcount++) {
            const int acc = tcount & 1; const uint32_t aph = (tcount >> 1) & 1;
            const int m0 = work_m0(w), n0 = work_n0(w);
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
#pragma unroll 1
            for (int c = (warp - 2) >> 2; c < GEMM_BN / 32; c += GEMM_EPI_WARPS / 4) {
                const int nb = n0 + c * 32;
                if (nb >= g.N) break;                      // warp-uniform
                float v[32];
                tmem_ld32(tmem_base + acc * GEMM_BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
                if (m < g.M && ksplit > 1) {                                 // raw partial sums of this k-range
                    float* pw = g.partial + ((size_t)(w % ksplit) * g.M + m) * g.N + nb;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (nb + j < g.N) *reinterpret_cast<float4*>(pw + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else if (m < g.M) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        if (nb + j >= g.N) break;
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; e++) x[e] = v[j + e] * g.inv_scale + (g.bias ? __ldg(g.bias + nb + j + e) : 0.f);
                        const size_t off = (size_t)m * g.N + nb + j;
                        if (EPI == GE_F32 || EPI == GE_F32_RESIDUAL || EPI == GE_F32_RESIDUAL_SPLIT) {
                            if (EPI == GE_F32_RESIDUAL || EPI == GE_F32_RESIDUAL_SPLIT) {
                                const float4 r0 = *reinterpret_cast<const float4*>(g.residual + off), r1 = *reinterpret_cast<const float4*>(g.residual + off + 4);
                                x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w; x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
                            }
                            *reinterpret_cast<float4*>(g.out_f32 + off) = make_float4(x[0], x[1], x[2], x[3]);
                            *reinterpret_cast<float4*>(g.out_f32 + off + 4) = make_float4(x[4], x[5], x[6], x[7]);
                            if (EPI == GE_F32_RESIDUAL_SPLIT) {
                                __align__(16) __half hi[8], lo[8];
#pragma unroll
                                for (int e = 0; e < 8; e++) split_f16(x[e], hi[e], lo[e]);
                                *reinterpret_cast<uint4*>(g.out_hi + off) = *reinterpret_cast<const uint4*>(hi);
                                *reinterpret_cast<uint4*>(g.out_lo + off) = *reinterpret_cast<const uint4*>(lo);
                            }
                        } else {
                            __align__(16) __half hi[8], lo[8];
#pragma unroll
                            for (int e = 0; e < 8; e++) split_f16(dswish_f(x[e]), hi[e], lo[e]);
                            *reinterpret_cast<uint4*>(g.out_hi + off) = *reinterpret_cast<const uint4*>(hi);
                            *reinterpret_cast<uint4*>(g.out_lo + off) = *reinterpret_cast<const uint4*>(lo);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CL == 2 && rank != 0) mbar_arrive_cluster(mapa_tc(tempty_bar(acc), 0)); else mbar_arrive(tempty_bar(acc)); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_tc();                        // no exit while the peer may still read my operands or signal my barriers
    if (warp == 1) { if (CL == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

// split-K second half: out = epilogue((partial[0] + partial[1] + ... in this order) * inv_scale + bias); 4 columns per thread
template <int EPI>
__global__ void gemm_splitk_finish_kernel(GemmArgs g) {
    const size_t total4 = (size_t)g.M * g.N / 4, plane = (size_t)g.M * g.N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t off = i * 4;
        const int n = (int)(off % (size_t)g.N);
        float4 a = *reinterpret_cast<const float4*>(g.partial + off);
        for (int ks = 1; ks < g.ksplit; ks++) {
            const float4 b = *reinterpret_cast<const float4*>(g.partial + (size_t)ks * plane + off);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        float x[4] = { a.x * g.inv_scale, a.y * g.inv_scale, a.z * g.inv_scale, a.w * g.inv_scale };
        if (g.bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n)); x[0] += b.x; x[1] += b.y; x[2] += b.z; x[3] += b.w; }
        if (EPI == GE_F32 || EPI == GE_F32_RESIDUAL || EPI == GE_F32_RESIDUAL_SPLIT) {
            if (EPI == GE_F32_RESIDUAL || EPI == GE_F32_RESIDUAL_SPLIT) {
                const float4 r = *reinterpret_cast<const float4*>(g.residual + off);
                x[0] += r.x; x[1] += r.y; x[2] += r.z; x[3] += r.w;
            }
            *reinterpret_cast<float4*>(g.out_f32 + off) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) x[e] = dswish_f(x[e]);
        }
        if (EPI == GE_F32_RESIDUAL_SPLIT || EPI == GE_DSWISH_SPLIT) {
            __align__(8) __half hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; e++) split_f16(x[e], hi[e], lo[e]);
            *reinterpret_cast<uint2*>(g.out_hi + off) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(g.out_lo + off) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// element-wise helpers that also produce the fp16 (hi, lo) operand copies for the next GEMM
// ------------------------------------------------------------------------------------------------
__global__ void split_rows_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += stride) {
        float4 v = *reinterpret_cast<const float4*>(x + i);
        __align__(8) __half h[4], l[4];
        split_f16(v.x, h[0], l[0]); split_f16(v.y, h[1], l[1]); split_f16(v.z, h[2], l[2]); split_f16(v.w, h[3], l[3]);
        *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
    }
}

// weights at load time: x * sc (a power of two: exact) as an fp16 (hi, lo) pair; the same roundings as the host would do
__global__ void split_scale_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, size_t n, float sc) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) split_f16(x[i] * sc, hi[i], lo[i]);
}

// y = a + b, plus split copies
__global__ void add_rows_split_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                      __half* __restrict__ hi, __half* __restrict__ lo, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += stride) {
        float4 p = *reinterpret_cast<const float4*>(a + i), q = *reinterpret_cast<const float4*>(b + i);
        float4 v = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
        *reinterpret_cast<float4*>(y + i) = v;
        __align__(8) __half h[4], l[4];
        split_f16(v.x, h[0], l[0]); split_f16(v.y, h[1], l[1]); split_f16(v.z, h[2], l[2]); split_f16(v.w, h[3], l[3]);
        *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
    }
}

// BasicNorm (one warp per row) writing fp32 and split copies
__global__ void basic_norm_split_kernel(const float* __restrict__ X, float* __restrict__ Y, __half* __restrict__ hi, __half* __restrict__ lo,
                                        int M, int D, float eps) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* x = X + (size_t)row * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s = fmaf(x[i], x[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float scale = 1.0f / sqrtf(s / (float)D + eps);
    for (int i = lane; i < D; i += 32) {
        float v = x[i] * scale;
        Y[(size_t)row * D + i] = v;
        __half h, l; split_f16(v, h, l);
        hi[(size_t)row * D + i] = h; lo[(size_t)row * D + i] = l;
    }
}

// ------------------------------------------------------------------------------------------------
// E1 front half for the tensor-core tier: conv1 + conv2 (fp32 SIMT, 0.3 MMAC per window) and the im2col of
// conv3's input written as fp16 (hi, lo) rows, so that conv3 (0.66 MMAC per window, 68 % of the subsampler)
// and the following 2304 -> D linear both run as tcgen05 GEMMs:
//   A3[(w * F3 + f)][ic * 9 + ky * 3 + kx] = c2[ic][ky][2 f + kx]        (conv3: kernel 3x3, stride 2, T3 == 1)
// conv3's GEMM output row (w, f) x C3 channels is, read as [w][f * C3 + c], the input of the linear with its
// weight columns permuted from (c * F3 + f) to (f * C3 + c) at load time (icefall Conv2dSubsampling:
// x.transpose(1, 2).view(b, t, c * f) [external]).
// One CTA per (session, step) window; 256 threads.
// ------------------------------------------------------------------------------------------------
struct Conv12Params {
    int SEG, MEL, STEP, C1, C2, T1, F1, T2, F2, F3;
    const float *w1, *b1, *w2t /*[C1*9][C2] transposed*/, *b2;
};

__global__ void __launch_bounds__(256)
conv12_im2col_kernel(Conv12Params P, const float* __restrict__ FB, int row_pitch, const int* __restrict__ nsteps, int t0, int B,
                     __half* __restrict__ A_hi, __half* __restrict__ A_lo) {
    extern __shared__ float smem_c[];
    const int b = blockIdx.x, t = blockIdx.y;
    const int K3 = P.C2 * 9;
    const size_t row0 = ((size_t)t * B + b) * P.F3;
    if (t0 + t >= nsteps[b]) {                                       // ragged tail: defined (zero) rows
        for (int i = threadIdx.x; i < P.F3 * K3; i += blockDim.x) { A_hi[row0 * K3 + i] = __float2half(0.f); A_lo[row0 * K3 + i] = __float2half(0.f); }
        return;
    }
    float* in = smem_c;                                   // [SEG][MEL]
    float* c1 = in + P.SEG * P.MEL;                       // [C1][T1][F1]
    float* c2 = c1 + P.C1 * P.T1 * P.F1;                  // [C2][T2][F2]
    float* w2s = c2 + P.C2 * P.T2 * P.F2;                 // [C1*9][C2]
    const float* src = FB + ((size_t)b * row_pitch + (size_t)(t0 + t) * P.STEP) * P.MEL;
    for (int i = threadIdx.x; i < P.SEG * P.MEL; i += blockDim.x) in[i] = src[i];
    for (int i = threadIdx.x; i < P.C1 * 9 * P.C2; i += blockDim.x) w2s[i] = P.w2t[i];
    __syncthreads();
    for (int i = threadIdx.x; i < P.C1 * P.T1 * P.F1; i += blockDim.x) {
        const int f = i % P.F1, y = (i / P.F1) % P.T1, c = i / (P.F1 * P.T1);
        const float* w = P.w1 + c * 9;
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) s = fmaf(in[(y + ky) * P.MEL + f + kx], __ldg(w + ky * 3 + kx), s);
        c1[i] = dswish_f(s + __ldg(P.b1 + c));
    }
    __syncthreads();
    // conv2: work item = (position p = y * F2 + x, group of 8 output channels): inputs loaded once, weights as float4
    const int npos = P.T2 * P.F2, ngrp = P.C2 / 8;
    for (int wi = threadIdx.x; wi < npos * ngrp; wi += blockDim.x) {
        const int p = wi % npos, cg = wi / npos;
        const int y = p / P.F2, x = p % P.F2;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = 0.f;
        for (int ic = 0; ic < P.C1; ic++) {
            const float* ip = c1 + (ic * P.T1 + 2 * y) * P.F1 + 2 * x;
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const float v = ip[(k / 3) * P.F1 + (k % 3)];
                const float4 wa = *reinterpret_cast<const float4*>(w2s + (ic * 9 + k) * P.C2 + cg * 8);
                const float4 wb = *reinterpret_cast<const float4*>(w2s + (ic * 9 + k) * P.C2 + cg * 8 + 4);
                acc[0] = fmaf(v, wa.x, acc[0]); acc[1] = fmaf(v, wa.y, acc[1]); acc[2] = fmaf(v, wa.z, acc[2]); acc[3] = fmaf(v, wa.w, acc[3]);
                acc[4] = fmaf(v, wb.x, acc[4]); acc[5] = fmaf(v, wb.y, acc[5]); acc[6] = fmaf(v, wb.z, acc[6]); acc[7] = fmaf(v, wb.w, acc[7]);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; e++) { const int oc = cg * 8 + e; c2[(oc * P.T2 + y) * P.F2 + x] = dswish_f(acc[e] + __ldg(P.b2 + oc)); }
    }
    __syncthreads();
    // im2col rows of conv3 (T3 == 1: ky indexes the 3 rows of c2 directly)
    for (int i = threadIdx.x; i < P.F3 * K3; i += blockDim.x) {
        const int f = i / K3, col = i % K3;
        const int ic = col / 9, ky = (col % 9) / 3, kx = col % 3;
        const float v = c2[(ic * P.T2 + ky) * P.F2 + 2 * f + kx];
        __half h, l; split_f16(v, h, l);
        A_hi[row0 * K3 + i] = h; A_lo[row0 * K3 + i] = l;
    }
}

}  // namespace tc
}  // namespace april
