// Persistent tensor-core LSTM scan, single-exchange formulation.
//
// The projected LSTM's recurrence  h' = W_hr m,  gates' = G + W_hh h'  needs two all-to-all exchanges per
// step when W_hh and W_hr live on different SMs (kernels_scan.cuh). Folding the projection into the
// next step's gate matrix,
//        gates(t) = G[t] + (W_hh W_hr) m(t-1),        W_eff = W_hh W_hr  in  [4H, H],
// makes m (H per session) the only recurrent quantity, so a step costs ONE grid-wide handover. The
// layer output h(t) = W_hr m(t) is then not on the critical path: it is one more time-batched GEMM
// over all steps (engine.cu). Mathematically identical to the reference's order of operations
// (icefall lstmp.py via extra/export-april.py:564 [external]); rounding differs at the 1e-7 level.
//
// Layout: H/8 CTAs (128 for H = 1024), CTA c owns hidden units [8c, 8c+8): 32 rows of W_eff, stacked
// [hi ; lo] as 64 fp16 rows x K = H, resident in shared memory as H/64 k-blocks of 8 KB (a 128-row UMMA
// over-reads the next k-block: TMEM lanes 64-127 are never read). Per step every CTA streams
// m(t-1) = [m_hi ; m_lo] for up to 64 sessions (N = 128) through a small ring of 5-D TMA boxes, issues
// H/16 tcgen05.mma, sums the four (row half, column half) quadrants, applies the cell update for its
// 8 units x sessions, writes its slice of m(t) (time-indexed rows: no write-after-read hazard) and
// bumps one global counter.
#pragma once
#include "kernels_scan.cuh"

namespace april {
namespace scan2 {

using namespace tc;
using scan::fast_sigmoid;
using scan::fast_tanh;
using scan::gtimer;
using scan::ld_acquire;
using scan::red_release;
using scan::wait_counter;
using scan::epi_bar_sync;

constexpr int THREADS = 320;               // warps 0-3: one TMA ring stage each (a spinning lane must not share a warp with another producer),
                                          // warp 4: weights + G loader, warp 5: MMA issuer, warps 6-9: epilogue
constexpr int UNITS = 8;                  // hidden units per CTA -> 32 gate rows, 64 stacked (hi ; lo)
constexpr int RSTAGES = 4;                // activation ring stages
constexpr int MAXG = 2;
// Up to two independent session groups are interleaved through the CTA (as in kernels_scan.cuh): the ring keeps
// streaming group 1's m rows while group 0 is in its epilogue / handover, and vice versa. An optional thread-block
// cluster (CS CTAs) fetches each copy once and multicasts it; measured on B200 this does not help (the ring, not L2,
// bounds the ingest), so the default is CS = 1.

struct Scan2Args {
    int T, t0;             // steps in this launch, absolute index of the first one
    int NG;                // interleaved groups
    int Bg[MAXG];          // sessions per group
    int goff[MAXG];        // first session of the group relative to g0
    int Bp;                // padded group size (multiple of 16, <= 64 for NG = 1, <= 32 for NG = 2)
    int Bs, g0;            // sessions in the whole pass = row stride of G / M buffers; first session of this launch
    int H;
    int GC;                // CTAs = H / 8
    int KBG;               // k-blocks per TMA copy
    float inv_scale;
    const int* nsteps;     // [Bs] (not offset)
    float* c_state;        // [Bs, H] fp32 pass-layout state of this layer (read at start, written at end)
    float* m_state;        // [Bs, H] fp32: m of the last live step (written when a session ends or the launch ends)
    __half* Mx;            // [(T+1)][2 (hi|lo)][Bs][H]: exchange rows; row 0 = initial m (prepared by the caller)
    __half* Mall_hi;       // [T * Bs, H] : m(t) rows for the layer-output GEMM
    __half* Mall_lo;
    unsigned* counter;     // [MAXG], zeroed by the caller
    unsigned long long* trace;
};

#define TRACE2(slot_) do { if (a.trace) a.trace[((size_t)blockIdx.x * a.T + t) * 8 + (slot_)] = gtimer(); } while (0)

// one L2 read, delivered to the same shared-memory offset (and mbarrier) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_5d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// tcgen05.commit arriving on the barrier at the same offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
lstm_scan2_kernel(const __grid_constant__ CUtensorMap mapW,    // 2-D [GC*64, H] stacked W_eff slices, box {64, 64}
                  const __grid_constant__ CUtensorMap mapMx,   // 5-D {64, Bs, 2, H/64, T+1}, box {64, Bp, 2, KBG, 1}
                  const __grid_constant__ CUtensorMap mapG,    // 2-D fp32 [T*Bs, 4H], box {8, Bp}
                  Scan2Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NG = a.NG;

    const int KB_W = a.H / 64;
    const uint32_t w_kb_bytes = 64u * 128u, WBYTES = (uint32_t)KB_W * w_kb_bytes;
    const int Nact = 2 * a.Bp;                                            // UMMA N: [m_hi rows ; m_lo rows]
    const uint32_t KBB = (uint32_t)Nact * 128u, SB = (uint32_t)a.KBG * KBB;
    const uint32_t ring0 = base + WBYTES;
    const int ld = a.Bp + 1;
    const uint32_t exch0 = ring0 + RSTAGES * SB;
    const uint32_t gbuf0 = (exch0 + 64u * (uint32_t)ld * 4u + 127u) & ~127u;
    const uint32_t GB = (uint32_t)a.Bp * UNITS * 4u * 4u;                 // 4 gates x [Bp x 8] fp32, per group
    const uint32_t cbuf = gbuf0 + (uint32_t)NG * GB;
    const uint32_t CB = (uint32_t)a.Bp * UNITS * 4u;                      // per group
    const uint32_t bar0 = (cbuf + (uint32_t)NG * CB + 15u) & ~15u;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (RSTAGES + s); };
    const uint32_t wfull_bar = bar0 + 8u * (2 * RSTAGES);
    auto tfull_bar = [&](int g) { return wfull_bar + 8u + 8u * g; };
    auto tempty_bar = [&](int g) { return wfull_bar + 8u + 8u * (MAXG + g); };
    auto gfull_bar = [&](int g) { return wfull_bar + 8u + 8u * (2 * MAXG + g); };
    auto gfree_bar = [&](int g) { return wfull_bar + 8u + 8u * (3 * MAXG + g); };
    const uint32_t tmem_slot = wfull_bar + 8u + 8u * (4 * MAXG);
    float* exch = reinterpret_cast<float*>(gptr(exch0));
    float* cst = reinterpret_cast<float*>(gptr(cbuf));

    const uint32_t crank = cluster_ctarank_(), CS = cluster_nctarank_();
    const uint16_t cmask = (uint16_t)((1u << CS) - 1u);
    if (threadIdx.x == 0) {
        for (int s = 0; s < RSTAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CS); }
        mbar_init(wfull_bar, 1);
        for (int g = 0; g < MAXG; g++) { mbar_init(tfull_bar(g), 1); mbar_init(tempty_bar(g), 1); mbar_init(gfull_bar(g), 1); mbar_init(gfree_bar(g), 1); }
        fence_barrier_init();
    }
    const uint32_t tmem_cols = (uint32_t)(NG * Nact) <= 32u ? 32u : ((uint32_t)(NG * Nact) <= 64u ? 64u : 128u);
    if (warp == RSTAGES + 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_();                                          // peers' barriers are initialised before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));
    const uint32_t idesc = make_idesc_f16(128, Nact);
    const int u0 = blockIdx.x * UNITS;
    const int copies = KB_W / a.KBG;                                      // TMA copies per (step, group)

    if (warp < RSTAGES) {
        if (lane == 0) {
            const int stage = warp;
            tma_prefetch_desc(&mapMx);
            const long long total = (long long)a.T * NG * copies;
            int opened_t = -1, opened_g = -1;
            for (long long idx = stage; idx < total; idx += RSTAGES) {
                const int j = (int)(idx % copies);
                const int tg = (int)(idx / copies), g = tg % NG, t = tg / NG;
                if (t != opened_t || g != opened_g) {
                    if (t > 0) wait_counter(a.counter + g, (unsigned)t * (unsigned)a.GC);     // m(t-1) of the group complete everywhere
                    opened_t = t; opened_g = g;
                    if (j == 0) TRACE2(0);
                }
                const uint32_t use = (uint32_t)(idx / RSTAGES);
                mbar_wait(empty_bar(stage), (use & 1) ^ 1);              // stage free (in every CTA of the cluster)
                mbar_expect_tx(full_bar(stage), SB);
                if (CS == 1) tma_load_5d(ring0 + stage * SB, &mapMx, full_bar(stage), 0, a.g0 + a.goff[g], 0, j * a.KBG, t);   // row t holds m(t-1)
                else if ((uint32_t)j % CS == crank) tma_load_5d_mc(ring0 + stage * SB, &mapMx, full_bar(stage), 0, a.g0 + a.goff[g], 0, j * a.KBG, t, cmask);
                if (j == copies - 1) TRACE2(1);
            }
        }
    } else if (warp == RSTAGES) {
        if (lane == 0) {
            tma_prefetch_desc(&mapG);
            mbar_expect_tx(wfull_bar, WBYTES);
            for (int kb = 0; kb < KB_W; kb++) tma_load_2d(base + kb * w_kb_bytes, &mapW, wfull_bar, kb * 64, blockIdx.x * 64);
            for (int t = 0; t < a.T; t++)
                for (int g = 0; g < NG; g++) {
                    if (t > 0) mbar_wait(gfree_bar(g), (t - 1) & 1);
                    mbar_expect_tx(gfull_bar(g), GB);
                    for (int g4 = 0; g4 < 4; g4++) tma_load_2d(gbuf0 + g * GB + g4 * (GB / 4), &mapG, gfull_bar(g), g4 * a.H + u0, t * a.Bs + a.g0 + a.goff[g]);
                }
        }
    } else if (warp == RSTAGES + 1) {
        if (lane == 0) {
            mbar_wait(wfull_bar, 0);
            long long idx = 0;
            for (int t = 0; t < a.T; t++)
                for (int g = 0; g < NG; g++) {
                    mbar_wait(tempty_bar(g), (t & 1) ^ 1);
                    tc_fence_after();
                    for (int j = 0; j < copies; j++, idx++) {
                        const int s = (int)(idx % RSTAGES);
                        const uint32_t use = (uint32_t)(idx / RSTAGES);
                        mbar_wait(full_bar(s), use & 1);
                        if (j == 0) TRACE2(2);
                        tc_fence_after();
                        for (int kk = 0; kk < a.KBG; kk++) {
                            const uint64_t ad = make_sw128_desc(base + (j * a.KBG + kk) * w_kb_bytes), bd = make_sw128_desc(ring0 + s * SB + kk * KBB);
#pragma unroll
                            for (int k = 0; k < 4; k++) umma_f16(tmem_base + g * Nact, ad + 2 * k, bd + 2 * k, idesc, (j | kk | k) != 0);
                        }
                        if (CS == 1) umma_commit(empty_bar(s)); else umma_commit_mc(empty_bar(s), cmask);
                    }
                    umma_commit(tfull_bar(g));
                    TRACE2(3);
                }
        }
    } else {
        const int q = warp & 3, et = (warp - (RSTAGES + 2)) * 32 + lane;   // et: 0..127
        // cell phase: Bp <= 32: session b = et / 4, units uq*2, uq*2+1 ; Bp == 64 (NG == 1): session b = et / 2, units uq*4 .. +3
        const bool wide = a.Bp > 32;
        const int upt = wide ? 4 : 2;                                      // units per thread
        const int my_b = wide ? (et >> 1) : (et >> 2), uq = wide ? (et & 1) : (et & 3);
        int my_n[MAXG];
        for (int g = 0; g < MAXG; g++) {
            const int sess = a.g0 + (g < NG ? a.goff[g] : 0) + my_b;
            my_n[g] = (g < NG && my_b < a.Bg[g]) ? a.nsteps[sess] : 0;
            if (g < NG && my_b < a.Bp)
                for (int uu = 0; uu < upt; uu++)
                    cst[g * (a.Bp * UNITS) + my_b * UNITS + uq * upt + uu] = my_b < a.Bg[g] ? a.c_state[(size_t)sess * a.H + u0 + uq * upt + uu] : 0.f;
        }
        for (int t = 0; t < a.T; t++)
            for (int g = 0; g < NG; g++) {
                mbar_wait(tfull_bar(g), t & 1);
                if (et == 0) TRACE2(4);
                tc_fence_after();
                if (q < 2) {                                              // TMEM lanes 0-31: W_hi rows, 32-63: W_lo rows
                    const uint32_t tb = tmem_base + g * Nact + ((uint32_t)(q * 32) << 16);
                    if (a.Bp == 16) {
                        float v[32];
                        tmem_ld32(tb, v);
#pragma unroll
                        for (int j = 0; j < 16; j++) exch[(q * 32 + lane) * ld + j] = v[j] + v[16 + j];
                    } else {
                        for (int c0 = 0; c0 < a.Bp; c0 += 32) {
                            float v[32], w2[32];
                            tmem_ld32(tb + c0, v);
                            tmem_ld32(tb + a.Bp + c0, w2);
#pragma unroll
                            for (int j = 0; j < 32; j++) if (c0 + j < a.Bp) exch[(q * 32 + lane) * ld + c0 + j] = v[j] + w2[j];
                        }
                    }
                }
                tc_fence_before();
                epi_bar_sync();
                if (et == 0) { mbar_arrive(tempty_bar(g)); TRACE2(5); }
                mbar_wait(gfull_bar(g), t & 1);
                if (my_b < a.Bp) {
                    const float* gb = reinterpret_cast<const float*>(gptr(gbuf0 + g * GB));
                    float* cs = cst + g * (a.Bp * UNITS) + my_b * UNITS + uq * upt;
                    const bool live = a.t0 + t < my_n[g];
                    const int sess = a.g0 + a.goff[g] + my_b;
                    float mval[4];
#pragma unroll
                    for (int uu = 0; uu < 4; uu++) {
                        if (uu < upt) {
                            const int u = uq * upt + uu;
                            float gate[4];
#pragma unroll
                            for (int g4 = 0; g4 < 4; g4++)
                                gate[g4] = (exch[(g4 * 8 + u) * ld + my_b] + exch[(32 + g4 * 8 + u) * ld + my_b]) * a.inv_scale + gb[(g4 * a.Bp + my_b) * UNITS + u];
                            const float cn = fast_sigmoid(gate[1]) * cs[uu] + fast_sigmoid(gate[0]) * fast_tanh(gate[2]);
                            if (live) cs[uu] = cn;
                            mval[uu] = fast_sigmoid(gate[3]) * fast_tanh(cn);
                        }
                    }
                    if (live) {
                        __align__(8) __half mh[4], ml[4];
#pragma unroll
                        for (int uu = 0; uu < 4; uu++) if (uu < upt) split_f16(mval[uu], mh[uu], ml[uu]);
                        const size_t col = (size_t)u0 + uq * upt;
                        __half* mx = a.Mx + (size_t)(t + 1) * 2 * a.Bs * a.H;
                        const size_t row = (size_t)t * a.Bs + sess;
                        const bool endw = (a.t0 + t == my_n[g] - 1) || (t == a.T - 1);
                        if (wide) {
                            *reinterpret_cast<uint2*>(mx + (size_t)sess * a.H + col) = *reinterpret_cast<const uint2*>(mh);
                            *reinterpret_cast<uint2*>(mx + ((size_t)a.Bs + sess) * a.H + col) = *reinterpret_cast<const uint2*>(ml);
                            *reinterpret_cast<uint2*>(a.Mall_hi + row * a.H + col) = *reinterpret_cast<const uint2*>(mh);
                            *reinterpret_cast<uint2*>(a.Mall_lo + row * a.H + col) = *reinterpret_cast<const uint2*>(ml);
                            if (endw) *reinterpret_cast<float4*>(a.m_state + (size_t)sess * a.H + col) = make_float4(mval[0], mval[1], mval[2], mval[3]);
                        } else {
                            *reinterpret_cast<uint32_t*>(mx + (size_t)sess * a.H + col) = *reinterpret_cast<const uint32_t*>(mh);
                            *reinterpret_cast<uint32_t*>(mx + ((size_t)a.Bs + sess) * a.H + col) = *reinterpret_cast<const uint32_t*>(ml);
                            *reinterpret_cast<uint32_t*>(a.Mall_hi + row * a.H + col) = *reinterpret_cast<const uint32_t*>(mh);
                            *reinterpret_cast<uint32_t*>(a.Mall_lo + row * a.H + col) = *reinterpret_cast<const uint32_t*>(ml);
                            if (endw) *reinterpret_cast<float2*>(a.m_state + (size_t)sess * a.H + col) = make_float2(mval[0], mval[1]);
                        }
                    }
                }
                if (et == 0) TRACE2(6);
                asm volatile("fence.proxy.async;" ::: "memory");
                epi_bar_sync();
                if (et == 0) { red_release(a.counter + g); mbar_arrive(gfree_bar(g)); TRACE2(7); }
            }
        for (int g = 0; g < NG; g++)
            if (my_b < a.Bg[g]) {
                const int sess = a.g0 + a.goff[g] + my_b;
                for (int uu = 0; uu < upt; uu++) a.c_state[(size_t)sess * a.H + u0 + uq * upt + uu] = cst[g * (a.Bp * UNITS) + my_b * UNITS + uq * upt + uu];
            }
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_();                                          // nobody exits while peers may still signal its barriers
    if (warp == RSTAGES + 1) tmem_dealloc(tmem_base, tmem_cols);
}

// row 0 of the exchange buffer (initial m of the launch's sessions) from the fp32 state; finished sessions too (harmless)
__global__ void scan2_prepare_kernel(const float* __restrict__ m_state, __half* __restrict__ Mx, int Bs, int H) {
    const size_t n = (size_t)Bs * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        __half hi, lo; split_f16(m_state[i], hi, lo);
        Mx[i] = hi; Mx[n + i] = lo;
    }
}

// W_eff = W_hh W_hr with fp64 accumulation: C[i][j] = sum_k A[i][k] B[k][j]  (A [M,K], B [K,N] row-major)
__global__ void weff_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, int M, int N, int K) {
    __shared__ double As[16][17], Bs2[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
    double acc = 0.0;
    for (int k0 = 0; k0 < K; k0 += 16) {
        As[ty][tx] = (row < M && k0 + tx < K) ? (double)A[(size_t)row * K + k0 + tx] : 0.0;
        Bs2[ty][tx] = (k0 + ty < K && col < N) ? (double)Bm[(size_t)(k0 + ty) * N + col] : 0.0;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) acc += As[ty][kk] * Bs2[kk][tx];
        __syncthreads();
    }
    if (row < M && col < N) C[(size_t)row * N + col] = (float)acc;
}

}  // namespace scan2
}  // namespace april
