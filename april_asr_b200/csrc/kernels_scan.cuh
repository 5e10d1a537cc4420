// Persistent tensor-core LSTM scan (E2 recurrent part): one launch walks all T time steps of one
// layer for a group of <= 64 sessions. The step-to-step dependency makes this latency-bound, so the
// design goal is to touch as little as possible per step:
//   * recurrent weights never leave shared memory: each of the H/16 "gate CTAs" keeps its 64 gate rows
//     of W_hh (fp16 hi and lo stacked as 128 UMMA rows, 128 KB), each of the (D/32) x BSP "projection
//     CTAs" keeps 32 rows of W_hr (hi+lo stacked as 64 rows, 128 KB);
//   * per step only the activations move: h (B x D fp16 hi/lo) to the gate CTAs, m (B x H) to the
//     projection CTAs, both by TMA from L2, and the precomputed input gates G[t] (fp32) prefetched one
//     step ahead;
//   * the two CTA groups hand over through two monotonic counters in global memory (release/acquire),
//     two handovers per step.
// Numerics: A = [W_hi ; W_lo] stacked along M, B = x_hi then x_lo along K, fp32 accumulation in TMEM;
// adding TMEM lanes r and r + (rows/2) gives (W_hi + W_lo) (x_hi + x_lo), i.e. all four partial
// products of the fp16x2 split (one more than the GEMM tier's three).
// Cell math follows icefall lstmp.py / torch LSTM gate order i,f,g,o [external], as restated in
// oracle/fake_ort and kernels_simt.cuh (lstm_gates_kernel).
#pragma once
#include "kernels_tc.cuh"

namespace april {
namespace scan {

using namespace tc;

constexpr int THREADS = 192;              // warp 0: TMA producer + grid handover, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int UNITS = 16;                 // hidden units per gate CTA -> 64 gate rows (x2 stacked = 128 UMMA rows)
constexpr int PROWS = 32;                 // W_hr rows per projection CTA (x2 stacked = 64 real UMMA rows)
constexpr int NSTAGE = 6;

struct ScanArgs {
    int T;                 // steps in this launch
    int t0;                // absolute index of the first step (for the ragged mask)
    int B, Bp;             // sessions of this group, padded to a multiple of 16 (<= 64)
    int Bs, g0;            // sessions in the whole pass (row stride of G / Hout) and this group's first session
    int D, H;
    int GC;                // gate CTAs = H / 16
    int RS, BSP, Np;       // projection CTAs = RS x BSP, Np = Bp / BSP sessions each
    float inv_scale_hh, inv_scale_hr;
    const int* nsteps;     // [B]   (all per-session pointers are pre-offset to the group)
    float* c_state;        // [B, H] fp32 (read at start, written at end)
    float* h_state;        // [B, D] fp32 (written every live step)
    __half* h_hi; __half* h_lo;     // [Bp, D] exchange buffers (initialised by the caller from h_state)
    __half* m_hi; __half* m_lo;     // [Bp, H]
    float* Hout;           // [T, Bs, D] layer output rows of this launch, pre-offset by g0 * D
    unsigned* counters;    // [2]: gates done, projections done (zeroed by the caller)
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
    const long long t_start = clock64();
    unsigned spin = 0;
    while (ld_acquire(p) < target) {
        if ((++spin & 0xFFF) == 0 && clock64() - t_start > 4000000000LL) { printf("libapril: scan handover timeout (block %d)\n", (int)blockIdx.x); __trap(); }
    }
    asm volatile("fence.proxy.async;" ::: "memory");      // later TMA (async proxy) reads must see the peer's generic-proxy stores
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// SW128 K-major descriptor for a tile whose 8-row groups are 1024 B apart (same as the GEMM tier)
__device__ __forceinline__ uint64_t desc_at(uint32_t addr) { return make_sw128_desc(addr); }

__global__ void __launch_bounds__(THREADS, 1)
lstm_scan_tc_kernel(const __grid_constant__ CUtensorMap mapWg,   // [GC*128, D]  stacked W_hh slices
                    const __grid_constant__ CUtensorMap mapWp,   // [RS*64, H]   stacked W_hr slices
                    const __grid_constant__ CUtensorMap mapHhi, const __grid_constant__ CUtensorMap mapHlo,   // [Bp, D], box {64, Bp}
                    const __grid_constant__ CUtensorMap mapMhi, const __grid_constant__ CUtensorMap mapMlo,   // [Bp, H], box {64, Np}
                    const __grid_constant__ CUtensorMap mapG,    // fp32 [T*B (+pad), 4H], box {16, Bp}
                    ScanArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_gate = (int)blockIdx.x < a.GC;

    // ---- shared memory carve-up ---------------------------------------------------------------------
    // [0, WBYTES)            resident weight tile
    // [.., +NSTAGE*SB)       activation stages (aliased by the epilogue exchange buffer)
    // [.., +2*GB)            G prefetch double buffer (gate CTAs)
    // [.., +CB)              c state (gate CTAs)
    // barriers
    const int KB_W = is_gate ? a.D / 64 : a.H / 64;                       // k-blocks of the resident tile == k-blocks per activation half
    const uint32_t w_kb_bytes = is_gate ? 128u * 128u : 64u * 128u;      // bytes per k-block of the weight tile
    const uint32_t WBYTES = (uint32_t)KB_W * w_kb_bytes;
    const int Nact = is_gate ? a.Bp : a.Np;                              // UMMA N
    const uint32_t SB = (uint32_t)Nact * 128u;                           // bytes per activation stage ([N x 64] fp16)
    // (a projection CTA's 128-row UMMA over-reads 64 rows past each 64-row k-block: they land in the next k-block /
    //  the stage region, and the TMEM lanes they produce, 64-127, are never read)
    const uint32_t stage0 = base + ((WBYTES + 1023u) & ~1023u);
    const uint32_t exch_bytes = 128u * (uint32_t)(a.Bp + 1) * 4u;
    const uint32_t stage_region = (NSTAGE * SB > exch_bytes ? NSTAGE * SB : exch_bytes);
    const uint32_t gbuf0 = (stage0 + stage_region + 1023u) & ~1023u;
    const uint32_t GB = (uint32_t)a.Bp * 16u * 4u * 4u;                  // 4 gates x [Bp x 16] fp32
    const uint32_t cbuf = gbuf0 + 2u * GB;
    const uint32_t CB = (uint32_t)a.B * UNITS * 4u;
    const uint32_t bar0 = (cbuf + CB + 15u) & ~15u;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    const uint32_t wfull_bar = bar0 + 8u * (2 * NSTAGE);
    const uint32_t tfull_bar = wfull_bar + 8u, tempty_bar = wfull_bar + 16u;
    const uint32_t gfull_bar0 = wfull_bar + 24u;                         // 2 barriers
    const uint32_t tmem_slot = wfull_bar + 40u;
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    float* exch = reinterpret_cast<float*>(gptr(stage0));
    float* cst = reinterpret_cast<float*>(gptr(cbuf));

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(wfull_bar, 1); mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 1);
        mbar_init(gfull_bar0, 1); mbar_init(gfull_bar0 + 8u, 1);
        fence_barrier_init();
    }
    const uint32_t tmem_cols = Nact <= 32 ? 32u : 64u;
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));
    const uint32_t idesc = make_idesc_f16(128, Nact);

    if (is_gate) {
        // =============================================================================================
        // gate CTA: units [u0, u0+16): gates = G[t] + (W_hh h) ; c' ; m
        // =============================================================================================
        const int u0 = blockIdx.x * UNITS;
        if (warp == 0) {
            if (lane == 0) {
                // resident weights
                mbar_expect_tx(wfull_bar, WBYTES);
                for (int kb = 0; kb < KB_W; kb++) tma_load_2d(base + kb * w_kb_bytes, &mapWg, wfull_bar, kb * 64, blockIdx.x * 128);
                // G for the first step
                mbar_expect_tx(gfull_bar0, GB);
                for (int g4 = 0; g4 < 4; g4++) tma_load_2d(gbuf0 + g4 * (GB / 4), &mapG, gfull_bar0, g4 * a.H + u0, a.g0);
                uint32_t it = 0;
                for (int t = 0; t < a.T; t++) {
                    if (t > 0) wait_counter(a.counters + 1, (unsigned)(t * a.RS * a.BSP));      // h(t-1) complete (=> every gate epilogue of t-1 is done)
                    if (t + 1 < a.T) {                                   // prefetch next step's input gates into the buffer step t-1 used
                        const uint32_t gb = gfull_bar0 + 8u * ((t + 1) & 1);
                        mbar_expect_tx(gb, GB);
                        for (int g4 = 0; g4 < 4; g4++) tma_load_2d(gbuf0 + ((t + 1) & 1) * GB + g4 * (GB / 4), &mapG, gb, g4 * a.H + u0, (t + 1) * a.Bs + a.g0);
                    }
                    for (int half = 0; half < 2; half++)
                        for (int kb = 0; kb < KB_W; kb++, it++) {
                            const int s = it % NSTAGE; const uint32_t ph = (it / NSTAGE) & 1;
                            mbar_wait(empty_bar(s), ph ^ 1);
                            mbar_expect_tx(full_bar(s), SB);
                            tma_load_2d(stage0 + s * SB, half ? &mapHlo : &mapHhi, full_bar(s), kb * 64, 0);
                        }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                mbar_wait(wfull_bar, 0);
                uint32_t it = 0;
                for (int t = 0; t < a.T; t++) {
                    mbar_wait(tempty_bar, (t & 1) ^ 1);
                    tc_fence_after();
                    for (int half = 0; half < 2; half++)
                        for (int kb = 0; kb < KB_W; kb++, it++) {
                            const int s = it % NSTAGE; const uint32_t ph = (it / NSTAGE) & 1;
                            mbar_wait(full_bar(s), ph);
                            tc_fence_after();
                            const uint64_t ad = desc_at(base + kb * w_kb_bytes), bd = desc_at(stage0 + s * SB);
#pragma unroll
                            for (int k = 0; k < 4; k++) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (half | kb | k) != 0);
                            umma_commit(empty_bar(s));
                        }
                    umma_commit(tfull_bar);
                }
            }
        } else {
            const int q = warp & 3, et = (warp - 2) * 32 + lane;          // et: 0..127
            // c state for this CTA: cst[b * 16 + u]
            for (int i = et; i < a.B * UNITS; i += 128) cst[i] = a.c_state[(size_t)(i / UNITS) * a.H + u0 + (i % UNITS)];
            const int ld = a.Bp + 1;
            for (int t = 0; t < a.T; t++) {
                mbar_wait(tfull_bar, t & 1);
                tc_fence_after();
                // TMEM lane r = q*32 + lane ; columns = sessions
                for (int c0 = 0; c0 < a.Bp; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + c0 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
                    for (int j = 0; j < 32; j++) if (c0 + j < a.Bp) exch[(q * 32 + lane) * ld + c0 + j] = v[j];
                }
                tc_fence_before();
                epi_bar_sync();
                if (et == 0) mbar_arrive(tempty_bar);                    // accumulator drained: next step's MMAs may overwrite it
                mbar_wait(gfull_bar0 + 8u * (t & 1), (t >> 1) & 1);
                const float* gb = reinterpret_cast<const float*>(gptr(gbuf0 + (t & 1) * GB));
                for (int i = et; i < a.B * UNITS; i += 128) {
                    const int b = i / UNITS, u = i % UNITS;
                    if (a.t0 + t >= a.nsteps[b]) continue;               // ragged: finished sessions keep their state
                    float gate[4];
#pragma unroll
                    for (int g4 = 0; g4 < 4; g4++)
                        gate[g4] = (exch[(g4 * 16 + u) * ld + b] + exch[(64 + g4 * 16 + u) * ld + b]) * a.inv_scale_hh + gb[(g4 * a.Bp + b) * 16 + u];
                    const float cn = fast_sigmoid(gate[1]) * cst[i] + fast_sigmoid(gate[0]) * fast_tanh(gate[2]);
                    cst[i] = cn;
                    const float m = fast_sigmoid(gate[3]) * fast_tanh(cn);
                    __half mh, ml; split_f16(m, mh, ml);
                    a.m_hi[(size_t)b * a.H + u0 + u] = mh;
                    a.m_lo[(size_t)b * a.H + u0 + u] = ml;
                }
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
                epi_bar_sync();
                if (et == 0) red_release(a.counters + 0);                // m(t) slice published
            }
            for (int i = et; i < a.B * UNITS; i += 128) a.c_state[(size_t)(i / UNITS) * a.H + u0 + (i % UNITS)] = cst[i];
        }
    } else {
        // =============================================================================================
        // projection CTA: rows [r0, r0+32) of h' = W_hr m for sessions [b0, b0+Np)
        // =============================================================================================
        const int pid = blockIdx.x - a.GC;
        const int rs = pid % a.RS, bs = pid / a.RS;
        const int r0 = rs * PROWS, b0 = bs * a.Np;
        if (warp == 0) {
            if (lane == 0) {
                mbar_expect_tx(wfull_bar, WBYTES);
                for (int kb = 0; kb < KB_W; kb++) tma_load_2d(base + kb * w_kb_bytes, &mapWp, wfull_bar, kb * 64, rs * 64);
                uint32_t it = 0;
                for (int t = 0; t < a.T; t++) {
                    wait_counter(a.counters + 0, (unsigned)((t + 1) * a.GC));                   // m(t) complete
                    for (int half = 0; half < 2; half++)
                        for (int kb = 0; kb < KB_W; kb++, it++) {
                            const int s = it % NSTAGE; const uint32_t ph = (it / NSTAGE) & 1;
                            mbar_wait(empty_bar(s), ph ^ 1);
                            mbar_expect_tx(full_bar(s), SB);
                            tma_load_2d(stage0 + s * SB, half ? &mapMlo : &mapMhi, full_bar(s), kb * 64, b0);
                        }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                mbar_wait(wfull_bar, 0);
                uint32_t it = 0;
                for (int t = 0; t < a.T; t++) {
                    mbar_wait(tempty_bar, (t & 1) ^ 1);
                    tc_fence_after();
                    for (int half = 0; half < 2; half++)
                        for (int kb = 0; kb < KB_W; kb++, it++) {
                            const int s = it % NSTAGE; const uint32_t ph = (it / NSTAGE) & 1;
                            mbar_wait(full_bar(s), ph);
                            tc_fence_after();
                            const uint64_t ad = desc_at(base + kb * w_kb_bytes), bd = desc_at(stage0 + s * SB);
#pragma unroll
                            for (int k = 0; k < 4; k++) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (half | kb | k) != 0);
                            umma_commit(empty_bar(s));
                        }
                    umma_commit(tfull_bar);
                }
            }
        } else {
            const int q = warp & 3, et = (warp - 2) * 32 + lane;
            const int ld = a.Np + 1;
            for (int t = 0; t < a.T; t++) {
                mbar_wait(tfull_bar, t & 1);
                tc_fence_after();
                if (q < 2) {                                             // lanes 0-31: W_hi rows, 32-63: W_lo rows; 64-127 unused
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
                    for (int j = 0; j < 32; j++) if (j < a.Np) exch[(q * 32 + lane) * ld + j] = v[j];
                }
                tc_fence_before();
                epi_bar_sync();
                if (et == 0) mbar_arrive(tempty_bar);
                // 32 rows x Np sessions; consecutive threads -> consecutive rows (coalesced global stores)
                for (int i = et; i < PROWS * a.Np; i += 128) {
                    const int r = i % PROWS, j = i / PROWS, b = b0 + j;
                    if (b >= a.B) continue;
                    const bool live = a.t0 + t < a.nsteps[b];
                    const float hv = (exch[r * ld + j] + exch[(32 + r) * ld + j]) * a.inv_scale_hr;
                    float* ho = a.Hout + ((size_t)t * a.Bs + b) * a.D + r0 + r;
                    if (live) {
                        *ho = hv;
                        a.h_state[(size_t)b * a.D + r0 + r] = hv;
                        __half hh, hl; split_f16(hv, hh, hl);
                        a.h_hi[(size_t)b * a.D + r0 + r] = hh;
                        a.h_lo[(size_t)b * a.D + r0 + r] = hl;
                    } else {
                        *ho = 0.f;
                    }
                }
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
                epi_bar_sync();
                if (et == 0) red_release(a.counters + 1);                // h(t) slice published
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// fp32 state [L? no: B, D] -> padded fp16 (hi, lo) exchange buffers [Bp, D] (rows >= B zeroed)
__global__ void scan_prepare_kernel(const float* __restrict__ h, __half* __restrict__ hi, __half* __restrict__ lo, int B, int Bp, int D) {
    const int n = Bp * D;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = i / D;
        float v = b < B ? h[i] : 0.f;
        __half a, c; split_f16(v, a, c);
        hi[i] = a; lo[i] = c;
    }
}

}  // namespace scan
}  // namespace april
