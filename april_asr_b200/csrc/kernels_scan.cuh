// Persistent tensor-core LSTM scan (E2 recurrent part): one launch walks all T time steps of one
// layer for up to 64 sessions, organised as NG <= 2 independent groups of <= 32 sessions that are
// interleaved through the same resident weights. The step-to-step dependency makes this
// latency-bound, so the design goal is to touch as little as possible per step and to hide the
// unavoidable handover latency of one group behind the other group's work:
//   * recurrent weights never leave shared memory: each of the H/16 "gate CTAs" keeps its 64 gate rows
//     of W_hh (fp16 hi and lo stacked as 128 UMMA rows, 128 KB); each projection CTA keeps 32 rows of
//     W_hr (hi+lo stacked as 64 rows, 128 KB) and serves 16 sessions of one group;
//   * per step only the activations move: h (B x D) to the gate CTAs, m (B x H) to the projection
//     CTAs, as fp16 [hi rows ; lo rows] tiles fetched by 4-D TMA boxes (a TMA copy costs ~150-300
//     issue cycles whatever its size, so copies are few and issued from several lanes), and the
//     precomputed input gates G[t] (fp32) prefetched one step ahead;
//   * gate CTAs alternate between the groups: while group 0's m(t) travels to its projection CTAs
//     and h(t) travels back (two release/acquire handovers through global counters, ~1 us each),
//     the gate CTAs work on group 1, and vice versa.
// Numerics: A = [W_hi ; W_lo] stacked along M, B = [x_hi ; x_lo] stacked along N, fp32 accumulation
// in TMEM; adding the four (row half, column half) entries gives (W_hi + W_lo)(x_hi + x_lo), i.e. all
// four partial products of the fp16x2 split (one more than the GEMM tier's three).
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
constexpr int GSTAGES = 3;                // gate CTA activation ring (stages of KBG k-blocks each)
constexpr int PCOPIES = 4;                // projection CTA: m arrives in at most this many TMA copies
constexpr int MAXG = 2;

struct ScanArgs {
    int T;                 // steps in this launch
    int t0;                // absolute index of the first step (for the ragged mask)
    int NG;                // interleaved session groups (1 or 2)
    int Bg[MAXG];          // sessions per group (<= 32)
    int goff[MAXG];        // first session of the group inside this launch's session range
    int Bp;                // padded group size (16 or 32), same for all groups
    int Bs, g0;            // sessions in the whole pass (row stride of G / Hout) and this launch's first session
    int D, H;
    int GC;                // gate CTAs = H / 16
    int RS, BSP;           // projection CTAs per group = RS x BSP (BSP = Bp / 16)
    int KBG;               // k-blocks (of 64) per gate-CTA TMA copy
    int PC;                // TMA copies per projection-CTA step (divides H / 64)
    float inv_scale_hh, inv_scale_hr;
    const int* nsteps;     // [sessions of the launch]   (all per-session pointers are pre-offset to the launch's first session)
    float* c_state;        // [., H] fp32 (read at start, written at end)
    float* h_state;        // [., D] fp32
    __half* h_x[MAXG];     // [2 (hi|lo)][Bp][D] exchange buffers (initialised by the caller from h_state)
    __half* m_x[MAXG];     // [2][Bp][H]
    float* Hout;           // [T, Bs, D] layer output rows of this launch, pre-offset by g0 * D
    unsigned* counters;    // [MAXG][2]: gates done, projections done (zeroed by the caller)
    unsigned long long* trace;   // optional [grid][T][8] globaltimer stamps (debug)
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
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TRACE(slot_) do { if (a.trace) a.trace[((size_t)blockIdx.x * a.T + t) * 8 + (slot_)] = gtimer(); } while (0)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
lstm_scan_tc_kernel(const __grid_constant__ CUtensorMap mapWg,   // 2-D [GC*128, D]  stacked W_hh slices, box {64, 128}
                    const __grid_constant__ CUtensorMap mapWp,   // 2-D [RS*64, H]   stacked W_hr slices, box {64, 64}
                    const __grid_constant__ CUtensorMap mapHx0, const __grid_constant__ CUtensorMap mapHx1,   // 4-D {64, Bp, 2 (hi|lo), D/64}, box {64, Bp, 2, KBG}
                    const __grid_constant__ CUtensorMap mapMx0, const __grid_constant__ CUtensorMap mapMx1,   // 4-D {64, Bp, 2, H/64}, box {64, 16, 2, H/64/PCOPIES}
                    const __grid_constant__ CUtensorMap mapG,    // 2-D fp32 [T*Bs, 4H], box {16, Bp}
                    ScanArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_gate = (int)blockIdx.x < a.GC;
    const int NG = a.NG;

    if (is_gate) {
        // =============================================================================================
        // gate CTA: units [u0, u0+16) of every group: gates = G[t] + (W_hh h) ; c' ; m
        // =============================================================================================
        const int KB_W = a.D / 64;
        const uint32_t w_kb_bytes = 128u * 128u, WBYTES = (uint32_t)KB_W * w_kb_bytes;
        const int Nact = 2 * a.Bp;                                        // UMMA N: [h_hi rows ; h_lo rows]
        const uint32_t KBB = (uint32_t)Nact * 128u, SB = (uint32_t)a.KBG * KBB;
        const uint32_t ring0 = base + WBYTES;
        const uint32_t exch0 = ring0 + GSTAGES * SB;
        const int ld = a.Bp + 1;
        const uint32_t gbuf0 = (exch0 + 128u * (uint32_t)ld * 4u + 127u) & ~127u;
        const uint32_t GB = (uint32_t)a.Bp * 16u * 4u * 4u;               // 4 gates x [Bp x 16] fp32, per group
        const uint32_t cbuf = gbuf0 + MAXG * GB;
        const uint32_t CB = (uint32_t)a.Bp * UNITS * 4u;                  // per group
        const uint32_t bar0 = (cbuf + MAXG * CB + 15u) & ~15u;
        auto full_bar = [&](int s) { return bar0 + 8u * s; };
        auto empty_bar = [&](int s) { return bar0 + 8u * (GSTAGES + s); };
        const uint32_t wfull_bar = bar0 + 8u * (2 * GSTAGES);
        auto tfull_bar = [&](int g) { return wfull_bar + 8u + 8u * g; };
        auto tempty_bar = [&](int g) { return wfull_bar + 8u + 8u * (MAXG + g); };
        auto gfull_bar = [&](int g) { return wfull_bar + 8u + 8u * (2 * MAXG + g); };
        auto gfree_bar = [&](int g) { return wfull_bar + 8u + 8u * (3 * MAXG + g); };
        const uint32_t tmem_slot = wfull_bar + 8u + 8u * (4 * MAXG);
        float* exch = reinterpret_cast<float*>(gptr(exch0));

        if (threadIdx.x == 0) {
            for (int s = 0; s < GSTAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            mbar_init(wfull_bar, 1);
            for (int g = 0; g < MAXG; g++) { mbar_init(tfull_bar(g), 1); mbar_init(tempty_bar(g), 1); mbar_init(gfull_bar(g), 1); mbar_init(gfree_bar(g), 1); }
            fence_barrier_init();
        }
        const uint32_t tmem_cols = (uint32_t)(NG * Nact) <= 32u ? 32u : ((uint32_t)(NG * Nact) <= 64u ? 64u : 128u);
        if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));
        const uint32_t idesc = make_idesc_f16(128, Nact);
        const int u0 = blockIdx.x * UNITS;
        const int copies = KB_W / a.KBG;                                  // TMA copies per (step, group)
        const unsigned nproj = (unsigned)(a.RS * a.BSP);                  // projection CTAs per group

        if (warp == 0) {
            if (lane == 0) {
                tma_prefetch_desc(&mapHx0); tma_prefetch_desc(&mapHx1); tma_prefetch_desc(&mapG);
                mbar_expect_tx(wfull_bar, WBYTES);                        // resident weights
                for (int kb = 0; kb < KB_W; kb++) tma_load_2d(base + kb * w_kb_bytes, &mapWg, wfull_bar, kb * 64, blockIdx.x * 128);
            }
            if (lane < GSTAGES) {
                // lane l feeds ring stage l: copies whose running index is congruent to l. The TMA issue cost is per thread.
                const long long total = (long long)a.T * NG * copies;
                int opened_t = -1, opened_g = -1;
                for (long long idx = lane; idx < total; idx += GSTAGES) {
                    const int j = (int)(idx % copies);
                    const int tg = (int)(idx / copies), g = tg % NG, t = tg / NG;
                    if (t != opened_t || g != opened_g) {                 // first copy of this lane for (t, g): h(t-1) of the group must be complete
                        if (t > 0) wait_counter(a.counters + g * 2 + 1, (unsigned)t * nproj);
                        opened_t = t; opened_g = g;
                        if (j == 0) TRACE(0);
                    }
                    const uint32_t use = (uint32_t)(idx / GSTAGES);
                    mbar_wait(empty_bar(lane), (use & 1) ^ 1);
                    mbar_expect_tx(full_bar(lane), SB);
                    tma_load_4d(ring0 + lane * SB, g ? &mapHx1 : &mapHx0, full_bar(lane), 0, 0, 0, j * a.KBG);
                    if (j == copies - 1) TRACE(1);
                }
            } else if (lane == GSTAGES) {
                // input gates G[t] of each group, one step ahead of the epilogue
                for (int t = 0; t < a.T; t++)
                    for (int g = 0; g < NG; g++) {
                        if (t > 0) mbar_wait(gfree_bar(g), (t - 1) & 1);
                        mbar_expect_tx(gfull_bar(g), GB);
                        for (int g4 = 0; g4 < 4; g4++)
                            tma_load_2d(gbuf0 + g * GB + g4 * (GB / 4), &mapG, gfull_bar(g), g4 * a.H + u0, t * a.Bs + a.g0 + a.goff[g]);
                    }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                mbar_wait(wfull_bar, 0);
                long long idx = 0;
                for (int t = 0; t < a.T; t++)
                    for (int g = 0; g < NG; g++) {
                        mbar_wait(tempty_bar(g), (t & 1) ^ 1);
                        tc_fence_after();
                        for (int j = 0; j < copies; j++, idx++) {
                            const int s = (int)(idx % GSTAGES);
                            const uint32_t use = (uint32_t)(idx / GSTAGES);
                            mbar_wait(full_bar(s), use & 1);
                            if (j == 0) TRACE(2);
                            tc_fence_after();
                            for (int kk = 0; kk < a.KBG; kk++) {
                                const uint64_t ad = make_sw128_desc(base + (j * a.KBG + kk) * w_kb_bytes), bd = make_sw128_desc(ring0 + s * SB + kk * KBB);
#pragma unroll
                                for (int k = 0; k < 4; k++) umma_f16(tmem_base + g * Nact, ad + 2 * k, bd + 2 * k, idesc, (j | kk | k) != 0);
                            }
                            umma_commit(empty_bar(s));
                        }
                        umma_commit(tfull_bar(g));
                        TRACE(3);
                    }
            }
        } else {
            const int q = warp & 3, et = (warp - 2) * 32 + lane;           // et: 0..127
            float* cst = reinterpret_cast<float*>(gptr(cbuf));
            // this thread's work item in the cell phase: session b = et / 4 of the group, units uq*4 .. uq*4+3
            const int my_b = et >> 2, uq = et & 3;
            int my_n[MAXG];
            for (int g = 0; g < MAXG; g++) {
                my_n[g] = (g < NG && my_b < a.Bg[g]) ? a.nsteps[a.goff[g] + my_b] : 0;
                if (g < NG && my_b < a.Bp) {
#pragma unroll
                    for (int uu = 0; uu < 4; uu++)
                        cst[g * (a.Bp * UNITS) + my_b * UNITS + uq * 4 + uu] = my_b < a.Bg[g] ? a.c_state[(size_t)(a.goff[g] + my_b) * a.H + u0 + uq * 4 + uu] : 0.f;
                }
            }
            for (int t = 0; t < a.T; t++)
                for (int g = 0; g < NG; g++) {
                    mbar_wait(tfull_bar(g), t & 1);
                    if (et == 0) TRACE(4);
                    tc_fence_after();
                    // TMEM lane r = q*32 + lane ; columns [0,Bp) = W.h_hi, [Bp,2Bp) = W.h_lo per session: pre-add the two
                    {
                        float v[32];
                        tmem_ld32(tmem_base + g * Nact + ((uint32_t)(q * 32) << 16), v);
                        if (a.Bp == 16) {
#pragma unroll
                            for (int j = 0; j < 16; j++) exch[(q * 32 + lane) * ld + j] = v[j] + v[16 + j];
                        } else {                                          // Bp == 32
                            float w2[32];
                            tmem_ld32(tmem_base + g * Nact + 32 + ((uint32_t)(q * 32) << 16), w2);
#pragma unroll
                            for (int j = 0; j < 32; j++) exch[(q * 32 + lane) * ld + j] = v[j] + w2[j];
                        }
                    }
                    tc_fence_before();
                    epi_bar_sync();
                    if (et == 0) { mbar_arrive(tempty_bar(g)); TRACE(5); } // accumulator drained: the group's next step may overwrite it
                    mbar_wait(gfull_bar(g), t & 1);
                    if (my_b < a.Bp) {
                        const float* gb = reinterpret_cast<const float*>(gptr(gbuf0 + g * GB));
                        float* cs = cst + g * (a.Bp * UNITS) + my_b * UNITS + uq * 4;
                        const bool live = a.t0 + t < my_n[g];             // ragged: finished / padding sessions keep their state
                        float mval[4];
#pragma unroll
                        for (int uu = 0; uu < 4; uu++) {
                            const int u = uq * 4 + uu;
                            float gate[4];
#pragma unroll
                            for (int g4 = 0; g4 < 4; g4++)
                                gate[g4] = (exch[(g4 * 16 + u) * ld + my_b] + exch[(64 + g4 * 16 + u) * ld + my_b]) * a.inv_scale_hh + gb[(g4 * a.Bp + my_b) * 16 + u];
                            const float cn = fast_sigmoid(gate[1]) * cs[uu] + fast_sigmoid(gate[0]) * fast_tanh(gate[2]);
                            if (live) cs[uu] = cn;
                            mval[uu] = fast_sigmoid(gate[3]) * fast_tanh(cn);
                        }
                        if (live) {
                            __align__(8) __half mh[4], ml[4];
#pragma unroll
                            for (int uu = 0; uu < 4; uu++) split_f16(mval[uu], mh[uu], ml[uu]);
                            __half* mx = a.m_x[g];
                            *reinterpret_cast<uint2*>(mx + (size_t)my_b * a.H + u0 + uq * 4) = *reinterpret_cast<const uint2*>(mh);
                            *reinterpret_cast<uint2*>(mx + (size_t)(a.Bp + my_b) * a.H + u0 + uq * 4) = *reinterpret_cast<const uint2*>(ml);
                        }
                    }
                    if (et == 0) TRACE(6);
                    asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy stores -> visible to the peers' TMA reads
                    epi_bar_sync();                                      // CTA-scope ordering; the release below is cumulative over it
                    if (et == 0) { red_release(a.counters + g * 2 + 0); mbar_arrive(gfree_bar(g)); TRACE(7); }   // m(t) slice published; G buffer reusable
                }
            for (int g = 0; g < NG; g++)
                if (my_b < a.Bg[g]) {
#pragma unroll
                    for (int uu = 0; uu < 4; uu++) a.c_state[(size_t)(a.goff[g] + my_b) * a.H + u0 + uq * 4 + uu] = cst[g * (a.Bp * UNITS) + my_b * UNITS + uq * 4 + uu];
                }
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
    } else {
        // =============================================================================================
        // projection CTA of group g: rows [r0, r0+32) of h' = W_hr m for sessions [b0, b0+16) of the group
        // =============================================================================================
        const int pid = blockIdx.x - a.GC;
        const int per_group = a.RS * a.BSP;
        const int g = pid / per_group, rs = (pid % per_group) % a.RS, bs = (pid % per_group) / a.RS;
        const int r0 = rs * PROWS, b0 = bs * 16;
        const int KB_W = a.H / 64;
        const uint32_t w_kb_bytes = 64u * 128u, WBYTES = (uint32_t)KB_W * w_kb_bytes;
        const int Nact = 32;                                              // [m_hi rows (16) ; m_lo rows (16)]
        const uint32_t KBB = (uint32_t)Nact * 128u;
        const int kb_per_copy = KB_W / a.PC;
        const uint32_t SB = (uint32_t)kb_per_copy * KBB;
        // (the 128-row UMMA over-reads 64 rows past each 64-row k-block of the weight tile: they land in the next k-block /
        //  the activation buffer, and the TMEM lanes they produce, 64-127, are never read)
        const uint32_t act0 = base + WBYTES;
        const uint32_t exch0 = act0 + (uint32_t)KB_W * KBB;
        const int ld = 17;
        const uint32_t bar0 = (exch0 + 64u * (uint32_t)ld * 4u + 15u) & ~15u;
        auto full_bar = [&](int s) { return bar0 + 8u * s; };             // PCOPIES
        const uint32_t wfull_bar = bar0 + 8u * PCOPIES, tfull_bar = wfull_bar + 8u, tempty_bar = wfull_bar + 16u, tmem_slot = wfull_bar + 24u;
        float* exch = reinterpret_cast<float*>(gptr(exch0));
        const CUtensorMap* mapMx = g ? &mapMx1 : &mapMx0;

        if (threadIdx.x == 0) {
            for (int s = 0; s < PCOPIES; s++) mbar_init(full_bar(s), 1);
            mbar_init(wfull_bar, 1); mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 1);
            fence_barrier_init();
        }
        if (warp == 1) tmem_alloc(tmem_slot, 32);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));
        const uint32_t idesc = make_idesc_f16(128, Nact);

        if (warp == 0) {
            if (lane == 0) {
                tma_prefetch_desc(mapMx);
                mbar_expect_tx(wfull_bar, WBYTES);
                for (int kb = 0; kb < KB_W; kb++) tma_load_2d(base + kb * w_kb_bytes, &mapWp, wfull_bar, kb * 64, rs * 64);
            }
            __syncwarp();
            if (lane < a.PC) {
                for (int t = 0; t < a.T; t++) {
                    wait_counter(a.counters + g * 2 + 0, (unsigned)((t + 1) * a.GC));   // m(t) of my group complete
                    if (lane == 0) TRACE(0);
                    // the previous step's MMAs are long finished (this CTA's epilogue of t-1 ran before the gate CTAs could produce m(t))
                    mbar_expect_tx(full_bar(lane), SB);
                    tma_load_4d(act0 + lane * SB, mapMx, full_bar(lane), 0, b0, 0, lane * kb_per_copy);
                    if (lane == 0) TRACE(1);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                mbar_wait(wfull_bar, 0);
                for (int t = 0; t < a.T; t++) {
                    mbar_wait(tempty_bar, (t & 1) ^ 1);
                    tc_fence_after();
                    for (int c = 0; c < a.PC; c++) {
                        mbar_wait(full_bar(c), t & 1);
                        if (c == 0) TRACE(2);
                        tc_fence_after();
                        for (int kk = 0; kk < kb_per_copy; kk++) {
                            const int kb = c * kb_per_copy + kk;
                            const uint64_t ad = make_sw128_desc(base + kb * w_kb_bytes), bd = make_sw128_desc(act0 + kb * KBB);
#pragma unroll
                            for (int k = 0; k < 4; k++) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(tfull_bar);
                    TRACE(3);
                }
            }
        } else {
            const int q = warp & 3, et = (warp - 2) * 32 + lane;
            // this thread's outputs: row r = et % 32, sessions j = et / 32 + 4 k
            const int my_r = et % PROWS, my_j0 = et / PROWS;
            int my_n[4];
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) { const int b = b0 + my_j0 + 4 * k4; my_n[k4] = b < a.Bg[g] ? a.nsteps[a.goff[g] + b] : -1; }
            __half* hx = a.h_x[g];
            for (int t = 0; t < a.T; t++) {
                mbar_wait(tfull_bar, t & 1);
                if (et == 0) TRACE(4);
                tc_fence_after();
                if (q < 2) {                                              // lanes 0-31: W_hi rows, 32-63: W_lo rows; 64-127 unused
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
                    for (int j = 0; j < 16; j++) exch[(q * 32 + lane) * ld + j] = v[j] + v[16 + j];      // cols [0,16) = m_hi, [16,32) = m_lo
                }
                tc_fence_before();
                epi_bar_sync();
                if (et == 0) { mbar_arrive(tempty_bar); TRACE(5); }
                // 32 rows x 16 sessions; consecutive threads -> consecutive rows (coalesced global stores)
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) {
                    if (my_n[k4] < 0) continue;                           // padding session
                    const int j = my_j0 + 4 * k4, b = b0 + j;
                    const float hv = (exch[my_r * ld + j] + exch[(32 + my_r) * ld + j]) * a.inv_scale_hr;
                    const int sess = a.goff[g] + b;
                    float* ho = a.Hout + ((size_t)t * a.Bs + sess) * a.D + r0 + my_r;
                    if (a.t0 + t < my_n[k4]) {
                        *ho = hv;
                        if (a.t0 + t == my_n[k4] - 1 || t == a.T - 1) a.h_state[(size_t)sess * a.D + r0 + my_r] = hv;   // state survives the launch
                        __half hh, hl; split_f16(hv, hh, hl);
                        hx[(size_t)b * a.D + r0 + my_r] = hh;
                        hx[(size_t)(a.Bp + b) * a.D + r0 + my_r] = hl;
                    } else {
                        *ho = 0.f;
                    }
                }
                if (et == 0) TRACE(6);
                asm volatile("fence.proxy.async;" ::: "memory");
                epi_bar_sync();
                if (et == 0) { red_release(a.counters + g * 2 + 1); TRACE(7); }   // h(t) slice published
            }
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_base, 32);
    }
}

// fp32 state [B, D] -> padded fp16 exchange buffer [2 (hi|lo)][Bp][D] (rows >= B zeroed)
__global__ void scan_prepare_kernel(const float* __restrict__ h, __half* __restrict__ hx, int B, int Bp, int D) {
    const int n = Bp * D;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = i / D;
        float v = b < B ? h[i] : 0.f;
        __half a, c; split_f16(v, a, c);
        hx[i] = a; hx[(size_t)n + i] = c;
    }
}

}  // namespace scan
}  // namespace april
