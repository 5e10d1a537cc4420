// Persistent tensor-core LSTM scan (E2 recurrent part): one launch walks all T time steps of one
// layer for a group of <= 64 sessions. The step-to-step dependency makes this latency-bound, so the
// design goal is to touch as little as possible per step:
//   * recurrent weights never leave shared memory: each of the H/16 "gate CTAs" keeps its 64 gate rows
//     of W_hh (fp16 hi and lo stacked as 128 UMMA rows, 128 KB), each of the (D/32) x BSP "projection
//     CTAs" keeps 32 rows of W_hr (hi+lo stacked as 64 rows, 128 KB);
//   * per step only the activations move: h (B x D fp16 hi/lo) to the gate CTAs, m (B x H) to the
//     projection CTAs, by 3-D TMA boxes (several 64-wide k-blocks per copy: a TMA copy costs ~150-300
//     issue cycles whatever its size — profiles/r01_tma_probe.md), and the precomputed input gates
//     G[t] (fp32) prefetched one step ahead;
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
constexpr int GSTAGES = 4;                // gate CTA activation ring (stages of KBG k-blocks each)

struct ScanArgs {
    int T;                 // steps in this launch
    int t0;                // absolute index of the first step (for the ragged mask)
    int B, Bp;             // sessions of this group, padded to a multiple of 16 (<= 64)
    int Bs, g0;            // sessions in the whole pass (row stride of G / Hout) and this group's first session
    int D, H;
    int GC;                // gate CTAs = H / 16
    int RS, BSP, Np;       // projection CTAs = RS x BSP, Np = Bp / BSP sessions each
    int KBG;               // k-blocks (of 64) per gate-CTA TMA copy
    float inv_scale_hh, inv_scale_hr;
    const int* nsteps;     // [B]   (all per-session pointers are pre-offset to the group)
    float* c_state;        // [B, H] fp32 (read at start, written at end)
    float* h_state;        // [B, D] fp32 (written every live step)
    __half* h_hi; __half* h_lo;     // [Bp, D] exchange buffers (initialised by the caller from h_state)
    __half* m_hi; __half* m_lo;     // [Bp, H]
    float* Hout;           // [T, Bs, D] layer output rows of this launch, pre-offset by g0 * D
    unsigned* counters;    // [2]: gates done, projections done (zeroed by the caller)
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

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
lstm_scan_tc_kernel(const __grid_constant__ CUtensorMap mapWg,   // 2-D [GC*128, D]  stacked W_hh slices, box {64, 128}
                    const __grid_constant__ CUtensorMap mapWp,   // 2-D [RS*64, H]   stacked W_hr slices, box {64, 64}
                    const __grid_constant__ CUtensorMap mapHhi, const __grid_constant__ CUtensorMap mapHlo,   // 3-D {64, Bp, D/64}, box {64, Bp, KBG}
                    const __grid_constant__ CUtensorMap mapMhi, const __grid_constant__ CUtensorMap mapMlo,   // 3-D {64, Bp, H/64}, box {64, Np, H/64}
                    const __grid_constant__ CUtensorMap mapG,    // 2-D fp32 [T*Bs, 4H], box {16, Bp}
                    ScanArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_gate = (int)blockIdx.x < a.GC;

    // ---- shared memory carve-up ---------------------------------------------------------------------
    // [0, WBYTES)            resident weight tile (k-block kb at kb * w_kb_bytes)
    // [.., +ACT)             activation buffer: gate = GSTAGES stages of KBG k-blocks, projection = both halves of m
    //                        (aliased by the epilogue exchange buffer)
    // [.., +GB)              G buffer (gate CTAs)      [.., +CB)  c state (gate CTAs)      barriers
    const int KB_W = is_gate ? a.D / 64 : a.H / 64;                       // k-blocks of the resident tile == k-blocks per activation half
    const uint32_t w_kb_bytes = is_gate ? 128u * 128u : 64u * 128u;      // bytes per k-block of the weight tile
    const uint32_t WBYTES = (uint32_t)KB_W * w_kb_bytes;
    const int Nact = is_gate ? a.Bp : a.Np;                              // UMMA N
    const uint32_t KBB = (uint32_t)Nact * 128u;                          // bytes per activation k-block ([N x 64] fp16)
    // (a projection CTA's 128-row UMMA over-reads 64 rows past each 64-row k-block: they land in the next k-block /
    //  the activation buffer, and the TMEM lanes they produce, 64-127, are never read)
    const uint32_t act0 = base + ((WBYTES + 1023u) & ~1023u);
    const uint32_t SB = is_gate ? (uint32_t)a.KBG * KBB : (uint32_t)KB_W * KBB;     // bytes per TMA copy
    const uint32_t act_bytes = is_gate ? GSTAGES * SB : 2u * SB;
    const uint32_t exch_bytes = 128u * (uint32_t)(a.Bp + 1) * 4u;
    const uint32_t act_region = act_bytes > exch_bytes ? act_bytes : exch_bytes;
    const uint32_t gbuf0 = (act0 + act_region + 1023u) & ~1023u;
    const uint32_t GB = (uint32_t)a.Bp * 16u * 4u * 4u;                  // 4 gates x [Bp x 16] fp32
    const uint32_t cbuf = gbuf0 + GB;
    const uint32_t CB = (uint32_t)a.B * UNITS * 4u;
    const uint32_t bar0 = (cbuf + CB + 15u) & ~15u;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };                // GSTAGES (gate) / 2 (projection)
    auto empty_bar = [&](int s) { return bar0 + 8u * (GSTAGES + s); };
    const uint32_t wfull_bar = bar0 + 8u * (2 * GSTAGES);
    const uint32_t tfull_bar = wfull_bar + 8u, tempty_bar = wfull_bar + 16u;
    const uint32_t gfull_bar = wfull_bar + 24u;                          // G buffer full
    const uint32_t gfree_bar = wfull_bar + 32u;                          // G buffer consumed by the epilogue
    const uint32_t tmem_slot = wfull_bar + 40u;
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    float* exch = reinterpret_cast<float*>(gptr(act0));
    float* cst = reinterpret_cast<float*>(gptr(cbuf));

    if (threadIdx.x == 0) {
        for (int s = 0; s < GSTAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(wfull_bar, 1); mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 1);
        mbar_init(gfull_bar, 1); mbar_init(gfree_bar, 1);
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
        const int copies_half = KB_W / a.KBG;                            // TMA copies per activation half
        const int copies = 2 * copies_half;                              // per step
        if (warp == 0) {
            if (lane == 0) {
                tma_prefetch_desc(&mapHhi); tma_prefetch_desc(&mapHlo); tma_prefetch_desc(&mapG);
                mbar_expect_tx(wfull_bar, WBYTES);                       // resident weights
                for (int kb = 0; kb < KB_W; kb++) tma_load_2d(base + kb * w_kb_bytes, &mapWg, wfull_bar, kb * 64, blockIdx.x * 128);
                mbar_expect_tx(gfull_bar, GB);                           // G for the first step
                for (int g4 = 0; g4 < 4; g4++) tma_load_2d(gbuf0 + g4 * (GB / 4), &mapG, gfull_bar, g4 * a.H + u0, a.g0);
            }
            __syncwarp();
            for (int t = 0; t < a.T; t++) {
                if (t > 0) {
                    if (lane == 0) {
                        // the epilogue has finished reading G(t-1): refill the buffer with G(t) (lands long before it is needed)
                        mbar_wait(gfree_bar, (t - 1) & 1);
                        mbar_expect_tx(gfull_bar, GB);
                        for (int g4 = 0; g4 < 4; g4++) tma_load_2d(gbuf0 + g4 * (GB / 4), &mapG, gfull_bar, g4 * a.H + u0, t * a.Bs + a.g0);
                        wait_counter(a.counters + 1, (unsigned)(t * a.RS * a.BSP));              // h(t-1) complete
                    }
                    __syncwarp();
                }
                if (lane == 0) TRACE(0);
                // lanes 0..GSTAGES-1 each feed one ring stage (copies j = lane, lane + GSTAGES, ...): the TMA issue cost is per thread
                if (lane < GSTAGES) {
                    for (int j = lane; j < copies; j += GSTAGES) {
                        const uint32_t it = (uint32_t)t * copies + j;    // host guarantees copies % GSTAGES == 0 or copies < GSTAGES
                        const uint32_t use = copies >= GSTAGES ? it / GSTAGES : (uint32_t)t;   // fills of this stage so far
                        const int s = j % GSTAGES;
                        const int half = j / copies_half, c = j % copies_half;
                        mbar_wait(empty_bar(s), (use & 1) ^ 1);
                        mbar_expect_tx(full_bar(s), SB);
                        tma_load_3d(act0 + s * SB, half ? &mapHlo : &mapHhi, full_bar(s), 0, 0, c * a.KBG);
                    }
                }
                __syncwarp();
                if (lane == 0) TRACE(1);
            }
        } else if (warp == 1) {
            if (lane == 0) {
                mbar_wait(wfull_bar, 0);
                for (int t = 0; t < a.T; t++) {
                    mbar_wait(tempty_bar, (t & 1) ^ 1);
                    tc_fence_after();
                    for (int j = 0; j < copies; j++) {
                        const uint32_t it = (uint32_t)t * copies + j;
                        const uint32_t use = copies >= GSTAGES ? it / GSTAGES : (uint32_t)t;
                        const int s = j % GSTAGES;
                        const int c = j % copies_half;
                        mbar_wait(full_bar(s), use & 1);
                        if (j == 0) TRACE(2);
                        tc_fence_after();
                        for (int kk = 0; kk < a.KBG; kk++) {
                            const uint64_t ad = make_sw128_desc(base + (c * a.KBG + kk) * w_kb_bytes), bd = make_sw128_desc(act0 + s * SB + kk * KBB);
#pragma unroll
                            for (int k = 0; k < 4; k++) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (j | kk | k) != 0);
                        }
                        umma_commit(empty_bar(s));
                    }
                    umma_commit(tfull_bar);
                    TRACE(3);
                }
            }
        } else {
            const int q = warp & 3, et = (warp - 2) * 32 + lane;          // et: 0..127
            // c state for this CTA: cst[b * 16 + u]
            for (int i = et; i < a.B * UNITS; i += 128) cst[i] = a.c_state[(size_t)(i / UNITS) * a.H + u0 + (i % UNITS)];
            // this thread's (session, unit) pairs: i = et + 128 k  ->  b = et/16 + 8k, u = et%16; nsteps never changes during the launch
            const int my_u = et % UNITS, my_b0 = et / UNITS;
            int my_n[8];
#pragma unroll
            for (int k8 = 0; k8 < 8; k8++) { const int b = my_b0 + 8 * k8; my_n[k8] = b < a.B ? a.nsteps[b] : 0; }
            const int ld = a.Bp + 1;
            for (int t = 0; t < a.T; t++) {
                mbar_wait(tfull_bar, t & 1);
                if (et == 0) TRACE(4);
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
                if (et == 0) { mbar_arrive(tempty_bar); TRACE(5); }       // accumulator drained: next step's MMAs may overwrite it
                mbar_wait(gfull_bar, t & 1);
                const float* gb = reinterpret_cast<const float*>(gptr(gbuf0));
#pragma unroll
                for (int k8 = 0; k8 < 8; k8++) {
                    const int b = my_b0 + 8 * k8;
                    if (a.t0 + t >= my_n[k8]) continue;                  // ragged: finished / padding sessions keep their state
                    const int i = b * UNITS + my_u;
                    float gate[4];
#pragma unroll
                    for (int g4 = 0; g4 < 4; g4++)
                        gate[g4] = (exch[(g4 * 16 + my_u) * ld + b] + exch[(64 + g4 * 16 + my_u) * ld + b]) * a.inv_scale_hh + gb[(g4 * a.Bp + b) * 16 + my_u];
                    const float cn = fast_sigmoid(gate[1]) * cst[i] + fast_sigmoid(gate[0]) * fast_tanh(gate[2]);
                    cst[i] = cn;
                    const float m = fast_sigmoid(gate[3]) * fast_tanh(cn);
                    __half mh, ml; split_f16(m, mh, ml);
                    a.m_hi[(size_t)b * a.H + u0 + my_u] = mh;
                    a.m_lo[(size_t)b * a.H + u0 + my_u] = ml;
                }
                if (et == 0) TRACE(6);
                asm volatile("fence.proxy.async;" ::: "memory");         // generic-proxy stores -> visible to the peers' TMA reads
                epi_bar_sync();                                          // CTA-scope ordering; the release below is cumulative over it
                if (et == 0) { red_release(a.counters + 0); mbar_arrive(gfree_bar); TRACE(7); }   // m(t) slice published; G buffer reusable
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
                tma_prefetch_desc(&mapMhi); tma_prefetch_desc(&mapMlo);
                mbar_expect_tx(wfull_bar, WBYTES);
                for (int kb = 0; kb < KB_W; kb++) tma_load_2d(base + kb * w_kb_bytes, &mapWp, wfull_bar, kb * 64, rs * 64);
            }
            __syncwarp();
            for (int t = 0; t < a.T; t++) {
                if (lane == 0) { wait_counter(a.counters + 0, (unsigned)((t + 1) * a.GC)); TRACE(0); }   // m(t) complete
                __syncwarp();
                // the previous step's MMAs are long finished (this CTA's epilogue of t-1 ran before the gate CTAs could produce m(t)),
                // so both halves can be overwritten immediately: one 3-D copy each, issued by two lanes in parallel
                if (lane < 2) {
                    mbar_expect_tx(full_bar(lane), SB);
                    tma_load_3d(act0 + lane * SB, lane ? &mapMlo : &mapMhi, full_bar(lane), 0, b0, 0);
                }
                __syncwarp();
                if (lane == 0) TRACE(1);
            }
        } else if (warp == 1) {
            if (lane == 0) {
                mbar_wait(wfull_bar, 0);
                for (int t = 0; t < a.T; t++) {
                    mbar_wait(tempty_bar, (t & 1) ^ 1);
                    tc_fence_after();
                    for (int half = 0; half < 2; half++) {
                        mbar_wait(full_bar(half), t & 1);
                        if (half == 0) TRACE(2);
                        tc_fence_after();
                        for (int kb = 0; kb < KB_W; kb++) {
                            const uint64_t ad = make_sw128_desc(base + kb * w_kb_bytes), bd = make_sw128_desc(act0 + half * SB + kb * KBB);
#pragma unroll
                            for (int k = 0; k < 4; k++) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (half | kb | k) != 0);
                        }
                    }
                    umma_commit(tfull_bar);
                    TRACE(3);
                }
            }
        } else {
            const int q = warp & 3, et = (warp - 2) * 32 + lane;
            const int ld = a.Np + 1;
            // this thread's outputs: i = et + 128 k -> row r = et % 32, session j = et / 32 + 4 k
            const int my_r = et % PROWS, my_j0 = et / PROWS;
            int my_n[4];
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) { const int b = b0 + my_j0 + 4 * k4; my_n[k4] = (my_j0 + 4 * k4 < a.Np && b < a.B) ? a.nsteps[b] : -1; }
            for (int t = 0; t < a.T; t++) {
                mbar_wait(tfull_bar, t & 1);
                if (et == 0) TRACE(4);
                tc_fence_after();
                if (q < 2) {                                             // lanes 0-31: W_hi rows, 32-63: W_lo rows; 64-127 unused
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
                    for (int j = 0; j < 32; j++) if (j < a.Np) exch[(q * 32 + lane) * ld + j] = v[j];
                }
                tc_fence_before();
                epi_bar_sync();
                if (et == 0) { mbar_arrive(tempty_bar); TRACE(5); }
                // 32 rows x Np sessions; consecutive threads -> consecutive rows (coalesced global stores)
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) {
                    if (my_n[k4] < 0) continue;                          // padding session
                    const int j = my_j0 + 4 * k4, b = b0 + j;
                    const float hv = (exch[my_r * ld + j] + exch[(32 + my_r) * ld + j]) * a.inv_scale_hr;
                    float* ho = a.Hout + ((size_t)t * a.Bs + b) * a.D + r0 + my_r;
                    if (a.t0 + t < my_n[k4]) {
                        *ho = hv;
                        a.h_state[(size_t)b * a.D + r0 + my_r] = hv;
                        __half hh, hl; split_f16(hv, hh, hl);
                        a.h_hi[(size_t)b * a.D + r0 + my_r] = hh;
                        a.h_lo[(size_t)b * a.D + r0 + my_r] = hl;
                    } else {
                        *ho = 0.f;
                    }
                }
                if (et == 0) TRACE(6);
                asm volatile("fence.proxy.async;" ::: "memory");
                epi_bar_sync();
                if (et == 0) { red_release(a.counters + 1); TRACE(7); }   // h(t) slice published
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// fp32 state [B, D] -> padded fp16 (hi, lo) exchange buffers [Bp, D] (rows >= B zeroed)
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
