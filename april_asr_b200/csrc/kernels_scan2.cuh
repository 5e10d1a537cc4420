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
// cell non-linearities: ex2/rcp forms. Measured on the full model: full-precision expf / tanhf here change neither the
// encoder-output error against the oracle (7.2e-5 abs either way) nor any token, and cost 0.3 us per layer-step.
__device__ __forceinline__ float cell_sigmoid(float x) { return fast_sigmoid(x); }
__device__ __forceinline__ float cell_tanh(float x) { return fast_tanh(x); }
using scan::gtimer;
using scan::ld_acquire;
using scan::red_release;
using scan::wait_counter;
using scan::epi_bar_sync;

constexpr int UNITS = 8;                  // hidden units per CTA -> 32 gate rows, 64 stacked (hi ; lo)
constexpr int RSTAGES = 4;                // activation ring stages
constexpr int MAXG = 4;
constexpr int EPI_WARPS = 8;
constexpr int MMAW = 4;                   // MMA-issuing warps = accumulators per group
// warps 0-3: one TMA ring stage each (a spinning lane must not share a warp with another producer), warp 4: weights + G loader,
// warps 5-8: MMA issuers, warps 9-16: epilogue (TMEM drain + cell update), warps 17-20: one releaser per session group (the
// store-acknowledge wait of the release is ~1.4 us: it must not hold the epilogue team, which serves the groups in turn).
// Why four MMA warps: measured on B200 (probe_mma.cu), one thread issues a tcgen05.mma every >= ~84 clk whatever M and N
// (N <= 128), so a group's H/16 = 64 K-slices cost 2.7 us from one thread. The limit is per issuing thread: four warps,
// each owning K-slice (k mod 4) of every k-block and its own TMEM accumulator, issue concurrently (~24-32 clk apart).
constexpr int THREADS = (RSTAGES + 1 + MMAW + EPI_WARPS + MAXG) * 32;
// Up to MAXG independent session groups are interleaved through the CTA (as in kernels_scan.cuh): the ring keeps
// streaming another group's m rows while one group is in its epilogue / handover. A group's step is a fixed chain
// (TMA landing, MMA tail, drain, cell, release, poll: ~7 us measured) plus its own ingest, so small groups (16 sessions)
// shorten the chain and more of them fill the ring. An optional thread-block
// cluster (CS CTAs) fetches each copy once and multicasts it; measured on B200 this does not help (the ring, not L2,
// bounds the ingest), so the default is CS = 1.

struct Scan2Args {
    int T, t0;             // steps in this launch, absolute index of the first one
    int NG;                // interleaved groups
    int Bg[MAXG];          // sessions per group
    int goff[MAXG];        // first session of the group relative to g0
    int Bp;                // padded group size (multiple of 16, NG * Bp <= 64)
    int Bs, g0;            // sessions in the whole pass = row stride of G / M buffers; first session of this launch
    int H;
    int GC;                // CTAs = H / 8
    int KBG;               // k-blocks per TMA copy
    int m64;               // 1: UMMA M = 64 (no over-read of 64 unused rows; accumulator rows sit in lanes 0-15 of each TMEM quarter)
    int dbg;               // option bits: 4 = single handover poller per CTA (scan3)
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void epi2_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
lstm_scan2_kernel(const __grid_constant__ CUtensorMap mapW,    // 2-D [GC*64, H] stacked W_eff slices, box {64, 64}
                  const __grid_constant__ CUtensorMap mapMx,   // 5-D {64, Bs, 2, H/64, T+1}, box {64, Bp, 2, KBG, 1}
                  const __grid_constant__ CUtensorMap mapG,    // 3-D fp32 view of [T*Bs, 4H]: {unit, gate, row}, box {8, 4, Bp}
                  Scan2Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NG = a.NG;
    constexpr int NACC = MMAW;

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
    const uint32_t nbuf = cbuf + (uint32_t)NG * CB;                       // [MAXG][64] int: live steps per session
    const uint32_t stg0 = nbuf + MAXG * 64 * 4;                           // [2][256] fp16 staging: 8 units of a session -> one 16-byte store
    const uint32_t bar0 = (stg0 + 2u * EPI_WARPS * 32u * 2u + 15u) & ~15u;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (RSTAGES + s); };
    const uint32_t wfull_bar = bar0 + 8u * (2 * RSTAGES);
    auto tfull_bar = [&](int g) { return wfull_bar + 8u + 8u * g; };
    auto tempty_bar = [&](int g) { return wfull_bar + 8u + 8u * (MAXG + g); };
    auto gfull_bar = [&](int g) { return wfull_bar + 8u + 8u * (2 * MAXG + g); };
    auto gfree_bar = [&](int g) { return wfull_bar + 8u + 8u * (3 * MAXG + g); };
    auto rel_bar = [&](int g) { return wfull_bar + 8u + 8u * (4 * MAXG + g); };
    const uint32_t tmem_slot = wfull_bar + 8u + 8u * (5 * MAXG);
    float* exch = reinterpret_cast<float*>(gptr(exch0));
    float* cst = reinterpret_cast<float*>(gptr(cbuf));

    const uint32_t crank = cluster_ctarank_(), CS = cluster_nctarank_();
    const uint16_t cmask = (uint16_t)((1u << CS) - 1u);
    if (threadIdx.x == 0) {
        for (int s = 0; s < RSTAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MMAW * CS); }
        mbar_init(wfull_bar, 1);
        for (int g = 0; g < MAXG; g++) { mbar_init(tfull_bar(g), MMAW); mbar_init(tempty_bar(g), 1); mbar_init(gfull_bar(g), 1); mbar_init(gfree_bar(g), 1); mbar_init(rel_bar(g), 1); }
        fence_barrier_init();
    }
    uint32_t tmem_cols = 32; while (tmem_cols < (uint32_t)(NG * NACC * Nact)) tmem_cols <<= 1;
    if (warp == RSTAGES + 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_();                                          // peers' barriers are initialised before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));
    const uint32_t idesc = make_idesc_f16(a.m64 ? 64 : 128, Nact);
    const int u0 = blockIdx.x * UNITS;
    const int copies = KB_W / a.KBG;                                      // TMA copies per (step, group)

    if (warp < RSTAGES) {
        // convergent warp; lane 0 polls the handover counter, one elected lane issues the copy
        const int stage = warp;
        if (lane == 0) tma_prefetch_desc(&mapMx);
        const long long total = (long long)a.T * NG * copies;
        int opened_t = -1, opened_g = -1;
        for (long long idx = stage; idx < total; idx += RSTAGES) {
            const int j = (int)(idx % copies);
            const int tg = (int)(idx / copies), g = tg % NG, t = tg / NG;
            if (t != opened_t || g != opened_g) {
                if (t > 0) {
                    if (lane == 0) wait_counter(a.counter + g * 64, (unsigned)t * (unsigned)a.GC);     // m(t-1) of the group complete everywhere
                    __syncwarp();
                    asm volatile("fence.proxy.async;" ::: "memory");     // whichever lane is elected below issues its TMA read after the acquire
                }
                opened_t = t; opened_g = g;
                if (j == 0 && lane == 0) TRACE2(0);
            }
            const uint32_t use = (uint32_t)(idx / RSTAGES);
            mbar_wait(empty_bar(stage), (use & 1) ^ 1);                  // stage free (in every CTA of the cluster)
            if (elect_one()) {
                mbar_expect_tx(full_bar(stage), SB);
                if (CS == 1) tma_load_5d(ring0 + stage * SB, &mapMx, full_bar(stage), 0, a.g0 + a.goff[g], 0, j * a.KBG, t);   // row t holds m(t-1)
                else if ((uint32_t)j % CS == crank) tma_load_5d_mc(ring0 + stage * SB, &mapMx, full_bar(stage), 0, a.g0 + a.goff[g], 0, j * a.KBG, t, cmask);
            }
            __syncwarp();
            if (j == copies - 1 && lane == 0) TRACE2(1);
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
                    tma_load_3d(gbuf0 + g * GB, &mapG, gfull_bar(g), u0, 0, t * a.Bs + a.g0 + a.goff[g]);          // lands as [session][gate][8 units]
                }
        }
    } else if (warp < RSTAGES + 1 + MMAW) {
        // MMA warp mw issues K-slice mw of every k-block into its own accumulator. The whole warp walks the loop
        // (convergent, so descriptors live in uniform registers); one elected lane issues the tcgen05 instructions.
        const int mw = warp - (RSTAGES + 1);
        mbar_wait(wfull_bar, 0);
        long long idx = 0;
        for (int t = 0; t < a.T; t++)
            for (int g = 0; g < NG; g++) {
                mbar_wait(tempty_bar(g), (t & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (g * NACC + mw) * Nact;
                for (int j = 0; j < copies; j++, idx++) {
                    const int s = (int)(idx % RSTAGES);
                    const uint32_t use = (uint32_t)(idx / RSTAGES);
                    mbar_wait(full_bar(s), use & 1);
                    if (j == 0 && mw == 0 && lane == 0) TRACE2(2);
                    tc_fence_after();
                    if (elect_one()) {
                        for (int kk = 0; kk < a.KBG; kk++) {
                            const uint64_t ad = make_sw128_desc(base + (j * a.KBG + kk) * w_kb_bytes), bd = make_sw128_desc(ring0 + s * SB + kk * KBB);
                            umma_f16(d_tmem, ad + 2 * mw, bd + 2 * mw, idesc, (j | kk) != 0);
                        }
                        if (CS == 1) umma_commit(empty_bar(s)); else umma_commit_mc(empty_bar(s), cmask);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(tfull_bar(g));
                __syncwarp();
                if (mw == 0 && lane == 0) TRACE2(3);
            }
    } else if (warp < RSTAGES + 1 + MMAW + EPI_WARPS) {
        const int ew = warp - (RSTAGES + 1 + MMAW), et = ew * 32 + lane;         // et: 0..255
        const int q = warp & 3, half = ew >> 2;                           // TMEM lane quarter of this warp; the two warps of a quarter split the columns
        const int cells = a.Bp * UNITS;                                   // (session, unit) pairs of a group: cell idx = b * 8 + u, thread et owns idx = et (+256)
        int* sn = reinterpret_cast<int*>(gptr(nbuf));
        __half* stg_h = reinterpret_cast<__half*>(gptr(stg0));
        __half* stg_l = stg_h + EPI_WARPS * 32;
        for (int g = 0; g < NG; g++)
            for (int idx = et; idx < cells; idx += EPI_WARPS * 32) {
                const int b = idx >> 3, u = idx & 7, sess = a.g0 + a.goff[g] + b;
                cst[g * cells + idx] = b < a.Bg[g] ? a.c_state[(size_t)sess * a.H + u0 + u] : 0.f;
                if (u == 0) sn[g * 64 + b] = b < a.Bg[g] ? a.nsteps[sess] : 0;
            }
        epi2_bar_sync();
        for (int t = 0; t < a.T; t++)
            for (int g = 0; g < NG; g++) {
                mbar_wait(tfull_bar(g), t & 1);
                if (et == 0) TRACE2(4);
                tc_fence_after();
                if (a.m64 || q < 2) {                                     // rows 0-31: W_hi, 32-63: W_lo; columns [0,Bp): m_hi, [Bp,2Bp): m_lo
                    // M = 128: row r in lane r (quarters 0, 1). M = 64: row r in lane 32 * (r / 16) + r % 16 (16 rows per quarter)
                    const uint32_t tb = tmem_base + g * NACC * Nact + ((uint32_t)(q * 32) << 16);
                    const int row = a.m64 ? q * 16 + lane : q * 32 + lane;
                    const bool valid = !a.m64 || lane < 16;
                    const int nc = a.Bp >> 1, c0 = half * nc;
                    for (int c = c0; c < c0 + nc; c += 8) {
                        uint32_t v[4][8], w2[4][8];
#pragma unroll
                        for (int ac = 0; ac < 4; ac++)
                            if (ac < NACC) { tmem_ld8_nowait(tb + ac * Nact + c, v[ac]); tmem_ld8_nowait(tb + ac * Nact + a.Bp + c, w2[ac]); }
                        tc_wait_ld();
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                float sum = __uint_as_float(v[0][j]) + __uint_as_float(w2[0][j]);
#pragma unroll
                                for (int ac = 1; ac < 4; ac++) if (ac < NACC) sum += __uint_as_float(v[ac][j]) + __uint_as_float(w2[ac][j]);
                                exch[row * ld + c + j] = sum;
                            }
                        }
                    }
                }
                tc_fence_before();
                epi2_bar_sync();
                if (et == 0) { mbar_arrive(tempty_bar(g)); TRACE2(5); }
                mbar_wait(gfull_bar(g), t & 1);
                const float* gb = reinterpret_cast<const float*>(gptr(gbuf0 + g * GB));
                for (int idx = et; idx < cells; idx += EPI_WARPS * 32) {  // warp-uniform trip count (cells is a multiple of 128)
                    const int b = idx >> 3, u = idx & 7;
                    const int n_live = sn[g * 64 + b];
                    const bool live = a.t0 + t < n_live;                  // ragged: finished / padding sessions keep their state
                    const int sess = a.g0 + a.goff[g] + b;
                    float gate[4];
#pragma unroll
                    for (int g4 = 0; g4 < 4; g4++)
                        gate[g4] = (exch[(g4 * 8 + u) * ld + b] + exch[(32 + g4 * 8 + u) * ld + b]) * a.inv_scale + gb[(b * 4 + g4) * UNITS + u];
                    const float cn = cell_sigmoid(gate[1]) * cst[g * cells + idx] + cell_sigmoid(gate[0]) * cell_tanh(gate[2]);
                    const float mval = cell_sigmoid(gate[3]) * cell_tanh(cn);
                    if (live) cst[g * cells + idx] = cn;
                    __half mh, ml;
                    split_f16(mval, mh, ml);
                    stg_h[et] = mh; stg_l[et] = ml;
                    if (live && ((a.t0 + t == n_live - 1) || (t == a.T - 1))) a.m_state[(size_t)sess * a.H + u0 + u] = mval;
                    __syncwarp();
                    if (live && u < 4) {                                  // lanes 8k..8k+7 hold one session: four of them store its 16-byte rows
                        const uint4 val = *reinterpret_cast<const uint4*>((u & 1) ? (stg_l + (et - u)) : (stg_h + (et - u)));
                        __half* dst = (u < 2) ? a.Mx + ((size_t)(t + 1) * 2 + u) * a.Bs * a.H + (size_t)sess * a.H
                                              : ((u == 2) ? a.Mall_hi : a.Mall_lo) + ((size_t)t * a.Bs + sess) * a.H;
                        *reinterpret_cast<uint4*>(dst + u0) = val;
                    }
                    __syncwarp();
                }
                epi2_bar_sync();                                          // every store of this (step, group) is issued; exch / G buffer reusable
                if (et == 0) { mbar_arrive(rel_bar(g)); mbar_arrive(gfree_bar(g)); TRACE2(6); }
            }
        for (int g = 0; g < NG; g++)
            for (int idx = et; idx < cells; idx += EPI_WARPS * 32) {
                const int b = idx >> 3, u = idx & 7;
                if (b < a.Bg[g]) a.c_state[(size_t)(a.g0 + a.goff[g] + b) * a.H + u0 + u] = cst[g * cells + idx];
            }
    } else {
        // releaser of group g: CTA-scope ordering came through the mbarrier; the gpu-scope release is cumulative over it
        const int g = warp - (RSTAGES + 1 + MMAW + EPI_WARPS);
        if (g < NG && lane == 0)
            for (int t = 0; t < a.T; t++) {
                mbar_wait(rel_bar(g), t & 1);
                asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy stores -> the peers' TMA (async proxy) reads
                red_release(a.counter + g * 64);        // one 256-byte stride per group: the groups' polls and reds do not share an L2 line
                TRACE2(7);
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
