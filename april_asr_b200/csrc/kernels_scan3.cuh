// Persistent tensor-core LSTM scan for SMALL batches (<= 64 sessions): the latency-optimised kernel of round 1, kept for the regime
// it was tuned for. kernels_scan4.cuh (sessions on the UMMA M dimension, up to 512 sessions per launch) wins as soon as a pass
// carries more than 64 sessions; below that the step is one dependent chain (handover -> TMA landing -> MMAs -> TMEM drain ->
// pair exchange -> cell -> release) and this kernel's two interleaved groups of 32 sessions, four MMA-issuing warps and gate rows
// prefetched by TMA make that chain 7.6 us per layer-step against 9.3-9.8 us there (measured on B200, 64 sessions x 30 s).
//
// Same recurrence (gates(t) = G[t] + W_eff m(t-1), W_eff = W_hh W_hr), different decomposition: two CTAs form a cluster, the pair
// owns 16 hidden units, CTA r of the pair multiplies the pair's 128 stacked rows ([W_hi ; W_lo] x 16 units x 4 gates, a full
// M = 128 UMMA) by K-half r of m(t-1) for up to 64 session columns ([m_hi ; m_lo] stacked along N). The two partial sums meet through
// distributed shared memory: a CTA keeps the rows of its own 8 units and stores the other 32 + 32 rows straight from registers into
// its peer's exchange buffer, then signals the peer's mbarrier (release.cluster). Every CTA publishes 8 units x sessions of m(t).
//
// Layout per CTA: W slice 128 rows x H/2 as H/128 k-blocks of 16 KB (SWIZZLE_128B), resident; ring of 4 stages;
// stacked row rho = hl * 64 + half * 32 + gate * 8 + u  (hl: W_hi / W_lo, half: which CTA of the pair owns unit u).
// TMEM lane = rho, so quarter q = rho / 32 = hl * 2 + half: quarters with half == my rank stay local.
#pragma once
#include "kernels_scan4.cuh"

namespace april {
namespace scan3 {

using namespace tc;
using scan4::gtimer;
using scan4::ld_acquire;
using scan4::red_release;
using scan4::wait_counter;
using scan4::cluster_ctarank_;
using scan4::cluster_sync_;
__device__ __forceinline__ float cell_sigmoid(float x) { return scan4::fast_sigmoid(x); }
__device__ __forceinline__ float cell_tanh(float x) { return scan4::fast_tanh(x); }

constexpr int UNITS = 8;                  // hidden units per CTA
constexpr int RSTAGES = 4;                // activation ring stages
constexpr int MAXG = 4;
constexpr int EPI_WARPS = 8;
constexpr int MMAW = 4;                   // MMA-issuing warps = accumulators per group (one thread issues a tcgen05.mma every >= ~84 clk)

struct Scan2Args {
    int T, t0;             // steps in this launch, absolute index of the first one
    int NG;                // interleaved groups
    int Bg[MAXG];          // sessions per group
    int goff[MAXG];        // first session of the group relative to g0
    int Bp;                // padded group size (multiple of 16, <= 32)
    int Bs, g0;            // sessions in the whole pass = row stride of G / M buffers; first session of this launch
    int H;
    int GC;                // CTAs = H / 8
    int KBG;               // k-blocks per TMA copy
    int dbg;               // option bits: 4 = single handover poller per CTA
    float inv_scale;
    const int* nsteps;     // [Bs] (not offset)
    float* c_state;        // [Bs, H] fp32 pass-layout state of this layer (read at start, written at end)
    float* m_state;        // [Bs, H] fp32: m of the last live step
    __half* Mx;            // [2 (hi | lo)][(T + 1) * Bs][H]: m rows, written once; row block 0 = initial m, block t + 1 = m(t). Exchange buffer of
                           // the scan AND A operand of the layer-output GEMM (the layout kernels_scan4.cuh uses)
    size_t plane;          // elements between the hi and the lo plane
    unsigned* counter;     // [MAXG * 64], zeroed by the caller
    unsigned long long* trace;
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ uint32_t mapa_(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_remote_v4(uint32_t raddr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory"); }
// bounded wait with cluster-scope acquire (the peer's st.shared::cluster stores must be visible afterwards)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t_start = clock64();
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if ((spin & 0xFFFF) == 0xFFFF && clock64() - t_start > 4000000000LL) { printf("libapril: pair mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x); __trap(); }
    }
}

// Epilogue teams of 8 warps: team k serves groups k, k + TEAMS, ... Measured on B200 (64 sessions): one team with two
// groups of 32 sessions = 7.7 us per layer-step; two teams with four groups of 16 = 8.8 us (the in-order ring and the
// pair skew eat what the second team gains), so one team it is. The code stays parametrised.
constexpr int TEAMS = 1;
constexpr int THREADS3 = (RSTAGES + 1 + MMAW + TEAMS * EPI_WARPS + MAXG) * 32;
__device__ __forceinline__ void team_bar_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "n"(EPI_WARPS * 32) : "memory"); }
#define TRACE3(slot_) do { if (a.trace && g == 0) a.trace[((size_t)blockIdx.x * a.T + t) * 8 + (slot_)] = gtimer(); } while (0)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS3, 1)
lstm_scan3_kernel(const __grid_constant__ CUtensorMap mapW,    // 2-D [GC*128, H/2] stacked W_eff slices, box {64, 128}
                  const __grid_constant__ CUtensorMap mapMx,   // 5-D {64, Bs, 2 (plane), H/64, T+1} over the two m planes, box {64, Bp, 2, KBG, 1}
                  const __grid_constant__ CUtensorMap mapG,    // 3-D fp32 view of [T*Bs, 4H]: {unit, gate, row}, box {8, 4, Bp}
                  Scan2Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NG = a.NG;
    constexpr int NACC = MMAW;
    const uint32_t rank = cluster_ctarank_(), peer = rank ^ 1u;

    const int KB_W = a.H / 128;                                           // k-blocks of my K half
    const uint32_t w_kb_bytes = 128u * 128u, WBYTES = (uint32_t)KB_W * w_kb_bytes;
    const int Nact = 2 * a.Bp;                                            // UMMA N: [m_hi rows ; m_lo rows]
    const uint32_t KBB = (uint32_t)Nact * 128u, SB = (uint32_t)a.KBG * KBB;
    const uint32_t ring0 = base + WBYTES;
    const int ld = a.Bp + 1;
    const int ldx = a.Bp + 4;                                             // 16-byte aligned rows for the blocks that cross the pair
    const uint32_t XB = 64u * (uint32_t)ldx * 4u;
    const uint32_t LB = (64u * (uint32_t)ld * 4u + 15u) & ~15u;
    const uint32_t exl0 = ring0 + RSTAGES * SB;                           // per team [2 hl][32 rows][ld]: my partial sums for my units
    const uint32_t exr0 = exl0 + TEAMS * LB;                              // per team [2 hl][32 rows][ldx], written by the peer: its partial sums for my units
    const uint32_t gbuf0 = (exr0 + TEAMS * XB + 127u) & ~127u;
    const uint32_t GB = (uint32_t)a.Bp * UNITS * 4u * 4u;                 // [Bp][4 gates][8] fp32, per group
    const uint32_t cbuf = gbuf0 + (uint32_t)NG * GB;
    const uint32_t CB = (uint32_t)a.Bp * UNITS * 4u;
    const uint32_t nbuf = cbuf + (uint32_t)NG * CB;
    const uint32_t stg0 = nbuf + MAXG * 64 * 4;
    const uint32_t bar0 = (stg0 + TEAMS * 2u * EPI_WARPS * 32u * 2u + 15u) & ~15u;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (RSTAGES + s); };
    const uint32_t wfull_bar = bar0 + 8u * (2 * RSTAGES);
    auto tfull_bar = [&](int g) { return wfull_bar + 8u + 8u * g; };
    auto tempty_bar = [&](int g) { return wfull_bar + 8u + 8u * (MAXG + g); };
    auto gfull_bar = [&](int g) { return wfull_bar + 8u + 8u * (2 * MAXG + g); };
    auto gfree_bar = [&](int g) { return wfull_bar + 8u + 8u * (3 * MAXG + g); };
    auto rel_bar = [&](int g) { return wfull_bar + 8u + 8u * (4 * MAXG + g); };
    auto prfull_bar = [&](int k) { return wfull_bar + 8u + 8u * (5 * MAXG + k); };           // the peer's partials for my units have landed in exr (team k)
    auto prfree_bar = [&](int k) { return wfull_bar + 8u + 8u * (5 * MAXG + TEAMS + k); };   // the peer has consumed what I stored into ITS exr
    auto opened_bar = [&](int g) { return wfull_bar + 8u + 8u * (5 * MAXG + 2 * TEAMS + g); };   // producer warp 0 saw the group's handover counter
    const uint32_t tmem_slot = wfull_bar + 8u + 8u * (6 * MAXG + 2 * TEAMS);
    float* cst = reinterpret_cast<float*>(gptr(cbuf));

    if (threadIdx.x == 0) {
        for (int s = 0; s < RSTAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MMAW); }
        mbar_init(wfull_bar, 1);
        for (int g = 0; g < MAXG; g++) { mbar_init(tfull_bar(g), MMAW); mbar_init(tempty_bar(g), 1); mbar_init(gfull_bar(g), 1); mbar_init(gfree_bar(g), 1); mbar_init(rel_bar(g), 1); }
        for (int k = 0; k < TEAMS; k++) { mbar_init(prfull_bar(k), 1); mbar_init(prfree_bar(k), 1); }
        for (int g = 0; g < MAXG; g++) mbar_init(opened_bar(g), 1);
        fence_barrier_init();
    }
    uint32_t tmem_cols = 32; while (tmem_cols < (uint32_t)(NG * NACC * Nact)) tmem_cols <<= 1;
    if (warp == RSTAGES + 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                                                      // the peer's barriers exist before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));
    const uint32_t idesc = make_idesc_f16(128, Nact);
    const int u0 = blockIdx.x * UNITS;                                    // my 8 units: pair * 16 + rank * 8
    const int copies = KB_W / a.KBG;                                      // TMA copies per (step, group)
    const int kb0 = (int)rank * KB_W;                                     // first k-block of my K half

    if (warp < RSTAGES) {
        const int stage = warp;
        const bool relay = (a.dbg & 4) && (copies % RSTAGES == 0);   // every producer warp then visits every (step, group): barrier phases stay in step
        if (lane == 0) tma_prefetch_desc(&mapMx);
        const long long total = (long long)a.T * NG * copies;
        int opened_t = -1, opened_g = -1;
        for (long long idx = stage; idx < total; idx += RSTAGES) {
            const int j = (int)(idx % copies);
            const int tg = (int)(idx / copies), g = tg % NG, t = tg / NG;
            if (t != opened_t || g != opened_g) {
                if (t > 0) {
                    if (relay && stage != 0) {
                        mbar_wait(opened_bar(g), (t - 1) & 1);           // one poller per CTA: warp 0 relays the handover
                    } else {
                        if (lane == 0) { wait_counter(a.counter + g * 64, (unsigned)t * (unsigned)a.GC); if (relay) mbar_arrive(opened_bar(g)); }   // m(t-1) of the group complete everywhere
                        __syncwarp();
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                opened_t = t; opened_g = g;
                if (j == 0 && lane == 0) TRACE3(0);
            }
            const uint32_t use = (uint32_t)(idx / RSTAGES);
            mbar_wait(empty_bar(stage), (use & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(full_bar(stage), SB);
                tma_load_5d(ring0 + stage * SB, &mapMx, full_bar(stage), 0, a.g0 + a.goff[g], 0, kb0 + j * a.KBG, t);   // row t holds m(t-1)
            }
            __syncwarp();
            if (j == copies - 1 && lane == 0) TRACE3(1);
        }
    } else if (warp == RSTAGES) {
        if (lane == 0) {
            tma_prefetch_desc(&mapG);
            mbar_expect_tx(wfull_bar, WBYTES);
            for (int kb = 0; kb < KB_W; kb++) tma_load_2d(base + kb * w_kb_bytes, &mapW, wfull_bar, kb * 64, blockIdx.x * 128);
            for (int t = 0; t < a.T; t++)
                for (int g = 0; g < NG; g++) {
                    if (t > 0) mbar_wait(gfree_bar(g), (t - 1) & 1);
                    mbar_expect_tx(gfull_bar(g), GB);
                    tma_load_3d(gbuf0 + g * GB, &mapG, gfull_bar(g), u0, 0, t * a.Bs + a.g0 + a.goff[g]);
                }
        }
    } else if (warp < RSTAGES + 1 + MMAW) {
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
                    if (j == 0 && mw == 0 && lane == 0) TRACE3(2);
                    tc_fence_after();
                    if (elect_one()) {
                        for (int kk = 0; kk < a.KBG; kk++) {
                            const uint64_t ad = make_sw128_desc(base + (j * a.KBG + kk) * w_kb_bytes), bd = make_sw128_desc(ring0 + s * SB + kk * KBB);
                            umma_f16(d_tmem, ad + 2 * mw, bd + 2 * mw, idesc, (j | kk) != 0);
                        }
                        umma_commit(empty_bar(s));
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(tfull_bar(g));
                __syncwarp();
                if (mw == 0 && lane == 0) TRACE3(3);
            }
    } else if (warp < RSTAGES + 1 + MMAW + TEAMS * EPI_WARPS) {
        const int team = (warp - (RSTAGES + 1 + MMAW)) / EPI_WARPS;
        const int ew = (warp - (RSTAGES + 1 + MMAW)) % EPI_WARPS, et = ew * 32 + lane;   // et: 0..255 within the team
        const int q = warp & 3, hc = ew >> 2;                             // TMEM lane quarter; the two warps of a quarter split the columns
        const int hl = q >> 1;
        const bool mine = (uint32_t)(q & 1) == rank;                      // this quarter's rows belong to my units
        const int cells = a.Bp * UNITS;
        int* sn = reinterpret_cast<int*>(gptr(nbuf));
        float* exl = reinterpret_cast<float*>(gptr(exl0 + team * LB));
        float* exr = reinterpret_cast<float*>(gptr(exr0 + team * XB));
        __half* stg_h = reinterpret_cast<__half*>(gptr(stg0)) + team * 2 * EPI_WARPS * 32;
        __half* stg_l = stg_h + EPI_WARPS * 32;
        const uint32_t exr_peer = mapa_(exr0 + team * XB, peer), prfull_peer = mapa_(prfull_bar(team), peer), prfree_peer = mapa_(prfree_bar(team), peer);
        for (int g = team; g < NG; g += TEAMS)
            for (int idx = et; idx < cells; idx += EPI_WARPS * 32) {
                const int b = idx >> 3, u = idx & 7, sess = a.g0 + a.goff[g] + b;
                cst[g * cells + idx] = b < a.Bg[g] ? a.c_state[(size_t)sess * a.H + u0 + u] : 0.f;
                if (u == 0) sn[g * 64 + b] = b < a.Bg[g] ? a.nsteps[sess] : 0;
            }
        team_bar_sync(team);
        uint32_t n = 0;                                                   // group-step counter of this team (the peer's same team counts alike)
        for (int t = 0; t < a.T; t++)
            for (int g = team; g < NG; g += TEAMS, n++) {
                mbar_wait(tfull_bar(g), t & 1);
                if (et == 0) TRACE3(4);
                tc_fence_after();
                if (!mine && n > 0) mbar_wait_cluster(prfree_bar(team), (n - 1) & 1);   // the peer is done with what I stored last time
                {
                    const uint32_t tb = tmem_base + g * NACC * Nact + ((uint32_t)(q * 32) << 16);
                    const int row = hl * 32 + lane;                       // row of the [2][32] exchange block
                    const int nc = a.Bp >> 1, c0 = hc * nc;
                    for (int c = c0; c < c0 + nc; c += 8) {
                        uint32_t v[NACC][8], w2[NACC][8];
#pragma unroll
                        for (int ac = 0; ac < NACC; ac++) { tmem_ld8_nowait(tb + ac * Nact + c, v[ac]); tmem_ld8_nowait(tb + ac * Nact + a.Bp + c, w2[ac]); }
                        tc_wait_ld();
                        float sum[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            sum[j] = __uint_as_float(v[0][j]) + __uint_as_float(w2[0][j]);
#pragma unroll
                            for (int ac = 1; ac < NACC; ac++) sum[j] += __uint_as_float(v[ac][j]) + __uint_as_float(w2[ac][j]);
                        }
                        if (mine) {
#pragma unroll
                            for (int j = 0; j < 8; j++) exl[row * ld + c + j] = sum[j];
                        } else {                                          // the peer's rows: 16-byte stores straight into its exchange block
                            const uint32_t ra = exr_peer + (uint32_t)(row * ldx + c) * 4u;
                            st_remote_v4(ra, make_float4(sum[0], sum[1], sum[2], sum[3]));
                            st_remote_v4(ra + 16u, make_float4(sum[4], sum[5], sum[6], sum[7]));
                        }
                    }
                }
                tc_fence_before();
                team_bar_sync(team);
                if (et == 0) { mbar_arrive(tempty_bar(g)); mbar_arrive_remote(prfull_peer); TRACE3(5); }
                mbar_wait(gfull_bar(g), t & 1);
                mbar_wait_cluster(prfull_bar(team), n & 1);                     // the peer's half of the K sum for my units
                const float* gb = reinterpret_cast<const float*>(gptr(gbuf0 + g * GB));
                for (int idx = et; idx < cells; idx += EPI_WARPS * 32) {  // warp-uniform trip count (cells is a multiple of 128)
                    const int b = idx >> 3, u = idx & 7;
                    const int n_live = sn[g * 64 + b];
                    const bool live = a.t0 + t < n_live;
                    const int sess = a.g0 + a.goff[g] + b;
                    float gate[4];
#pragma unroll
                    for (int g4 = 0; g4 < 4; g4++) {
                        const int r0 = (g4 * 8 + u) * ld + b, r1 = (32 + g4 * 8 + u) * ld + b;
                        const int x0 = (g4 * 8 + u) * ldx + b, x1 = (32 + g4 * 8 + u) * ldx + b;
                        gate[g4] = ((exl[r0] + exl[r1]) + (exr[x0] + exr[x1])) * a.inv_scale + gb[(b * 4 + g4) * UNITS + u];
                    }
                    const float cn = cell_sigmoid(gate[1]) * cst[g * cells + idx] + cell_sigmoid(gate[0]) * cell_tanh(gate[2]);
                    const float mval = cell_sigmoid(gate[3]) * cell_tanh(cn);
                    if (live) cst[g * cells + idx] = cn;
                    __half mh, ml;
                    split_f16(mval, mh, ml);
                    stg_h[et] = mh; stg_l[et] = ml;
                    if (live && ((a.t0 + t == n_live - 1) || (t == a.T - 1))) a.m_state[(size_t)sess * a.H + u0 + u] = mval;
                    __syncwarp();
                    if (live && u < 2) {                                  // lanes 8k..8k+7 hold one session: two of them store its 16-byte rows (hi plane, lo plane)
                        const uint4 val = *reinterpret_cast<const uint4*>(u ? (stg_l + (et - u)) : (stg_h + (et - u)));
                        __half* dst = a.Mx + (size_t)u * a.plane + ((size_t)(t + 1) * a.Bs + sess) * a.H;
                        *reinterpret_cast<uint4*>(dst + u0) = val;
                    }
                    __syncwarp();
                }
                team_bar_sync(team);                                      // stores issued; exchange blocks and G buffer reusable
                if (et == 0) { mbar_arrive(rel_bar(g)); mbar_arrive(gfree_bar(g)); mbar_arrive_remote(prfree_peer); TRACE3(6); }
            }
        for (int g = team; g < NG; g += TEAMS)
            for (int idx = et; idx < cells; idx += EPI_WARPS * 32) {
                const int b = idx >> 3, u = idx & 7;
                if (b < a.Bg[g]) a.c_state[(size_t)(a.g0 + a.goff[g] + b) * a.H + u0 + u] = cst[g * cells + idx];
            }
    } else {
        const int g = warp - (RSTAGES + 1 + MMAW + TEAMS * EPI_WARPS);
        if (g < NG && lane == 0)
            for (int t = 0; t < a.T; t++) {
                mbar_wait(rel_bar(g), t & 1);
                asm volatile("fence.proxy.async;" ::: "memory");
                red_release(a.counter + g * 64);        // one 256-byte stride per group: the groups' polls and reds do not share an L2 line
                TRACE3(7);
            }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                                                      // nobody exits while its peer may still store into it or signal its barriers
    if (warp == RSTAGES + 1) tmem_dealloc(tmem_base, tmem_cols);
}


// row 0 of the exchange buffer (initial m of the launch's sessions) from the fp32 state
// packs W_eff (scaled by `sc`) into this kernel's per-CTA layout: CTA c = pair * 2 + r holds 128 stacked rows
// rho = hl * 64 + half * 32 + gate * 8 + u of units pair * 16 + half * 8 + u, K half r:  dst[(c * 128 + rho) * (H / 2) + k]
__global__ void pack_weff3_kernel(const float* __restrict__ weff, __half* __restrict__ dst, int H, float sc) {
    const int KH = H / 2;
    const int c = blockIdx.x, rho = blockIdx.y;
    const int hl = rho >> 6, half = (rho >> 5) & 1, g4 = (rho >> 3) & 3, u = rho & 7, r = c & 1, pair = c >> 1;
    const float* src = weff + ((size_t)g4 * H + (size_t)pair * 16 + half * 8 + u) * H + (size_t)r * KH;
    __half* d = dst + ((size_t)c * 128 + rho) * KH;
    for (int k = threadIdx.x; k < KH; k += blockDim.x) {
        const float x = src[k] * sc;
        const __half hh = __float2half_rn(x);
        d[k] = hl ? __float2half_rn(x - __half2float(hh)) : hh;
    }
}

}  // namespace scan3
}  // namespace april
