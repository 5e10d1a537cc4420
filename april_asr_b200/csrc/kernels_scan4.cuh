// Persistent tensor-core LSTM scan, wide-batch formulation ("scan4").
//
// Recurrence (single exchange per step, as derived in DESIGN.md):  gates(t) = G[t] + W_eff m(t-1),  W_eff = W_hh W_hr  [4H, H],
// m (H per session) the only recurrent quantity; the layer output h = W_hr m is a time-batched GEMM afterwards.
// This replaces the three g_ort->Run-per-step encoder recurrences of april_session.c:131-148 for up to 512 sessions per launch.
//
// Decomposition (what changed against the 64-session kernels of round 1):
//  * SESSIONS sit on the UMMA M dimension: a group of up to 128 sessions is the A operand ([128 x K] tiles of m_hi / m_lo streamed
//    by TMA), the resident W_eff slice is the B operand (N = W rows). TMEM lane = session, so one epilogue thread owns one session
//    and finds all four gates of its units in its own lane: no shared-memory transpose, the cell update runs from registers.
//  * three products per k-slice into ONE accumulator: m_hi W_hi + m_hi W_lo + m_lo W_hi (no lo*lo quadrant, no quadrant sums).
//  * K is split over a cluster of CS CTAs (CS = 2 or 4). The cluster owns 8*CS hidden units = 32*CS W rows (N = 64 or 128);
//    CTA r multiplies them by K-slice r of m, so a CTA ingests H/CS columns of m per step. Partial sums cross through distributed
//    shared memory straight from the TMEM drain registers (st.shared::cluster, 16 B per lane, conflict-free layout); every CTA
//    then finishes the cell update of its own 8 units x 128 sessions and publishes its slice of m(t).
//  * up to MAXG = 4 groups (512 sessions) are interleaved through one launch: while a group waits for its grid-wide handover the
//    tensor pipe works on the others. G is read by the epilogue threads directly (one 16-byte load per gate and thread), c lives
//    in registers for the whole launch, m(t) rows are written once ([(T+1) * Bs, H] hi and lo planes) and are both the
//    exchange buffer and the A operand of the W_hr layer-output GEMM.
//
// Shared memory per CTA (H = 1024): W slice hi+lo 128 KB (resident) + ring of 16 KB activation tiles + exchange blocks.
#pragma once
#include "kernels_tc.cuh"

namespace april {
namespace scan4 {

using namespace tc;

constexpr int UNITS = 8;          // hidden units whose cell update a CTA owns
constexpr int GROUP = 128;        // sessions per group = UMMA M
constexpr int MAXG = 4;           // interleaved groups per launch (512 sessions)
constexpr int EPI_WARPS = 8;      // warps per epilogue team: two per TMEM lane quarter, each takes 4 of the CTA's 8 units
constexpr int TEAMS = 2;          // epilogue teams: team k serves groups k, k + 2 (measured on B200 with one team: the TMEM drain,
                                  // pair exchange and cell update of an item keep a team busy ~5.5 us, longer than the item's MMAs)
constexpr uint32_t TILE_BYTES = GROUP * 128u;   // [128 sessions x 64 k] fp16, 128B-swizzled

struct Args {
    int T, t0;             // steps in this launch, absolute index of the first one
    int NG;                // groups
    int Bg[MAXG];          // sessions per group (<= 128)
    int goff[MAXG];        // first session of the group relative to g0
    int Bp;                // rows per TMA box: max Bg rounded up to 8
    int Bs, g0;            // sessions in the whole pass (row stride of G / m buffers); first session of this launch
    int H, GC;             // hidden size; CTAs = H / 8
    float inv_scale;
    const int* nsteps;     // [Bs]
    const float* G;        // [T * Bs, 4H] fp32 input gates (bias included)
    float* c_state;        // [Bs, H] fp32 (read at start, written at end)
    float* m_state;        // [Bs, H] fp32: m of the last live step
    __half* Mx_hi;         // [(T + 1) * Bs, H]: row block 0 = initial m, block t + 1 = m(t)
    __half* Mx_lo;
    unsigned* counter;     // [MAXG * 64], zeroed by the caller
    unsigned long long* trace;
};

// ---- helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
    const long long t_start = clock64();
    unsigned spin = 0;
    while (ld_acquire(p) < target) {
        if ((++spin & 0xFFF) == 0 && clock64() - t_start > 4000000000LL) { printf("libapril: scan handover timeout (block %d)\n", (int)blockIdx.x); __trap(); }
    }
}
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }
__device__ __forceinline__ uint32_t cluster_ctarank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void st_remote_v4(uint32_t raddr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory"); }
// bounded wait with cluster-scope acquire (the peers' st.shared::cluster stores must be visible afterwards)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t_start = clock64();
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if ((spin & 0xFFFF) == 0xFFFF && clock64() - t_start > 4000000000LL) { printf("libapril: cluster mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}

template <int CS, int MMAW> struct Cfg {
    static constexpr int NC = 32 * CS;                          // W rows of the cluster = UMMA N: [owner rank p][unit half hc][gate][4 units]
    static constexpr int RST = CS == 2 ? 4 : 3;                 // ring stages (16 KB tiles)
    static constexpr int XBUF = CS == 2 ? 2 : 1;                // exchange block buffers
    static constexpr int NSLOT = (512 / (MMAW * NC)) > 4 ? 4 : (512 / (MMAW * NC));   // accumulator sets in TMEM
    static constexpr int WARPS = 1 + MMAW + TEAMS * EPI_WARPS + 1;   // producer, MMA issuers, epilogue teams, releaser (20 warps with MMAW = 2: five per
                                                                     // scheduler, i.e. 96 registers per thread)
    static constexpr int THREADS = WARPS * 32;
    static constexpr uint32_t XSRC_BYTES = 2u * 4u * GROUP * 16u;   // one source rank's block: [hc][16-byte chunk][session] = 16 KB
    static constexpr uint32_t XB_BYTES = (CS - 1) * XSRC_BYTES;
    static size_t smem_bytes(int H) { return 1024 + (size_t)H * 128 + (size_t)RST * TILE_BYTES + (size_t)XBUF * XB_BYTES + 512; }
};

#define TRACE4(slot_) do { if (a.trace && g == 0) a.trace[((size_t)blockIdx.x * a.T + t) * 8 + (slot_)] = gtimer(); } while (0)

// grid = H / 8 CTAs in clusters of CS (launch attribute), cooperative (all CTAs co-resident: they spin on each other's counters)
template <int CS, int MMAW>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(Cfg<CS, MMAW>::THREADS, 1)
lstm_scan4_kernel(const __grid_constant__ CUtensorMap mapW,     // 2-D [GC * 2 * NC, H / CS] fp16: per CTA [hi rows ; lo rows], box {64, NC}
                  const __grid_constant__ CUtensorMap mapMhi,   // 2-D [(T + 1) * Bs, H] fp16, box {64, Bp}
                  const __grid_constant__ CUtensorMap mapMlo,
                  Args a) {
    using C = Cfg<CS, MMAW>;
    constexpr int NC = C::NC, RST = C::RST, XBUF = C::XBUF, NSLOT = C::NSLOT;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NG = a.NG;
    const uint32_t rank = cluster_ctarank_();

    const int Kc = a.H / CS, KB = Kc / 64;                                // my K slice, in 64-wide k-blocks
    const uint32_t WT = (uint32_t)NC * 128u;                              // one W k-block tile
    const uint32_t whi0 = base, wlo0 = base + (uint32_t)KB * WT, WBYTES = 2u * (uint32_t)KB * WT;
    const uint32_t ring0 = base + WBYTES;
    const uint32_t exr0 = ring0 + RST * TILE_BYTES;
    const uint32_t bar0 = exr0 + XBUF * C::XB_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (RST + s); };
    const uint32_t wfull_bar = bar0 + 8u * (2 * RST);
    auto tfull_bar = [&](int s) { return wfull_bar + 8u + 8u * s; };
    auto tempty_bar = [&](int s) { return wfull_bar + 8u + 8u * (4 + s); };
    auto rel_bar = [&](int g) { return wfull_bar + 8u + 8u * (8 + g); };
    auto prfull_bar = [&](int x) { return wfull_bar + 8u + 8u * (12 + x); };   // every peer's partial sums for my units have landed in exr[x]
    auto prfree_bar = [&](int x) { return wfull_bar + 8u + 8u * (14 + x); };   // every peer has consumed what was stored into ITS exr[x]
    const uint32_t tmem_slot = wfull_bar + 8u + 8u * 16;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RST; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MMAW); }
        mbar_init(wfull_bar, 1);
        for (int s = 0; s < 4; s++) { mbar_init(tfull_bar(s), MMAW); mbar_init(tempty_bar(s), EPI_WARPS); }
        for (int g = 0; g < MAXG; g++) mbar_init(rel_bar(g), EPI_WARPS);
        for (int x = 0; x < 2; x++) { mbar_init(prfull_bar(x), EPI_WARPS * (CS - 1)); mbar_init(prfree_bar(x), EPI_WARPS * (CS - 1)); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                                                      // the peers' barriers exist before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));
    const int u0 = blockIdx.x * UNITS;                                    // my 8 units: cluster * 8 * CS + rank * 8

    if (warp == 0) {
        // ===================== TMA producer: W slice once, then the m tiles of every (step, group) in order =====================
        if (lane == 0) { tma_prefetch_desc(&mapW); tma_prefetch_desc(&mapMhi); tma_prefetch_desc(&mapMlo); }
        if (elect_one()) {
            mbar_expect_tx(wfull_bar, WBYTES);
            for (int kb = 0; kb < KB; kb++) {
                tma_load_2d(whi0 + kb * WT, &mapW, wfull_bar, kb * 64, ((int)blockIdx.x * 2 + 0) * NC);
                tma_load_2d(wlo0 + kb * WT, &mapW, wfull_bar, kb * 64, ((int)blockIdx.x * 2 + 1) * NC);
            }
        }
        __syncwarp();
        const uint32_t tile_tx = (uint32_t)a.Bp * 128u;
        uint32_t n = 0;
        for (int t = 0; t < a.T; t++)
            for (int g = 0; g < NG; g++) {
                if (t > 0) {
                    if (lane == 0) wait_counter(a.counter + g * 64, (unsigned)t * (unsigned)a.GC);     // m(t-1) of the group complete everywhere
                    __syncwarp();
                    asm volatile("fence.proxy.async;" ::: "memory");     // whichever lane is elected below issues its TMA reads after the acquire
                }
                if (lane == 0) TRACE4(0);
                const int row = t * a.Bs + a.g0 + a.goff[g];             // row block t holds m(t-1)
                for (int kb = 0; kb < KB; kb++)
                    for (int hl = 0; hl < 2; hl++, n++) {
                        const int s = (int)(n % RST);
                        mbar_wait(empty_bar(s), ((n / RST) & 1) ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(full_bar(s), tile_tx);
                            tma_load_2d(ring0 + s * TILE_BYTES, hl ? &mapMlo : &mapMhi, full_bar(s), (int)rank * Kc + kb * 64, row);
                        }
                        __syncwarp();
                    }
                if (lane == 0) TRACE4(1);
            }
    } else if (warp <= MMAW) {
        // ===================== MMA issuers: warp mw owns K-slices (ks % MMAW == mw) of every tile and its own accumulator =====================
        const int mw = warp - 1;
        constexpr uint32_t idesc = make_idesc_f16(GROUP, NC);
        mbar_wait(wfull_bar, 0);
        uint32_t n = 0, item = 0;
        for (int t = 0; t < a.T; t++)
            for (int g = 0; g < NG; g++, item++) {
                const int slot = (int)(item % NSLOT);
                mbar_wait(tempty_bar(slot), ((item / NSLOT) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)((slot * MMAW + mw) * NC);
                for (int kb = 0; kb < KB; kb++)
                    for (int hl = 0; hl < 2; hl++, n++) {
                        const int s = (int)(n % RST);
                        mbar_wait(full_bar(s), (n / RST) & 1);
                        if (kb == 0 && hl == 0 && mw == 0 && lane == 0) TRACE4(2);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t ad = make_sw128_desc(ring0 + s * TILE_BYTES);
                            const uint64_t bh = make_sw128_desc(whi0 + kb * WT), bl = make_sw128_desc(wlo0 + kb * WT);
#pragma unroll
                            for (int ks = mw; ks < 4; ks += MMAW) {
                                if (hl == 0) {
                                    umma_f16(d_tmem, ad + 2 * ks, bh + 2 * ks, idesc, (kb != 0 || ks != mw) ? 1u : 0u);   // m_hi W_hi
                                    umma_f16(d_tmem, ad + 2 * ks, bl + 2 * ks, idesc, 1u);                               // m_hi W_lo
                                } else {
                                    umma_f16(d_tmem, ad + 2 * ks, bh + 2 * ks, idesc, 1u);                               // m_lo W_hi
                                }
                            }
                            umma_commit(empty_bar(s));
                        }
                        __syncwarp();
                    }
                if (elect_one()) umma_commit(tfull_bar(slot));
                __syncwarp();
                if (mw == 0 && lane == 0) TRACE4(3);
            }
    } else if (warp < 1 + MMAW + TEAMS * EPI_WARPS) {
        // ===================== epilogue teams: one thread = one session x 4 of my units (and the same 4-unit column block of every peer) =====
        const int team = (warp - (1 + MMAW)) / EPI_WARPS;
        const int ew = (warp - (1 + MMAW)) % EPI_WARPS;
        const int q = warp & 3, hc = ew >> 2;                             // TMEM lane quarter this warp may read; which half of a CTA's units
        const int sl = q * 32 + lane;                                     // session lane within the group
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int ucol = u0 + hc * 4;                                     // first of my 4 units
        constexpr int GPT = MAXG / TEAMS;                                 // groups per team: slot gi serves group gi * TEAMS + team
        float cst[GPT][4];
        int nlive[GPT];
#pragma unroll
        for (int gi = 0; gi < GPT; gi++) {
            const int g = gi * TEAMS + team;
            const bool valid = g < NG && sl < a.Bg[g];
            const int sess = a.g0 + a.goff[g] + sl;
            nlive[gi] = valid ? a.nsteps[sess] : 0;
            float4 c4 = valid ? *reinterpret_cast<const float4*>(a.c_state + (size_t)sess * a.H + ucol) : make_float4(0.f, 0.f, 0.f, 0.f);
            cst[gi][0] = c4.x; cst[gi][1] = c4.y; cst[gi][2] = c4.z; cst[gi][3] = c4.w;
        }
        for (int t = 0; t < a.T; t++) {
#pragma unroll
            for (int gi = 0; gi < GPT; gi++) {
                const int g = gi * TEAMS + team;
                if (g >= NG) break;
                const uint32_t item = (uint32_t)(t * NG + g);
                const int slot = (int)(item % NSLOT);
                const int xb = (int)(item % XBUF);
                const uint32_t xuse = item / XBUF;
                const bool valid = sl < a.Bg[g];
                const bool warp_valid = q * 32 < a.Bg[g];
                const int sess = a.g0 + a.goff[g] + sl;
                mbar_wait(tfull_bar(slot), (item / NSLOT) & 1);
                if (ew == 0 && lane == 0) TRACE4(4);
                tc_fence_after();
                if (CS > 1 && xuse > 0) mbar_wait_cluster(prfree_bar(xb), (xuse - 1) & 1);   // the peers are done with what they received two items ago
                float mine[16];
#pragma unroll
                for (int j = 0; j < 16; j++) mine[j] = 0.f;
                if (warp_valid) {
#pragma unroll
                    for (int pp = 0; pp < CS; pp++) {
                        const uint32_t p = (rank + pp) % CS;              // pp == 0: my own columns
                        float v[16];
                        const uint32_t tcol = tmem_base + lane_off + (uint32_t)(slot * MMAW * NC) + p * 32u + (uint32_t)hc * 16u;
                        {
                            uint32_t r0[16];
                            tmem_ld16_nowait(tcol, r0);
                            tc_wait_ld();
#pragma unroll
                            for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r0[j]);
                        }
#pragma unroll
                        for (int ac = 1; ac < MMAW; ac++) {
                            uint32_t r1[16];
                            tmem_ld16_nowait(tcol + (uint32_t)(ac * NC), r1);
                            tc_wait_ld();
#pragma unroll
                            for (int j = 0; j < 16; j++) v[j] += __uint_as_float(r1[j]);
                        }
                        if (pp == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j++) mine[j] = v[j];
                        } else if (valid) {
                            const uint32_t src_slot = rank < p ? rank : rank - 1u;   // my block among the peer's CS - 1 sources
                            const uint32_t dst = mapa_(exr0 + (uint32_t)xb * C::XB_BYTES + src_slot * C::XSRC_BYTES + (uint32_t)hc * (4u * GROUP * 16u) + (uint32_t)sl * 16u, p);
#pragma unroll
                            for (int c = 0; c < 4; c++) st_remote_v4(dst + (uint32_t)c * (GROUP * 16u), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(tempty_bar(slot));                        // accumulators drained (this warp's part)
#pragma unroll
                    for (int pp = 1; pp < CS; pp++) mbar_arrive_remote(mapa_(prfull_bar(xb), (rank + pp) % CS));
                }
                if (ew == 0 && lane == 0) TRACE4(5);
                // input gates of my 4 units: in flight while the peers' partial sums arrive
                float4 gin[4];
                if (valid) {
                    const float* gp = a.G + ((size_t)t * a.Bs + sess) * 4 * a.H + ucol;
#pragma unroll
                    for (int g4 = 0; g4 < 4; g4++) gin[g4] = __ldcs(reinterpret_cast<const float4*>(gp + (size_t)g4 * a.H));
                }
                if (CS > 1) {
                    mbar_wait_cluster(prfull_bar(xb), xuse & 1);          // the other K-slices' sums for my units
                    if (valid) {
                        const float4* xr = reinterpret_cast<const float4*>(gptr(exr0 + (uint32_t)xb * C::XB_BYTES + (uint32_t)hc * (4u * GROUP * 16u) + (uint32_t)sl * 16u));
#pragma unroll
                        for (int sp = 0; sp < CS - 1; sp++)
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const float4 x = xr[(size_t)sp * (C::XSRC_BYTES / 16) + (size_t)c * GROUP];
                                mine[4 * c] += x.x; mine[4 * c + 1] += x.y; mine[4 * c + 2] += x.z; mine[4 * c + 3] += x.w;
                            }
                    }
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int pp = 1; pp < CS; pp++) mbar_arrive_remote(mapa_(prfree_bar(xb), (rank + pp) % CS));
                    }
                }
                if (ew == 0 && lane == 0) TRACE4(6);
                if (valid) {
                    const bool live = a.t0 + t < nlive[gi];                // ragged: finished sessions keep their state
                    const float xi[4] = { gin[0].x, gin[0].y, gin[0].z, gin[0].w }, xf[4] = { gin[1].x, gin[1].y, gin[1].z, gin[1].w },
                                xg[4] = { gin[2].x, gin[2].y, gin[2].z, gin[2].w }, xo[4] = { gin[3].x, gin[3].y, gin[3].z, gin[3].w };
                    float mval[4];
                    __align__(8) __half mh[4], ml[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const float pi = mine[u] * a.inv_scale + xi[u], pf = mine[4 + u] * a.inv_scale + xf[u];
                        const float pg = mine[8 + u] * a.inv_scale + xg[u], po = mine[12 + u] * a.inv_scale + xo[u];
                        const float cn = fast_sigmoid(pf) * cst[gi][u] + fast_sigmoid(pi) * fast_tanh(pg);
                        mval[u] = fast_sigmoid(po) * fast_tanh(cn);
                        if (live) cst[gi][u] = cn;
                        split_f16(mval[u], mh[u], ml[u]);
                    }
                    const size_t orow = ((size_t)(t + 1) * a.Bs + sess) * a.H + ucol;
                    *reinterpret_cast<uint2*>(a.Mx_hi + orow) = *reinterpret_cast<const uint2*>(mh);
                    *reinterpret_cast<uint2*>(a.Mx_lo + orow) = *reinterpret_cast<const uint2*>(ml);
                    if (live && ((a.t0 + t == nlive[gi] - 1) || (t == a.T - 1)))
                        *reinterpret_cast<float4*>(a.m_state + (size_t)sess * a.H + ucol) = make_float4(mval[0], mval[1], mval[2], mval[3]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(rel_bar(g));                   // this warp's stores of (t, g) are issued
            }
        }
#pragma unroll
        for (int gi = 0; gi < GPT; gi++) {
            const int g = gi * TEAMS + team;
            if (g < NG && sl < a.Bg[g])
                *reinterpret_cast<float4*>(a.c_state + (size_t)(a.g0 + a.goff[g] + sl) * a.H + ucol) = make_float4(cst[gi][0], cst[gi][1], cst[gi][2], cst[gi][3]);
        }
    } else {
        // ===================== releaser: the gpu-scope release (store acknowledgement ~1 us) must not hold an epilogue team ====
        if (lane == 0)
            for (int t = 0; t < a.T; t++)
                for (int g = 0; g < NG; g++) {
                    mbar_wait(rel_bar(g), t & 1);
                    asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy stores -> the other CTAs' TMA (async proxy) reads
                    red_release(a.counter + g * 64);                      // one 256-byte stride per group
                    TRACE4(7);
                }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                                                      // nobody exits while a peer may still store into it or signal its barriers
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// row block 0 of the m planes (initial m of every session of the pass) from the fp32 state
__global__ void scan4_prepare_kernel(const float* __restrict__ m_state, __half* __restrict__ Mx_hi, __half* __restrict__ Mx_lo, size_t n) {
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (size_t)gridDim.x * blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(m_state + i);
        __align__(8) __half h[4], l[4];
        split_f16(v.x, h[0], l[0]); split_f16(v.y, h[1], l[1]); split_f16(v.z, h[2], l[2]); split_f16(v.w, h[3], l[3]);
        *reinterpret_cast<uint2*>(Mx_hi + i) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(Mx_lo + i) = *reinterpret_cast<const uint2*>(l);
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

// packs W_eff (scaled by `sc`, a power of two) into the per-CTA [hi rows ; lo rows] x K-slice layout the kernel's mapW reads:
// dst[((c * 2 + hl) * NC + n) * Kc + k],  n = p * 32 + hc * 16 + gate * 4 + u4  ->  unit = (c / CS) * 8 * CS + p * 8 + hc * 4 + u4
__global__ void pack_weff_kernel(const float* __restrict__ weff, __half* __restrict__ dst, int H, int CS, float sc) {
    const int NC = 32 * CS, Kc = H / CS;
    const int c = blockIdx.x, n = blockIdx.y;
    const int p = n >> 5, hc = (n >> 4) & 1, g4 = (n >> 2) & 3, u4 = n & 3;
    const int unit = (c / CS) * 8 * CS + p * 8 + hc * 4 + u4, r = c % CS;
    const float* src = weff + ((size_t)g4 * H + unit) * H + (size_t)r * Kc;
    __half* dh = dst + (((size_t)c * 2 + 0) * NC + n) * Kc;
    __half* dl = dst + (((size_t)c * 2 + 1) * NC + n) * Kc;
    for (int k = threadIdx.x; k < Kc; k += blockDim.x) {
        const float x = src[k] * sc;
        const __half h = __float2half_rn(x);
        dh[k] = h; dl[k] = __float2half_rn(x - __half2float(h));
    }
}

}  // namespace scan4
}  // namespace april
