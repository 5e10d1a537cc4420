// Persistent tensor-core LSTM scan for LARGE batches: kernels_scan4.cuh's formulation (sessions on the UMMA M dimension, three products
// per k-slice into one accumulator, K split across CTAs with a DSMEM partial-sum exchange) on CTA PAIRS (tcgen05 cta_group::2).
//
// Why: measured on B200 (profiles/r02*), scan4 at 512 sessions is paced by what a CTA ingests — every CTA streams m(t-1) of ALL sessions
// for its K half (1 MB per layer-step and CTA at ~28 B/clk), the tensor pipe idles. A pair MMA reads its B operand from BOTH CTAs' shared
// memory, so with the same 128 KB of resident W_eff per CTA a pair covers twice the W rows (N = 128: 32 hidden units) and each CTA of the
// pair streams only ITS half of the sessions (M = 256 = 2 x 128): per layer-step a CTA ingests half the bytes.
//
// Cluster of 4 = 2 pairs: rank = 2 p + q; pair p multiplies K half p of m(t-1); CTA q of a pair holds W rows [64 q, 64 q + 64) of the
// cluster's 128 (as hi and lo tiles) and the sessions [128 q, 128 q + 128) of a 256-session item as its A rows. The leader (q = 0) issues
// M = 256, N = 128 MMAs; every CTA's TMEM receives [its 128 sessions] x [128 columns]. The K halves meet as in scan4: CTA (p, q) finishes
// the units [16 p, 16 p + 16) of the cluster's 32 for its sessions, and sends the other 64 columns to CTA (1 - p, q) over DSMEM.
// Everything else (time-indexed m planes written once, per-item handover counters, gate rows read by the epilogue threads, c in registers,
// releaser warp) is scan4's. Used for passes of more than 128 sessions; items are 256 sessions, up to 2 per launch.
#pragma once
#include "kernels_scan4.cuh"

namespace april {
namespace scan5 {

using namespace tc;
using scan4::Args;              // same launch arguments: NG = items (<= 2), Bg = sessions per item (<= 256), goff = first session of the item
using scan4::cluster_ctarank_;
using scan4::cluster_sync_;
using scan4::fast_sigmoid;
using scan4::fast_tanh;
using scan4::gtimer;
using scan4::mapa_;
using scan4::mbar_arrive_remote;
using scan4::mbar_wait_cluster;
using scan4::red_release;
using scan4::st_remote_v4;
using scan4::tmem_ld16_nowait;
using scan4::wait_counter;

constexpr int CL = 4;               // cluster size
constexpr int NCT = 128;            // W rows of the cluster = UMMA N: [owner 0..3][unit half hc][gate][4 units]; owner o = units [8 o, 8 o + 8)
constexpr int NB = 64;              // ... of which one CTA holds 64 (owners 2 q, 2 q + 1)
constexpr int HALF = 128;           // sessions per CTA of a pair = its half of M = 256
constexpr int ITEM = 256;           // sessions per item
constexpr int MAXI = 2;             // items per launch (512 sessions)
constexpr int EPI_WARPS = 8;        // two per TMEM lane quarter: each thread finishes 8 units of its session (measured: 16 warps x 4 units are
                                    // slower at every batch size, 20.4 vs 18.2 us per layer-step at 512 sessions: profiles/r02g_pair_epilogue_sweep.log)
constexpr int MMAW = 2;             // issuing warps (leader only) = accumulators per item
constexpr int RST = 4;              // ring stages of one 16 KB tile
constexpr int NSLOT = 512 / (MMAW * NCT);   // 2 accumulator sets
constexpr uint32_t TILE = HALF * 128u;      // [128 sessions x 64 k] fp16
constexpr uint32_t WT = NB * 128u;          // one W k-block tile: 8 KB
constexpr uint32_t XB_BYTES = 2u * 8u * HALF * 16u;   // exchange block: [which of my two owners][16-byte chunk of its 32 columns][session] = 32 KB
constexpr int WARPS = 1 + MMAW + EPI_WARPS + 1;
constexpr int THREADS = WARPS * 32;
static inline size_t smem_bytes(int H) { return 1024 + (size_t)H * 128 + (size_t)RST * TILE + XB_BYTES + 512; }

#define TRACE5(slot_) do { if (a.trace && g == 0) a.trace[((size_t)blockIdx.x * a.T + t) * 8 + (slot_)] = gtimer(); } while (0)

// grid = H / 8 CTAs in clusters of 4 (compile-time cluster dims: ncu cannot replay a launch that carries both the cluster and the cooperative attribute)
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
lstm_scan5_kernel(const __grid_constant__ CUtensorMap mapW,     // 2-D [GC * 2 * 64, H / 2] fp16: per CTA [hi rows ; lo rows], box {64, 64}
                  const __grid_constant__ CUtensorMap mapMhi,   // 2-D [(T + 1) * Bs, H] fp16, box {64, 128}
                  const __grid_constant__ CUtensorMap mapMlo,
                  Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NG = a.NG;
    const uint32_t rank = cluster_ctarank_(), p = rank >> 1, q = rank & 1, leader = rank & ~1u, kpeer = rank ^ 2u;
    const uint16_t pair_mask = (uint16_t)(0x3u << (2 * p));

    const int Kc = a.H / 2, KB = Kc / 64;
    const uint32_t whi0 = base, wlo0 = base + (uint32_t)KB * WT, WBYTES = 2u * (uint32_t)KB * WT;
    const uint32_t ring0 = base + WBYTES;
    const uint32_t exr0 = ring0 + RST * TILE;
    const uint32_t bar0 = exr0 + XB_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };                  // used on the leader: both CTAs' tile bytes
    auto empty_bar = [&](int s) { return bar0 + 8u * (RST + s); };         // in every CTA: the leader's commits are multicast to the pair
    const uint32_t wfull_bar = bar0 + 8u * (2 * RST);                      // local: my W slice
    const uint32_t wpair_bar = wfull_bar + 8u;                             // leader: the partner's W slice is resident too
    auto tfull_bar = [&](int s) { return wpair_bar + 8u + 8u * s; };       // in every CTA (multicast commit)
    auto tempty_bar = [&](int s) { return wpair_bar + 8u + 8u * (2 + s); };   // leader: both CTAs' epilogue warps
    auto rel_bar = [&](int g) { return wpair_bar + 8u + 8u * (4 + g); };
    const uint32_t prfull_bar = wpair_bar + 8u + 8u * 6, prfree_bar = prfull_bar + 8u;
    const uint32_t tmem_slot = prfree_bar + 8u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RST; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MMAW); }
        mbar_init(wfull_bar, 1); mbar_init(wpair_bar, 1);
        for (int s = 0; s < NSLOT; s++) { mbar_init(tfull_bar(s), MMAW); mbar_init(tempty_bar(s), 2 * EPI_WARPS); }
        for (int g = 0; g < MAXI; g++) mbar_init(rel_bar(g), EPI_WARPS);
        mbar_init(prfull_bar, EPI_WARPS); mbar_init(prfree_bar, EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                                                       // the peers' barriers exist before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer: my W slice once, then my session half of every item's m tiles =====================
        if (lane == 0) { tma_prefetch_desc(&mapW); tma_prefetch_desc(&mapMhi); tma_prefetch_desc(&mapMlo); }
        if (elect_one()) {
            mbar_expect_tx(wfull_bar, WBYTES);
            for (int kb = 0; kb < KB; kb++) {
                tma_load_2d(whi0 + kb * WT, &mapW, wfull_bar, kb * 64, ((int)blockIdx.x * 2 + 0) * NB);
                tma_load_2d(wlo0 + kb * WT, &mapW, wfull_bar, kb * 64, ((int)blockIdx.x * 2 + 1) * NB);
            }
        }
        __syncwarp();
        if (q == 1) {                                                      // tell the leader when my half of the B operand has landed
            mbar_wait(wfull_bar, 0);
            if (lane == 0) mbar_arrive_remote(mapa_(wpair_bar, leader));
            __syncwarp();
        }
        uint32_t n = 0;
        for (int t = 0; t < a.T; t++)
            for (int g = 0; g < NG; g++) {
                if (t > 0) {
                    if (lane == 0) wait_counter(a.counter + g * 64, (unsigned)t * (unsigned)a.GC);     // m(t-1) of the item complete everywhere
                    __syncwarp();
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                if (lane == 0) TRACE5(0);
                const int row = t * a.Bs + a.g0 + a.goff[g] + (int)q * HALF;     // my half of the item's sessions; row block t holds m(t-1)
                for (int kb = 0; kb < KB; kb++)
                    for (int hl = 0; hl < 2; hl++, n++) {
                        const int s = (int)(n % RST);
                        mbar_wait(empty_bar(s), ((n / RST) & 1) ^ 1);
                        if (elect_one()) {
                            if (q == 0) mbar_expect_tx(full_bar(s), 2u * TILE);                       // both CTAs' tiles land on the leader's barrier
                            tma_load_2d_pair(ring0 + s * TILE, hl ? &mapMlo : &mapMhi, mapa_(full_bar(s), leader), (int)p * Kc + kb * 64, row);
                        }
                        __syncwarp();
                    }
                if (lane == 0) TRACE5(1);
            }
    } else if (warp <= MMAW) {
        // ===================== MMA issuers (pair leader only): warp mw owns K-slices (ks % MMAW == mw) and its own accumulator =====================
        if (q == 0) {
            const int mw = warp - 1;
            constexpr uint32_t idesc = make_idesc_f16(2 * HALF, NCT);
            mbar_wait(wfull_bar, 0);
            mbar_wait_cluster(wpair_bar, 0);
            uint32_t n = 0, item = 0;
            for (int t = 0; t < a.T; t++)
                for (int g = 0; g < NG; g++, item++) {
                    const int slot = (int)(item % NSLOT);
                    mbar_wait_cluster(tempty_bar(slot), ((item / NSLOT) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)((slot * MMAW + mw) * NCT);
                    for (int kb = 0; kb < KB; kb++)
                        for (int hl = 0; hl < 2; hl++, n++) {
                            const int s = (int)(n % RST);
                            mbar_wait(full_bar(s), (n / RST) & 1);
                            if (kb == 0 && hl == 0 && mw == 0 && lane == 0) TRACE5(2);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t ad = make_sw128_desc(ring0 + s * TILE);
                                const uint64_t bh = make_sw128_desc(whi0 + kb * WT), bl = make_sw128_desc(wlo0 + kb * WT);
#pragma unroll
                                for (int ks = mw; ks < 4; ks += MMAW) {
                                    if (hl == 0) {
                                        umma_f16_pair(d_tmem, ad + 2 * ks, bh + 2 * ks, idesc, (kb != 0 || ks != mw) ? 1u : 0u);   // m_hi W_hi
                                        umma_f16_pair(d_tmem, ad + 2 * ks, bl + 2 * ks, idesc, 1u);                               // m_hi W_lo
                                    } else {
                                        umma_f16_pair(d_tmem, ad + 2 * ks, bh + 2 * ks, idesc, 1u);                               // m_lo W_hi
                                    }
                                }
                                umma_commit_pair(empty_bar(s), pair_mask);                                                        // stage free in both CTAs
                            }
                            __syncwarp();
                        }
                    if (elect_one()) umma_commit_pair(tfull_bar(slot), pair_mask);
                    __syncwarp();
                    if (mw == 0 && lane == 0) TRACE5(3);
                }
        }
    } else if (warp < 1 + MMAW + EPI_WARPS) {
        // ===================== epilogue: one thread = one of my 128 sessions x 8 units (one owner block of 32 columns) =====================
        const int ew = warp - (1 + MMAW);
        const int qt = warp & 3, hf = ew >> 2;                             // TMEM lane quarter; which of my two owner blocks
        const int sl = qt * 32 + lane;                                     // session lane within my half
        const uint32_t lane_off = (uint32_t)(qt * 32) << 16;
        const uint32_t own_mine = 2u * p + (uint32_t)hf, own_send = 2u * (1u - p) + (uint32_t)hf;   // owner blocks: columns [32 o, 32 o + 32)
        const int ucol = ((int)blockIdx.x / CL) * 32 + (int)own_mine * 8;  // first of my 8 units
        float cst[MAXI][8];
        int nlive[MAXI];
#pragma unroll
        for (int g = 0; g < MAXI; g++) {
            const int si = (int)q * HALF + sl;                             // session index within the item
            const bool valid = g < NG && si < a.Bg[g];
            const int sess = a.g0 + a.goff[g] + si;
            nlive[g] = valid ? a.nsteps[sess] : 0;
#pragma unroll
            for (int h4 = 0; h4 < 2; h4++) {
                const float4 c4 = valid ? *reinterpret_cast<const float4*>(a.c_state + (size_t)sess * a.H + ucol + 4 * h4) : make_float4(0.f, 0.f, 0.f, 0.f);
                cst[g][4 * h4] = c4.x; cst[g][4 * h4 + 1] = c4.y; cst[g][4 * h4 + 2] = c4.z; cst[g][4 * h4 + 3] = c4.w;
            }
        }
        const uint32_t tempty_leader0 = mapa_(tempty_bar(0), leader), prfull_peer = mapa_(prfull_bar, kpeer), prfree_peer = mapa_(prfree_bar, kpeer);
        for (int t = 0; t < a.T; t++) {
#pragma unroll
            for (int g = 0; g < MAXI; g++) {
                if (g >= NG) break;
                const uint32_t item = (uint32_t)(t * NG + g);
                const int slot = (int)(item % NSLOT);
                const int si = (int)q * HALF + sl;
                const bool valid = si < a.Bg[g];
                const bool warp_valid = (int)q * HALF + qt * 32 < a.Bg[g];
                const int sess = a.g0 + a.goff[g] + si;
                mbar_wait(tfull_bar(slot), (item / NSLOT) & 1);
                if (ew == 0 && lane == 0) TRACE5(4);
                tc_fence_after();
                if (item > 0) mbar_wait_cluster(prfree_bar, (item - 1) & 1);                 // the K partner is done with what it received last item
                float mine[32];
#pragma unroll
                for (int j = 0; j < 32; j++) mine[j] = 0.f;
                if (warp_valid) {
                    const uint32_t tb = tmem_base + lane_off + (uint32_t)(slot * MMAW * NCT);
#pragma unroll
                    for (int h16 = 0; h16 < 2; h16++) {                    // the partner's columns first: they have the longer way
                        uint32_t r0[16], r1[16];
                        tmem_ld16_nowait(tb + own_send * 32u + (uint32_t)h16 * 16u, r0);
                        tmem_ld16_nowait(tb + (uint32_t)NCT + own_send * 32u + (uint32_t)h16 * 16u, r1);
                        tc_wait_ld();
                        if (valid) {
                            const uint32_t dst = mapa_(exr0 + (uint32_t)((hf * 8 + h16 * 4) * HALF + sl) * 16u, kpeer);
#pragma unroll
                            for (int c = 0; c < 4; c++)
                                st_remote_v4(dst + (uint32_t)c * (HALF * 16u), __uint_as_float(r0[4 * c]) + __uint_as_float(r1[4 * c]), __uint_as_float(r0[4 * c + 1]) + __uint_as_float(r1[4 * c + 1]),
                                             __uint_as_float(r0[4 * c + 2]) + __uint_as_float(r1[4 * c + 2]), __uint_as_float(r0[4 * c + 3]) + __uint_as_float(r1[4 * c + 3]));
                        }
                    }
#pragma unroll
                    for (int h16 = 0; h16 < 2; h16++) {
                        uint32_t r0[16], r1[16];
                        tmem_ld16_nowait(tb + own_mine * 32u + (uint32_t)h16 * 16u, r0);
                        tmem_ld16_nowait(tb + (uint32_t)NCT + own_mine * 32u + (uint32_t)h16 * 16u, r1);
                        tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; j++) mine[16 * h16 + j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_remote(tempty_leader0 + 8u * (uint32_t)slot);          // accumulators drained (this warp's part): the leader's barrier counts both CTAs
                    mbar_arrive_remote(prfull_peer);
                }
                if (ew == 0 && lane == 0) TRACE5(5);
                // input gates of my 8 units: in flight while the partner's partial sums arrive. Column block layout: [hc][gate][u4]
                float4 gin[2][4];
                if (valid) {
                    const float* gp = a.G + ((size_t)t * a.Bs + sess) * 4 * a.H + ucol;
#pragma unroll
                    for (int hc = 0; hc < 2; hc++)
#pragma unroll
                        for (int g4 = 0; g4 < 4; g4++) gin[hc][g4] = __ldcs(reinterpret_cast<const float4*>(gp + (size_t)g4 * a.H + 4 * hc));
                }
                mbar_wait_cluster(prfull_bar, item & 1);                   // the other K half's sums for my units
                if (valid) {
                    const float4* xr = reinterpret_cast<const float4*>(gptr(exr0 + (uint32_t)((hf * 8) * HALF + sl) * 16u));
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const float4 x = xr[(size_t)c * HALF];
                        mine[4 * c] += x.x; mine[4 * c + 1] += x.y; mine[4 * c + 2] += x.z; mine[4 * c + 3] += x.w;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(prfree_peer);
                if (ew == 0 && lane == 0) TRACE5(6);
                if (valid) {
                    const bool live = a.t0 + t < nlive[g];
                    __align__(16) __half mh[8], ml[8];
                    float mval[8];
#pragma unroll
                    for (int hc = 0; hc < 2; hc++) {
                        const float xi[4] = { gin[hc][0].x, gin[hc][0].y, gin[hc][0].z, gin[hc][0].w }, xf[4] = { gin[hc][1].x, gin[hc][1].y, gin[hc][1].z, gin[hc][1].w },
                                    xg[4] = { gin[hc][2].x, gin[hc][2].y, gin[hc][2].z, gin[hc][2].w }, xo[4] = { gin[hc][3].x, gin[hc][3].y, gin[hc][3].z, gin[hc][3].w };
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float* mm = mine + 16 * hc;              // [gate][u4]
                            const float pi = mm[u] * a.inv_scale + xi[u], pf = mm[4 + u] * a.inv_scale + xf[u];
                            const float pg = mm[8 + u] * a.inv_scale + xg[u], po = mm[12 + u] * a.inv_scale + xo[u];
                            const float cn = fast_sigmoid(pf) * cst[g][4 * hc + u] + fast_sigmoid(pi) * fast_tanh(pg);
                            mval[4 * hc + u] = fast_sigmoid(po) * fast_tanh(cn);
                            if (live) cst[g][4 * hc + u] = cn;
                            split_f16(mval[4 * hc + u], mh[4 * hc + u], ml[4 * hc + u]);
                        }
                    }
                    const size_t orow = ((size_t)(t + 1) * a.Bs + sess) * a.H + ucol;
                    *reinterpret_cast<uint4*>(a.Mx_hi + orow) = *reinterpret_cast<const uint4*>(mh);
                    *reinterpret_cast<uint4*>(a.Mx_lo + orow) = *reinterpret_cast<const uint4*>(ml);
                    if (live && ((a.t0 + t == nlive[g] - 1) || (t == a.T - 1))) {
                        *reinterpret_cast<float4*>(a.m_state + (size_t)sess * a.H + ucol) = make_float4(mval[0], mval[1], mval[2], mval[3]);
                        *reinterpret_cast<float4*>(a.m_state + (size_t)sess * a.H + ucol + 4) = make_float4(mval[4], mval[5], mval[6], mval[7]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(rel_bar(g));
            }
        }
#pragma unroll
        for (int g = 0; g < MAXI; g++) {
            const int si = (int)q * HALF + sl;
            if (g < NG && si < a.Bg[g]) {
                float* cp = a.c_state + (size_t)(a.g0 + a.goff[g] + si) * a.H + ucol;
                *reinterpret_cast<float4*>(cp) = make_float4(cst[g][0], cst[g][1], cst[g][2], cst[g][3]);
                *reinterpret_cast<float4*>(cp + 4) = make_float4(cst[g][4], cst[g][5], cst[g][6], cst[g][7]);
            }
        }
    } else {
        // ===================== releaser: the gpu-scope release (store acknowledgement ~1 us) must not hold the epilogue =====================
        if (lane == 0)
            for (int t = 0; t < a.T; t++)
                for (int g = 0; g < NG; g++) {
                    mbar_wait(rel_bar(g), t & 1);
                    asm volatile("fence.proxy.async;" ::: "memory");
                    red_release(a.counter + g * 64);
                    TRACE5(7);
                }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                                                       // nobody exits while a peer may still store into it, read its tiles or signal its barriers
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// packs W_eff (scaled by `sc`) into the pair kernel's layout: CTA c = 4 cl + 2 p + q holds, for K half p, the rows n = 64 q + j of the
// cluster's 128:  owner o = n / 32 (units cl * 32 + 8 o ...), hc = (n / 16) % 2, gate = (n / 4) % 4, u4 = n % 4:
// dst[((c * 2 + hl) * 64 + j) * (H / 2) + k]
__global__ void pack_weff5_kernel(const float* __restrict__ weff, __half* __restrict__ dst, int H, float sc) {
    const int Kc = H / 2;
    const int c = blockIdx.x, j = blockIdx.y;
    const int cl = c >> 2, p = (c >> 1) & 1, q = c & 1;
    const int n = 64 * q + j, o = n >> 5, hc = (n >> 4) & 1, g4 = (n >> 2) & 3, u4 = n & 3;
    const int unit = cl * 32 + o * 8 + hc * 4 + u4;
    const float* src = weff + ((size_t)g4 * H + unit) * H + (size_t)p * Kc;
    __half* dh = dst + (((size_t)c * 2 + 0) * NB + j) * Kc;
    __half* dl = dst + (((size_t)c * 2 + 1) * NB + j) * Kc;
    for (int k = threadIdx.x; k < Kc; k += blockDim.x) {
        const float x = src[k] * sc;
        const __half h = __float2half_rn(x);
        dh[k] = h; dl[k] = __float2half_rn(x - __half2float(h));
    }
}

}  // namespace scan5
}  // namespace april
