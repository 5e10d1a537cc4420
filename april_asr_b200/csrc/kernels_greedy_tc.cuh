// J1 + G1 on tensor cores: joiner (logits = W_o tanh(enc + dec) + b, april_session.c:166-179; raw logits, no softmax: SURVEY.md section 9)
// and the part of the greedy search that feeds back into the model (april_session.c:306-429,431-476) for up to 32 sessions per cluster.
//
// Why a second search kernel: the SIMT cluster kernel (kernels_simt.cuh) multiplies register-resident joiner rows by z for 4-8
// sessions per cluster and spends ~14 us per joiner pass in barriers, table loads and the matvec; with 1 200 dependent passes per
// 30 s pass that is 16.6 ms at 64 sessions and 67 ms (3.5 waves of clusters) at 512. Here a cluster of 8 CTAs walks 32 sessions:
//  * sessions sit on the UMMA M dimension (TMEM lane = session), CTA r keeps vocabulary rows [64 r, 64 r + 64) of W_o resident in
//    shared memory as the B operand (fp16 hi / lo, pre-scaled by a power of two), N = 64, K = J;
//  * every CTA builds the whole A operand itself: z = tanh(e + d) for its 32 sessions x J, split into fp16 (hi, lo) and written
//    straight into 128B-swizzled K-major tiles (the layout TMA would produce) — 16 K tanh per pass and CTA, no exchange of z;
//  * three products per k-slice (z_hi W_hi + z_hi W_lo + z_lo W_hi) into fp32 TMEM accumulators, four issuing warps;
//  * one warp (TMEM lanes 0-31 = the sessions) drains its session's 64 logits, takes the local arg-max over non-blank ids (strict >,
//    ascending id: lowest index wins ties, april_session.c:311-320) and publishes (best, index, blank logit) to all 8 CTAs over
//    distributed shared memory; after one cluster barrier every CTA evaluates the same greedy rule on the same 8 partials
//    (replicated, deterministic state), exactly as the SIMT cluster kernel does;
//  * the stateless decoder is the same [V, V, J] table lookup.
// 18 clusters x 32 sessions = 576 sessions advance in one wave.
#pragma once
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"

namespace april {
namespace gt {

using namespace tc;
using k::DecisionRecord;
using k::GreedyState;
using k::JoinerParams;

constexpr int CLUSTER = 8;
constexpr int RV = 64;                 // vocabulary rows per CTA = UMMA N
constexpr int MAX_S = 32;              // sessions per cluster = lanes of the draining warp
constexpr int BUILD_WARPS = 8;
constexpr int MMA_WARPS = 4;
constexpr int THREADS = (BUILD_WARPS + MMA_WARPS) * 32;
constexpr uint32_t A_TILE = MAX_S * 128u;          // [32 sessions x 64 k] fp16, 128B-swizzled: 4 KB (the M = 128 MMA over-reads 12 KB of whatever follows: dead lanes)
constexpr uint32_t W_TILE = RV * 128u;             // 8 KB

struct Partial { float best_val, best_idx, blank_val, has_blank; };

struct Args {
    int J, V, blank;
    int B, Tcap, t_base, chunk_T, stride_ms, last_chunk, S;    // S: sessions per cluster (<= 32)
    float inv_scale;
    const float* bias;            // [V]
    const uint8_t* vflags;        // [V]
    const float* dtab;            // [V][V][J]
    const float* enc_out;         // [chunk_T * B, J], time-major
    const int* nsteps;            // [B]
    const int* slot;              // [B]
    const int* flush_flag;        // [B]
    GreedyState* gstate;          // [slots]
    DecisionRecord* rec;          // [B][Tcap][3]
    int* npass;                   // [B][Tcap]
    float* tap_logits;            // [B][Tcap][3][V] or null
    unsigned long long* trace;    // dev: [passes][8] globaltimer stamps of cluster 0 / rank 0, or null
    int trace_cap;
};

static inline size_t smem_bytes(int J, int V) {
    const size_t kb = (size_t)J / 64;
    return 1024 + 2 * kb * A_TILE + 2 * kb * W_TILE + 2 * CLUSTER * MAX_S * 16 /* inbox */ + MAX_S * sizeof(GreedyState) + 1024 /* per-session flags, row pointers */ + 128 /* barriers */ + (size_t)V + 64;
}

__device__ __forceinline__ uint32_t cluster_rank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_remote_v4(uint32_t laddr, uint32_t rank, float a, float b, float c, float d) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer_() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define GT_TRACE(slot_) do { if (a.trace && blockIdx.x == 0 && (int)pc < a.trace_cap) a.trace[(size_t)pc * 8 + (slot_)] = gtimer_(); } while (0)
__device__ __forceinline__ void build_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(BUILD_WARPS * 32) : "memory"); }

// grid = clusters * 8 CTAs, cluster dims (8,1,1); mapW: 2-D [8 * 2 * 64, J] fp16 = per rank [hi rows ; lo rows], box {64, 64}
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(THREADS, 1)
joiner_greedy_tc_kernel(const __grid_constant__ CUtensorMap mapW, Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto gptr = [&](uint32_t saddr) { return base_ptr + (saddr - base); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank_();
    const int J = a.J, V = a.V, KB = J / 64;
    const int cid = blockIdx.x / CLUSTER;
    const int s0 = cid * a.S;
    const int ns = min(a.S, a.B - s0);                                    // cluster-uniform

    const uint32_t a_hi0 = base, a_lo0 = base + (uint32_t)KB * A_TILE;    // A tiles first: their over-read lands in the W region
    const uint32_t w_hi0 = base + 2u * (uint32_t)KB * A_TILE, w_lo0 = w_hi0 + (uint32_t)KB * W_TILE;
    const uint32_t inbox0 = w_lo0 + (uint32_t)KB * W_TILE;                // [2 parity][8 ranks][32 sessions] x 16 B = 8 KB
    const uint32_t st0 = inbox0 + 2u * CLUSTER * MAX_S * 16u;             // GreedyState[32] (176 B each)
    const uint32_t misc0 = st0 + MAX_S * (uint32_t)sizeof(GreedyState);
    const uint32_t bar0 = (misc0 + 1024u + 7u) & ~7u;
    const uint32_t vf0 = bar0 + 64u;
    GreedyState* st = reinterpret_cast<GreedyState*>(gptr(st0));
    int* s_done = reinterpret_cast<int*>(gptr(misc0));                    // [32]
    int* s_np = s_done + MAX_S;
    int* s_n = s_np + MAX_S;
    const float** s_drow = reinterpret_cast<const float**>(s_n + 2 * MAX_S);     // [32] decoder rows of the current contexts (8-byte aligned: misc0 is)
    uint8_t* vf_s = gptr(vf0);
    const uint32_t wfull_bar = bar0, afull_bar = bar0 + 8u, tfull_bar = bar0 + 16u, tmem_slot = bar0 + 24u;

    if (threadIdx.x == 0) {
        mbar_init(wfull_bar, 1); mbar_init(afull_bar, 1); mbar_init(tfull_bar, MMA_WARPS);
        fence_barrier_init();
    }
    if (warp == BUILD_WARPS) tmem_alloc(tmem_slot, 256);
    for (int i = threadIdx.x; i < V; i += blockDim.x) vf_s[i] = a.vflags[i];
    if (threadIdx.x < MAX_S) {
        const int s = threadIdx.x;
        if (s < ns) { st[s] = a.gstate[a.slot[s0 + s]]; s_n[s] = a.nsteps[s0 + s]; s_drow[s] = a.dtab + ((size_t)st[s].ctx0 * V + st[s].ctx1) * J; }
        else s_n[s] = 0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gptr(tmem_slot));
    if (warp == BUILD_WARPS && elect_one()) {                            // my 64 vocabulary rows, resident for the whole launch
        mbar_expect_tx(wfull_bar, 2u * (uint32_t)KB * W_TILE);
        for (int kb = 0; kb < KB; kb++) {
            tma_load_2d(w_hi0 + kb * W_TILE, &mapW, wfull_bar, kb * 64, ((int)rank * 2 + 0) * RV);
            tma_load_2d(w_lo0 + kb * W_TILE, &mapW, wfull_bar, kb * 64, ((int)rank * 2 + 1) * RV);
        }
    }
    int t_end = 0;
    for (int s = 0; s < ns; s++) t_end = max(t_end, min(s_n[s], a.t_base + a.chunk_T));
    const int row0 = (int)rank * RV;
    uint32_t pc = 0;                                                      // joiner passes done by this cluster: barrier phase = pc & 1
    constexpr uint32_t idesc = make_idesc_f16(128, RV);

    for (int t = a.t_base; t < t_end; t++) {
        if (threadIdx.x < MAX_S) {
            const int s = threadIdx.x;
            const bool act = s < ns && t < s_n[s];
            s_done[s] = act ? 0 : 1; s_np[s] = 0;
            if (act) st[s].current_ms += a.stride_ms;                     // april_session.c:442-443
        }
        __syncthreads();
        for (int pass = 0; pass < 3; pass++, pc++) {
            if (threadIdx.x == 0) GT_TRACE(0);
            if (warp < BUILD_WARPS) {
                // ---- A: z = tanh(e + d) for the sessions still in this step, as fp16 (hi, lo) 128B-swizzled K-major rows
                const int cps = J / 8;                                    // 16-byte chunks per session row
                const float* e_t = a.enc_out + ((size_t)(t - a.t_base) * a.B + s0) * J;
                for (int idx = threadIdx.x; idx < ns * cps; idx += BUILD_WARPS * 32) {
                    const int s = idx / cps, c = idx - s * cps;
                    if (s_done[s]) continue;
                    const float4 e0 = __ldg(reinterpret_cast<const float4*>(e_t + (size_t)s * J + 8 * c)), e1 = __ldg(reinterpret_cast<const float4*>(e_t + (size_t)s * J + 8 * c + 4));
                    const float* dr = s_drow[s] + 8 * c;
                    const float4 d0 = __ldg(reinterpret_cast<const float4*>(dr)), d1 = __ldg(reinterpret_cast<const float4*>(dr + 4));
                    const float z[8] = { tanhf(e0.x + d0.x), tanhf(e0.y + d0.y), tanhf(e0.z + d0.z), tanhf(e0.w + d0.w),
                                         tanhf(e1.x + d1.x), tanhf(e1.y + d1.y), tanhf(e1.z + d1.z), tanhf(e1.w + d1.w) };
                    __align__(16) __half h[8], l[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) split_f16(z[i], h[i], l[i]);
                    const int kb = c >> 3, cc = c & 7;
                    const uint32_t off = (uint32_t)kb * A_TILE + (uint32_t)(s >> 3) * 1024u + (uint32_t)(s & 7) * 128u + (uint32_t)((cc ^ (s & 7)) << 4);
                    *reinterpret_cast<uint4*>(gptr(a_hi0 + off)) = *reinterpret_cast<const uint4*>(h);
                    *reinterpret_cast<uint4*>(gptr(a_lo0 + off)) = *reinterpret_cast<const uint4*>(l);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic-proxy tile stores -> the tensor core's (async proxy) reads
                tc_fence_before();                                        // and my tcgen05.ld of the previous pass precede the next MMAs
                if (threadIdx.x == 0) GT_TRACE(1);
                build_bar_sync();
                if (threadIdx.x == 0) { mbar_arrive(afull_bar); GT_TRACE(2); }
            } else {
                // ---- B: logits of my 64 rows for all 32 session lanes: warp mw issues K-slice mw of every k-block into its accumulator
                const int mw = warp - BUILD_WARPS;
                if (pc == 0) mbar_wait(wfull_bar, 0);
                mbar_wait(afull_bar, pc & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(mw * RV);
                    for (int kb = 0; kb < KB; kb++) {
                        const uint64_t ah = make_sw128_desc(a_hi0 + kb * A_TILE) + 2 * mw, al = make_sw128_desc(a_lo0 + kb * A_TILE) + 2 * mw;
                        const uint64_t bh = make_sw128_desc(w_hi0 + kb * W_TILE) + 2 * mw, bl = make_sw128_desc(w_lo0 + kb * W_TILE) + 2 * mw;
                        umma_f16(d_tmem, ah, bh, idesc, kb != 0 ? 1u : 0u);
                        umma_f16(d_tmem, ah, bl, idesc, 1u);
                        umma_f16(d_tmem, al, bh, idesc, 1u);
                    }
                    umma_commit(tfull_bar);
                    if (mw == 0) GT_TRACE(3);
                }
                __syncwarp();
            }
            // ---- C: warp 0 (TMEM lanes 0-31 = my sessions): drain, local arg-max over my rows, publish to all 8 CTAs
            const int parity = (int)(pc & 1);
            if (warp == 0) {
                mbar_wait(tfull_bar, pc & 1);
                if (lane == 0) GT_TRACE(4);
                tc_fence_after();
                const int s = lane;
                float bv = (float)(-9999999999.0); int bi = -1; float blank_v = 0.f; int has_blank = 0;     // april_session.c:311-312
                const bool mine = s < ns && !s_done[s];
                float* tl = (a.tap_logits && mine) ? a.tap_logits + (((size_t)(s0 + s) * a.Tcap + t) * 3 + pass) * V : nullptr;
#pragma unroll
                for (int c16 = 0; c16 < RV / 16; c16++) {
                    uint32_t r0[16], r1[16], r2[16], r3[16];
                    tmem_ld16(tmem_base + (uint32_t)(0 * RV + c16 * 16), r0);
                    tmem_ld16(tmem_base + (uint32_t)(1 * RV + c16 * 16), r1);
                    tmem_ld16(tmem_base + (uint32_t)(2 * RV + c16 * 16), r2);
                    tmem_ld16(tmem_base + (uint32_t)(3 * RV + c16 * 16), r3);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int v = row0 + c16 * 16 + j;                // ascending id inside and across CTAs
                        if (v < V) {
                            const float x = ((__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + (__uint_as_float(r2[j]) + __uint_as_float(r3[j]))) * a.inv_scale + __ldg(a.bias + v);
                            if (tl) tl[v] = x;
                            if (v == a.blank) { blank_v = x; has_blank = 1; }
                            else if (x > bv) { bv = x; bi = v; }
                        }
                    }
                }
                tc_fence_before();
                if (mine) {
                    const uint32_t slot_addr = inbox0 + (uint32_t)((parity * CLUSTER + (int)rank) * MAX_S + s) * 16u;
#pragma unroll
                    for (uint32_t r = 0; r < CLUSTER; r++) st_remote_v4(slot_addr, r, bv, (float)bi, blank_v, has_blank ? 1.f : 0.f);
                }
            }
            if (threadIdx.x == 0) GT_TRACE(5);
            cluster_sync_();                                              // every CTA's partials have landed everywhere
            if (threadIdx.x == 0) GT_TRACE(6);
            // ---- D: every CTA evaluates the same rule on the same 8 partials (replicated state)
            if (threadIdx.x < ns && !s_done[threadIdx.x]) {
                const int s = threadIdx.x;
                const Partial* inbox = reinterpret_cast<const Partial*>(gptr(inbox0));
                float mv = (float)(-9999999999.0); int mi = -1; float blank_v = 0.f;
                for (int c = 0; c < CLUSTER; c++) {                       // rank order == ascending vocabulary order
                    const Partial p = inbox[(parity * CLUSTER + c) * MAX_S + s];
                    if (p.has_blank != 0.f) blank_v = p.blank_val;
                    if ((int)p.best_idx >= 0 && p.best_val > mv) { mv = p.best_val; mi = (int)p.best_idx; }
                }
                bool ctx_changed;
                const bool is_blank = k::greedy_decide(st[s], vf_s, a.blank, pass == 0 ? 1.0f : 0.0f, mi, mv, blank_v, ctx_changed);
                if (rank == 0) {
                    DecisionRecord r; r.max_idx = mi; r.is_blank = is_blank ? 1 : 0; r.max_val = mv; r.blank_val = blank_v;
                    a.rec[((size_t)(s0 + s) * a.Tcap + t) * 3 + pass] = r;
                }
                s_np[s] = pass + 1;
                if (ctx_changed) s_drow[s] = a.dtab + ((size_t)st[s].ctx0 * V + st[s].ctx1) * J;   // aas_run_decoder = table row of the new context
                if (is_blank) s_done[s] = 1;
            }
            __syncthreads();
            if (threadIdx.x == 0) GT_TRACE(7);
            bool all_done = true;
            for (int s = 0; s < ns; s++) all_done &= (s_done[s] != 0);
            if (all_done) { pc++; break; }
        }
        if (rank == 0 && threadIdx.x < ns && t < s_n[threadIdx.x]) a.npass[(size_t)(s0 + threadIdx.x) * a.Tcap + t] = s_np[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < ns) {
        const int s = threadIdx.x;
        if (a.last_chunk && a.flush_flag[s0 + s]) {                       // _aas_flush tail, april_session.c:561-563
            st[s].head = 0;
            if (st[s].ctx0 != a.blank) { st[s].ctx0 = a.blank; st[s].ctx1 = a.blank; }
        }
        if (rank == 0) a.gstate[a.slot[s0 + s]] = st[s];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_();                                                      // no CTA may exit while peers can still write its inbox
    if (warp == BUILD_WARPS) tmem_dealloc(tmem_base, 256);
}

// packs the joiner output matrix [V, J] (scaled by `sc`, a power of two) as per-rank [hi rows ; lo rows] of 64 vocabulary rows; rows >= V are zero
__global__ void pack_joiner_kernel(const float* __restrict__ w, __half* __restrict__ dst, int V, int J, float sc) {
    const int r = blockIdx.x, n = blockIdx.y, v = r * RV + n;
    __half* dh = dst + (((size_t)r * 2 + 0) * RV + n) * J;
    __half* dl = dst + (((size_t)r * 2 + 1) * RV + n) * J;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        const float x = v < V ? w[(size_t)v * J + j] * sc : 0.f;
        const __half h = __float2half_rn(x);
        dh[j] = h; dl[j] = __float2half_rn(x - __half2float(h));
    }
}

}  // namespace gt
}  // namespace april
