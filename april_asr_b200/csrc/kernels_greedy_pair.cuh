// J1 + G1 on CTA PAIRS: the joiner matrix streamed from L2 instead of held in registers.
//
// greedy_cluster_kernel (kernels_simt.cuh) keeps the [V, J] joiner matrix in the registers of an 8-CTA cluster: a joiner pass costs
// ~10 us of barriers and dependent loads whatever the batch, and every 8 sessions occupy 8 SMs — 512 sessions need four waves of
// clusters, and the search cannot run beside the encoder. Measured on B200: an SM ingests ~28 B/clk, so a CTA that STREAMS half of the
// matrix (V/2 rows x J floats = 512 KB for V = 500, J = 512) through shared memory needs the same ~10 us per pass. This kernel does
// that: a pair of CTAs (one TPC) serves up to 8 sessions; CTA r multiplies rows [r * Rh, r * Rh + Rh) for all of them. 512 sessions
// are one wave of 128 CTAs; 64 sessions occupy 16 SMs, which leaves the rest of the GPU to the encoder of the next time chunk
// (engine.cu: run_pass).
//
// Same arithmetic as greedy_cluster_kernel, bit for bit: lane l holds z[l + 32 i], accumulates its 16 products per row in order,
// the 32 lanes are summed by the same butterfly, the bias is added last; arg-max ties go to the lowest index; the greedy rules are
// greedy_decide (april_session.c:306-429), evaluated redundantly by both CTAs on the same two partials.
//
// Eight warps (4 rows x 8 sessions per warp and stage); one 1-D bulk copy of 32 rows (64 KB) per stage, two stages: thread 0 issues
// the copy of stage n + 1 before the CTA multiplies stage n, across passes too, so the first 64 KB of the next pass land during the
// arg-max / decision phase. (A ninth, producing warp would put three warps on one scheduler and cap the kernel at 168 registers.)
#pragma once
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"

namespace april {
namespace k {

constexpr int GP_MAX_S = 8;           // sessions per CTA pair
constexpr int GP_ROWS = 32;           // joiner rows per stage: 8 consumer warps x 4 rows
constexpr int GP_STAGES = 2;
constexpr int GP_CONSUMERS = 256;
constexpr int GP_THREADS = GP_CONSUMERS;
constexpr int GP_MAX_JPL = 16;        // J / 32

static inline size_t greedy_pair_smem(int J, int V, int Rh) {
    const size_t Rp = (size_t)((Rh + 3) & ~3);
    return sizeof(float) * ((size_t)GP_STAGES * GP_ROWS * J + (size_t)4 * GP_MAX_S * J + (size_t)GP_MAX_S * Rp + (size_t)Rh) + sizeof(GreedyPartial) * 2 * 2 * GP_MAX_S + (size_t)((V + 15) & ~15) + 128;
}

__device__ __forceinline__ void gp_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void gp_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GP_CONSUMERS) : "memory"); }
__device__ __forceinline__ uint32_t gp_mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void gp_arrive_cluster(uint32_t cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory"); }
__device__ __forceinline__ void gp_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t_start = clock64();
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if ((spin & 0xFFFF) == 0xFFFF && clock64() - t_start > 4000000000LL) { printf("libapril: search pair exchange timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x); __trap(); }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GP_THREADS, 1)
greedy_pair_kernel(JoinerParams P, const float* __restrict__ dtab, const float* __restrict__ enc_out /*[T*B, J] chunk, time-major*/,
                   const int* __restrict__ nsteps, const int* __restrict__ slot, const int* __restrict__ flush_flag, int B, int Tcap, int t_base,
                   int stride_ms, GreedyState* __restrict__ gstate, DecisionRecord* __restrict__ rec, int* __restrict__ npass,
                   float* __restrict__ tap_logits, int last_chunk, int chunk_T, int S, int Rh) {
    extern __shared__ __align__(128) float gp_smem[];
    const int J = P.J, V = P.V;
    const uint32_t rank = cluster_rank();
    const int cid = blockIdx.x / 2;
    const int s0 = cid * S;
    const int ns = min(S, B - s0);                         // sessions of this pair (pair-uniform)
    const int Rp = (Rh + 3) & ~3;
    float* wst = gp_smem;                                 // [2][32][J] joiner rows in flight
    float* es = wst + (size_t)GP_STAGES * GP_ROWS * J;     // [2][S][J] encoder rows: current step / next step
    float* ds = es + (size_t)2 * GP_MAX_S * J;             // [S][J] decoder rows of the current contexts
    float* zs = ds + (size_t)GP_MAX_S * J;                 // [S][J] tanh(e + d)
    float* logit_s = zs + (size_t)GP_MAX_S * J;            // [S][Rp]
    float* bias_s = logit_s + (size_t)GP_MAX_S * Rp;       // [Rh]
    GreedyPartial* inbox = reinterpret_cast<GreedyPartial*>(bias_s + Rh + ((4 - (Rh & 3)) & 3));   // [2 parity][2 ranks][S]
    uint8_t* vf_s = reinterpret_cast<uint8_t*>(inbox + 2 * 2 * GP_MAX_S);                            // [V]
    __shared__ GreedyState st[GP_MAX_S];
    __shared__ int s_done[GP_MAX_S], s_np[GP_MAX_S], s_newctx[GP_MAX_S], s_n[GP_MAX_S];
    __shared__ __align__(8) uint64_t bars[2 * GP_STAGES + 2 + 2 + 1];
    const uint32_t bar0 = tc::smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (GP_STAGES + s); };
    auto efull_bar = [&](int s) { return bar0 + 8u * (2 * GP_STAGES + s); };
    auto xbar = [&](int s) { return bar0 + 8u * (2 * GP_STAGES + 2 + s); };       // both CTAs' partials of a pass have landed (parity s)
    const uint32_t exit_bar = bar0 + 8u * (2 * GP_STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (int)rank * Rh;
    const int rows_end = min(V, row0 + Rh);
    const int nrows = max(0, rows_end - row0);
    const int nst = (nrows + GP_ROWS - 1) / GP_ROWS;      // stages per pass (>= 1: the host launches with Rh <= V)
    const int jpl = J / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GP_STAGES; s++) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), GP_CONSUMERS / 32); }
        for (int s = 0; s < 2; s++) { tc::mbar_init(efull_bar(s), 1); tc::mbar_init(xbar(s), 2); }
        tc::mbar_init(exit_bar, 2);
        tc::fence_barrier_init();
    }
    if (threadIdx.x < GP_CONSUMERS) {
        for (int i = threadIdx.x; i < V; i += GP_CONSUMERS) vf_s[i] = P.vflags[i];
        for (int i = threadIdx.x; i < nrows; i += GP_CONSUMERS) bias_s[i] = P.join_b[row0 + i];
        if (threadIdx.x < ns) { st[threadIdx.x] = gstate[slot[s0 + threadIdx.x]]; s_newctx[threadIdx.x] = 1; s_n[threadIdx.x] = nsteps[s0 + threadIdx.x]; }
    }
    __syncthreads();
    cluster_sync_all();                                    // the peer's barriers exist before anyone signals them

    // stage n of the row stream = rows (n mod nst) * 32 ... of my half, into buffer n mod 2 (called by thread 0 only)
    auto issue_stage = [&](uint32_t n) {
        const int sb = (int)(n % GP_STAGES);
        tc::mbar_wait(empty_bar(sb), ((n / GP_STAGES) & 1) ^ 1);               // every warp is done with the buffer's previous rows
        const int kst = (int)(n % (uint32_t)nst), r0 = row0 + kst * GP_ROWS, nr = min(GP_ROWS, rows_end - r0);
        const uint32_t bytes = (uint32_t)nr * (uint32_t)J * 4u;
        tc::mbar_expect_tx(full_bar(sb), bytes);
        gp_bulk_g2s(tc::smem_u32(wst + (size_t)sb * GP_ROWS * J), P.join_w + (size_t)r0 * J, bytes, full_bar(sb));
    };
    {
        int t_end = 0;
        for (int s = 0; s < ns; s++) t_end = max(t_end, min(s_n[s], t_base + chunk_T));
        const uint32_t ebytes = (uint32_t)ns * (uint32_t)J * 4u;
        if (threadIdx.x == 0 && t_base < t_end) {          // encoder rows of the first step (rows of sessions s0.. are contiguous)
            tc::mbar_expect_tx(efull_bar(0), ebytes);
            gp_bulk_g2s(tc::smem_u32(es), enc_out + (size_t)s0 * J, ebytes, efull_bar(0));
        }
        uint32_t nstage = 0, xround = 0;
        if (threadIdx.x == 0) issue_stage(0);
        for (int t = t_base; t < t_end; t++) {
            const int eb = (t - t_base) & 1;
            const float* e_cur = es + (size_t)eb * GP_MAX_S * J;
            tc::mbar_wait(efull_bar(eb), ((uint32_t)(t - t_base) >> 1) & 1);
            if (threadIdx.x == 0 && t + 1 < t_end) {       // next step's rows land while this step is searched (its buffer was last read in step t - 1)
                tc::mbar_expect_tx(efull_bar(eb ^ 1), ebytes);
                gp_bulk_g2s(tc::smem_u32(es + (size_t)(eb ^ 1) * GP_MAX_S * J), enc_out + ((size_t)(t + 1 - t_base) * B + s0) * J, ebytes, efull_bar(eb ^ 1));
            }
            if (threadIdx.x < ns) {
                const int s = threadIdx.x;
                const bool act = t < s_n[s];
                s_done[s] = act ? 0 : 1; s_np[s] = 0;
                if (act) st[s].current_ms += stride_ms;                       // april_session.c:442-443
            }
            gp_consumer_sync();
            for (int pass = 0; pass < 3; pass++) {
                // --- A: decoder rows for changed contexts (table lookup = aas_run_decoder), z = tanh(e + d)
                {
                    float d_new[GP_MAX_S][2];                                 // all sessions' table rows requested before any is used
#pragma unroll
                    for (int s = 0; s < GP_MAX_S; s++) {
                        const bool nc = s < ns && !s_done[s] && s_newctx[s] != 0;
                        const float* drow = nc ? dtab + ((size_t)st[s].ctx0 * V + st[s].ctx1) * J : dtab;
#pragma unroll
                        for (int u = 0; u < 2; u++) { const int j = threadIdx.x + u * GP_CONSUMERS; d_new[s][u] = (nc && j < J) ? __ldg(drow + j) : 0.f; }
                    }
#pragma unroll
                    for (int s = 0; s < GP_MAX_S; s++) {
                        if (s >= ns || s_done[s]) continue;
                        const bool nc = s_newctx[s] != 0;
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            const int j = threadIdx.x + u * GP_CONSUMERS;
                            if (j < J) {
                                float dv;
                                if (nc) { dv = d_new[s][u]; ds[(size_t)s * J + j] = dv; } else dv = ds[(size_t)s * J + j];
                                zs[(size_t)s * J + j] = tanhf(e_cur[(size_t)s * J + j] + dv);
                            }
                        }
                    }
                }
                gp_consumer_sync();
                // only a session that took part in this pass has consumed its flag (see greedy_cluster_kernel)
                if (threadIdx.x < ns && !s_done[threadIdx.x]) s_newctx[threadIdx.x] = 0;
                // --- B: logits of my rows: z in registers, rows streamed through shared memory
                {
                    float z[GP_MAX_S][GP_MAX_JPL];
#pragma unroll
                    for (int s = 0; s < GP_MAX_S; s++)
#pragma unroll
                        for (int i = 0; i < GP_MAX_JPL; i++) z[s][i] = (s < ns && i < jpl) ? zs[(size_t)s * J + lane + 32 * i] : 0.f;
                    for (int kst = 0; kst < nst; kst++, nstage++) {
                        const int sg = (int)(nstage % GP_STAGES);
                        if (threadIdx.x == 0) issue_stage(nstage + 1);                // the next rows (of the next pass, after the last stage) fly meanwhile
                        tc::mbar_wait(full_bar(sg), (nstage / GP_STAGES) & 1);
                        const float* wr = wst + ((size_t)sg * GP_ROWS + warp * 4) * J + lane;
                        float acc[32];
#pragma unroll
                        for (int p = 0; p < 32; p++) acc[p] = 0.f;
#pragma unroll
                        for (int i = 0; i < GP_MAX_JPL; i++) {
                            if (i < jpl) {
#pragma unroll
                                for (int r = 0; r < 4; r++) {
                                    const float wv = wr[(size_t)r * J + 32 * i];
#pragma unroll
                                    for (int s = 0; s < GP_MAX_S; s++) acc[r * GP_MAX_S + s] = fmaf(wv, z[s][i], acc[r * GP_MAX_S + s]);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(empty_bar(sg));        // the stage's rows are in registers / consumed
                        const float tot = warp_reduce_scatter32(acc, lane);   // lane -> (row lane / 8, session lane % 8)
                        const int lr = kst * GP_ROWS + warp * 4 + lane / GP_MAX_S, sq = lane % GP_MAX_S;
                        if (lr < nrows && sq < ns) logit_s[sq * Rp + lr] = tot + bias_s[lr];
                    }
                }
                gp_consumer_sync();
                // --- C: warp s = arg-max of session s over my rows (ties -> lowest index), published to both CTAs
                const int par = (int)(xround & 1);
                if (warp < ns && !s_done[warp]) {
                    const int s = warp;
                    float bv = (float)(-9999999999.0); int bi = -1; float blank_v = 0.f; int has_blank = 0;
                    float* tl = tap_logits ? tap_logits + (((size_t)(s0 + s) * Tcap + t) * 3 + pass) * V : nullptr;
                    const float* lg = logit_s + s * Rp;
                    for (int r = lane; r < nrows; r += 32) {
                        const int v = row0 + r;
                        const float x = lg[r];
                        if (tl) tl[v] = x;
                        if (v == P.blank) { blank_v = x; has_blank = 1; }
                        else if (x > bv) { bv = x; bi = v; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        const float ob = __shfl_xor_sync(0xffffffffu, blank_v, o);
                        const int oh = __shfl_xor_sync(0xffffffffu, has_blank, o);
                        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
                        if (oh) { blank_v = ob; has_blank = 1; }
                    }
                    if (lane < 2) {
                        GreedyPartial* slot_local = &inbox[(par * 2 + rank) * GP_MAX_S + s];
                        dsmem_store_f32(&slot_local->best_val, lane, bv);
                        dsmem_store_f32(&slot_local->best_idx, lane, (float)bi);
                        dsmem_store_f32(&slot_local->blank_val, lane, blank_v);
                        dsmem_store_f32(&slot_local->has_blank, lane, has_blank ? 1.f : 0.f);
                    }
                }
                gp_consumer_sync();
                if (threadIdx.x < 2) gp_arrive_cluster(gp_mapa(xbar(par), threadIdx.x));   // my partials are out: tell both CTAs
                gp_wait_cluster(xbar(par), (xround >> 1) & 1);
                xround++;
                // --- D: both CTAs evaluate the same rule on the same two partials (replicated state)
                if (threadIdx.x < ns && !s_done[threadIdx.x]) {
                    const int s = threadIdx.x;
                    float mv = (float)(-9999999999.0); int mi = -1; float blank_v = 0.f;
                    for (int c = 0; c < 2; c++) {                             // rank order == ascending vocabulary order
                        const GreedyPartial& p = inbox[(par * 2 + c) * GP_MAX_S + s];
                        if (p.has_blank != 0.f) blank_v = p.blank_val;
                        if ((int)p.best_idx >= 0 && p.best_val > mv) { mv = p.best_val; mi = (int)p.best_idx; }
                    }
                    bool ctx_changed;
                    const bool is_blank = greedy_decide(st[s], vf_s, P.blank, pass == 0 ? 1.0f : 0.0f, mi, mv, blank_v, ctx_changed);
                    if (rank == 0) {
                        DecisionRecord r; r.max_idx = mi; r.is_blank = is_blank ? 1 : 0; r.max_val = mv; r.blank_val = blank_v;
                        rec[((size_t)(s0 + s) * Tcap + t) * 3 + pass] = r;
                    }
                    s_np[s] = pass + 1;
                    if (ctx_changed) s_newctx[s] = 1;
                    if (is_blank) s_done[s] = 1;
                }
                gp_consumer_sync();
                bool all_done = true;
                for (int s = 0; s < ns; s++) all_done &= (s_done[s] != 0);
                if (all_done) break;
            }
            if (rank == 0 && threadIdx.x < ns && t < s_n[threadIdx.x]) npass[(size_t)(s0 + threadIdx.x) * Tcap + t] = s_np[threadIdx.x];
        }
        gp_consumer_sync();
        if (threadIdx.x < ns) {
            const int s = threadIdx.x;
            if (last_chunk && flush_flag[s0 + s]) {                           // _aas_flush tail, april_session.c:561-563
                st[s].head = 0;
                if (st[s].ctx0 != P.blank) { st[s].ctx0 = P.blank; st[s].ctx1 = P.blank; }
            }
            if (rank == 0) gstate[slot[s0 + s]] = st[s];
        }
        tc::mbar_wait(full_bar((int)(nstage % GP_STAGES)), (nstage / GP_STAGES) & 1);   // the one stage issued ahead has landed: nothing in flight at exit
        // nobody exits while the peer can still store into its inbox or signal its barriers
        if (threadIdx.x < 2) gp_arrive_cluster(gp_mapa(exit_bar, threadIdx.x));
        gp_wait_cluster(exit_bar, 0);
    }
}

}  // namespace k
}  // namespace april
