// Engine implementation: device memory layout, pass planning (the integer logic of fbank.c's ring:
// fbank.c:174-226,308-349) and kernel orchestration. See engine.h and DESIGN.md.
#include "engine.h"
#include "plan.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_scan4.cuh"
#include "kernels_scan3.cuh"
#include "kernels_scan5.cuh"
#include "kernels_greedy_pair.cuh"

namespace april {

// CUDA errors abort the process (the reference's ORT_ABORT_ON_ERROR policy, src/ort_util.h:29-38) — except while a small pass is being
// captured into a CUDA graph: there a failure only cancels the capture and the pass runs eagerly.
static thread_local bool tl_soft_errors = false, tl_soft_failed = false;
static std::atomic<uint64_t> g_buf_gen{ 0 };       // bumped whenever a device / pinned work buffer is reallocated: cached graphs hold their addresses

#define CK(x)                                                                                              \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) {                                                                           \
            if (tl_soft_errors) { tl_soft_failed = true; break; }                                          \
            fprintf(stderr, "libapril: (engine) [ERROR] CUDA: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            abort();                                                                                       \
        }                                                                                                  \
    } while (0)

namespace {

// ---- TMA descriptors (driver entry point fetched through the runtime; no -lcuda at link time) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "libapril: (engine) [ERROR] cuTensorMapEncodeTiled unavailable\n"); abort(); }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// [rows, cols] fp16 row-major (cols contiguous = K); box = 64 x box_rows, 128-byte swizzle
static CUtensorMap make_map_f16(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = { cols, rows };
    cuuint64_t strides[1] = { cols * sizeof(__half) };
    cuuint32_t box[2] = { 64, box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "libapril: (engine) [ERROR] cuTensorMapEncodeTiled failed: %d (rows %llu cols %llu)\n", (int)r, (unsigned long long)rows, (unsigned long long)cols); abort(); }
    return m;
}

// generic 2-D map: [rows, cols] row-major of `esize`-byte elements, box = box_cols x box_rows
static CUtensorMap make_map_2d(const void* ptr, CUtensorMapDataType dt, size_t esize, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows, bool swizzle128) {
    CUtensorMap m;
    cuuint64_t dims[2] = { cols, rows };
    cuuint64_t strides[1] = { cols * esize };
    cuuint32_t box[2] = { box_cols, box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = get_encode_tiled()(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "libapril: (engine) [ERROR] cuTensorMapEncodeTiled failed: %d\n", (int)r); abort(); }
    return m;
}

// 3-D view of the gate pre-activations G [rows, 4H] fp32: {unit, gate, row}; one box = 8 units x 4 gates x box_rows rows
static CUtensorMap make_map_g3(const void* ptr, uint64_t rows, uint64_t H, uint32_t box_units, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[3] = { H, 4, rows };
    cuuint64_t strides[2] = { H * sizeof(float), 4 * H * sizeof(float) };
    cuuint32_t box[3] = { box_units, 4, box_rows };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "libapril: (engine) [ERROR] cuTensorMapEncodeTiled (G 3-D) failed: %d\n", (int)r); abort(); }
    return m;
}

// 5-D view of the two m planes [2][(T + 1) * Bs][H] fp16: {64-wide k-block element, session, hi | lo plane, k-block, time}
static CUtensorMap make_map_mx(const void* ptr, uint64_t Bs, uint64_t H, uint64_t Trows, uint64_t plane_elems, uint32_t box_sessions, uint32_t box_kb) {
    CUtensorMap m;
    cuuint64_t dims[5] = { 64, Bs, 2, H / 64, Trows };
    cuuint64_t strides[4] = { H * sizeof(__half), plane_elems * sizeof(__half), 64 * sizeof(__half), Bs * H * sizeof(__half) };
    cuuint32_t box[5] = { 64, box_sessions, 2, box_kb, 1 };
    cuuint32_t estr[5] = { 1, 1, 1, 1, 1 };
    CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "libapril: (engine) [ERROR] cuTensorMapEncodeTiled (5-D) failed: %d\n", (int)r); abort(); }
    return m;
}

struct TcWeight {                   // fp16 (hi, lo) copies of a [N, K] weight, pre-scaled by a power of two
    __half *hi = nullptr, *lo = nullptr;
    float inv_scale = 1.f;
    int N = 0, K = 0;
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CK(cudaFree(p));
        size_t want = bytes + bytes / 8 + 256;
        CK(cudaMalloc(&p, want));
        cap = want;
        g_buf_gen++;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
    ~DevBuf() { if (p) cudaFree(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CK(cudaFreeHost(p));
        size_t want = bytes + bytes / 4 + 4096;
        CK(cudaMallocHost(&p, want));
        cap = want;
        g_buf_gen++;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
    ~PinBuf() { if (p) cudaFreeHost(p); }
};

constexpr int CARRY_CAP = 16;       // feature rows a drained stream can still hold (< SEG after any pass)
constexpr int FLUSH_ZEROS = 6400;   // 2 x SEGSIZE zero samples, april_session.c:500,555-556

struct PassMeta {                   // per-session integers of one pass, uploaded as one block
    int slot, carry, n_new, pad_before, n_frames, nsteps, flush_flag, keep_first, n_keep, pad0;
    long long pcm_off;
};

struct DeviceLayer {
    float *w_ih, *b_gate, *w_hh, *w_hr, *ff1_w, *ff1_b, *ff2_w, *ff2_b;
    float norm_eps;
};

class EngineImpl : public Engine {
public:
    ModelDims d;
    Params prm;
    int dev = 0;
    int num_sms = 148;
    int sm_budget = 148;              // SMs the persistent GEMM grids may count on: num_sms minus those lent to the search kernel running beside the encoder
    cudaStream_t st2 = nullptr;       // the search of time chunk c runs here, under the encoder of chunk c + 1 (run_pass)
    bool overlap_search = true;       // APRIL_B200_OVERLAP=0 switches it off, =N (>= 2) sets the number of time chunks
    int overlap_chunks = 3;           // measured at 64 x 30 s (profiles/r02k): 105.5 ms without, 96.1 / 94.6 / 95.6 / 97.2 ms with 2 / 3 / 4 / 5 halving chunks
    cudaStream_t st = nullptr;
    std::mutex mu;
    EngineStats stat;
    TapFn tap = nullptr;
    void* tap_ud = nullptr;
    bool profile = false;
    // per-kernel-class timing with CUDA events on the launching stream (no host syncs inside the pass)
    struct Span { cudaEvent_t a, b; int cls; };
    std::vector<cudaEvent_t> ev_pool;
    std::vector<Span> spans;
    cudaEvent_t ev_get() { if (ev_pool.empty()) { cudaEvent_t e; CK(cudaEventCreate(&e)); return e; } cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    void span_begin(int cls, cudaStream_t on = nullptr) { if (!profile) return; Span sp{ ev_get(), ev_get(), cls }; CK(cudaEventRecord(sp.a, on ? on : st)); spans.push_back(sp); }
    void span_end(cudaStream_t on = nullptr) { if (!profile) return; CK(cudaEventRecord(spans.back().b, on ? on : st)); }
    void spans_collect() {
        for (auto& sp : spans) { float ms = 0; CK(cudaEventElapsedTime(&ms, sp.a, sp.b)); if (sp.cls == 0) stat.gemm_ms += ms; else if (sp.cls == 1) stat.lstm_scan_ms += ms; else if (sp.cls == 2) stat.greedy_ms += ms; else if (sp.cls == 3) { stat.gemm_ms += ms; stat.lstm_gemm_ms += ms; } else stat.frontend_ms += ms; ev_pool.push_back(sp.a); ev_pool.push_back(sp.b); }
        spans.clear();
    }

    // weights
    std::vector<void*> owned;
    float *conv1_w, *conv1_b, *conv2_w, *conv2_b, *conv3_w, *conv3_b, *out_w, *out_b, *enc_proj_w, *enc_proj_b;
    float out_norm_eps = 0.f;
    std::vector<DeviceLayer> layers;
    int npass = 3;                  // 0: exact fp32 SIMT tier; 1: fp16 tensor cores; 3: fp16x3 split (default)
    TcWeight tw_out, tw_proj, tw_conv3, tw_out_perm;
    tc::Conv12Params c12{};
    bool conv_tc = false;
    DevBuf b_a3_h, b_a3_l;
    struct TcLayer { TcWeight w_ih, ff1, ff2; };
    std::vector<TcLayer> tlayers;
    bool gemm_attr[16] = {};
    // wide-batch scan (kernels_scan4.cuh): sessions on the UMMA M dimension, K split over clusters of scan4_cs CTAs
    struct Scan4Layer { __half* w = nullptr; float inv = 1.f; CUtensorMap mapW; TcWeight w_hr; __half* w3 = nullptr; CUtensorMap mapW3; __half* w5 = nullptr; CUtensorMap mapW5; };   // w3 / w5: the same W_eff in the small-batch / pair kernel's layout
    std::vector<Scan4Layer> s4layers;
    bool scan4 = false;
    int scan4_cs = 2, scan4_mmaw = 2;
    size_t scan4_smem = 0;
    bool scan3 = false;             // small-batch kernel (kernels_scan3.cuh) available: used for passes of <= 64 sessions
    bool pair_greedy = false;       // kernels_greedy_pair.cuh available
    int pair_greedy_min = 257;      // passes of at least this many sessions use it (measured, profiles/r02j: 160 sessions 31.3 vs 26.9 ms for the 8-CTA clusters,
                                    // 512 sessions 40.6 vs 74.9 ms); APRIL_B200_GREEDY=pair|cluster forces one of them
    int gp_Rh = 0;
    size_t gp_smem = 0;
    bool scan5 = false;             // pair kernel (kernels_scan5.cuh) available: used for passes of >= scan5_min sessions
    int scan5_min = 257;
    size_t scan5_smem = 0;
    size_t scan3_smem = 0;
    size_t mall_rows = 0;           // rows per m plane of the current pass chunking: (Tc + 1) * B
    float* m_store = nullptr;       // [slot][L][H]
    DevBuf b_mall, b_mw;            // b_mall: the two m planes [2][(Tc + 1) * B][H] fp16
    unsigned* scan_counters = nullptr;
    DevBuf b_trace;
    DevBuf b_xe_h, b_xe_l, b_x_h, b_x_l, b_x1_h, b_x1_l, b_f1_h, b_f1_l;
    k::JoinerParams jp{};
    k::FbankParams fp{};
    k::ConvParams cp{};
    float* dout_blank = nullptr;
    float* dtab = nullptr;            // decoder output table [V][V][J] (fast greedy path)
    bool fast_greedy = false;
    int gr_R = 0;
    size_t gr_smem = 0;
    int frame_N = 0, frame_shift = 0;

    // CUDA graphs of small passes (streaming: every tick enqueues the same ~110 launches with the same geometry; launching them one by
    // one is CPU-bound at ~3 ms per pass). Key = everything the enqueue depends on besides device-resident per-session integers.
    struct GraphKey {
        int B = -1, Tmax = 0, max_new = 0, row_pitch = 0, any_flush = 0; size_t pcm_total = 0; uint64_t gen = 0;
        bool operator==(const GraphKey& o) const { return B == o.B && Tmax == o.Tmax && max_new == o.max_new && row_pitch == o.row_pitch && any_flush == o.any_flush && pcm_total == o.pcm_total && gen == o.gen; }
    };
    struct CachedGraph { GraphKey key; cudaGraphExec_t exec = nullptr; uint64_t launches = 0, h2d = 0, d2h = 0; size_t rec_bytes = 0, np_bytes = 0; };
    std::vector<CachedGraph> graphs;      // a few most recent
    GraphKey pending_key;                 // geometry of the previous eager small pass: captured when it repeats
    bool graphs_ok = true, capturing = false;

    // per-slot state
    int slot_cap = 0;
    std::vector<int> free_slots;
    float *h_store = nullptr, *c_store = nullptr, *dout_store = nullptr, *carry_store = nullptr;
    k::GreedyState* gstate = nullptr;

    // pass buffers
    DevBuf b_splitk;                  // partial tiles of split-K GEMMs
    bool splitk_ok = true;            // APRIL_B200_SPLITK=0: dev switch
    DevBuf b_pcm, b_meta, b_fb, b_xe, b_x, b_x1, b_x2, b_g, b_hout, b_f1, b_m, b_hw, b_cw, b_eo, b_rec, b_np, b_tap;
    DevBuf b_nsteps, b_slot, b_carry, b_nnew, b_padb, b_nfr, b_flush, b_keepf, b_nkeep, b_pcmoff;
    PinBuf p_pcm, p_meta, p_rec, p_np, p_eo, p_tap;

    std::vector<void*> load_scratch;      // device staging of make_tc_weight, freed once the stream has drained
    void release_load_scratch() { CK(cudaStreamSynchronize(st)); for (void* p : load_scratch) cudaFree(p); load_scratch.clear(); }

    template <class T> T* upload(const std::vector<T>& v) {
        T* p = nullptr;
        CK(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
        // on the engine's stream: `st` is non-blocking, so a copy on the legacy stream (whose DMA may still be in flight when
        // cudaMemcpy returns from a pageable source) would not be ordered before the kernels that read the weights
        if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
        owned.push_back(p);
        return p;
    }

    bool init(const LoadedModel& m, int device, std::string& err) {
        dev = device;
        d = m.w.d;
        prm = m.params;
        if (cudaSetDevice(dev) != cudaSuccess) { err = "cudaSetDevice failed"; return false; }
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10) { err = "this build targets sm_100a (B200); device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor); return false; }
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        num_sms = prop.multiProcessorCount; sm_budget = num_sms;
        { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CK(cudaStreamCreateWithPriority(&st2, cudaStreamNonBlocking, hi)); }   // its few CTAs are placed before the encoder's next kernel fills the GPU
        if (const char* o = getenv("APRIL_B200_SPLITK")) splitk_ok = atoi(o) != 0;
        if (const char* o = getenv("APRIL_B200_OVERLAP")) { overlap_search = atoi(o) > 0; if (atoi(o) > 1) overlap_chunks = std::min(atoi(o), 8); }
        profile = getenv("APRIL_B200_TIMING") != nullptr;
        if (d.D % 16 || d.H % 16 || d.F % 16 || d.J % 16 || (d.C3 * d.F3) % 16) { err = "model dims must be multiples of 16"; return false; }

        const ModelWeights& w = m.w;
        conv1_w = upload(w.conv1_w); conv1_b = upload(w.conv1_b);
        conv2_w = upload(w.conv2_w); conv2_b = upload(w.conv2_b);
        conv3_w = upload(w.conv3_w); conv3_b = upload(w.conv3_b);
        out_w = upload(w.out_w); out_b = upload(w.out_b); out_norm_eps = w.out_norm_eps;
        for (auto& lw : w.layers) {
            DeviceLayer dl{};
            std::vector<float> bsum(lw.b_ih.size());
            for (size_t i = 0; i < bsum.size(); i++) bsum[i] = lw.b_ih[i] + lw.b_hh[i];
            dl.w_ih = upload(lw.w_ih); dl.b_gate = upload(bsum); dl.w_hh = upload(lw.w_hh); dl.w_hr = upload(lw.w_hr);
            dl.ff1_w = upload(lw.ff1_w); dl.ff1_b = upload(lw.ff1_b); dl.ff2_w = upload(lw.ff2_w); dl.ff2_b = upload(lw.ff2_b);
            dl.norm_eps = lw.norm_eps;
            layers.push_back(dl);
        }
        enc_proj_w = upload(w.enc_proj_w); enc_proj_b = upload(w.enc_proj_b);
        if (const char* pm = getenv("APRIL_B200_PRECISION")) {
            if (!strcmp(pm, "fp32")) npass = 0; else if (!strcmp(pm, "fp16")) npass = 1; else if (!strcmp(pm, "fp16x3")) npass = 3;
            else { err = "APRIL_B200_PRECISION must be fp32, fp16 or fp16x3"; return false; }
        }
        if (npass) {
            if (d.D % 8 || d.H % 8 || d.F % 8 || d.J % 8) { err = "tensor-core tier needs dims that are multiples of 8"; return false; }
            tw_out = make_tc_weight(w.out_w, d.D, d.C3 * d.F3);
            tw_proj = make_tc_weight(w.enc_proj_w, d.J, d.D);
            // conv3 and the subsampler's linear as GEMMs (kernels_tc.cuh: conv12_im2col_kernel)
            conv_tc = (d.C2 % 8 == 0) && ((d.C2 * 9) % 8 == 0) && (d.C3 % 8 == 0) && ((d.SEG - 2 - 3) / 2 + 1 == 3);
            if (conv_tc) {
                tw_conv3 = make_tc_weight(w.conv3_w, d.C3, d.C2 * 9);                    // [oc][ic*9 + ky*3 + kx]
                std::vector<float> perm((size_t)d.D * d.C3 * d.F3);
                for (int n = 0; n < d.D; n++) for (int c = 0; c < d.C3; c++) for (int f = 0; f < d.F3; f++)
                    perm[(size_t)n * d.C3 * d.F3 + (size_t)f * d.C3 + c] = w.out_w[(size_t)n * d.C3 * d.F3 + (size_t)c * d.F3 + f];
                tw_out_perm = make_tc_weight(perm, d.D, d.C3 * d.F3);
                std::vector<float> w2t((size_t)d.C1 * 9 * d.C2);
                for (int oc = 0; oc < d.C2; oc++) for (int ic = 0; ic < d.C1; ic++) for (int k9 = 0; k9 < 9; k9++)
                    w2t[((size_t)ic * 9 + k9) * d.C2 + oc] = w.conv2_w[((size_t)oc * d.C1 + ic) * 9 + k9];
                c12.w2t = upload(w2t);
            }
            for (auto& lw : w.layers) {
                TcLayer tl;
                tl.w_ih = make_tc_weight(lw.w_ih, 4 * d.H, d.D);
                tl.ff1 = make_tc_weight(lw.ff1_w, d.F, d.D);
                tl.ff2 = make_tc_weight(lw.ff2_w, d.D, d.F);
                tlayers.push_back(tl);
            }
            const char* scan_sel = getenv("APRIL_B200_SCAN");
            if (!(scan_sel && !strcmp(scan_sel, "simt"))) init_scan4(w);   // "simt": per-step fp32 recurrence inside the tensor-core tier (precision studies, tests)
        }

        // vocabulary flags (april_session.c:338-351)
        std::vector<uint8_t> vf((size_t)d.V, 0);
        for (int i = 0; i < d.V; i++) {
            const std::string& t = prm.tokens[(size_t)i];
            uint8_t f = 0;
            if (!t.empty() && t[0] == ' ') f |= k::VF_SPACE;
            if (t.size() == 1 && (t[0] == '.' || t[0] == '!' || t[0] == '?')) f |= k::VF_EOS;
            if (t.size() == 1 && t[0] == ',') f |= k::VF_COMMA;
            if (!t.empty() && t[0] >= '0' && t[0] <= '9') f |= k::VF_DIGIT;
            if (t.size() == 1 && t[0] == '.') f |= k::VF_DOT;
            vf[(size_t)i] = f;
        }
        jp.J = d.J; jp.E = d.E; jp.V = d.V; jp.blank = prm.blank_id;
        jp.emb = upload(w.emb); jp.dec_conv = upload(w.dec_conv_w); jp.dec_w = upload(w.dec_proj_w); jp.dec_b = upload(w.dec_proj_b);
        jp.join_w = upload(w.join_w); jp.join_b = upload(w.join_b); jp.vflags = upload(vf);

        // front-end tables: generate_povey_window fbank.c:49-55, generate_banks fbank.c:65-95
        int window_size = prm.frame_length_ms * prm.sample_rate / 1000;
        frame_shift = prm.frame_shift_ms * prm.sample_rate / 1000;
        int N = window_size;
        if (prm.round_pow2) { N -= 1; N |= N >> 1; N |= N >> 2; N |= N >> 4; N |= N >> 8; N |= N >> 16; N += 1; }
        if ((N & (N - 1)) != 0 || N < 64 || N > 1024) { err = "front-end: padded window must be a power of two in [64, 1024]"; return false; }
        frame_N = N;
        std::vector<float> win((size_t)N);
        for (int i = 0; i < N; i++) win[(size_t)i] = (float)pow(0.5 - 0.5 * cos((double)i / (double)N * 6.283185307), 0.85);
        std::vector<double2> tw((size_t)N / 2);
        for (int i = 0; i < N / 2; i++) { double a = 2.0 * M_PI * (double)i / (double)N; tw[(size_t)i] = make_double2(cos(a), -sin(a)); }
        int nb = N / 2, mel = prm.mel_features;
        std::vector<float> bank((size_t)mel * nb, 0.f);
        {
            int high = prm.mel_high == 0 ? prm.sample_rate / 2 : prm.mel_high;
            auto mel_scale = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
            volatile float bw = (float)prm.sample_rate / (float)N;
            volatile float mel_low = (float)mel_scale((double)prm.mel_low), mel_high = (float)mel_scale((double)high);
            volatile float delta = (mel_high - mel_low) / ((float)mel + 1.0f);
            for (int i = 0; i < mel; i++) {
                volatile float left = mel_low + (float)i * delta;
                volatile float center = left + delta;
                volatile float right = center + delta;
                for (int j = 0; j < nb; j++) {
                    volatile float freq = bw * (float)j;
                    volatile float mf = (float)mel_scale((double)freq);
                    float wgt = 0.f;
                    if (mf > left && mf < right) {
                        if (mf <= center) { volatile float a = mf - left, c = center - left; wgt = a / c; }
                        else { volatile float a = right - mf, c = right - center; wgt = a / c; }
                    }
                    bank[(size_t)i * nb + j] = wgt;
                }
            }
        }
        std::vector<int> ms((size_t)mel, 0), mc((size_t)mel, 0), mo((size_t)mel, 0);
        std::vector<float> mw;
        for (int i = 0; i < mel; i++) {
            int lo = -1, hi = -1;
            for (int j = 0; j < nb; j++) if (bank[(size_t)i * nb + j] != 0.f) { if (lo < 0) lo = j; hi = j; }
            mo[(size_t)i] = (int)mw.size();
            if (lo >= 0) { ms[(size_t)i] = lo; mc[(size_t)i] = hi - lo + 1; for (int j = lo; j <= hi; j++) mw.push_back(bank[(size_t)i * nb + j]); }
        }
        fp.N = N; fp.log2N = 0; while ((1 << fp.log2N) < N) fp.log2N++;
        fp.shift = frame_shift; fp.mel = mel; fp.nbins = nb;
        fp.pad_value = (float)log((double)1.1920928955078125e-07f);
        fp.window = upload(win); fp.twiddle = upload(tw); fp.mel_start = upload(ms); fp.mel_count = upload(mc); fp.mel_off = upload(mo); fp.mel_w = upload(mw);
        CK(cudaFuncSetAttribute(k::fbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * N * (int)sizeof(double)));

        cp.SEG = d.SEG; cp.MEL = d.MEL; cp.STEP = d.STEP; cp.C1 = d.C1; cp.C2 = d.C2; cp.C3 = d.C3;
        cp.T1 = d.SEG - 2; cp.F1 = d.MEL - 2; cp.T2 = (cp.T1 - 3) / 2 + 1; cp.F2 = (cp.F1 - 3) / 2 + 1; cp.F3 = d.F3;
        cp.w1 = conv1_w; cp.b1 = conv1_b; cp.w2 = conv2_w; cp.b2 = conv2_b; cp.w3 = conv3_w; cp.b3 = conv3_b;
        CK(cudaFuncSetAttribute(k::conv_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem()));
        c12.SEG = cp.SEG; c12.MEL = cp.MEL; c12.STEP = cp.STEP; c12.C1 = cp.C1; c12.C2 = cp.C2; c12.T1 = cp.T1; c12.F1 = cp.F1; c12.T2 = cp.T2; c12.F2 = cp.F2; c12.F3 = cp.F3;
        c12.w1 = conv1_w; c12.b1 = conv1_b; c12.b2 = conv2_b;
        CK(cudaFuncSetAttribute(tc::conv12_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem()));

        CK(cudaMalloc(&dout_blank, sizeof(float) * (size_t)d.J));
        k::decoder_kernel<<<1, 256, (size_t)(d.E + d.J) * sizeof(float), st>>>(jp, prm.blank_id, prm.blank_id, dout_blank);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        init_fast_greedy();
        release_load_scratch();
        grow_slots(64);
        return true;
    }

    // cluster greedy kernel prerequisites: joiner rows fit 8 CTAs' shared memory, decoder table fits in HBM budget
    void init_fast_greedy() {
        const int V = d.V, J = d.J;
        gr_R = (V + k::GC_CLUSTER - 1) / k::GC_CLUSTER;
        gr_smem = sizeof(float) * ((size_t)4 * k::GC_MAX_S * J + (size_t)k::GC_MAX_S * ((gr_R + 3) & ~3)) + sizeof(k::GreedyPartial) * 2 * k::GC_CLUSTER * k::GC_MAX_S + (size_t)V + 64;
        const size_t tab_bytes = sizeof(float) * (size_t)V * V * J;
        if (J % 32 != 0 || J / 32 > k::GC_MAX_JPL || J > 512 || gr_R > 8 * k::GC_RPW || gr_smem > 200 * 1024 || tab_bytes > ((size_t)3 << 30)) return;
        CK(cudaMalloc(&dtab, tab_bytes));
        const int CH = std::max(1, std::min(V, 32768 / std::max(V, 1)));   // contexts per launch: ~32 K table rows at a time
        float* Y = nullptr;
        CK(cudaMalloc(&Y, sizeof(float) * (size_t)CH * V * d.E));
        for (int c0 = 0; c0 < V; c0 += CH) {
            const int n = std::min(CH, V - c0);
            k::decoder_table_y_kernel<<<dim3((unsigned)V, (unsigned)n), 128, 0, st>>>(jp, c0, Y);
            gemm_launch(k::EPI_BIAS, Y, jp.dec_w, jp.dec_b, nullptr, dtab + (size_t)c0 * V * J, n * V, J, d.E);
        }
        CK(cudaStreamSynchronize(st));
        CK(cudaFree(Y));
        CK(cudaFuncSetAttribute(k::greedy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gr_smem));
        fast_greedy = true;
        // the CTA-pair search kernel (joiner rows streamed from L2): needs the same table and J <= 512
        gp_Rh = (V + 1) / 2;
        gp_smem = k::greedy_pair_smem(J, V, gp_Rh);
        pair_greedy = V >= 2 && gp_smem <= 227 * 1024 && cudaFuncSetAttribute(k::greedy_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gp_smem) == cudaSuccess;
        cudaGetLastError();
        if (const char* g = getenv("APRIL_B200_GREEDY")) { if (!strcmp(g, "cluster")) pair_greedy_min = 1 << 30; else if (!strcmp(g, "pair")) pair_greedy_min = 1; }
    }

    size_t conv_smem() const {
        return sizeof(float) * ((size_t)cp.SEG * cp.MEL + (size_t)cp.C1 * cp.T1 * cp.F1 + (size_t)cp.C2 * cp.T2 * cp.F2 + (size_t)cp.C2 * cp.C1 * 9);
    }

    ~EngineImpl() override {
        cudaSetDevice(dev);
        if (st) cudaStreamSynchronize(st);
        for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
        for (void* p : owned) cudaFree(p);
        cudaFree(dout_blank); cudaFree(dtab); cudaFree(m_store); cudaFree(h_store); cudaFree(c_store); cudaFree(dout_store); cudaFree(carry_store); cudaFree(gstate);
        if (st2) cudaStreamDestroy(st2);
        if (st) cudaStreamDestroy(st);
    }

    template <class T> void regrow(T*& p, size_t old_n, size_t new_n) {
        T* q = nullptr;
        CK(cudaMalloc(&q, new_n * sizeof(T)));
        CK(cudaMemsetAsync(q, 0, new_n * sizeof(T), st));
        if (p && old_n) CK(cudaMemcpyAsync(q, p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        if (p) CK(cudaFree(p));
        p = q;
    }

    void grow_slots(int new_cap) {
        size_t o = (size_t)slot_cap, n = (size_t)new_cap;
        regrow(h_store, o * d.L * d.D, n * d.L * d.D);
        regrow(m_store, o * d.L * d.H, n * d.L * d.H);
        regrow(c_store, o * d.L * d.H, n * d.L * d.H);
        regrow(dout_store, o * d.J, n * d.J);
        regrow(carry_store, o * CARRY_CAP * d.MEL, n * CARRY_CAP * d.MEL);
        regrow(gstate, o, n);
        for (int i = new_cap - 1; i >= slot_cap; i--) free_slots.push_back(i);
        slot_cap = new_cap;
    }

    int alloc_slot() override {
        std::lock_guard<std::mutex> lk(mu);
        CK(cudaSetDevice(dev));
        if (free_slots.empty()) grow_slots(slot_cap * 2);
        int s = free_slots.back();
        free_slots.pop_back();
        // zero h/c (calloc'ed tensors, ort_util.h:48-50), context = blanks + decoder output (april_session.c:432-438)
        CK(cudaMemsetAsync(h_store + (size_t)s * d.L * d.D, 0, sizeof(float) * (size_t)d.L * d.D, st));
        CK(cudaMemsetAsync(c_store + (size_t)s * d.L * d.H, 0, sizeof(float) * (size_t)d.L * d.H, st));
        CK(cudaMemsetAsync(m_store + (size_t)s * d.L * d.H, 0, sizeof(float) * (size_t)d.L * d.H, st));
        k::GreedyState g{};
        g.ctx0 = g.ctx1 = prm.blank_id;
        CK(cudaMemcpyAsync(gstate + s, &g, sizeof g, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dout_store + (size_t)s * d.J, dout_blank, sizeof(float) * (size_t)d.J, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        return s;
    }

    void free_slot(int slot) override {
        std::lock_guard<std::mutex> lk(mu);
        if (slot >= 0) free_slots.push_back(slot);
    }

    EngineStats stats(bool reset) override {
        std::lock_guard<std::mutex> lk(mu);
        EngineStats s = stat;
        if (reset) stat = EngineStats();
        return s;
    }
    void set_tap(TapFn fn, void* ud) override { std::lock_guard<std::mutex> lk(mu); tap = fn; tap_ud = ud; }
    void set_timing(bool on) override { std::lock_guard<std::mutex> lk(mu); profile = on; }
    int stride_ms() const override { return prm.segment_step * prm.frame_shift_ms; }   // fbank.c:359-361
    int device() const override { return dev; }

    TcWeight make_tc_weight(const std::vector<float>& w, int N, int K) {
        TcWeight t; t.N = N; t.K = K;
        float mx = 0.f;
        for (float v : w) if (std::isfinite(v)) mx = std::max(mx, std::fabs(v));   // a damaged file may hold inf/nan: they stay what they are
        int e = mx > 0.f ? (int)std::floor(std::log2(1024.0 / (double)mx)) : 0;      // max |w| * 2^e in [512, 1024]
        e = std::max(-24, std::min(24, e));
        const float sc = std::ldexp(1.0f, e);
        t.inv_scale = std::ldexp(1.0f, -e);
        // the split runs on the device (57 M weights: ~1.5 s of host conversions otherwise); the fp32 staging copy is freed by release_load_scratch()
        const size_t n = w.size();
        float* tmp = nullptr;
        CK(cudaMalloc(&tmp, std::max<size_t>(n, 1) * sizeof(float))); load_scratch.push_back(tmp);
        CK(cudaMalloc(&t.hi, std::max<size_t>(n, 1) * sizeof(__half))); owned.push_back(t.hi);
        CK(cudaMalloc(&t.lo, std::max<size_t>(n, 1) * sizeof(__half))); owned.push_back(t.lo);
        if (n) {
            CK(cudaMemcpyAsync(tmp, w.data(), n * sizeof(float), cudaMemcpyHostToDevice, st));
            tc::split_scale_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, (size_t)148 * 16), 256, 0, st>>>(tmp, t.hi, t.lo, n, sc);
            CK(cudaGetLastError());
        }
        return t;
    }

    static float pow2_scale(const std::vector<float>& w, float& inv) {
        float mx = 0.f;
        for (float v : w) if (std::isfinite(v)) mx = std::max(mx, std::fabs(v));   // a damaged file may hold inf/nan: they stay what they are
        int e = mx > 0.f ? (int)std::floor(std::log2(1024.0 / (double)mx)) : 0;
        e = std::max(-24, std::min(24, e));
        inv = std::ldexp(1.0f, -e);
        return std::ldexp(1.0f, e);
    }

    void dump_trace(int grid, int T, int GC) {
        std::vector<unsigned long long> h((size_t)grid * T * 8);
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(h.data(), b_trace.p, h.size() * 8, cudaMemcpyDeviceToHost));
        auto at = [&](int cta, int t, int k) { return (double)h[((size_t)cta * T + t) * 8 + k]; };
        // averages over steps 8..T-8 for gate CTA 0 and projection CTA GC; times relative to the gate CTA's stamp 0 of the same step
        double acc_g[8] = {0}, acc_p[8] = {0}; int n = 0;
        double step = 0;
        for (int t = 8; t < T - 8; t++, n++) {
            double t0g = at(0, t, 0);
            for (int k2 = 0; k2 < 8; k2++) { acc_g[k2] += at(0, t, k2) - t0g; acc_p[k2] += at(GC, t, k2) - t0g; }
            step += at(0, t + 1, 0) - t0g;
        }
        fprintf(stderr, "scan trace (ns, rel. to gate stamp0, avg of %d steps; step = %.0f ns)\n  gate: ", n, step / n);
        for (int k2 = 0; k2 < 8; k2++) fprintf(stderr, "%d:%.0f ", k2, acc_g[k2] / n);
        fprintf(stderr, "\n  proj: ");
        for (int k2 = 0; k2 < 8; k2++) fprintf(stderr, "%d:%.0f ", k2, acc_p[k2] / n);
        // spread across CTAs of stamp 7 (release) at one step
        double mn = 1e30, mx = 0; for (int c = 0; c < GC; c++) { double v = at(c, T / 2, 7) - at(0, T / 2, 0); mn = std::min(mn, v); mx = std::max(mx, v); }
        double mn2 = 1e30, mx2 = 0; for (int c = GC; c < grid; c++) { double v = at(c, T / 2, 7) - at(0, T / 2, 0); mn2 = std::min(mn2, v); mx2 = std::max(mx2, v); }
        fprintf(stderr, "\n  gate release spread [%.0f, %.0f]  proj release spread [%.0f, %.0f]\n", mn, mx, mn2, mx2);
    }

    // ---- wide-batch scan -------------------------------------------------------------------------------------
    template <int CS, int MW> bool scan4_prepare_variant(int GC) {
        const size_t smem = scan4::Cfg<CS, MW>::smem_bytes(d.H);
        if (smem > 227 * 1024) return false;
        if (cudaFuncSetAttribute(scan4::lstm_scan4_kernel<CS, MW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)GC); cfg.blockDim = dim3(scan4::Cfg<CS, MW>::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        int ncl = 0;                                               // cluster dims are compiled in
        if (cudaOccupancyMaxActiveClusters(&ncl, scan4::lstm_scan4_kernel<CS, MW>, &cfg) != cudaSuccess || ncl * CS < GC) { cudaGetLastError(); return false; }   // every cluster must be co-resident
        scan4_smem = smem;
        return true;
    }
    template <int CS, int MW> void scan4_launch(const Scan4Layer& sl, const CUtensorMap& mh, const CUtensorMap& ml, const scan4::Args& a) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)a.GC); cfg.blockDim = dim3(scan4::Cfg<CS, MW>::THREADS); cfg.dynamicSmemBytes = scan4_smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = capturing ? 0 : 1;          // graph capture: plain launch (co-residency was checked at init: 1 CTA per SM, grid <= SMs)
        CK(cudaLaunchKernelEx(&cfg, scan4::lstm_scan4_kernel<CS, MW>, sl.mapW, mh, ml, a));
    }

    void init_scan4(const ModelWeights& w) {
        const int D = d.D, H = d.H;
        if (H % 128 || H / scan4::UNITS > num_sms || D % 8) return;     // K slices are whole 64-wide k-blocks
        const int GC = H / scan4::UNITS;
        scan4_cs = 2; scan4_mmaw = 2;
        const bool ok = scan4_prepare_variant<2, 2>(GC);
        if (!ok) return;
        {   // the small-batch kernel: all pairs co-resident, groups of <= 32 sessions
            scan3_smem = 1024 + (size_t)(H / 128) * 16384 + (size_t)scan3::RSTAGES * 16384 + scan3::TEAMS * (64 * 33 * 4 + 64 * 36 * 4) + 256 + (size_t)64 * scan3::UNITS * 16 + (size_t)64 * scan3::UNITS * 4 + scan3::MAXG * 64 * 4 + scan3::TEAMS * 2 * scan3::EPI_WARPS * 32 * 2 + 512;
            scan3 = scan3_smem <= 227 * 1024 && cudaFuncSetAttribute(scan3::lstm_scan3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan3_smem) == cudaSuccess;
            if (scan3) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3((unsigned)GC); cfg.blockDim = dim3(scan3::THREADS3); cfg.dynamicSmemBytes = scan3_smem; cfg.stream = st;
                int ncl = 0;
                if (cudaOccupancyMaxActiveClusters(&ncl, scan3::lstm_scan3_kernel, &cfg) != cudaSuccess || ncl * 2 < GC) scan3 = false;
            }
            cudaGetLastError();
        }
        {   // the pair kernel: clusters of 4, all co-resident
            const char* sel = getenv("APRIL_B200_SCAN");
            const bool off = sel && !strcmp(sel, "nopair");                 // dev: scan4 for every large batch
            scan5_min = (sel && !strcmp(sel, "pair")) ? 129 : 257;          // measured on B200 (profiles/r02e, r02g): a tie at 256 sessions, 18 % faster at 512; "pair" forces it from 129
            scan5_smem = scan5::smem_bytes(H);
            scan5 = !off && H % 128 == 0 && GC % scan5::CL == 0 && scan5_smem <= 227 * 1024 &&
                    cudaFuncSetAttribute(scan5::lstm_scan5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan5_smem) == cudaSuccess;
            if (scan5) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3((unsigned)GC); cfg.blockDim = dim3(scan5::THREADS); cfg.dynamicSmemBytes = scan5_smem; cfg.stream = st;
                int ncl = 0;
                if (cudaOccupancyMaxActiveClusters(&ncl, scan5::lstm_scan5_kernel, &cfg) != cudaSuccess || ncl * scan5::CL < GC) scan5 = false;
            }
            cudaGetLastError();
        }
        const int CS = scan4_cs, NC = 32 * CS, Kc = H / CS;
        float* dC = nullptr;
        CK(cudaMalloc(&dC, sizeof(float) * (size_t)4 * H * H));
        std::vector<float> weff((size_t)4 * H * H);
        for (size_t li = 0; li < w.layers.size(); li++) {
            const LayerWeights& lw = w.layers[li];
            Scan4Layer sl;
            // W_eff = W_hh W_hr from the device copies init() uploaded on this stream
            scan4::weff_kernel<<<dim3((unsigned)((H + 15) / 16), (unsigned)((4 * H + 15) / 16)), 256, 0, st>>>(layers[li].w_hh, layers[li].w_hr, dC, 4 * H, H, D);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(weff.data(), dC, sizeof(float) * weff.size(), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const float sc = pow2_scale(weff, sl.inv);
            CK(cudaMalloc(&sl.w, sizeof(__half) * (size_t)GC * 2 * NC * Kc)); owned.push_back(sl.w);
            scan4::pack_weff_kernel<<<dim3((unsigned)GC, (unsigned)NC), 128, 0, st>>>(dC, sl.w, H, CS, sc);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));
            sl.mapW = make_map_2d(sl.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)GC * 2 * NC, (uint64_t)Kc, 64, (uint32_t)NC, true);
            sl.w_hr = make_tc_weight(lw.w_hr, D, H);
            if (scan3) {
                CK(cudaMalloc(&sl.w3, sizeof(__half) * (size_t)GC * 128 * (H / 2))); owned.push_back(sl.w3);
                scan3::pack_weff3_kernel<<<dim3((unsigned)GC, 128), 128, 0, st>>>(dC, sl.w3, H, sc);
                CK(cudaGetLastError());
                CK(cudaStreamSynchronize(st));
                sl.mapW3 = make_map_2d(sl.w3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)GC * 128, (uint64_t)(H / 2), 64, 128, true);
            }
            if (scan5) {
                CK(cudaMalloc(&sl.w5, sizeof(__half) * (size_t)GC * 2 * scan5::NB * (H / 2))); owned.push_back(sl.w5);
                scan5::pack_weff5_kernel<<<dim3((unsigned)GC, scan5::NB), 128, 0, st>>>(dC, sl.w5, H, sc);
                CK(cudaGetLastError());
                CK(cudaStreamSynchronize(st));
                sl.mapW5 = make_map_2d(sl.w5, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)GC * 2 * scan5::NB, (uint64_t)(H / 2), 64, scan5::NB, true);
            }
            s4layers.push_back(sl);
        }
        CK(cudaFree(dC));
        if (!scan_counters) { CK(cudaMalloc(&scan_counters, 256 * sizeof(unsigned))); owned.push_back(scan_counters); }
        scan4 = true;
    }

    // small batches (<= 64 sessions): two interleaved groups of <= 32 sessions through the latency-optimised pair kernel
    void scan3_layer(const Scan4Layer& sl, const float* G, float* mw, float* cw, const int* d_nsteps, int t0, int tc, int B, __half* Mx, size_t plane, int l) {
        const int H = d.H;
        scan3::Scan2Args a{};
        a.T = tc; a.t0 = t0;
        a.NG = std::max((B + 31) / 32, std::min(2, (B + 15) / 16));        // groups of <= 32 sessions; two groups from 17 sessions on
        const int per = (B + a.NG - 1) / a.NG;
        for (int g = 0, o = 0; g < scan3::MAXG; g++) { a.goff[g] = o; a.Bg[g] = g < a.NG ? std::min(per, B - o) : 0; o += a.Bg[g]; }
        a.Bp = (per + 15) & ~15;
        a.Bs = B; a.g0 = 0; a.H = H; a.GC = H / scan3::UNITS;
        const int KBW = H / 128;                                            // k-blocks a CTA walks per step (its K half)
        a.KBG = 1; for (int kk = 1; kk <= KBW; kk++) if (KBW % kk == 0 && (size_t)kk * 2 * a.Bp * 128 <= 16384) a.KBG = kk;
        a.dbg = 4;
        a.inv_scale = sl.inv; a.nsteps = d_nsteps; a.c_state = cw; a.m_state = mw;
        a.Mx = Mx; a.plane = plane; a.counter = scan_counters;
        static const bool want_trace = getenv("APRIL_B200_SCAN_TRACE") != nullptr;
        const bool do_trace = want_trace && l == 0 && tc >= 64;
        if (do_trace) { b_trace.ensure(sizeof(unsigned long long) * (size_t)a.GC * tc * 8); CK(cudaMemsetAsync(b_trace.p, 0, sizeof(unsigned long long) * (size_t)a.GC * tc * 8, st)); a.trace = b_trace.as<unsigned long long>(); }
        CK(cudaMemsetAsync(scan_counters, 0, 256 * sizeof(unsigned), st));
        const CUtensorMap mMx = make_map_mx(Mx, (uint64_t)B, (uint64_t)H, (uint64_t)tc + 1, (uint64_t)plane, (uint32_t)a.Bp, (uint32_t)a.KBG);
        const CUtensorMap mG = make_map_g3(G, (uint64_t)tc * B, (uint64_t)H, scan3::UNITS, (uint32_t)a.Bp);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)a.GC); cfg.blockDim = dim3(scan3::THREADS3); cfg.dynamicSmemBytes = scan3_smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = capturing ? 0 : 1;                   // cluster dims are compiled in
        CK(cudaLaunchKernelEx(&cfg, scan3::lstm_scan3_kernel, sl.mapW3, mMx, mG, a));
        launch_check();
        if (do_trace) dump_trace(a.GC, tc, a.GC / 2);
    }

    // large batches (> 128 sessions): items of <= 256 sessions through the pair kernel, two items per launch
    void scan5_layer(const Scan4Layer& sl, const float* G, float* mw, float* cw, const int* d_nsteps, int t0, int tc, int B, __half* Mh, __half* Ml, int l) {
        const int H = d.H;
        const int per_launch = scan5::MAXI * scan5::ITEM;
        for (int g0 = 0; g0 < B; g0 += per_launch) {
            const int Bc = std::min(per_launch, B - g0);
            scan4::Args a{};
            a.T = tc; a.t0 = t0;
            a.NG = (Bc + scan5::ITEM - 1) / scan5::ITEM;
            const int per = ((Bc + a.NG - 1) / a.NG + 7) & ~7;
            for (int g = 0, o = 0; g < scan4::MAXG; g++) { a.goff[g] = o; a.Bg[g] = g < a.NG ? std::max(0, std::min(per, Bc - o)) : 0; o += a.Bg[g]; }
            a.Bp = scan5::HALF;
            a.Bs = B; a.g0 = g0; a.H = H; a.GC = H / scan4::UNITS;
            a.inv_scale = sl.inv; a.nsteps = d_nsteps; a.G = G; a.c_state = cw; a.m_state = mw;
            a.Mx_hi = Mh; a.Mx_lo = Ml; a.counter = scan_counters;
            static const bool want_trace = getenv("APRIL_B200_SCAN_TRACE") != nullptr;
            const bool do_trace = want_trace && l == 0 && g0 == 0 && tc >= 64;
            if (do_trace) { b_trace.ensure(sizeof(unsigned long long) * (size_t)a.GC * tc * 8); CK(cudaMemsetAsync(b_trace.p, 0, sizeof(unsigned long long) * (size_t)a.GC * tc * 8, st)); a.trace = b_trace.as<unsigned long long>(); }
            CK(cudaMemsetAsync(scan_counters, 0, 256 * sizeof(unsigned), st));
            const CUtensorMap mh = make_map_f16(Mh, (uint64_t)(tc + 1) * B, (uint64_t)H, scan5::HALF), ml = make_map_f16(Ml, (uint64_t)(tc + 1) * B, (uint64_t)H, scan5::HALF);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)a.GC); cfg.blockDim = dim3(scan5::THREADS); cfg.dynamicSmemBytes = scan5_smem; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
            cfg.attrs = at; cfg.numAttrs = capturing ? 0 : 1;      // cluster dims are compiled in
            CK(cudaLaunchKernelEx(&cfg, scan5::lstm_scan5_kernel, sl.mapW5, mh, ml, a));
            launch_check();
            if (do_trace) dump_trace(a.GC, tc, 2);
        }
    }

    // one layer, steps [t0, t0+tc) of all B sessions of the pass: up to 512 sessions (4 groups of <= 128) per launch
    void scan4_layer(int l, const float* G, float* mw, float* cw, const int* d_nsteps, int t0, int tc, int B) {
        const int H = d.H;
        const Scan4Layer& sl = s4layers[(size_t)l];
        const size_t plane = ((size_t)mall_rows) * H;                      // elements per m plane (sized by run_pass)
        __half *Mh = b_mall.as<__half>(), *Ml = Mh + plane;
        const size_t n0 = (size_t)B * H;
        scan4::scan4_prepare_kernel<<<(unsigned)std::min<size_t>((n0 / 4 + 255) / 256, (size_t)num_sms * 4), 256, 0, st>>>(mw, Mh, Ml, n0); launch_check();
        static const bool no_small = getenv("APRIL_B200_SCAN") && !strcmp(getenv("APRIL_B200_SCAN"), "nosmall");   // dev: scan4 for every size
        if (scan3 && B <= 64 && !no_small) { scan3_layer(sl, G, mw, cw, d_nsteps, t0, tc, B, Mh, plane, l); return; }
        if (scan5 && B >= scan5_min) { scan5_layer(sl, G, mw, cw, d_nsteps, t0, tc, B, Mh, Ml, l); return; }
        static const int groups_env = getenv("APRIL_B200_SCAN_GROUPS") ? atoi(getenv("APRIL_B200_SCAN_GROUPS")) : 0;
        const int per_launch = scan4::MAXG * scan4::GROUP;
        for (int g0 = 0; g0 < B; g0 += per_launch) {
            const int Bc = std::min(per_launch, B - g0);
            scan4::Args a{};
            a.T = tc; a.t0 = t0;
            // two epilogue teams (groups k, k + 2): an even number of groups keeps both busy (measured: 128 sessions as 2 x 64 take 10.8 us
            // per layer-step, as 1 x 128 11.8 us)
            a.NG = Bc > 2 * scan4::GROUP ? 4 : (Bc >= 16 ? 2 : 1);
            if (groups_env >= (Bc + scan4::GROUP - 1) / scan4::GROUP && groups_env <= scan4::MAXG && Bc >= 8 * groups_env) a.NG = groups_env;
            const int per = ((Bc + a.NG - 1) / a.NG + 7) & ~7;
            for (int g = 0, o = 0; g < scan4::MAXG; g++) { a.goff[g] = o; a.Bg[g] = g < a.NG ? std::max(0, std::min(per, Bc - o)) : 0; o += a.Bg[g]; }
            a.Bp = per;
            a.Bs = B; a.g0 = g0; a.H = H; a.GC = H / scan4::UNITS;
            a.inv_scale = sl.inv; a.nsteps = d_nsteps; a.G = G; a.c_state = cw; a.m_state = mw;
            a.Mx_hi = Mh; a.Mx_lo = Ml; a.counter = scan_counters;
            static const bool want_trace = getenv("APRIL_B200_SCAN_TRACE") != nullptr;
            const bool do_trace = want_trace && l == 0 && g0 == 0 && tc >= 64;
            if (do_trace) { b_trace.ensure(sizeof(unsigned long long) * (size_t)a.GC * tc * 8); CK(cudaMemsetAsync(b_trace.p, 0, sizeof(unsigned long long) * (size_t)a.GC * tc * 8, st)); a.trace = b_trace.as<unsigned long long>(); }
            CK(cudaMemsetAsync(scan_counters, 0, 256 * sizeof(unsigned), st));
            const CUtensorMap mh = make_map_f16(Mh, (uint64_t)(tc + 1) * B, (uint64_t)H, (uint32_t)a.Bp), ml = make_map_f16(Ml, (uint64_t)(tc + 1) * B, (uint64_t)H, (uint32_t)a.Bp);
            scan4_launch<2, 2>(sl, mh, ml, a);
            launch_check();
            if (do_trace) dump_trace(a.GC, tc, 1);
        }
    }

    // the radix-8 front-end serves the 512-sample window (every model of the reference: 25 ms at 16 kHz); other windows take the generic kernel
    bool fbank512() const { static const bool generic = getenv("APRIL_B200_FBANK") && !strcmp(getenv("APRIL_B200_FBANK"), "generic"); return frame_N == k::FB3_N && !generic; }
    static constexpr size_t kFbank512Smem = 4 * 2 * k::FB3_PITCH * sizeof(double);

    template <int NP, int EPI, int CL> void gemm_tc_launch(unsigned grid, const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh, const CUtensorMap& mBl, const tc::GemmArgs& g) {
        bool& attr_done = gemm_attr[(NP == 3 ? 8 : 0) + EPI * 2 + (CL - 1)];   // per engine (= per device) and instantiation
        if (!attr_done) { CK(cudaFuncSetAttribute(tc::gemm_tc_kernel<NP, EPI, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GemmCfg<NP, CL>::SMEM_BYTES)); attr_done = true; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc::GEMM_THREADS); cfg.dynamicSmemBytes = tc::GemmCfg<NP, CL>::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, tc::gemm_tc_kernel<NP, EPI, CL>, mAh, mAl, mBh, mBl, g));
    }
    template <int NP, int CL> void gemm_tc_epi(int epi, unsigned grid, const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh, const CUtensorMap& mBl, const tc::GemmArgs& g) {
        if (epi == tc::GE_F32) gemm_tc_launch<NP, tc::GE_F32, CL>(grid, mAh, mAl, mBh, mBl, g);
        else if (epi == tc::GE_F32_RESIDUAL) gemm_tc_launch<NP, tc::GE_F32_RESIDUAL, CL>(grid, mAh, mAl, mBh, mBl, g);
        else if (epi == tc::GE_F32_RESIDUAL_SPLIT) gemm_tc_launch<NP, tc::GE_F32_RESIDUAL_SPLIT, CL>(grid, mAh, mAl, mBh, mBl, g);
        else gemm_tc_launch<NP, tc::GE_DSWISH_SPLIT, CL>(grid, mAh, mAl, mBh, mBl, g);
    }

    // C = A * W^T (+bias, epilogue) on the tensor cores. A given as fp16 (hi, lo) [M, K].
    void gemm_tc(int epi, const __half* a_hi, const __half* a_lo, const TcWeight& w, const float* bias, const float* residual,
                 float* out_f32, __half* out_hi, __half* out_lo, int M, int cls = 0) {
        span_begin(cls);
        if (profile) { stat.gemm_flops += 2.0 * (double)M * (double)w.N * (double)w.K; stat.gemm_launches++; if (cls == 3) stat.lstm_gemm_flops += 2.0 * (double)M * (double)w.N * (double)w.K; }
        const int tiles_m = (M + tc::GEMM_BM - 1) / tc::GEMM_BM, tiles_n = (w.N + tc::GEMM_BN - 1) / tc::GEMM_BN;
        const bool pair = tiles_m >= 2;          // CTA pairs (cta_group::2): each CTA ingests its A rows and half of the B rows
        CUtensorMap mAh = make_map_f16(a_hi, (uint64_t)M, (uint64_t)w.K, tc::GEMM_BM), mAl = make_map_f16(a_lo, (uint64_t)M, (uint64_t)w.K, tc::GEMM_BM);
        const uint32_t bbox = pair ? tc::GEMM_BN / 2 : tc::GEMM_BN;
        CUtensorMap mBh = make_map_f16(w.hi, (uint64_t)w.N, (uint64_t)w.K, bbox), mBl = make_map_f16(w.lo, (uint64_t)w.N, (uint64_t)w.K, bbox);
        tc::GemmArgs g{ M, w.N, w.K, w.inv_scale, bias, residual, out_f32, out_hi, out_lo, 1, nullptr };
        // Few tiles (a streaming tick is M = sessions rows): a tile's CTA streaming ALL of K is paced by what one SM ingests (FF2 at
        // M = 256: 2 work items, 46 us) while the rest of the GPU idles. Cut K into ranges of >= 2 k-blocks so that the work items
        // about fill the GPU; the partial tiles are summed in order by a second small kernel (deterministic, unlike atomics).
        const int nkb = (w.K + tc::GEMM_BK - 1) / tc::GEMM_BK;
        const int work = pair ? ((tiles_m + 1) / 2) * tiles_n : tiles_m * tiles_n, units = pair ? sm_budget / 2 : sm_budget;
        if (splitk_ok && work * 2 <= units && nkb >= 4) {
            g.ksplit = std::max(1, std::min({ nkb / 2, units / work, 16 }));
            if (g.ksplit > 1) { b_splitk.ensure(sizeof(float) * (size_t)g.ksplit * M * w.N); g.partial = b_splitk.as<float>(); } else g.ksplit = 1;
        }
        if (pair) {
            const unsigned grid = 2u * (unsigned)std::min(work * g.ksplit, sm_budget / 2);
            if (npass == 3) gemm_tc_epi<3, 2>(epi, grid, mAh, mAl, mBh, mBl, g); else gemm_tc_epi<1, 2>(epi, grid, mAh, mAl, mBh, mBl, g);
        } else {
            const unsigned grid = (unsigned)std::min(work * g.ksplit, sm_budget);
            if (npass == 3) gemm_tc_epi<3, 1>(epi, grid, mAh, mAl, mBh, mBl, g); else gemm_tc_epi<1, 1>(epi, grid, mAh, mAl, mBh, mBl, g);
        }
        launch_check();
        if (g.ksplit > 1) {
            const unsigned fgrid = (unsigned)std::min<size_t>(((size_t)M * w.N / 4 + 255) / 256, (size_t)num_sms * 8);
            if (epi == tc::GE_F32) tc::gemm_splitk_finish_kernel<tc::GE_F32><<<fgrid, 256, 0, st>>>(g);
            else if (epi == tc::GE_F32_RESIDUAL) tc::gemm_splitk_finish_kernel<tc::GE_F32_RESIDUAL><<<fgrid, 256, 0, st>>>(g);
            else if (epi == tc::GE_F32_RESIDUAL_SPLIT) tc::gemm_splitk_finish_kernel<tc::GE_F32_RESIDUAL_SPLIT><<<fgrid, 256, 0, st>>>(g);
            else tc::gemm_splitk_finish_kernel<tc::GE_DSWISH_SPLIT><<<fgrid, 256, 0, st>>>(g);
            launch_check();
        }
        span_end();
    }

    // test hook: out[M,N] = A[M,K] W[N,K]^T + bias through the selected tier (0 = SIMT fp32, 1 = fp16, 3 = fp16x3)
    int debug_gemm(int tier, int M, int N, int K, const float* A, const float* W, const float* bias, float* out) override {
        std::lock_guard<std::mutex> lk(mu);
        CK(cudaSetDevice(dev));
        DevBuf dA, dW, dB, dO, dAh, dAl;
        dA.ensure(sizeof(float) * (size_t)M * K); dW.ensure(sizeof(float) * (size_t)N * K); dB.ensure(sizeof(float) * (size_t)N); dO.ensure(sizeof(float) * (size_t)M * N);
        CK(cudaMemcpyAsync(dA.p, A, sizeof(float) * (size_t)M * K, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dW.p, W, sizeof(float) * (size_t)N * K, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dB.p, bias, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, st));
        if (tier == 0) gemm_launch(k::EPI_BIAS, dA.as<float>(), dW.as<float>(), dB.as<float>(), nullptr, dO.as<float>(), M, N, K);
        else {
            if (K % 8 || N % 8) return -1;
            const int saved = npass;
            npass = tier;
            std::vector<float> wv(W, W + (size_t)N * K);
            TcWeight tw = make_tc_weight(wv, N, K);
            dAh.ensure(sizeof(__half) * (size_t)M * K); dAl.ensure(sizeof(__half) * (size_t)M * K);
            tc::split_rows_kernel<<<(unsigned)std::min<size_t>(((size_t)M * K / 4 + 255) / 256, 148 * 8), 256, 0, st>>>(dA.as<float>(), dAh.as<__half>(), dAl.as<__half>(), (size_t)M * K);
            launch_check();
            gemm_tc(tc::GE_F32, dAh.as<__half>(), dAl.as<__half>(), tw, dB.as<float>(), nullptr, dO.as<float>(), nullptr, nullptr, M);
            npass = saved;
        }
        CK(cudaMemcpyAsync(out, dO.p, sizeof(float) * (size_t)M * N, cudaMemcpyDeviceToHost, st));
        release_load_scratch();
        return 0;
    }

    int debug_fbank(const int16_t* pcm, size_t n, float* out_rows, int max_rows) override {
        std::lock_guard<std::mutex> lk(mu);
        CK(cudaSetDevice(dev));
        int nf = n >= (size_t)frame_N ? (int)((n - (size_t)frame_N) / (size_t)frame_shift) + 1 : 0;
        nf = std::min(nf, max_rows);
        if (nf <= 0) return 0;
        b_pcm.ensure(n * sizeof(int16_t));
        CK(cudaMemcpyAsync(b_pcm.p, pcm, n * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        int hm[4] = { 0, nf, nf, 0 };                     // pad_before, n_frames, n_new, carry
        long long off = 0;
        b_meta.ensure(64);
        CK(cudaMemcpyAsync(b_meta.p, hm, sizeof hm, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync((char*)b_meta.p + 32, &off, sizeof off, cudaMemcpyHostToDevice, st));
        b_fb.ensure(sizeof(float) * (size_t)nf * d.MEL);
        int* dm = b_meta.as<int>();
        if (fbank512()) k::fbank512_kernel<<<dim3((unsigned)((nf + 3) / 4), 1), 128, kFbank512Smem, st>>>(
            fp, b_pcm.as<int16_t>(), reinterpret_cast<long long*>((char*)b_meta.p + 32), dm, dm + 1, dm + 2, dm + 3, b_fb.as<float>(), nf);
        else k::fbank_kernel<<<dim3((unsigned)((nf + 3) / 4), 1), 128, (size_t)4 * 3 * frame_N * sizeof(double), st>>>(
            fp, b_pcm.as<int16_t>(), reinterpret_cast<long long*>((char*)b_meta.p + 32), dm, dm + 1, dm + 2, dm + 3, b_fb.as<float>(), nf);
        launch_check();
        CK(cudaMemcpyAsync(out_rows, b_fb.p, sizeof(float) * (size_t)nf * d.MEL, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return nf;
    }

    // ------------------------------------------------------------------------------------------
    void launch_check() { stat.kernel_launches++; CK(cudaGetLastError()); }

    void gemm(int epi, const float* A, const float* W, const float* bias, const float* R, float* C, int M, int N, int K) {
        span_begin(0);
        if (profile) { stat.gemm_flops += 2.0 * (double)M * (double)N * (double)K; stat.gemm_launches++; }
        gemm_launch(epi, A, W, bias, R, C, M, N, K);
        span_end();
    }
    void gemm_launch(int epi, const float* A, const float* W, const float* bias, const float* R, float* C, int M, int N, int K) {
        dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
        if (epi == k::EPI_BIAS) k::gemm_simt_kernel<k::EPI_BIAS><<<grid, 256, 0, st>>>(A, W, bias, R, C, M, N, K);
        else if (epi == k::EPI_BIAS_DSWISH) k::gemm_simt_kernel<k::EPI_BIAS_DSWISH><<<grid, 256, 0, st>>>(A, W, bias, R, C, M, N, K);
        else k::gemm_simt_kernel<k::EPI_BIAS_RESIDUAL><<<grid, 256, 0, st>>>(A, W, bias, R, C, M, N, K);
        launch_check();
    }

    double sync_ms_since(std::chrono::steady_clock::time_point& t0) {
        CK(cudaStreamSynchronize(st));
        auto t1 = std::chrono::steady_clock::now();
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
        return ms;
    }

    void run_pass(std::vector<PassItem>& items, const std::function<void(int, int)>* early) override {
        std::lock_guard<std::mutex> lk(mu);
        const int B = (int)items.size();
        if (B == 0) return;
        stat.passes++; stat.pass_sessions += (uint64_t)B;
        CK(cudaSetDevice(dev));
        const int SEG = d.SEG, STEP = d.STEP, N = frame_N, shift = frame_shift, MEL = d.MEL;

        // ---- plan (integer logic of fbank.c's ring + april_session.c:547-559) ---------------------
        std::vector<PassMeta> meta((size_t)B);
        size_t pcm_total = 0;
        int max_new = 0, max_rows = 0, Tmax = 0;
        for (int b = 0; b < B; b++) {
            PassItem& it = items[(size_t)b];
            StreamFraming& fr = *it.fr;
            PassMeta& m = meta[(size_t)b];
            m = PassMeta{};
            m.slot = fr.slot;
            PlanState ps; ps.avail = fr.avail; ps.avail_f = fr.avail_f; ps.tail = fr.tail.size();
            const PassPlan pl = plan_pass(ps, it.count, it.flush, SEG, STEP, N, shift, (size_t)FLUSH_ZEROS);   // plan.h: CPU-tested against the oracle
            const int steps = pl.nsteps, nf = pl.n_frames;
            const size_t n_buf = pl.n_buf;
            m.carry = pl.carry;
            m.pad_before = pl.pad_before; m.n_frames = nf; m.n_new = pl.n_new;
            m.nsteps = steps; m.flush_flag = it.flush ? 1 : 0;
            m.keep_first = pl.keep_first; m.n_keep = pl.n_keep;
            if (m.n_keep != (int)ps.avail || m.n_keep > CARRY_CAP || m.n_keep < 0) { fprintf(stderr, "libapril: (engine) [ERROR] internal: carry accounting\n"); abort(); }
            m.pcm_off = (long long)pcm_total;
            pcm_total += (n_buf + 7) & ~(size_t)7;
            max_new = std::max(max_new, m.n_new);
            max_rows = std::max(max_rows, m.carry + m.n_new);
            Tmax = std::max(Tmax, steps);
            fr.avail = ps.avail; fr.avail_f = ps.avail_f; fr.steps_total += steps;
            it.n_steps = steps;
            stat.frames += (uint64_t)nf;
            stat.encoder_steps += (uint64_t)steps;
        }
        const int row_pitch = std::max(max_rows, 1);

        // ---- stage PCM: [tail | new samples or zeros] per session ---------------------------------
        p_pcm.ensure(std::max<size_t>(pcm_total, 8) * sizeof(int16_t));
        int16_t* hp = p_pcm.as<int16_t>();
        auto stage_one = [&](int b) {
            PassItem& it = items[(size_t)b];
            StreamFraming& fr = *it.fr;
            int16_t* dst = hp + meta[(size_t)b].pcm_off;
            size_t nt = fr.tail.size(), n_in = it.flush ? (size_t)FLUSH_ZEROS : it.count;
            if (nt) memcpy(dst, fr.tail.data(), nt * sizeof(int16_t));
            if (it.flush) memset(dst + nt, 0, n_in * sizeof(int16_t));
            else if (n_in) memcpy(dst + nt, it.pcm, n_in * sizeof(int16_t));
            size_t n_buf = nt + n_in;
            size_t consumed = (size_t)meta[(size_t)b].n_frames * (size_t)shift;   // next frame starts here (fbank.c:198-224)
            if (consumed > n_buf) consumed = n_buf;
            fr.tail.assign(dst + consumed, dst + n_buf);
        };
        // Big passes: the copy into pinned memory is the host's largest cost and the H2D transfer the next one. Worker threads stage
        // the sessions in index order; as soon as a block of sessions is staged its bytes are handed to the copy engine, so the
        // transfer of block k runs under the staging of block k + 1 (the pass then waits only for the last block's transfer).
        cudaEvent_t ev0 = ev_get(), ev1 = ev_get(), evc = ev_get();
        bool pcm_uploaded = false;
        const int n_stage = (pcm_total * sizeof(int16_t) >= ((size_t)8 << 20)) ? std::min(B, 8) : 1;
        // Caller-pinned audio (aab_pin_host_memory, or any buffer page-locked for CUDA): no staging copy at all. The few hundred tail
        // samples of a session go through the pinned staging buffer, its new samples are DMA'd straight from the caller's memory.
        bool all_pinned = n_stage > 1;
        for (int b = 0; b < B && all_pinned; b++) {
            const PassItem& it = items[(size_t)b];
            if (it.flush || it.count == 0) continue;
            cudaPointerAttributes pa{};
            if (cudaPointerGetAttributes(&pa, it.pcm) != cudaSuccess || pa.type != cudaMemoryTypeHost) { cudaGetLastError(); all_pinned = false; }
        }
        if (all_pinned) {
            b_pcm.ensure(pcm_total * sizeof(int16_t));
            CK(cudaEventRecord(ev0, st));
            int16_t* dp = b_pcm.as<int16_t>();
            for (int b = 0; b < B; b++) {
                PassItem& it = items[(size_t)b];
                StreamFraming& fr = *it.fr;
                const size_t off = (size_t)meta[(size_t)b].pcm_off, nt = fr.tail.size(), n_in = it.flush ? (size_t)FLUSH_ZEROS : it.count, n_buf = nt + n_in;
                if (nt) { memcpy(hp + off, fr.tail.data(), nt * sizeof(int16_t)); CK(cudaMemcpyAsync(dp + off, hp + off, nt * sizeof(int16_t), cudaMemcpyHostToDevice, st)); }
                if (it.flush) CK(cudaMemsetAsync(dp + off + nt, 0, n_in * sizeof(int16_t), st));
                else if (n_in) CK(cudaMemcpyAsync(dp + off + nt, it.pcm, n_in * sizeof(int16_t), cudaMemcpyHostToDevice, st));
                size_t consumed = (size_t)meta[(size_t)b].n_frames * (size_t)shift;     // next frame starts here (fbank.c:198-224)
                if (consumed > n_buf) consumed = n_buf;
                std::vector<int16_t> rest; rest.reserve(n_buf - consumed);
                for (size_t i = consumed; i < n_buf; i++) rest.push_back(i < nt ? hp[off + i] : (it.flush ? (int16_t)0 : it.pcm[i - nt]));
                fr.tail.swap(rest);
            }
            stat.h2d_bytes += pcm_total * sizeof(int16_t);
            pcm_uploaded = true;
        } else if (n_stage > 1) {
            b_pcm.ensure(pcm_total * sizeof(int16_t));
            CK(cudaEventRecord(ev0, st));
            const int nblk = std::min(B, 8);
            std::vector<int> blk_end((size_t)nblk);
            for (int g = 0; g < nblk; g++) blk_end[(size_t)g] = (int)(((long long)B * (g + 1)) / nblk);
            std::vector<std::atomic<int>> blk_done((size_t)nblk);
            for (auto& x : blk_done) x.store(0);
            std::atomic<int> next{ 0 };
            auto block_of = [&](int b) { int g = 0; while (b >= blk_end[(size_t)g]) g++; return g; };
            std::vector<std::thread> th;
            for (int k = 0; k < n_stage; k++) th.emplace_back([&] { for (int b; (b = next.fetch_add(1)) < B;) { stage_one(b); blk_done[(size_t)block_of(b)].fetch_add(1, std::memory_order_release); } });
            for (int g = 0, b0 = 0; g < nblk; b0 = blk_end[(size_t)g], g++) {
                const int want = blk_end[(size_t)g] - b0;
                while (blk_done[(size_t)g].load(std::memory_order_acquire) < want) std::this_thread::yield();
                const size_t o0 = (size_t)meta[(size_t)b0].pcm_off, o1 = blk_end[(size_t)g] < B ? (size_t)meta[(size_t)blk_end[(size_t)g]].pcm_off : pcm_total;
                if (o1 > o0) CK(cudaMemcpyAsync(b_pcm.as<int16_t>() + o0, hp + o0, (o1 - o0) * sizeof(int16_t), cudaMemcpyHostToDevice, st));
            }
            for (auto& t : th) t.join();
            stat.h2d_bytes += pcm_total * sizeof(int16_t);
            pcm_uploaded = true;
        } else {
            for (int b = 0; b < B; b++) stage_one(b);
        }
        // per-session ints as struct-of-arrays
        const size_t nI = (size_t)B;
        const size_t meta_bytes = (9 * nI + (nI & 1)) * sizeof(int) + nI * sizeof(long long);
        p_meta.ensure(meta_bytes);
        int* hm = p_meta.as<int>();
        int *h_slot = hm, *h_carry = hm + nI, *h_nnew = hm + 2 * nI, *h_padb = hm + 3 * nI, *h_nfr = hm + 4 * nI, *h_nsteps = hm + 5 * nI,
            *h_flush = hm + 6 * nI, *h_keepf = hm + 7 * nI, *h_nkeep = hm + 8 * nI;
        long long* h_off = reinterpret_cast<long long*>(hm + 9 * nI + (nI & 1));
        for (int b = 0; b < B; b++) {
            const PassMeta& m = meta[(size_t)b];
            h_slot[b] = m.slot; h_carry[b] = m.carry; h_nnew[b] = m.n_new; h_padb[b] = m.pad_before; h_nfr[b] = m.n_frames; h_nsteps[b] = m.nsteps;
            h_flush[b] = m.flush_flag; h_keepf[b] = m.keep_first; h_nkeep[b] = m.n_keep; h_off[b] = m.pcm_off;
        }
        size_t rec_bytes = 0, np_bytes = 0;
        const bool want_tap = tap != nullptr;
        bool any_flush_k = false;
        for (int b = 0; b < B; b++) any_flush_k |= items[(size_t)b].flush;
        struct EarlySteps { cudaEvent_t ev; int t0, t1; };
        std::vector<EarlySteps> early_steps;        // time chunks whose decisions reach the host while later chunks are still computed
        int early_until = 0;                        // steps [0, early_until) are copied chunk by chunk; the final copy takes the rest
        // everything from the uploads to the result copies is enqueued by this lambda: eagerly, or once into a CUDA graph
        auto enqueue = [&](bool eager) {
        if (!pcm_uploaded) {
            b_pcm.ensure(std::max<size_t>(pcm_total, 8) * sizeof(int16_t));
            if (pcm_total) CK(cudaMemcpyAsync(b_pcm.p, hp, pcm_total * sizeof(int16_t), cudaMemcpyHostToDevice, st));
            stat.h2d_bytes += pcm_total * sizeof(int16_t);
        }

        b_meta.ensure(meta_bytes);
        CK(cudaMemcpyAsync(b_meta.p, hm, meta_bytes, cudaMemcpyHostToDevice, st));
        stat.h2d_bytes += meta_bytes;
        int* dm = b_meta.as<int>();
        int *d_slot = dm, *d_carry = dm + nI, *d_nnew = dm + 2 * nI, *d_padb = dm + 3 * nI, *d_nfr = dm + 4 * nI, *d_nsteps = dm + 5 * nI,
            *d_flush = dm + 6 * nI, *d_keepf = dm + 7 * nI, *d_nkeep = dm + 8 * nI;
        long long* d_off = reinterpret_cast<long long*>(dm + 9 * nI + (nI & 1));

        if (eager) CK(cudaEventRecord(evc, st));     // uploads done: what follows is the device-resident part of the pass
        // ---- front-end ------------------------------------------------------------------------------
        b_fb.ensure(sizeof(float) * (size_t)B * row_pitch * MEL);
        float* FB = b_fb.as<float>();
        span_begin(4);
        k::carry_in_kernel<<<B, 128, 0, st>>>(carry_store, d_slot, d_carry, FB, row_pitch, MEL, CARRY_CAP); launch_check();
        if (max_new > 0) {
            if (fbank512()) {
                dim3 grid((unsigned)((max_new + 3) / 4), (unsigned)B);
                k::fbank512_kernel<<<grid, 128, kFbank512Smem, st>>>(fp, b_pcm.as<int16_t>(), d_off, d_padb, d_nfr, d_nnew, d_carry, FB, row_pitch);
            } else {
                dim3 grid((unsigned)((max_new + 3) / 4), (unsigned)B);
                k::fbank_kernel<<<grid, 128, (size_t)4 * 3 * N * sizeof(double), st>>>(fp, b_pcm.as<int16_t>(), d_off, d_padb, d_nfr, d_nnew, d_carry, FB, row_pitch);
            }
            launch_check();
        }
        span_end();

        // ---- encoder + search, in time chunks -------------------------------------------------------
        if (Tmax > 0) {
            const size_t row_budget = 160 * 1024;                                  // rows per chunk (bounds the activations to a few GB)
            int Tc = (int)std::max<size_t>(1, std::min<size_t>((size_t)Tmax, row_budget / (size_t)B));
            // Search under the encoder: in the small-batch regime the scan occupies H / 8 SMs and is a latency chain; the search of time
            // chunk c (CTA pairs, 8 sessions each) runs on the SMs it leaves free while the encoder works on chunk c + 1. Only the last
            // chunk's search is exposed. The persistent GEMM grids are sized for the SMs that remain.
            const int ovl_pairs = (B + k::GP_MAX_S - 1) / k::GP_MAX_S;
            int ovl_n = overlap_chunks;                                           // fewer chunks for shorter passes: the last one keeps >= 24 steps
            while (ovl_n >= 2 && Tmax < 24 * ((1 << ovl_n) - 1)) ovl_n--;
            const bool ovl = eager && overlap_search && pair_greedy && fast_greedy && !want_tap && npass && scan4 && scan3 && B >= 24 && B <= 64 && ovl_n >= 2 &&   // below ~24 sessions the search is too short to be worth the chunking (8 sessions: 69.0 vs 65.6 ms)
                            
                             (size_t)Tmax * (size_t)B <= row_budget && d.H / scan3::UNITS + 2 * ovl_pairs <= num_sms;
            // chunk lengths: uniform without overlap; with it, halving (8 : 4 : 2 : 1 ...): the search of a chunk runs ~2.5x slower on its
            // few SMs than the encoder of the same steps on the rest, so each chunk's search must fit under the encoder of ALL later
            // chunks, and the last chunk (whose search is exposed) should be short
            std::vector<int> chunk_len;
            if (ovl) {
                const int n = ovl_n, wsum = (1 << n) - 1;
                for (int c = 0, left = Tmax; c < n; c++) {
                    const int len = c == n - 1 ? left : std::min(left, std::max(1, (int)(((long long)Tmax * (1 << (n - 1 - c)) + wsum / 2) / wsum)));
                    if (len > 0) chunk_len.push_back(len);
                    left -= len;
                }
                Tc = 0; for (int len : chunk_len) Tc = std::max(Tc, len);
            } else {
                // a listener for early steps and a wide batch: three time chunks even when memory does not force them, so that the result
                // replay of a chunk (5.7 ms of host time at 160 x 30 s) runs while the device computes the next
                if (early && eager && !want_tap && B > 64 && Tmax >= 96 && Tc >= Tmax) Tc = (Tmax + 2) / 3;
                for (int t0 = 0; t0 < Tmax; t0 += Tc) chunk_len.push_back(std::min(Tc, Tmax - t0));
            }
            const size_t Mc = (size_t)Tc * B;
            const int L = d.L, D = d.D, H = d.H, F = d.F, J = d.J, KE = d.C3 * d.F3;
            b_xe.ensure(sizeof(float) * Mc * KE); b_x.ensure(sizeof(float) * Mc * D); b_x1.ensure(sizeof(float) * Mc * D); b_x2.ensure(sizeof(float) * Mc * D);
            b_g.ensure(sizeof(float) * Mc * 4 * H); b_hout.ensure(sizeof(float) * Mc * D); b_f1.ensure(sizeof(float) * Mc * F);
            b_m.ensure(sizeof(float) * (size_t)B * H); b_hw.ensure(sizeof(float) * (size_t)L * B * D); b_cw.ensure(sizeof(float) * (size_t)L * B * H);
            b_eo.ensure(sizeof(float) * Mc * J * (ovl ? 2 : 1));          // double-buffered when the search trails the encoder by a chunk
            const bool use_s4 = npass && scan4;
            if (use_s4) { mall_rows = Mc + (size_t)B; b_mall.ensure(sizeof(__half) * 2 * mall_rows * H); }
            b_mw.ensure(sizeof(float) * (size_t)L * B * H);
            if (npass) {
                b_xe_h.ensure(sizeof(__half) * Mc * KE); b_xe_l.ensure(sizeof(__half) * Mc * KE);
                if (conv_tc) { b_a3_h.ensure(sizeof(__half) * Mc * d.F3 * d.C2 * 9); b_a3_l.ensure(sizeof(__half) * Mc * d.F3 * d.C2 * 9); }
                b_x_h.ensure(sizeof(__half) * Mc * D); b_x_l.ensure(sizeof(__half) * Mc * D);
                b_x1_h.ensure(sizeof(__half) * Mc * D); b_x1_l.ensure(sizeof(__half) * Mc * D);
                b_f1_h.ensure(sizeof(__half) * Mc * F); b_f1_l.ensure(sizeof(__half) * Mc * F);
            }
            __half *Xe_h = b_xe_h.as<__half>(), *Xe_l = b_xe_l.as<__half>(), *X_h = b_x_h.as<__half>(), *X_l = b_x_l.as<__half>(), *X1_h = b_x1_h.as<__half>(),
                   *X1_l = b_x1_l.as<__half>(), *F1_h = b_f1_h.as<__half>(), *F1_l = b_f1_l.as<__half>();
            rec_bytes = sizeof(k::DecisionRecord) * (size_t)B * Tmax * 3; np_bytes = sizeof(int) * (size_t)B * Tmax;
            b_rec.ensure(rec_bytes); b_np.ensure(np_bytes);
            p_rec.ensure(rec_bytes); p_np.ensure(np_bytes);
            if (want_tap) { b_tap.ensure(sizeof(float) * (size_t)B * Tmax * 3 * d.V); p_tap.ensure(sizeof(float) * (size_t)B * Tmax * 3 * d.V); p_eo.ensure(sizeof(float) * (size_t)Tmax * B * J); }
            float *Xe = b_xe.as<float>(), *X = b_x.as<float>(), *X1 = b_x1.as<float>(), *X2 = b_x2.as<float>(), *G = b_g.as<float>(), *Hout = b_hout.as<float>(),
                  *F1 = b_f1.as<float>(), *Ms = b_m.as<float>(), *Hw = b_hw.as<float>(), *Cw = b_cw.as<float>(), *EO = b_eo.as<float>();
            k::state_gather_kernel<<<dim3((unsigned)B, (unsigned)L), 128, 0, st>>>(h_store, d_slot, Hw, L, B, D); launch_check();
            k::state_gather_kernel<<<dim3((unsigned)B, (unsigned)L), 128, 0, st>>>(c_store, d_slot, Cw, L, B, H); launch_check();
            float* Mw = b_mw.as<float>();
            if (use_s4) { k::state_gather_kernel<<<dim3((unsigned)B, (unsigned)L), 128, 0, st>>>(m_store, d_slot, Mw, L, B, H); launch_check(); }
            const size_t greedy_smem = sizeof(float) * (size_t)(3 * J + d.E);
            std::vector<cudaEvent_t> ev_search;                                   // per chunk: its search has finished (overlap mode)
            float* const EO_base = EO;
            for (int t0 = 0, ci = 0; ci < (int)chunk_len.size(); t0 += chunk_len[(size_t)ci], ci++) {
                const int tc = chunk_len[(size_t)ci];
                const int M = tc * B;
                if (ovl) EO = EO_base + (size_t)(ci & 1) * Mc * J;
                const unsigned ew_grid = (unsigned)std::min<size_t>(((size_t)M * D / 4 + 255) / 256, (size_t)num_sms * 8);
                if (npass == 0) {
                    k::conv_embed_kernel<<<dim3((unsigned)B, (unsigned)tc), 256, conv_smem(), st>>>(cp, FB, row_pitch, d_nsteps, t0, B, Xe); launch_check();
                    gemm(k::EPI_BIAS, Xe, out_w, out_b, nullptr, X1, M, D, KE);
                    k::basic_norm_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(X1, X, M, D, out_norm_eps); launch_check();
                } else if (conv_tc) {
                    // conv1/conv2 + im2col (SIMT), conv3 as a GEMM with fused DoubleSwish + split, then the linear on permuted weights
                    tc::conv12_im2col_kernel<<<dim3((unsigned)B, (unsigned)tc), 256, conv_smem(), st>>>(c12, FB, row_pitch, d_nsteps, t0, B, b_a3_h.as<__half>(), b_a3_l.as<__half>()); launch_check();
                    gemm_tc(tc::GE_DSWISH_SPLIT, b_a3_h.as<__half>(), b_a3_l.as<__half>(), tw_conv3, conv3_b, nullptr, nullptr, Xe_h, Xe_l, M * d.F3);
                    gemm_tc(tc::GE_F32, Xe_h, Xe_l, tw_out_perm, out_b, nullptr, X1, nullptr, nullptr, M);
                    tc::basic_norm_split_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(X1, X, X_h, X_l, M, D, out_norm_eps); launch_check();
                } else {
                    k::conv_embed_kernel<<<dim3((unsigned)B, (unsigned)tc), 256, conv_smem(), st>>>(cp, FB, row_pitch, d_nsteps, t0, B, Xe); launch_check();
                    tc::split_rows_kernel<<<(unsigned)std::min<size_t>(((size_t)M * KE / 4 + 255) / 256, (size_t)num_sms * 8), 256, 0, st>>>(Xe, Xe_h, Xe_l, (size_t)M * KE); launch_check();
                    gemm_tc(tc::GE_F32, Xe_h, Xe_l, tw_out, out_b, nullptr, X1, nullptr, nullptr, M);
                    tc::basic_norm_split_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(X1, X, X_h, X_l, M, D, out_norm_eps); launch_check();
                }
                for (int l = 0; l < L; l++) {
                    const DeviceLayer& w = layers[(size_t)l];
                    if (npass == 0) gemm(k::EPI_BIAS, X, w.w_ih, w.b_gate, nullptr, G, M, 4 * H, D);
                    else gemm_tc(tc::GE_F32, X_h, X_l, tlayers[(size_t)l].w_ih, w.b_gate, nullptr, G, nullptr, nullptr, M, 3);
                    float* hw = Hw + (size_t)l * B * D;
                    float* cw = Cw + (size_t)l * B * H;
                    span_begin(1);
                    if (use_s4) scan4_layer(l, G, Mw + (size_t)l * B * H, cw, d_nsteps, t0, tc, B);
                    else for (int t = 0; t < tc; t++) {
                        k::lstm_gates_kernel<<<dim3((unsigned)(H / 16), (unsigned)((B + 63) / 64)), 256, 0, st>>>(
                            G + (size_t)t * B * 4 * H, hw, w.w_hh, cw, Ms, d_nsteps, t0 + t, B, D, H);
                        launch_check();
                        k::lstm_proj_kernel<<<dim3((unsigned)((D + 63) / 64), (unsigned)((B + 63) / 64)), 256, 0, st>>>(
                            Ms, w.w_hr, hw, Hout + (size_t)t * B * D, d_nsteps, t0 + t, B, D, H);
                        launch_check();
                    }
                    span_end();
                    if (npass == 0) {
                        k::add_rows_kernel<<<ew_grid, 256, 0, st>>>(reinterpret_cast<const float4*>(X), reinterpret_cast<const float4*>(Hout), reinterpret_cast<float4*>(X1), (size_t)M * D / 4);
                        launch_check();
                        gemm(k::EPI_BIAS_DSWISH, X1, w.ff1_w, w.ff1_b, nullptr, F1, M, F, D);
                        gemm(k::EPI_BIAS_RESIDUAL, F1, w.ff2_w, w.ff2_b, X1, X2, M, D, F);
                        k::basic_norm_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(X2, X, M, D, w.norm_eps); launch_check();
                    } else {
                        const TcLayer& tl = tlayers[(size_t)l];
                        const size_t moff = (size_t)B * H;                // row block 0 of the m planes is the initial state
                        if (use_s4) gemm_tc(tc::GE_F32_RESIDUAL_SPLIT, b_mall.as<__half>() + moff, b_mall.as<__half>() + mall_rows * H + moff, s4layers[(size_t)l].w_hr, nullptr, X, X1, X1_h, X1_l, M, 3);   // x + W_hr m
                        else { tc::add_rows_split_kernel<<<ew_grid, 256, 0, st>>>(X, Hout, X1, X1_h, X1_l, (size_t)M * D); launch_check(); }
                        gemm_tc(tc::GE_DSWISH_SPLIT, X1_h, X1_l, tl.ff1, w.ff1_b, nullptr, nullptr, F1_h, F1_l, M);
                        gemm_tc(tc::GE_F32_RESIDUAL, F1_h, F1_l, tl.ff2, w.ff2_b, X1, X2, nullptr, nullptr, M);
                        tc::basic_norm_split_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(X2, X, X_h, X_l, M, D, w.norm_eps); launch_check();
                    }
                }
                if (ovl && ci >= 2) CK(cudaStreamWaitEvent(st, ev_search[(size_t)ci - 2], 0));   // this chunk's encoder rows overwrite the buffer searched two chunks ago
                if (npass == 0) gemm(k::EPI_BIAS, X, enc_proj_w, enc_proj_b, nullptr, EO, M, J, D);
                else gemm_tc(tc::GE_F32, X_h, X_l, tw_proj, enc_proj_b, nullptr, EO, nullptr, nullptr, M);
                const int last = (t0 + tc >= Tmax) ? 1 : 0;
                if (ovl) {
                    cudaEvent_t e_enc = ev_get(), e_done = ev_get();
                    CK(cudaEventRecord(e_enc, st));
                    CK(cudaStreamWaitEvent(st2, e_enc, 0));
                    span_begin(2, st2);
                    const int npair = (B + k::GP_MAX_S - 1) / k::GP_MAX_S;
                    if (!last) {
                        // 8 sessions per CTA pair: as few SMs as possible
                        k::greedy_pair_kernel<<<npair * 2, k::GP_THREADS, gp_smem, st2>>>(jp, dtab, EO, d_nsteps, d_slot, d_flush, B, Tmax, t0, stride_ms(), gstate,
                                                                                           b_rec.as<k::DecisionRecord>(), b_np.as<int>(), nullptr, last, tc, k::GP_MAX_S, gp_Rh);
                    } else {
                        // the last chunk's search is exposed and has the GPU to itself: the 8-CTA clusters are the faster kernel at this size
                        const int maxc = std::max(1, num_sms / k::GC_CLUSTER);
                        const int S = std::min(k::GC_MAX_S, std::max(1, (B + maxc - 1) / maxc));
                        const int ncl = (B + S - 1) / S;
                        k::greedy_cluster_kernel<<<ncl * k::GC_CLUSTER, 256, gr_smem, st2>>>(jp, dtab, EO, d_nsteps, d_slot, d_flush, B, Tmax, t0, stride_ms(), gstate,
                                                                                             b_rec.as<k::DecisionRecord>(), b_np.as<int>(), nullptr, last, tc, S, gr_R);
                    }
                    launch_check();
                    span_end(st2);
                    CK(cudaEventRecord(e_done, st2));
                    ev_search.push_back(e_done); ev_pool.push_back(e_enc);       // (an event may be re-recorded once the waits on it are enqueued)
                    if (early && !last) {                                        // this chunk's decisions go to the host now: [B][Tmax][3] records, [B][Tmax] pass counts
                        const size_t rp = sizeof(k::DecisionRecord) * 3;
                        CK(cudaMemcpy2DAsync(p_rec.as<char>() + (size_t)t0 * rp, (size_t)Tmax * rp, b_rec.as<char>() + (size_t)t0 * rp, (size_t)Tmax * rp, (size_t)tc * rp, (size_t)B, cudaMemcpyDeviceToHost, st2));
                        CK(cudaMemcpy2DAsync(p_np.as<char>() + (size_t)t0 * sizeof(int), (size_t)Tmax * sizeof(int), b_np.as<char>() + (size_t)t0 * sizeof(int), (size_t)Tmax * sizeof(int), (size_t)tc * sizeof(int), (size_t)B, cudaMemcpyDeviceToHost, st2));
                        cudaEvent_t e_copy = ev_get();
                        CK(cudaEventRecord(e_copy, st2));
                        early_steps.push_back(EarlySteps{ e_copy, t0, t0 + tc });
                        early_until = t0 + tc;
                    }
                    sm_budget = last ? num_sms : ((num_sms - 2 * npair) & ~1);  // GEMMs of the next chunk: the SMs the search does not hold
                    if (last) CK(cudaStreamWaitEvent(st, e_done, 0));
                    continue;
                }
                span_begin(2);
                if (fast_greedy && pair_greedy && B >= pair_greedy_min) {
                    const int maxp = std::max(1, num_sms / 2);
                    const int S = std::min(k::GP_MAX_S, std::max(1, (B + maxp - 1) / maxp));     // one wave of CTA pairs
                    const int npair = (B + S - 1) / S;
                    k::greedy_pair_kernel<<<npair * 2, k::GP_THREADS, gp_smem, st>>>(jp, dtab, EO, d_nsteps, d_slot, d_flush, B, Tmax, t0, stride_ms(), gstate,
                                                                                      b_rec.as<k::DecisionRecord>(), b_np.as<int>(), want_tap ? b_tap.as<float>() : nullptr, last, tc, S, gp_Rh);
                } else if (fast_greedy) {
                    const int maxc = std::max(1, num_sms / k::GC_CLUSTER);
                    const int S = std::min(k::GC_MAX_S, std::max(1, (B + maxc - 1) / maxc));
                    const int ncl = (B + S - 1) / S;
                    k::greedy_cluster_kernel<<<ncl * k::GC_CLUSTER, 256, gr_smem, st>>>(jp, dtab, EO, d_nsteps, d_slot, d_flush, B, Tmax, t0, stride_ms(), gstate,
                                                                                        b_rec.as<k::DecisionRecord>(), b_np.as<int>(), want_tap ? b_tap.as<float>() : nullptr, last, tc, S, gr_R);
                } else {
                    k::greedy_kernel<<<B, 256, greedy_smem, st>>>(jp, EO, d_nsteps, d_slot, d_flush, B, Tmax, t0, stride_ms(), gstate, dout_store, dout_blank,
                                                                 b_rec.as<k::DecisionRecord>(), b_np.as<int>(), want_tap ? b_tap.as<float>() : nullptr, last, tc);
                }
                launch_check();
                span_end();
                if (early && eager && !last && !want_tap) {                      // any multi-chunk pass: this chunk's decisions go to the host now
                    const size_t rp = sizeof(k::DecisionRecord) * 3;
                    CK(cudaMemcpy2DAsync(p_rec.as<char>() + (size_t)t0 * rp, (size_t)Tmax * rp, b_rec.as<char>() + (size_t)t0 * rp, (size_t)Tmax * rp, (size_t)tc * rp, (size_t)B, cudaMemcpyDeviceToHost, st));
                    CK(cudaMemcpy2DAsync(p_np.as<char>() + (size_t)t0 * sizeof(int), (size_t)Tmax * sizeof(int), b_np.as<char>() + (size_t)t0 * sizeof(int), (size_t)Tmax * sizeof(int), (size_t)tc * sizeof(int), (size_t)B, cudaMemcpyDeviceToHost, st));
                    cudaEvent_t e_copy = ev_get();
                    CK(cudaEventRecord(e_copy, st));
                    early_steps.push_back(EarlySteps{ e_copy, t0, t0 + tc });
                    early_until = t0 + tc;
                }
                if (want_tap) { CK(cudaMemcpyAsync(p_eo.as<float>() + (size_t)t0 * B * J, EO, sizeof(float) * (size_t)M * J, cudaMemcpyDeviceToHost, st)); stat.d2h_bytes += sizeof(float) * (size_t)M * J; }
            }
            sm_budget = num_sms;
            for (cudaEvent_t e : ev_search) ev_pool.push_back(e);
            k::state_scatter_kernel<<<dim3((unsigned)B, (unsigned)L), 128, 0, st>>>(h_store, d_slot, Hw, L, B, D); launch_check();
            k::state_scatter_kernel<<<dim3((unsigned)B, (unsigned)L), 128, 0, st>>>(c_store, d_slot, Cw, L, B, H); launch_check();
            if (use_s4) { k::state_scatter_kernel<<<dim3((unsigned)B, (unsigned)L), 128, 0, st>>>(m_store, d_slot, Mw, L, B, H); launch_check(); }
            if (early_until == 0) {
                CK(cudaMemcpyAsync(p_rec.p, b_rec.p, rec_bytes, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(p_np.p, b_np.p, np_bytes, cudaMemcpyDeviceToHost, st));
            } else {                                                              // the host may be reading the earlier chunks' records: copy only the rest
                const size_t rp = sizeof(k::DecisionRecord) * 3; const int t0r = early_until, tcr = Tmax - early_until;
                CK(cudaMemcpy2DAsync(p_rec.as<char>() + (size_t)t0r * rp, (size_t)Tmax * rp, b_rec.as<char>() + (size_t)t0r * rp, (size_t)Tmax * rp, (size_t)tcr * rp, (size_t)B, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpy2DAsync(p_np.as<char>() + (size_t)t0r * sizeof(int), (size_t)Tmax * sizeof(int), b_np.as<char>() + (size_t)t0r * sizeof(int), (size_t)Tmax * sizeof(int), (size_t)tcr * sizeof(int), (size_t)B, cudaMemcpyDeviceToHost, st));
            }
            stat.d2h_bytes += rec_bytes + np_bytes;
            if (want_tap) { CK(cudaMemcpyAsync(p_tap.p, b_tap.p, sizeof(float) * (size_t)B * Tmax * 3 * d.V, cudaMemcpyDeviceToHost, st)); }
        } else {
            // no encoder step, but a flush must still clear the context on the device (april_session.c:561-563)
            bool any_flush = false;
            for (int b = 0; b < B; b++) any_flush |= items[(size_t)b].flush;
            if (any_flush) {
                b_rec.ensure(sizeof(k::DecisionRecord) * 3); b_np.ensure(sizeof(int));
                b_eo.ensure(sizeof(float) * (size_t)d.J);
                k::greedy_kernel<<<B, 256, sizeof(float) * (size_t)(3 * d.J + d.E), st>>>(jp, b_eo.as<float>(), d_nsteps, d_slot, d_flush, B, 1, 0, stride_ms(), gstate, dout_store,
                                                                                         dout_blank, b_rec.as<k::DecisionRecord>(), b_np.as<int>(), nullptr, 1, 1);
                launch_check();
            }
        }
        k::carry_out_kernel<<<B, 128, 0, st>>>(carry_store, d_slot, d_keepf, d_nkeep, FB, row_pitch, MEL, CARRY_CAP); launch_check();
        };
        // small passes with a geometry seen before replay a graph; the second occurrence of a geometry captures it
        static const bool graphs_env = !(getenv("APRIL_B200_GRAPHS") && !strcmp(getenv("APRIL_B200_GRAPHS"), "0"));
        const bool small = graphs_ok && graphs_env && !profile && !want_tap && Tmax <= 8 && (size_t)B * (size_t)std::max(Tmax, 1) <= 8192;
        GraphKey key; key.B = B; key.Tmax = Tmax; key.max_new = max_new; key.row_pitch = row_pitch; key.any_flush = any_flush_k ? 1 : 0; key.pcm_total = pcm_total; key.gen = g_buf_gen.load();
        CachedGraph* hit = nullptr;
        if (small) for (auto& g : graphs) if (g.key == key) { hit = &g; break; }
        bool timed_compute = false;
        if (!pcm_uploaded) CK(cudaEventRecord(ev0, st));
        if (hit) {
            CK(cudaGraphLaunch(hit->exec, st));
            stat.kernel_launches += hit->launches; stat.h2d_bytes += hit->h2d; stat.d2h_bytes += hit->d2h; rec_bytes = hit->rec_bytes; np_bytes = hit->np_bytes;
        } else if (small && pending_key == key) {
            const uint64_t l0 = stat.kernel_launches, h0 = stat.h2d_bytes, d0 = stat.d2h_bytes;
            cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
            tl_soft_errors = true; tl_soft_failed = false; capturing = true;
            bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) { enqueue(false); ok = cudaStreamEndCapture(st, &graph) == cudaSuccess && !tl_soft_failed && graph != nullptr; }
            if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            tl_soft_errors = false; capturing = false;
            if (graph) cudaGraphDestroy(graph);
            if (ok && key.gen == g_buf_gen.load()) {
                CachedGraph g; g.key = key; g.exec = exec; g.launches = stat.kernel_launches - l0; g.h2d = stat.h2d_bytes - h0; g.d2h = stat.d2h_bytes - d0; g.rec_bytes = rec_bytes; g.np_bytes = np_bytes;
                if (graphs.size() >= 4) { cudaGraphExecDestroy(graphs.front().exec); graphs.erase(graphs.begin()); }
                graphs.push_back(g);
                CK(cudaGraphLaunch(exec, st));
            } else {
                cudaGetLastError();
                if (exec) cudaGraphExecDestroy(exec);
                if (!ok) { graphs_ok = false; fprintf(stderr, "libapril: (engine) [WARNING] CUDA graph capture of small passes is not available; launching eagerly\n"); }
                stat.kernel_launches = l0; stat.h2d_bytes = h0; stat.d2h_bytes = d0;
                enqueue(true); timed_compute = true;
            }
        } else {
            enqueue(true); timed_compute = true;
            pending_key = key; pending_key.gen = g_buf_gen.load();      // buffers may have grown during this eager pass
            if (!small) pending_key.B = -1;
        }
        CK(cudaEventRecord(ev1, st));
        // ---- hand decisions back: chunk by chunk while the device works on (overlap mode with a listener), the rest at the end ----
        const k::DecisionRecord* hr = p_rec.as<k::DecisionRecord>();
        const int* hn = p_np.as<int>();
        for (int b = 0; b < B; b++) {
            PassItem& it = items[(size_t)b];
            it.npass.assign((size_t)it.n_steps, 0);
            it.decisions.assign((size_t)it.n_steps * 3, StepDecision{});
        }
        auto hand_back = [&](int t_from, int t_to) {
            for (int b = 0; b < B; b++) {
                PassItem& it = items[(size_t)b];
                for (int t = t_from; t < std::min(t_to, it.n_steps); t++) {
                    int np = hn[(size_t)b * Tmax + t];
                    it.npass[(size_t)t] = np;
                    stat.joiner_passes += (uint64_t)np;
                    for (int p = 0; p < np; p++) {
                        const k::DecisionRecord& r = hr[((size_t)b * Tmax + t) * 3 + p];
                        it.decisions[(size_t)t * 3 + p] = StepDecision{ r.max_idx, r.is_blank, r.max_val, r.blank_val };
                        if (!r.is_blank) stat.decoder_runs++;
                    }
                }
            }
        };
        for (const EarlySteps& e : early_steps) {
            CK(cudaEventSynchronize(e.ev));
            hand_back(e.t0, e.t1);
            (*early)(e.t0, e.t1);
            ev_pool.push_back(e.ev);
        }
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        stat.device_ms += ms;
        if (timed_compute) CK(cudaEventElapsedTime(&ms, evc, ev1));
        stat.compute_ms += ms;
        ev_pool.push_back(ev0); ev_pool.push_back(ev1); ev_pool.push_back(evc);
        spans_collect();

        hand_back(early_until, Tmax);
        if (want_tap && Tmax > 0) {
            const float* eo = p_eo.as<float>();
            const float* lg = p_tap.as<float>();
            for (int b = 0; b < B; b++) {
                PassItem& it = items[(size_t)b];
                for (int t = 0; t < it.n_steps; t++) {
                    tap(tap_ud, it.cookie, 0, eo + ((size_t)t * B + b) * d.J, (size_t)d.J);
                    for (int p = 0; p < it.npass[(size_t)t]; p++) tap(tap_ud, it.cookie, 2, lg + (((size_t)b * Tmax + t) * 3 + p) * d.V, (size_t)d.V);
                }
            }
        }
    }
};

}  // namespace

Engine* Engine::create(const LoadedModel& m, int device, std::string& err) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { err = "no CUDA device available (this library has no CPU fallback)"; return nullptr; }
    if (device < 0 || device >= n) { err = "CUDA device index out of range"; return nullptr; }
    EngineImpl* e = new EngineImpl();
    if (!e->init(m, device, err)) { delete e; return nullptr; }
    return e;
}

int host_pin(void* p, size_t bytes) {
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}
int host_unpin(void* p) {
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}

std::vector<int> cuda_devices_sm100() {
    std::vector<int> out;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return out; }
    for (int i = 0; i < n; i++) { cudaDeviceProp p; if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) out.push_back(i); }
    return out;
}

}  // namespace april
