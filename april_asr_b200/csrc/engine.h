// Device engine: owns the weights of one model on one B200 and runs batched passes of the hot path
// (fbank -> encoder -> decoder/joiner -> greedy decisions) over many sessions.
// Replaces the three g_ort->Run call sites (april_session.c:145,160,176), the tensor allocation
// (april_session.c:40-58) and the fbank_* calls (april_session.c:35,441-442,528,552-559).
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "model_loader.h"

namespace april {

struct StepDecision {           // mirrors k::DecisionRecord
    int32_t max_idx;
    int32_t is_blank;
    float max_val;
    float blank_val;
};

// Host-side framing state of one stream (what OnlineFBank_i keeps: fbank.c:98-127), minus the data
// that lives on the device.
struct StreamFraming {
    int slot = -1;
    std::vector<int16_t> tail;      // samples from the start of the next frame onwards (prev_leftover)
    long long avail = 0;            // temp_segment_avail
    long long avail_f = 0;          // temp_segment_avail_f
    long long steps_total = 0;
};

struct PassItem {
    StreamFraming* fr = nullptr;
    void* cookie = nullptr;         // handed back to the tap
    const int16_t* pcm = nullptr;
    size_t count = 0;
    bool flush = false;
    // outputs of the pass:
    int n_steps = 0;
    std::vector<int> npass;                 // [n_steps]
    std::vector<StepDecision> decisions;    // [n_steps][3]
    int delivered = 0;                      // steps the caller has already consumed through the early-steps callback
};

struct EngineStats {
    uint64_t encoder_steps = 0, joiner_passes = 0, decoder_runs = 0, frames = 0, kernel_launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    double device_ms = 0, lstm_scan_ms = 0, gemm_ms = 0, compute_ms = 0, gemm_flops = 0;
    uint64_t gemm_launches = 0;
    double greedy_ms = 0;
    uint64_t passes = 0, pass_sessions = 0;
    double lstm_gemm_ms = 0, lstm_gemm_flops = 0;   // the W_ih and W_hr (layer output) GEMMs: with lstm_scan_ms, the LSTM-GEMM subset of north_star
    double frontend_ms = 0;                          // fbank + carry rows
};

typedef void (*TapFn)(void* userdata, void* cookie, int net, const float* data, size_t n);

class Engine {
public:
    static Engine* create(const LoadedModel& m, int device, std::string& err);
    virtual ~Engine() {}
    virtual int alloc_slot() = 0;           // device-side per-stream state (h, c, context, carry rows); zeros / blanks
    virtual void free_slot(int slot) = 0;
    // aborts on CUDA errors (ORT_ABORT_ON_ERROR convention). `early`, if given, is called ON THE CALLING THREAD, WHILE THE ENGINE IS LOCKED
    // and the device still works on later time chunks, with [t_begin, t_end): the decisions of those steps of every item are final and
    // stored in the items (steps beyond an item's n_steps do not exist). Steps never announced are simply there when run_pass returns.
    virtual void run_pass(std::vector<PassItem>& items, const std::function<void(int, int)>* early = nullptr) = 0;
    virtual EngineStats stats(bool reset) = 0;
    virtual void set_tap(TapFn fn, void* userdata) = 0;
    virtual void set_timing(bool on) = 0;
    // test hook: log-mel rows of one standalone stream (frames [160 i, 160 i + N)), returns the row count
    virtual int debug_fbank(const int16_t* pcm, size_t n, float* out_rows, int max_rows) = 0;
    virtual int debug_gemm(int tier, int M, int N, int K, const float* A, const float* W, const float* bias, float* out) = 0;
    virtual int stride_ms() const = 0;
    virtual int device() const = 0;
};

std::vector<int> cuda_devices_sm100();     // CUDA ordinals of the usable sm_100 devices, in ordinal order
int host_pin(void* p, size_t bytes);       // cudaHostRegister (portable across the GPUs of the process) / cudaHostUnregister; 0 = ok
int host_unpin(void* p);

}  // namespace april
