// Host half of the session runtime: result shaping and callbacks.
// Re-implements the caller-visible behaviour of /root/reference/src/april_session.c:199-429
// (aas_finalize_tokens, aas_finalize_previous_words, aas_emit_silence, aas_emit_token,
// aas_clear_context, aas_process_logits) driven by per-pass decision summaries from the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/april_api.h"

namespace april {

constexpr int kMaxActiveTokens = 72;      // src/april_session.h:30

struct Vocab {
    std::vector<std::string> tokens;      // storage the AprilToken::token pointers refer to (lives with the model)
    int blank_id = 0;
};

class TokenShaper {
public:
    TokenShaper(const Vocab* vocab, AprilRecognitionResultHandler handler, void* userdata);

    // S1 step prologue: current_time_ms += stride (april_session.c:442-443)
    void begin_step(size_t stride_ms) { current_time_ms_ += stride_ms; }

    // One joiner pass (april_session.c:306-429). Returns is_blank as the reference computes it from
    // (max_idx, max_val, blank_val) and the session's own state.
    bool on_joiner_pass(float early_emit, int max_idx, float max_val, float blank_val);

    // Tail of _aas_flush (april_session.c:561-563).
    void on_flush_end();

    long long context(int i) const { return ctx_[i]; }
    size_t current_time_ms() const { return current_time_ms_; }

private:
    void finalize_tokens();
    void finalize_previous_words(const AprilToken& new_token);
    void emit_silence();
    bool emit_token(const AprilToken* new_token, bool force);
    void update_context(long long token);
    void clear_context();

    const Vocab* vocab_;
    AprilRecognitionResultHandler handler_;
    void* userdata_;
    long long ctx_[2];
    AprilToken active_[kMaxActiveTokens];
    size_t head_ = 0;
    size_t last_handler_call_head_ = 0;
    bool emitted_silence_ = true;          // april_session.c:64
    size_t current_time_ms_ = 0, last_emission_time_ms_ = 0;
};

}  // namespace april
