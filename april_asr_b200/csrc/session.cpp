#include "session.h"

#include <cstring>

namespace april {

TokenShaper::TokenShaper(const Vocab* vocab, AprilRecognitionResultHandler handler, void* userdata)
    : vocab_(vocab), handler_(handler), userdata_(userdata) {
    std::memset(active_, 0, sizeof active_);                       // calloc'ed session, april_session.c:26
    // first aas_infer pushes blank x context_size (april_session.c:432-438)
    ctx_[0] = ctx_[1] = vocab->blank_id;
}

void TokenShaper::update_context(long long token) {               // april_session.c:181-196 (context_size == 2)
    ctx_[0] = ctx_[1];
    ctx_[1] = token;
}

void TokenShaper::clear_context() {                                // april_session.c:296-301
    if (ctx_[0] == vocab_->blank_id) return;
    update_context(vocab_->blank_id);
    update_context(vocab_->blank_id);
}

void TokenShaper::finalize_tokens() {                              // april_session.c:199-211
    if (head_ == 0) return;
    handler_(userdata_, APRIL_RESULT_RECOGNITION_FINAL, head_, active_);
    last_handler_call_head_ = head_;
    head_ = 0;
}

void TokenShaper::finalize_previous_words(const AprilToken& new_token) {   // april_session.c:213-255
    if (head_ == 0) return;
    if (new_token.flags & APRIL_TOKEN_FLAG_WORD_BOUNDARY_BIT) { finalize_tokens(); return; }
    // the new piece continues a word: keep that word active, finalize what precedes it
    size_t start_of_word = kMaxActiveTokens;
    for (size_t i = head_ - 1; i > 2; i--) {
        if (active_[i].flags & APRIL_TOKEN_FLAG_WORD_BOUNDARY_BIT) { start_of_word = i; break; }
    }
    if (start_of_word == (size_t)kMaxActiveTokens) { finalize_tokens(); return; }
    handler_(userdata_, APRIL_RESULT_RECOGNITION_FINAL, start_of_word, active_);
    std::memmove(active_, &active_[start_of_word], sizeof(AprilToken) * (head_ - start_of_word));
    head_ -= start_of_word;
}

void TokenShaper::emit_silence() {                                 // april_session.c:257-268
    if (emitted_silence_) return;
    emitted_silence_ = true;
    handler_(userdata_, APRIL_RESULT_SILENCE, 0, nullptr);
}

bool TokenShaper::emit_token(const AprilToken* new_token, bool force) {    // april_session.c:270-294
    if (new_token != nullptr) {
        if (!force && last_handler_call_head_ == head_ + 1 && active_[head_].token == new_token->token) return false;
        active_[head_++] = *new_token;
    } else {
        if (!force && last_handler_call_head_ == head_) return false;
    }
    handler_(userdata_, APRIL_RESULT_RECOGNITION_PARTIAL, head_, active_);
    last_handler_call_head_ = head_;
    return true;
}

bool TokenShaper::on_joiner_pass(float early_emit, int max_idx, float max_val, float blank_val) {
    const long long blank = vocab_->blank_id;
    const bool was_context_cleared = ctx_[1] == blank;                      // :322
    const bool is_equal_to_previous = ctx_[1] == (long long)max_idx;        // :326
    if (is_equal_to_previous) early_emit = 0.0f;
    bool is_blank = (blank_val - early_emit) > max_val;                      // :330

    static const char kEmpty[2] = { 0, 0 };
    const char* text = (max_idx >= 0 && max_idx < (int)vocab_->tokens.size()) ? vocab_->tokens[(size_t)max_idx].c_str() : kEmpty;
    AprilToken token;
    std::memset(&token, 0, sizeof token);
    token.token = text;
    token.logprob = max_val;                                                 // :333 (a raw logit)
    token.time_ms = current_time_ms_;
    int flags = 0;
    if (text[0] == ' ') flags |= APRIL_TOKEN_FLAG_WORD_BOUNDARY_BIT;         // :338
    const bool single = text[0] != 0 && text[1] == 0;
    bool is_end_of_sentence = single && (text[0] == '.' || text[0] == '!' || text[0] == '?');
    bool is_punctuation = is_end_of_sentence || (single && text[0] == ',');
    if (is_punctuation && head_ > 0) {                                       // "10.5": '.' after a digit piece, :345-351
        const char* last = active_[head_ - 1].token;
        if (last[0] >= '0' && last[0] <= '9' && text[0] == '.') { is_end_of_sentence = false; is_punctuation = false; }
    }
    if (is_end_of_sentence) flags |= APRIL_TOKEN_FLAG_SENTENCE_END_BIT;
    token.flags = (AprilTokenFlagBits)flags;
    if (!was_context_cleared && is_punctuation && !is_equal_to_previous && max_val > (blank_val - 3.5f)) is_blank = false;   // :356-358
    if (max_idx < 0) is_blank = true;        // no finite non-blank logit (non-finite activations): nothing can be emitted; the reference would index tokens[-1]

    if (!is_blank) {
        last_emission_time_ms_ = current_time_ms_;
        update_context((long long)max_idx);
        bool is_final = head_ >= (size_t)(kMaxActiveTokens - 1);
        if (head_ > 0 && (flags & APRIL_TOKEN_FLAG_WORD_BOUNDARY_BIT)) {      // :369-388
            AprilToken& last = active_[head_ - 1];
            const bool last_single = last.token[0] != 0 && last.token[1] == 0;
            const bool last_eos = last_single && (last.token[0] == '.' || last.token[0] == '!' || last.token[0] == '?');
            if (last_eos && !(last.flags & APRIL_TOKEN_FLAG_SENTENCE_END_BIT))
                last.flags = (AprilTokenFlagBits)(last.flags | APRIL_TOKEN_FLAG_SENTENCE_END_BIT);
            if (last_eos) is_final = true;
        }
        if (is_final) finalize_previous_words(token);
        if (head_ >= (size_t)(kMaxActiveTokens - 1)) head_ = 0;              // :392-396 "no room left"
        emit_token(&token, true);
        emitted_silence_ = false;
    } else {
        const size_t since = current_time_ms_ - last_emission_time_ms_;      // :402
        max_val -= (float)since / 3000.0f;                                   // :406
        const bool reasonably_confident = !is_equal_to_previous && max_val > (blank_val - 4.0f);
        if (since >= 2200) {                                                 // :411-416
            finalize_tokens();
            clear_context();
            emit_silence();
        } else if (reasonably_confident) {                                   // tentative piece: shown, not committed
            token.logprob -= 8.0f;
            if (emit_token(&token, false)) head_--;
        } else {
            emit_token(nullptr, false);
        }
    }
    return is_blank;
}

void TokenShaper::on_flush_end() {
    finalize_tokens();
    clear_context();
    emit_silence();
}

}  // namespace april
