// .april container + ONNX initializer extraction (host side, C++17).
//
// Replaces, for the engine, what the reference does in
//   src/file/model_file.c:58-185 (container), src/params.c:46-112 (params + vocabulary),
//   src/april_model.c:57-102 (graph I/O dims + sanity checks) and ONNXRuntime's model loading.
// The three ONNX graphs are not executed: their weights are pulled out by walking the graph
// topology (initializer names of a real export are not known — SURVEY.md Appendix B).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace april {

struct ModelDims {
    int L = 0;        // encoder layers
    int D = 0;        // d_model == LSTM projection size (h dim)
    int H = 0;        // LSTM hidden (c dim)
    int F = 0;        // feed-forward dim
    int J = 0;        // joiner dim (encoder_out / decoder_out)
    int E = 0;        // decoder embedding dim
    int V = 0;        // vocabulary size
    int CTX = 0;      // decoder context size
    int MEL = 0, SEG = 0, STEP = 0;
    int C1 = 0, C2 = 0, C3 = 0, F3 = 0;   // conv channels, final freq bins
};

struct Params {       // src/params.h:26-46
    int batch_size, segment_size, segment_step, mel_features, sample_rate;
    int frame_shift_ms, frame_length_ms, round_pow2, mel_low, mel_high, snip_edges;
    int token_count, blank_id;
    std::vector<std::string> tokens;
};

struct LayerWeights {
    std::vector<float> w_ih, b_ih, w_hh, b_hh, w_hr, ff1_w, ff1_b, ff2_w, ff2_b;
    float norm_eps = 0.f;
};

struct ModelWeights {
    ModelDims d;
    std::vector<float> conv1_w, conv1_b, conv2_w, conv2_b, conv3_w, conv3_b;
    std::vector<float> out_w, out_b;
    float out_norm_eps = 0.f;
    std::vector<LayerWeights> layers;
    std::vector<float> enc_proj_w, enc_proj_b;
    std::vector<float> emb, dec_conv_w, dec_proj_w, dec_proj_b;
    std::vector<float> join_w, join_b;
};

struct LoadedModel {
    std::string name, description, language;
    Params params;
    ModelWeights w;
};

// Returns false and fills `err` on any malformed / unsupported input.
bool load_april_file(const std::string& path, LoadedModel& out, std::string& err);

// Test hook (CPU only): order-independent-of-graph-structure fingerprint of what the loader extracted: out[0..14] = the model
// dims (L, D, H, F, J, E, V, CTX, MEL, SEG, STEP, C1, C2, C3, F3), out[15] = FNV-1a over every weight array in canonical order
// (bit patterns) and the norm eps values. Two files holding the same network in differently structured graphs give equal digests.
void model_digest(const LoadedModel& m, uint64_t out[16]);

}  // namespace april
