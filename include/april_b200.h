/*
 * april_b200.h — additive C ABI of the B200 engine (not part of the reference's april_api.h).
 *
 * The reference processes one session per call (src/april_session.c:479-564, batch hard-wired to
 * 1: src/params.c:71, src/april_session.c:51-55). The GPU engine batches the per-frame hot loop
 * (fbank -> encoder -> decoder/joiner -> greedy search) over many sessions; these entry points
 * expose that batching for SYNCHRONOUS sessions without changing any reference type: the session
 * handles, AprilConfig, the handler and its calling convention are exactly april_api.h's.
 * Semantics: what calling aas_feed_pcm16 / aas_flush on each session would do — every session's
 * handler runs on the caller's thread before the call returns and sees that session's callbacks
 * in stream order (the order ACROSS sessions is described at aab_feed_pcm16_batch) — but all
 * sessions advance through ONE batched device pass. Plain pointers and sizes only.
 *
 * Reference-side binding a maintainer would add: see INTEGRATION.md.
 */
#ifndef APRIL_B200_H
#define APRIL_B200_H

#include "april_api.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Batched aas_feed_pcm16 (replaces N calls of src/april_session.c:479). All sessions must be
 * synchronous sessions of the same model. pcm16[i] may be NULL iff short_count[i] == 0.
 * Returns 0, or -1 on invalid arguments (nothing processed). Device errors abort(), as ORT errors do
 * in the reference (src/ort_util.h:29-38).
 * Handlers run on the calling thread before the call returns, each session's callbacks in stream order. Across the sessions of the
 * call the order is: time chunk by time chunk (big passes are computed in chunks and the callbacks of a finished chunk are delivered
 * while the device works on the next), session by session within a chunk. Handlers invoked from the two batch entry points must not
 * call back into the library (the engine is locked while they run); the reference's own entry points have no such restriction. */
APRIL_EXPORT int aab_feed_pcm16_batch(AprilASRSession* sessions, short** pcm16,
                                      const size_t* short_count, size_t n_sessions);

/* Batched aas_flush (replaces N calls of src/april_session.c:541). */
APRIL_EXPORT int aab_flush_batch(AprilASRSession* sessions, size_t n_sessions);

/* Number of usable sm_100 devices; 0 means the library cannot run (there is no CPU fallback). */
APRIL_EXPORT int aab_device_count(void);

/* Device placement (the reference has no device notion, src/april_session.c:25-93). By default
 * aam_create_model builds one engine per visible sm_100 device (weights replicated) and aas_create_session places each new
 * session on the least-loaded one: a drop-in C caller on an 8-GPU box uses all eight. One process per GPU (torchrun) is the
 * other supported layout: LOCAL_RANK or APRIL_B200_DEVICE pins every model of the process to that B200 (index among the
 * sm_100 devices, mapped to the CUDA ordinal internally), and aab_set_device(i) pins the models created afterwards on the
 * calling thread; aab_set_device(-1) returns to the default. APRIL_B200_MAX_DEVICES caps the number of engines. */
APRIL_EXPORT int aab_set_device(int device);
/* CUDA ordinal of the GPU a session was placed on. */
APRIL_EXPORT int aab_session_device(AprilASRSession session);

/* Optional: page-lock a long-lived audio buffer of the caller (a capture ring, a decoded file) so that batched feeds are DMA'd
 * straight out of it. Without it the library copies every feed into its own pinned staging buffer first (the largest host cost of
 * a big pass). Any memory that is already page-locked for CUDA (cudaHostAlloc / cudaHostRegister) is recognised too. Returns 0 on
 * success. The range must stay registered while a feed call that reads from it is in progress; unpin before freeing it. */
APRIL_EXPORT int aab_pin_host_memory(void* ptr, size_t bytes);
APRIL_EXPORT int aab_unpin_host_memory(void* ptr);

typedef struct AprilB200Stats {
    uint64_t encoder_steps;     /* session-steps run through the encoder */
    uint64_t joiner_passes;     /* joiner evaluations (1..3 per session-step) */
    uint64_t decoder_runs;      /* decoder evaluations on device */
    uint64_t frames;            /* fbank frames computed */
    uint64_t kernel_launches;   /* CUDA kernels launched by the engine */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    double   device_ms;         /* CUDA-event time of all engine passes */
    double   lstm_scan_ms;      /* ... of which the recurrent scan kernels */
    double   gemm_ms;           /* ... of which the time-batched GEMMs */
    double   compute_ms;        /* device_ms minus the host->device upload at the head of each pass */
    double   gemm_flops;        /* algorithmic FLOPs (2 M N K) of the GEMM launches timed in gemm_ms */
    uint64_t gemm_launches;
    double   greedy_ms;         /* ... of which the joiner + greedy search kernel */
    uint64_t passes;            /* batched device passes run */
    uint64_t pass_sessions;     /* sessions summed over those passes (pass_sessions / passes = mean batch) */
    double   lstm_gemm_ms;      /* part of gemm_ms spent in the W_ih and W_hr (layer output) GEMMs */
    double   lstm_gemm_flops;   /* their algorithmic FLOPs */
    double   frontend_ms;       /* fbank kernel + feature-row carry */
} AprilB200Stats;

/* Cumulative statistics of a model's engine; reset != 0 zeroes them after reading. lstm_scan_ms / gemm_ms /
 * gemm_flops are only collected while per-kernel-class CUDA-event timing is on (aab_set_timing). */
APRIL_EXPORT int aab_get_stats(AprilASRModel model, AprilB200Stats* out, int reset);
APRIL_EXPORT void aab_set_timing(AprilASRModel model, int enabled);

/* Debug/parity tap: after every engine pass, called once per session-step with the encoder output
 * (net = 0, n = joiner dim) and once per joiner pass with the raw logits (net = 2, n = vocab) —
 * the same quantities the oracle's fake_ort_set_tap reports for the reference's OrtRun calls
 * (src/april_session.c:145,176). Costs extra device->host traffic; NULL disables. */
typedef void (*AprilB200Tap)(void* userdata, AprilASRSession session, int net, const float* data, size_t n);
APRIL_EXPORT void aab_set_tap(AprilASRModel model, AprilB200Tap tap, void* userdata);

/* A handler (AprilRecognitionResultHandler signature) that only counts: userdata -> uint64_t[4]
 * {partial calls, final calls, silence calls, tokens delivered in final results}. */
APRIL_EXPORT void aab_counting_handler(void* userdata, AprilResultType type, size_t count, const AprilToken* tokens);

/* Test hook: run only the front-end kernel (F1/F2, src/fbank.c:174-306) on one standalone stream and
 * return its log-mel rows ([rows][mel_features] floats). Returns the number of rows written. */
APRIL_EXPORT int aab_debug_fbank(AprilASRModel model, const short* pcm16, size_t short_count, float* out_rows, int max_rows);

/* Test hook: out[M,N] = A[M,K] * W[N,K]^T + bias[N] through one GEMM tier of the engine
 * (tier 0 = fp32 SIMT, 1 = fp16 tcgen05, 3 = fp16x3 split tcgen05). Host pointers. */
APRIL_EXPORT int aab_debug_gemm(AprilASRModel model, int tier, int M, int N, int K, const float* A, const float* W, const float* bias, float* out);

/* Test hook, CPU only (no device needed): the host half of the greedy search and the result shaping
 * (april_session.c:199-429 as re-implemented in session.cpp) driven by a recorded stream of joiner-pass summaries.
 * Reads only the vocabulary / parameters of `model_path`. For each of n_steps encoder steps it consumes up to three
 * records (max_idx = arg-max over non-blank ids, max_val, blank_val) exactly as aas_infer does (early_emit 1, 0, 0; stop at
 * the first blank), calling `handler` like a session would; `flush` != 0 appends the tail of aas_flush. Returns the number
 * of records consumed, or -1. */
APRIL_EXPORT int aab_debug_replay(const char* model_path, int n_steps, int n_records, const int* max_idx, const float* max_val, const float* blank_val,
                                  int flush, AprilRecognitionResultHandler handler, void* userdata);

/* Test hook, CPU only: the pass planner (plan.h: frame counts, flush padding, encoder steps per call) for one stream fed
 * with n_calls calls; call i feeds counts[i] samples, or is an aas_flush when flush[i] != 0. Writes the encoder steps and
 * the frames computed per call. Returns 0, or -1 on a bad model file. */
APRIL_EXPORT int aab_debug_plan(const char* model_path, int n_calls, const size_t* counts, const int* flush, int* out_steps, int* out_frames);

/* Test hook, CPU only: loads a .april file with the product's loader (container, params, topology-driven weight extraction from the
 * three ONNX graphs: model_loader.cpp, replacing src/file/model_file.c:58-185 + ONNXRuntime's model loading) and writes a
 * fingerprint: out16[0..14] = the dims it found, out16[15] = a hash over every extracted weight array in canonical order.
 * Returns 0, or -1 when the loader rejects the file (nothing is written, nothing crashes). */
APRIL_EXPORT int aab_debug_model_digest(const char* model_path, uint64_t* out16);
APRIL_EXPORT const char* aab_engine_version(void);

#ifdef __cplusplus
}
#endif
#endif
