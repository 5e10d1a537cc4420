/*
 * april_api.h — public C ABI of libaprilasr.so (B200-native engine).
 *
 * This header re-declares, with identical names, argument meaning, struct layout and enum values,
 * the public interface of abb128/april-asr so that existing callers (example.cpp, the
 * Python/C#/Java bindings, downstream apps) link and run unmodified against this library:
 *
 *   entry point                      replaces reference declaration     reference implementation
 *   aam_api_init                     april_api.h:58                     src/init.c:33-50
 *   aam_create_model                 april_api.h:61                     src/april_model.c:24-107
 *   aam_get_name/description/lang    april_api.h:65-67                  src/april_model.c:109-111
 *   aam_get_sample_rate              april_api.h:70                     src/april_model.c:113-115
 *   aam_free                         april_api.h:74                     src/april_model.c:118-134
 *   aas_create_session               april_api.h:178                    src/april_session.c:25-93
 *   aas_feed_pcm16                   april_api.h:183                    src/april_session.c:479-538
 *   aas_flush                        april_api.h:186                    src/april_session.c:541-564
 *   aas_realtime_get_speedup         april_api.h:192                    src/april_session.c:95-97
 *   aas_free                         april_api.h:196                    src/april_session.c:99-119
 *
 * Behavioural contract kept from the reference: constructors return NULL on failure and log to
 * stderr (level from env APRIL_LOG_LEVEL); a synchronous session invokes the handler on the
 * caller's thread before aas_feed_pcm16 / aas_flush return; token strings live as long as the
 * model; the AprilToken array is valid only during the callback.
 * Differences: the forward path runs on an NVIDIA B200 (sm_100a) — there is no CPU fallback, and
 * model/session creation fails loudly if no such device is usable.
 */
#ifndef APRIL_API_H_B200
#define APRIL_API_H_B200

#include <stddef.h>
#include <stdint.h>

#if defined(_WIN32)
#define APRIL_EXPORT __declspec(dllexport)
#else
#define APRIL_EXPORT __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

struct AprilASRModel_i;
struct AprilASRSession_i;
typedef struct AprilASRModel_i* AprilASRModel;
typedef struct AprilASRSession_i* AprilASRSession;

#define APRIL_VERSION 1

/* Call once before anything else, with APRIL_VERSION. */
APRIL_EXPORT void aam_api_init(int version);

/* Load a .april file (extra/file-format.md). NULL on failure. */
APRIL_EXPORT AprilASRModel aam_create_model(const char* model_path);

/* Metadata strings; valid until aam_free(model). */
APRIL_EXPORT const char* aam_get_name(AprilASRModel model);
APRIL_EXPORT const char* aam_get_description(AprilASRModel model);
APRIL_EXPORT const char* aam_get_language(AprilASRModel model);

/* Sample rate the model expects, in Hz. */
APRIL_EXPORT size_t aam_get_sample_rate(AprilASRModel model);

/* Free every session of the model first. NULL-safe. */
APRIL_EXPORT void aam_free(AprilASRModel model);

/* Speaker identifier; accepted and ignored (as in the reference). */
typedef struct AprilSpeakerID {
    uint8_t data[16];
} AprilSpeakerID;

typedef enum AprilResultType {
    APRIL_RESULT_UNKNOWN = 0,
    APRIL_RESULT_RECOGNITION_PARTIAL = 1, /* text so far; a later call repeats it, updated */
    APRIL_RESULT_RECOGNITION_FINAL = 2,   /* text is final; later calls start from empty */
    APRIL_RESULT_ERROR_CANT_KEEP_UP = 3,  /* async only; count = 0, tokens = NULL */
    APRIL_RESULT_SILENCE = 4              /* emitted once per silence; count = 0, tokens = NULL */
} AprilResultType;

typedef enum AprilTokenFlagBits {
    APRIL_TOKEN_FLAG_WORD_BOUNDARY_BIT = 0x00000001, /* token starts a new word (token[0] == ' ') */
    APRIL_TOKEN_FLAG_SENTENCE_END_BIT = 0x00000002   /* token is ".", "!" or "?" ending a sentence */
} AprilTokenFlagBits;

typedef struct AprilToken {
    const char* token;        /* NUL-terminated piece, owned by the model */
    float logprob;            /* raw joiner logit of the piece (the reference applies no softmax) */
    AprilTokenFlagBits flags;
    size_t time_ms;           /* audio time at which the piece was emitted */
    void* reserved;
} AprilToken;

/* (userdata, result, count, tokens) — tokens valid only during the call. */
typedef void (*AprilRecognitionResultHandler)(void*, AprilResultType, size_t, const AprilToken*);

typedef enum AprilConfigFlagBits {
    APRIL_CONFIG_FLAG_ZERO_BIT = 0x00000000,
    /* Feed/flush return immediately; the handler runs on an engine thread. On a B200 the engine
       never needs to time-stretch audio, so this behaves like ASYNC_NO_RT and
       aas_realtime_get_speedup() reports 1.0. */
    APRIL_CONFIG_FLAG_ASYNC_RT_BIT = 0x00000001,
    /* As above; if the 3 s input ring overflows the chunk is dropped and the handler gets
       APRIL_RESULT_ERROR_CANT_KEEP_UP. */
    APRIL_CONFIG_FLAG_ASYNC_NO_RT_BIT = 0x00000002
} AprilConfigFlagBits;

typedef struct AprilConfig {
    AprilSpeakerID speaker;
    AprilRecognitionResultHandler handler; /* required */
    void* userdata;
    AprilConfigFlagBits flags;
} AprilConfig;

/* Many sessions may share one model. NULL if the handler is NULL. */
APRIL_EXPORT AprilASRSession aas_create_session(AprilASRModel model, AprilConfig config);

/* Mono PCM16 at aam_get_sample_rate(); short_count counts shorts, not bytes. */
APRIL_EXPORT void aas_feed_pcm16(AprilASRSession session, short* pcm16, size_t short_count);

/* Process what is buffered and produce a final result. */
APRIL_EXPORT void aas_flush(AprilASRSession session);

/* Always 1.0 (see APRIL_CONFIG_FLAG_ASYNC_RT_BIT). */
APRIL_EXPORT float aas_realtime_get_speedup(AprilASRSession session);

/* NULL-safe. */
APRIL_EXPORT void aas_free(AprilASRSession session);

#ifdef __cplusplus
}
#endif
#endif
