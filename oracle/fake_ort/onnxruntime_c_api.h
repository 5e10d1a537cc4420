/*
 * TEST INFRASTRUCTURE — stand-in for ONNXRuntime's <onnxruntime_c_api.h>.
 *
 * ONNXRuntime 1.13.1 (the third-party dependency that holds the reference's net arithmetic,
 * pinned by /root/reference/download_onnx_linux_x64.sh:8) is not in this image and cannot be
 * fetched. This header declares exactly the slice of the ORT C API that the reference's sources
 * touch (SURVEY.md §8c lists the 23 OrtApi entry points and their call sites), so that
 * /root/reference/src/{init,april_model,april_session,ort_util}.c compile UNMODIFIED. The
 * implementation behind it (fake_ort.c) is a small fp32 CPU interpreter of the ONNX graphs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load anything built from this directory. The product (libaprilasr.so) never does.
 */
#ifndef FAKE_ONNXRUNTIME_C_API_H
#define FAKE_ONNXRUNTIME_C_API_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>   /* the reference gets calloc/abort/getenv transitively from the real header */
#include <string.h>   /* ... and memmove/strcmp */

#ifdef __cplusplus
extern "C" {
#endif

#define ORT_API_VERSION 13

typedef struct OrtEnv OrtEnv;
typedef struct OrtStatus OrtStatus;
typedef struct OrtSession OrtSession;
typedef struct OrtSessionOptions OrtSessionOptions;
typedef struct OrtRunOptions OrtRunOptions;
typedef struct OrtValue OrtValue;
typedef struct OrtMemoryInfo OrtMemoryInfo;
typedef struct OrtTypeInfo OrtTypeInfo;
typedef struct OrtTensorTypeAndShapeInfo OrtTensorTypeAndShapeInfo;

typedef enum OrtLoggingLevel {
    ORT_LOGGING_LEVEL_VERBOSE, ORT_LOGGING_LEVEL_INFO, ORT_LOGGING_LEVEL_WARNING,
    ORT_LOGGING_LEVEL_ERROR, ORT_LOGGING_LEVEL_FATAL
} OrtLoggingLevel;

typedef enum OrtAllocatorType { OrtInvalidAllocator = -1, OrtDeviceAllocator = 0, OrtArenaAllocator = 1 } OrtAllocatorType;
typedef enum OrtMemType { OrtMemTypeCPUInput = -2, OrtMemTypeCPUOutput = -1, OrtMemTypeCPU = OrtMemTypeCPUOutput, OrtMemTypeDefault = 0 } OrtMemType;

typedef enum ONNXTensorElementDataType {
    ONNX_TENSOR_ELEMENT_DATA_TYPE_UNDEFINED = 0,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT = 1,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_INT32 = 6,
    ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64 = 7
} ONNXTensorElementDataType;

typedef struct OrtApi {
    OrtStatus* (*CreateEnv)(OrtLoggingLevel level, const char* logid, OrtEnv** out);
    OrtStatus* (*CreateSessionOptions)(OrtSessionOptions** out);
    OrtStatus* (*SetIntraOpNumThreads)(OrtSessionOptions* o, int n);
    OrtStatus* (*SetInterOpNumThreads)(OrtSessionOptions* o, int n);
    OrtStatus* (*CreateSessionFromArray)(const OrtEnv* env, const void* data, size_t len,
                                         const OrtSessionOptions* o, OrtSession** out);
    OrtStatus* (*SessionGetInputCount)(const OrtSession* s, size_t* out);
    OrtStatus* (*SessionGetOutputCount)(const OrtSession* s, size_t* out);
    OrtStatus* (*SessionGetInputTypeInfo)(const OrtSession* s, size_t idx, OrtTypeInfo** out);
    OrtStatus* (*SessionGetOutputTypeInfo)(const OrtSession* s, size_t idx, OrtTypeInfo** out);
    OrtStatus* (*CastTypeInfoToTensorInfo)(const OrtTypeInfo* ti, const OrtTensorTypeAndShapeInfo** out);
    OrtStatus* (*GetDimensionsCount)(const OrtTensorTypeAndShapeInfo* i, size_t* out);
    OrtStatus* (*GetDimensions)(const OrtTensorTypeAndShapeInfo* i, int64_t* dims, size_t n);
    void (*ReleaseTypeInfo)(OrtTypeInfo* ti);
    OrtStatus* (*CreateCpuMemoryInfo)(OrtAllocatorType a, OrtMemType m, OrtMemoryInfo** out);
    OrtStatus* (*CreateTensorWithDataAsOrtValue)(const OrtMemoryInfo* mi, void* data, size_t bytes,
                                                 const int64_t* shape, size_t ndim,
                                                 ONNXTensorElementDataType type, OrtValue** out);
    OrtStatus* (*Run)(OrtSession* s, const OrtRunOptions* ro,
                      const char* const* in_names, const OrtValue* const* in, size_t n_in,
                      const char* const* out_names, size_t n_out, OrtValue** out);
    const char* (*GetErrorMessage)(const OrtStatus* st);
    void (*ReleaseStatus)(OrtStatus* st);
    void (*ReleaseValue)(OrtValue* v);
    void (*ReleaseMemoryInfo)(OrtMemoryInfo* mi);
    void (*ReleaseSession)(OrtSession* s);
    void (*ReleaseSessionOptions)(OrtSessionOptions* o);
    void (*ReleaseEnv)(OrtEnv* e);
} OrtApi;

typedef struct OrtApiBase {
    const OrtApi* (*GetApi)(uint32_t version);
    const char* (*GetVersionString)(void);
} OrtApiBase;

const OrtApiBase* OrtGetApiBase(void);

/* ---- oracle-only extensions (not part of ORT) ------------------------------------------- */
/* Tap: called after every Run with the first output of the graph. `net` is the creation index of
 * the OrtSession inside its env (0 encoder, 1 decoder, 2 joiner — april_model.c:57-59). */
typedef void (*FakeOrtTap)(void* userdata, int net, const float* out0, size_t n);
void fake_ort_set_tap(FakeOrtTap tap, void* userdata);
/* Counters of Run calls per net since process start (index as above). */
void fake_ort_get_run_counts(uint64_t counts[3]);

#ifdef __cplusplus
}
#endif
#endif
