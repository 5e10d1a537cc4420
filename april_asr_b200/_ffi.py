"""ctypes view of the april C ABI (include/april_api.h, mirroring /root/reference/april_api.h:58-196)
plus this repo's additive batch entry points (include/april_b200.h).

The same loader is used by the tests against the parity oracle (oracle/_ref/libaprilasr_ref.so),
which exports the identical 12 reference symbols.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libaprilasr.so")


class AprilSpeakerID(ctypes.Structure):
    _fields_ = [("data", ctypes.c_uint8 * 16)]


class AprilToken(ctypes.Structure):                      # april_api.h:118-137, 32 bytes on LP64
    _fields_ = [("token", ctypes.c_char_p),
                ("logprob", ctypes.c_float),
                ("flags", ctypes.c_int),
                ("time_ms", ctypes.c_size_t),
                ("reserved", ctypes.c_void_p)]


HANDLER = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(AprilToken))


class AprilConfig(ctypes.Structure):                     # april_api.h:164-174, 40 bytes, passed BY VALUE
    _fields_ = [("speaker", AprilSpeakerID),
                ("handler", HANDLER),
                ("userdata", ctypes.c_void_p),
                ("flags", ctypes.c_int)]


TAP = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t)


class AprilB200Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("encoder_steps", "joiner_passes", "decoder_runs", "frames", "kernel_launches",
                                               "h2d_bytes", "d2h_bytes")] + \
               [(n, ctypes.c_double) for n in ("device_ms", "lstm_scan_ms", "gemm_ms", "compute_ms", "gemm_flops")] + [("gemm_launches", ctypes.c_uint64), ("greedy_ms", ctypes.c_double), ("passes", ctypes.c_uint64), ("pass_sessions", ctypes.c_uint64),
                  ("lstm_gemm_ms", ctypes.c_double), ("lstm_gemm_flops", ctypes.c_double), ("frontend_ms", ctypes.c_double)]


REFERENCE_SYMBOLS = [
    "aam_api_init", "aam_create_model", "aam_get_name", "aam_get_description", "aam_get_language",
    "aam_get_sample_rate", "aam_free", "aas_create_session", "aas_feed_pcm16", "aas_flush",
    "aas_realtime_get_speedup", "aas_free",
]
B200_SYMBOLS = [
    "aab_feed_pcm16_batch", "aab_flush_batch", "aab_device_count", "aab_set_device", "aab_session_device", "aab_pin_host_memory", "aab_unpin_host_memory", "aab_get_stats",
    "aab_set_tap", "aab_engine_version", "aab_debug_fbank", "aab_set_timing", "aab_counting_handler", "aab_debug_gemm", "aab_debug_replay", "aab_debug_plan", "aab_debug_model_digest",
]


class Lib:
    """Loads a library exporting the april C ABI and declares prototypes."""

    def __init__(self, path=None):
        self.path = path or os.environ.get("APRIL_LIB", DEFAULT_LIB)
        if not os.path.exists(self.path):
            raise OSError("april library not found: %s (run `python -c 'import __graft_entry__ as g; g.build()'`)" % self.path)
        self.c = c = ctypes.CDLL(self.path)
        vp, cp, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t
        c.aam_api_init.argtypes, c.aam_api_init.restype = [ctypes.c_int], None
        c.aam_create_model.argtypes, c.aam_create_model.restype = [cp], vp
        for n in ("aam_get_name", "aam_get_description", "aam_get_language"):
            getattr(c, n).argtypes, getattr(c, n).restype = [vp], cp
        c.aam_get_sample_rate.argtypes, c.aam_get_sample_rate.restype = [vp], sz
        c.aam_free.argtypes, c.aam_free.restype = [vp], None
        c.aas_create_session.argtypes, c.aas_create_session.restype = [vp, AprilConfig], vp
        c.aas_feed_pcm16.argtypes, c.aas_feed_pcm16.restype = [vp, ctypes.POINTER(ctypes.c_short), sz], None
        c.aas_flush.argtypes, c.aas_flush.restype = [vp], None
        c.aas_realtime_get_speedup.argtypes, c.aas_realtime_get_speedup.restype = [vp], ctypes.c_float
        c.aas_free.argtypes, c.aas_free.restype = [vp], None
        self.has_batch = hasattr(c, "aab_feed_pcm16_batch")
        if self.has_batch:
            c.aab_feed_pcm16_batch.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(ctypes.POINTER(ctypes.c_short)),
                                               ctypes.POINTER(sz), sz]
            c.aab_feed_pcm16_batch.restype = ctypes.c_int
            c.aab_flush_batch.argtypes, c.aab_flush_batch.restype = [ctypes.POINTER(vp), sz], ctypes.c_int
            c.aab_device_count.argtypes, c.aab_device_count.restype = [], ctypes.c_int
            c.aab_set_device.argtypes, c.aab_set_device.restype = [ctypes.c_int], ctypes.c_int
            c.aab_session_device.argtypes, c.aab_session_device.restype = [ctypes.c_void_p], ctypes.c_int
            c.aab_pin_host_memory.argtypes, c.aab_pin_host_memory.restype = [ctypes.c_void_p, ctypes.c_size_t], ctypes.c_int
            c.aab_unpin_host_memory.argtypes, c.aab_unpin_host_memory.restype = [ctypes.c_void_p], ctypes.c_int
            c.aab_engine_version.argtypes, c.aab_engine_version.restype = [], cp
            c.aab_debug_fbank.argtypes = [vp, ctypes.POINTER(ctypes.c_short), sz, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
            c.aab_debug_fbank.restype = ctypes.c_int
            c.aab_set_tap.argtypes, c.aab_set_tap.restype = [vp, TAP, vp], None
            fp = ctypes.POINTER(ctypes.c_float)
            c.aab_debug_gemm.argtypes, c.aab_debug_gemm.restype = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, fp, fp], ctypes.c_int
            c.aab_debug_replay.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), fp, fp, ctypes.c_int, HANDLER, vp]
            c.aab_debug_replay.restype = ctypes.c_int
            c.aab_debug_plan.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(sz), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
            c.aab_debug_plan.restype = ctypes.c_int
            c.aab_set_timing.argtypes, c.aab_set_timing.restype = [vp, ctypes.c_int], None
            c.aab_debug_model_digest.argtypes, c.aab_debug_model_digest.restype = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64)], ctypes.c_int
            c.aab_get_stats.argtypes, c.aab_get_stats.restype = [vp, ctypes.POINTER(AprilB200Stats), ctypes.c_int], ctypes.c_int
        c.aam_api_init(1)
