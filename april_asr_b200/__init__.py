"""april_asr_b200 — Python mirror of the reference's Python binding
(/root/reference/bindings/python/april_asr/_april.py: Result, Token, Model, Session) over this
repo's B200-native `libaprilasr.so`, plus the additive batched entry points.

PyTorch is not needed to use the library; the engine owns its CUDA context.
"""
import ctypes
from enum import IntEnum

from ._ffi import HANDLER, AprilConfig, AprilToken, Lib  # noqa: F401

_default_lib = None


def default_lib():
    global _default_lib
    if _default_lib is None:
        _default_lib = Lib()
    return _default_lib


class Result(IntEnum):                     # april_api.h:86-106
    UNKNOWN = 0
    PARTIAL_RECOGNITION = 1
    FINAL_RECOGNITION = 2
    ERROR_CANT_KEEP_UP = 3
    SILENCE = 4


class Token:
    """One decoded piece (april_api.h:118-137)."""
    __slots__ = ("token", "logprob", "word_boundary", "sentence_end", "time", "flags", "time_ms")

    def __init__(self, t):
        self.token = t.token.decode("utf-8")
        self.logprob = float(t.logprob)
        self.flags = int(t.flags)
        self.word_boundary = bool(t.flags & 1)
        self.sentence_end = bool(t.flags & 2)
        self.time_ms = int(t.time_ms)
        self.time = self.time_ms / 1000.0

    def key(self):
        return (self.token, self.flags, self.time_ms)


class Model:
    def __init__(self, path, lib=None):
        self.lib = lib or default_lib()
        self.handle = self.lib.c.aam_create_model(path.encode("utf-8"))
        if not self.handle:
            raise RuntimeError("Failed to load model " + path)

    def get_name(self):
        return self.lib.c.aam_get_name(self.handle).decode("utf-8")

    def get_description(self):
        return self.lib.c.aam_get_description(self.handle).decode("utf-8")

    def get_language(self):
        return self.lib.c.aam_get_language(self.handle).decode("utf-8")

    def get_sample_rate(self):
        return int(self.lib.c.aam_get_sample_rate(self.handle))

    def close(self):
        if self.handle:
            self.lib.c.aam_free(self.handle)
            self.handle = None


class Session:
    """callback(result: Result, tokens: list[Token])"""

    def __init__(self, model, callback, asynchronous=False, no_rt=False):
        self.model = model
        self.lib = model.lib
        self.callback = callback

        def _tramp(userdata, rtype, count, tokens):
            self.callback(Result(rtype), [Token(tokens[i]) for i in range(count)])

        self._handler = HANDLER(_tramp)                  # keep alive
        cfg = AprilConfig()
        cfg.handler = self._handler
        cfg.userdata = None
        cfg.flags = (2 if no_rt else 1) if asynchronous else 0
        self.handle = self.lib.c.aas_create_session(model.handle, cfg)
        if not self.handle:
            raise RuntimeError("Failed to create session")

    def get_rt_speedup(self):
        return float(self.lib.c.aas_realtime_get_speedup(self.handle))

    def feed_pcm16(self, data):
        """data: bytes / bytearray / numpy int16 array"""
        if hasattr(data, "ctypes"):
            ptr, n = data.ctypes.data_as(ctypes.POINTER(ctypes.c_short)), data.size
        else:
            buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
            ptr, n = ctypes.cast(buf, ctypes.POINTER(ctypes.c_short)), len(data) // 2
        self.lib.c.aas_feed_pcm16(self.handle, ptr, n)

    def flush(self):
        self.lib.c.aas_flush(self.handle)

    def close(self):
        if self.handle:
            self.lib.c.aas_free(self.handle)
            self.handle = None


def feed_pcm16_batch(sessions, arrays):
    """Additive: one batched engine pass over many sessions (include/april_b200.h: aab_feed_pcm16_batch).
    arrays: list of C-contiguous numpy int16 arrays (one per session; may be ragged)."""
    lib = sessions[0].lib
    n = len(sessions)
    hs = (ctypes.c_void_p * n)(*[s.handle for s in sessions])
    ps = (ctypes.POINTER(ctypes.c_short) * n)(*[a.ctypes.data_as(ctypes.POINTER(ctypes.c_short)) for a in arrays])
    cs = (ctypes.c_size_t * n)(*[a.size for a in arrays])
    rc = lib.c.aab_feed_pcm16_batch(hs, ps, cs, n)
    if rc != 0:
        raise RuntimeError("aab_feed_pcm16_batch failed: %d" % rc)


def flush_batch(sessions):
    lib = sessions[0].lib
    n = len(sessions)
    hs = (ctypes.c_void_p * n)(*[s.handle for s in sessions])
    rc = lib.c.aab_flush_batch(hs, n)
    if rc != 0:
        raise RuntimeError("aab_flush_batch failed: %d" % rc)
