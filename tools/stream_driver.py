"""ctypes wrapper of tools/stream_driver.cpp (april_asr_b200/lib/libaprilstream.so): native threads driving synchronous sessions
through the reference entry points of ANY library exporting april_api.h. Bench/test tooling."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "april_asr_b200", "lib", "libaprilstream.so")


class AsdApi(ctypes.Structure):
    _fields_ = [("create_session", ctypes.c_void_p), ("feed_pcm16", ctypes.c_void_p), ("flush", ctypes.c_void_p), ("free_session", ctypes.c_void_p)]


_drv = None


def _driver():
    global _drv
    if _drv is None:
        d = ctypes.CDLL(SO)
        pp = ctypes.POINTER(ctypes.POINTER(ctypes.c_short))
        d.asd_stream.argtypes = [ctypes.POINTER(AsdApi), ctypes.c_void_p, ctypes.c_int, pp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong)]
        d.asd_stream.restype = ctypes.c_int
        d.asd_bulk.argtypes = [ctypes.POINTER(AsdApi), ctypes.c_void_p, ctypes.c_int, pp, ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
        d.asd_bulk.restype = ctypes.c_double
        _drv = d
    return _drv


def _api(lib):
    a = AsdApi()
    for f, n in (("create_session", "aas_create_session"), ("feed_pcm16", "aas_feed_pcm16"), ("flush", "aas_flush"), ("free_session", "aas_free")):
        setattr(a, f, ctypes.cast(getattr(lib.c, n), ctypes.c_void_p).value)
    return a


def _ptrs(pcms):
    return (ctypes.POINTER(ctypes.c_short) * len(pcms))(*[p.ctypes.data_as(ctypes.POINTER(ctypes.c_short)) for p in pcms])


COUNTER_NAMES = ("partial_callbacks", "final_callbacks", "silence_callbacks", "final_tokens", "error_callbacks")


def stream(lib, model_handle, pcms, chunk=640, warm_ticks=20, ticks=250, pace_us=0):
    """Returns (latencies_ms [ticks, n_sessions], counters dict)."""
    n = len(pcms)
    lat = np.zeros((ticks, n), np.float64)
    cnt = (ctypes.c_ulonglong * 5)()
    api = _api(lib)
    rc = _driver().asd_stream(ctypes.byref(api), model_handle, n, _ptrs(pcms), min(p.size for p in pcms), chunk, warm_ticks, ticks, pace_us,
                              lat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), cnt)
    if rc != 0:
        raise RuntimeError("asd_stream failed: %d" % rc)
    return lat, dict(zip(COUNTER_NAMES, [int(x) for x in cnt]))


def bulk(lib, model_handle, pcms, chunk=0):
    """Every session fed from its own native thread (whole clip, or `chunk` samples per call) and flushed. Returns (wall_ms, counters)."""
    n = len(pcms)
    cnt = (ctypes.c_ulonglong * 5)()
    sizes = (ctypes.c_size_t * n)(*[p.size for p in pcms])
    api = _api(lib)
    wall = _driver().asd_bulk(ctypes.byref(api), model_handle, n, _ptrs(pcms), sizes, chunk, cnt)
    if wall < 0:
        raise RuntimeError("asd_bulk failed: %g" % wall)
    return wall, dict(zip(COUNTER_NAMES, [int(x) for x in cnt]))
