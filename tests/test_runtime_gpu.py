"""GPU tests (-m gpu) of the session runtime around the engine: synchronous callers on many threads (coalesced into batched
passes), asynchronous sessions (per-GPU batcher), overload, and in-library placement on several GPUs. Every expectation is the
ORACLE's callback log for the same PCM (oracle/_ref = the reference's own host code), never the product's own output."""
import threading
import time

import numpy as np
import pytest

import april_asr_b200 as A
import common as C
from april_asr_b200 import _ffi

pytestmark = pytest.mark.gpu


def _oracle_many(model, pcms, threads=8, chunk=None):
    out = [None] * len(pcms)

    def work(k):
        for i in range(k, len(pcms), threads):
            out[i] = C.run_oracle(model, pcms[i], chunk=chunk)
    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    for t in ts: t.start()
    for t in ts: t.join()
    return out


def _wait_for(pred, timeout):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return True
        time.sleep(0.01)
    return pred()


def test_sync_sessions_on_many_threads_are_coalesced_and_match_the_oracle():
    """The reference's documented usage (april-docs/src/concepts.md:41-58): one model, many synchronous sessions, each driven
    from its own thread through plain aas_feed_pcm16 / aas_flush. Callbacks must equal the oracle's for every session, and the
    calls must have been merged into far fewer device passes than there were calls."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    n_sess, chunk = 48, 6400
    pcms = [C.gen_stream(300 + i, 16000 * 6 + 1234 * (i % 5)) for i in range(n_sess)]
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    cols = [C.Collected() for _ in pcms]
    sessions = [A.Session(m, c.cb) for c in cols]
    start = threading.Barrier(n_sess)
    calls = [0] * n_sess

    def drive(i):
        start.wait()
        for part in C.chunk_iter(pcms[i], chunk):
            sessions[i].feed_pcm16(np.ascontiguousarray(part)); calls[i] += 1
        sessions[i].flush(); calls[i] += 1
    st = _ffi.AprilB200Stats()
    lib.c.aab_get_stats(m.handle, st, 1)
    ts = [threading.Thread(target=drive, args=(i,)) for i in range(n_sess)]
    for t in ts: t.start()
    for t in ts: t.join()
    lib.c.aab_get_stats(m.handle, st, 1)
    want = _oracle_many(C.TINY_MODEL, pcms, chunk=chunk)
    bad = [i for i in range(n_sess) if cols[i].keys() != want[i].keys()]
    for s in sessions: s.close()
    m.close()
    assert not bad, "sessions whose callbacks differ from the oracle: %s" % bad
    print("sync coalescing: %d calls -> %d passes, mean batch %.1f" % (sum(calls), st.passes, st.pass_sessions / max(1, st.passes)))
    assert st.pass_sessions == sum(calls)
    assert st.passes < sum(calls)            # Python threads share the GIL: the strong merging figure is asserted with native threads below


def test_native_threads_through_the_reference_entry_points_are_merged_into_batched_passes():
    """64 native threads (tools/stream_driver.cpp: nothing but aas_create_session / aas_feed_pcm16 / aas_flush / aas_free), one
    synchronous session each, 0.4 s per call. The same driver binary drives the ORACLE library with the same PCM: the summed
    callback counters (partials, finals, silences, tokens in finals) must be equal, and the product must have merged the calls
    into passes of many sessions."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    import stream_driver as SD
    n_sess, chunk = 64, 6400
    pcms = [C.gen_stream(800 + i, 16000 * 5) for i in range(n_sess)]
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    st = _ffi.AprilB200Stats()
    lib.c.aab_get_stats(m.handle, st, 1)
    wall, got = SD.bulk(lib, m.handle, pcms, chunk)
    lib.c.aab_get_stats(m.handle, st, 1)
    m.close()
    olib = C.oracle_lib()
    om = A.Model(C.TINY_MODEL, olib)
    _, want = SD.bulk(olib, om.handle, pcms, chunk)
    om.close()
    calls = n_sess * (len(range(0, 16000 * 5, chunk)) + 1)
    print("native sync callers: %d calls -> %d passes (mean batch %.1f) in %.0f ms; counters %s" % (calls, st.passes, st.pass_sessions / max(1, st.passes), wall, got))
    assert got == want
    assert st.pass_sessions == calls
    assert st.pass_sessions >= 8 * st.passes, "calls were not merged: mean batch %.1f" % (st.pass_sessions / max(1, st.passes))


def test_single_sync_caller_pays_no_coalescing_wait():
    """One session, one thread: every call is its own pass and must not sit in a coalescing window."""
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    col = C.Collected()
    s = A.Session(m, col.cb)
    pcm = C.gen_stream(0, 96000)
    s.feed_pcm16(pcm[:3200])
    t0 = time.perf_counter()
    for i in range(3200, 96000, 3200):
        s.feed_pcm16(pcm[i:i + 3200])
    per_call = (time.perf_counter() - t0) / 29
    s.flush()
    g = C.load_golden("tiny_callbacks.json")["stream0"]
    assert col.keys() == [(r, [(t, f, ms) for t, f, ms, lp in toks]) for r, toks in g["log"]]
    s.close(); m.close()
    print("single sync caller: %.2f ms per 200 ms chunk" % (per_call * 1e3))
    assert per_call < 0.05


def _async_run(m, pcms, chunk, gap_s):
    cols = [C.Collected() for _ in pcms]
    ss = [A.Session(m, c.cb, asynchronous=True, no_rt=True) for c in cols]
    n = max(len(p) for p in pcms)
    for i in range(0, n, chunk):
        for s, p in zip(ss, pcms):
            if i < len(p):
                s.feed_pcm16(np.ascontiguousarray(p[i:i + chunk]))
        time.sleep(gap_s)
    return cols, ss


def test_async_sessions_match_the_oracle_log():
    """Async sessions (APRIL_CONFIG_FLAG_ASYNC_NO_RT): feed returns at once, the per-GPU batcher thread invokes the handlers.
    Features are chunking-invariant, so whatever batches the batcher forms the callbacks must equal the oracle's."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    pcms = [C.gen_stream(30 + i, 48000 + 800 * i) for i in range(6)]
    cols, ss = _async_run(m, pcms, 1600, 0.002)
    for s in ss: s.flush()
    want = _oracle_many(C.TINY_MODEL, pcms, threads=6)
    ok = _wait_for(lambda: all(len(c.log) >= len(w.log) for c, w in zip(cols, want)), 20.0)
    for s in ss: s.close()
    m.close()
    assert ok, "async callbacks did not arrive"
    for i, (c, w) in enumerate(zip(cols, want)):
        assert 3 not in [k[0] for k in c.keys()], "session %d could not keep up" % i
        assert c.keys() == w.keys(), "async session %d differs from the oracle" % i


def test_async_flush_is_served_while_other_sessions_keep_streaming():
    """A flush is decided per session (the reference has one thread per session, april_session.c:567-585): session A's
    final/silence callbacks must arrive although session B never stops feeding."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    a_pcm, b_pcm = C.gen_stream(60, 32000), C.gen_stream(61, 16000 * 8)
    ca, cb = C.Collected(), C.Collected()
    sa = A.Session(m, ca.cb, asynchronous=True, no_rt=True)
    sb = A.Session(m, cb.cb, asynchronous=True, no_rt=True)
    stop = threading.Event()

    def stream_b():
        i = 0
        while not stop.is_set() and i < len(b_pcm):
            sb.feed_pcm16(np.ascontiguousarray(b_pcm[i:i + 800])); i += 800
            time.sleep(0.001)
    tb = threading.Thread(target=stream_b); tb.start()
    for i in range(0, len(a_pcm), 1600):
        sa.feed_pcm16(np.ascontiguousarray(a_pcm[i:i + 1600])); time.sleep(0.001)
    sa.flush()
    want = C.run_oracle(C.TINY_MODEL, a_pcm)
    served = _wait_for(lambda: len(ca.log) >= len(want.log), 3.0)
    still_streaming = tb.is_alive()
    stop.set(); tb.join()
    sa.close(); sb.close(); m.close()
    assert served, "the flush of session A was starved (%d of %d callbacks)" % (len(ca.log), len(want.log))
    assert ca.keys() == want.keys()
    print("flush served while the other session was %s" % ("still streaming" if still_streaming else "already done"))


def test_async_ring_overflow_reports_cant_keep_up_and_recovers():
    """audio_provider.c:59-64 / april_session.c:482-492: a chunk that would overflow the 48 000-sample ring is dropped whole and the
    handler is called with ERROR_CANT_KEEP_UP (count 0) from the feeding thread; later audio is recognised normally."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    col = C.Collected()
    s = A.Session(m, col.cb, asynchronous=True, no_rt=True)
    s.feed_pcm16(C.gen_stream(50, 100000))                # cannot fit whatever the batcher has drained
    assert col.keys() == [(3, [])]                         # delivered synchronously, nothing else yet
    s.feed_pcm16(C.gen_stream(51, 48000))                  # exactly the ring size: the reference's test is >=, dropped too
    assert col.keys() == [(3, []), (3, [])]
    pcm = C.gen_stream(52, 40000)
    s.feed_pcm16(pcm)
    s.flush()
    want = C.run_oracle(C.TINY_MODEL, pcm)
    ok = _wait_for(lambda: len(col.log) >= 2 + len(want.log), 10.0)
    s.close(); m.close()
    assert ok
    assert col.keys() == [(3, []), (3, [])] + want.keys()


def test_sessions_are_placed_on_every_gpu_of_the_box():
    """one model, one process, all B200s: engines replicated per device, sessions placed on the least loaded."""
    lib = C.product_lib()
    n = lib.c.aab_device_count()
    if n < 2:
        pytest.skip("needs >= 2 B200s (run under gpurun --gpus 2)")
    import os
    if os.environ.get("APRIL_B200_DEVICE") or os.environ.get("LOCAL_RANK"):
        pytest.skip("process is pinned to one device")
    m = A.Model(C.TINY_MODEL, lib)
    pcms = [C.gen_stream(500 + i, 16000 * 4 + 321 * i) for i in range(4 * n)]
    cols = [C.Collected() for _ in pcms]
    ss = [A.Session(m, c.cb) for c in cols]
    devs = [lib.c.aab_session_device(s.handle) for s in ss]
    assert sorted(set(devs)) == list(range(n)) or len(set(devs)) == n
    assert max(devs.count(d) for d in set(devs)) == 4
    A.feed_pcm16_batch(ss, pcms)                          # one call, passes on all GPUs at once
    A.flush_batch(ss)
    want = _oracle_many(C.TINY_MODEL, pcms)
    for s in ss: s.close()
    m.close()
    for i, (c, w) in enumerate(zip(cols, want)):
        assert c.keys() == w.keys(), "session %d (cuda:%d) differs from the oracle" % (i, devs[i])


def test_damaged_model_files_never_crash_the_full_load_path(tmp_path):
    """The CPU suite feeds truncated and bit-damaged `.april` files to the parser alone; here the same damage goes through
    `aam_create_model` (src/april_model.c:24-107's role: parse, upload, fp16 splits, W_eff fold, decoder table). Truncations must
    give NULL; random damage may load (weights are just numbers) or not, but the call must return and a loaded model must be
    freeable. Run in a subprocess so that a crash is a failing return code."""
    import os, subprocess, sys
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import common as C
lib = C.product_lib()
good = open(C.TINY_MODEL, "rb").read()
out = sys.argv[1]
def load(data):
    p = os.path.join(out, "m.april"); open(p, "wb").write(data)
    h = lib.c.aam_create_model(p.encode())
    if h: lib.c.aam_free(h)
    return bool(h)
assert load(good)
n_bad = sum(not load(good[:cut]) for cut in [0, 7, 12, 64, 1000, len(good) // 3, len(good) // 2, len(good) - 1000, len(good) - 1])
rs = np.random.RandomState(11)
loaded = 0
for trial in range(60):
    b = bytearray(good)
    region = [(0, 64), (0, 4096), (0, len(good)), (len(good) - 20000, len(good)), (len(good) // 2 - 4096, len(good) // 2 + 4096)][trial %% 5]
    for _ in range(1 + trial %% 7):
        b[rs.randint(region[0], region[1])] = rs.randint(0, 256)
    loaded += load(bytes(b))
print("truncations rejected:", n_bad, "damaged files that still loaded:", loaded)
assert n_bad >= 8
assert load(good)                       # and the library is still usable afterwards
''' % (C.ROOT, os.path.join(C.ROOT, "tests"), os.path.join(C.ROOT, "tools"))
    r = subprocess.run([sys.executable, "-c", code, str(tmp_path)], capture_output=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-1200:])


def test_feeding_from_caller_pinned_buffers_gives_the_same_callbacks():
    """aab_pin_host_memory (include/april_b200.h): big batched feeds are DMA'd straight from page-locked caller memory instead of
    being copied into the library's staging buffer first. Both upload paths must deliver the same samples — including the
    carried-over tail between two feeds, which the pinned path rebuilds from the caller's buffer — so the callback logs must be
    identical, and equal to the oracle's. 32 sessions x ~25 s in two feeds (each above the 8 MB threshold of the block-wise upload)."""
    import ctypes
    path = C.model_path("mid")
    lib = C.product_lib()
    pcms = [C.gen_stream(3300 + i, 16000 * 25 + 1234 * i) for i in range(32)]
    plain = C.run_product_batch(path, pcms, chunk=16000 * 14)
    for p in pcms:
        assert lib.c.aab_pin_host_memory(ctypes.c_void_p(p.ctypes.data), p.nbytes) == 0
    try:
        pinned = C.run_product_batch(path, pcms, chunk=16000 * 14)
    finally:
        for p in pcms:
            assert lib.c.aab_unpin_host_memory(ctypes.c_void_p(p.ctypes.data)) == 0
    assert [c.keys() for c in pinned] == [c.keys() for c in plain]
    want = _oracle_many(path, pcms[:3], threads=3, chunk=16000 * 14)
    C.check_streams(path, pcms, pinned, want, [0, 1, 2], "pinned upload path", chunk=16000 * 14)
