"""GPU parity tests (-m gpu): CUDA path through the C ABI vs the oracle / golden fixtures."""
import ctypes
import os

import numpy as np
import pytest

import april_asr_b200 as A
import common as C

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-3      # stated fp32 tolerance on raw joiner logits / encoder_out (values are O(1..10))


def test_fbank_rows_match_reference_golden():
    g = np.load(os.path.join(C.GOLDEN, "fbank_stream0_3s.npz"))
    pcm = C.gen_stream(0, 48000)
    assert C.pcm_digest(pcm) == str(g["pcm_sha"])
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    out = np.zeros((400, 80), np.float32)
    n = lib.c.aab_debug_fbank(m.handle, pcm.ctypes.data_as(ctypes.POINTER(ctypes.c_short)), pcm.size,
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 400)
    ref = g["rows"]
    assert n == ref.shape[0]
    diff = np.abs(out[:n] - ref)
    mism = int((out[:n].view(np.uint32) != ref.view(np.uint32)).sum())
    print("fbank: max abs diff %.3g, bit-mismatching values %d / %d" % (diff.max(), mism, ref.size))
    assert diff.max() <= 2e-6           # fp64 FFT on device: expected bit-exact up to rare double-rounding
    assert mism <= ref.size // 1000
    m.close()


@pytest.mark.parametrize("key", ["stream0", "stream1", "stream2", "stream3", "stream4"])
def test_tokens_match_golden(key):
    g = C.load_golden("tiny_callbacks.json")[key]
    pcm = C.gen_stream(g["stream"], g["samples"]) if g["samples"] else np.zeros(0, np.int16)
    assert C.pcm_digest(pcm) == g["pcm_sha"]
    col = C.run_product_batch(C.TINY_MODEL, [pcm], chunk=g["chunk"])[0]
    want = [(r, [(t, f, ms) for t, f, ms, lp in toks]) for r, toks in g["log"]]
    assert col.keys() == want
    got_lp = [lp for r, toks in col.log for _, _, _, lp in toks]
    want_lp = [lp for r, toks in g["log"] for _, _, _, lp in toks]
    assert np.allclose(got_lp, want_lp, atol=LOGIT_TOL)


def test_batched_ragged_streams_match_oracle():
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    lens = [48000, 31000, 64000, 500, 0, 16000, 40000]
    pcms = [C.gen_stream(20 + i, n) if n else np.zeros(0, np.int16) for i, n in enumerate(lens)]
    cols = C.run_product_batch(C.TINY_MODEL, pcms, chunk=6400)
    for p, c in zip(pcms, cols):
        assert c.keys() == C.run_oracle(C.TINY_MODEL, p).keys()


def test_encoder_out_and_logits_within_tolerance():
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    pcm = C.gen_stream(3, 80000)
    o = C.run_oracle(C.TINY_MODEL, pcm, tap=True)
    p = C.run_product_batch(C.TINY_MODEL, [pcm], tap=True)[0]
    assert len(o.enc) == len(p.enc) and len(o.logits) == len(p.logits)
    e = max(float(np.abs(a - b).max()) for a, b in zip(o.enc, p.enc))
    l = max(float(np.abs(a - b).max()) for a, b in zip(o.logits, p.logits))
    print("encoder_out max abs err %.3g, logits max abs err %.3g" % (e, l))
    assert e < LOGIT_TOL and l < LOGIT_TOL
    assert o.keys() == p.keys()


def test_chunk_size_invariance_and_streaming_equals_bulk():
    pcm = C.gen_stream(9, 64000)
    base = C.run_product_batch(C.TINY_MODEL, [pcm])[0].keys()
    for chunk in (640, 3200, 7001):
        assert C.run_product_batch(C.TINY_MODEL, [pcm], chunk=chunk)[0].keys() == base


def test_single_session_reference_api_and_flush_idempotence():
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    assert m.get_sample_rate() == 16000 and m.get_language() == "en-us"
    col = C.Collected()
    s = A.Session(m, col.cb)
    pcm = C.gen_stream(0, 96000)
    s.feed_pcm16(pcm)
    s.flush()
    n = len(col.log)
    s.flush()                                   # april_session.c:548-550
    assert len(col.log) == n
    g = C.load_golden("tiny_callbacks.json")["stream0"]
    assert col.keys() == [(r, [(t, f, ms) for t, f, ms, lp in toks]) for r, toks in g["log"]]
    assert s.get_rt_speedup() == 1.0
    s.close()
    m.close()


def test_flush_on_fresh_session_and_feed_after_flush():
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    pcm = C.gen_stream(11, 40000)

    def scenario(lib_kind):
        lib = C.oracle_lib() if lib_kind == "o" else C.product_lib()
        m = A.Model(C.TINY_MODEL, lib)
        col = C.Collected()
        s = A.Session(m, col.cb)
        s.flush()                               # fresh flush: 25 padded steps
        s.feed_pcm16(pcm[:20000]); s.flush()
        s.feed_pcm16(pcm[20000:20100]); s.flush()   # < 512 new samples: avail_f stays below -27 until the zeros arrive
        s.feed_pcm16(pcm[20100:]); s.flush()
        s.close(); m.close()
        return col.keys()
    assert scenario("p") == scenario("o")


def test_full_size_model_matches_oracle_tokens():
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    path = C.model_path("full")
    pcms = [C.gen_stream(100 + i, 64000) for i in range(2)]
    cols = C.run_product_batch(path, pcms, tap=True)
    for pcm, c in zip(pcms, cols):
        o = C.run_oracle(path, pcm, tap=True)
        e = max(float(np.abs(a - b).max()) for a, b in zip(o.enc, c.enc))
        print("full model: encoder_out max abs err %.3g over %d steps" % (e, len(o.enc)))
        assert e < LOGIT_TOL
        assert c.keys() == o.keys()


def test_async_session_batcher():
    import time
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    cols = [C.Collected() for _ in range(4)]
    ss = [A.Session(m, c.cb, asynchronous=True, no_rt=True) for c in cols]
    pcms = [C.gen_stream(30 + i, 48000) for i in range(4)]
    for i in range(0, 48000, 1600):
        for s, p in zip(ss, pcms):
            s.feed_pcm16(p[i:i + 1600])
        time.sleep(0.002)
    time.sleep(0.3)
    for s in ss:
        s.flush()
    time.sleep(0.5)
    sync = C.run_product_batch(C.TINY_MODEL, pcms, model=m)
    for a, b in zip(cols, sync):
        assert [k for k in a.keys() if k[0] != 3] == b.keys()
    for s in ss:
        s.close()
    m.close()


@pytest.mark.parametrize("prog", ["example", "example_srt"])
def test_reference_example_cpp_unmodified_against_this_library(tmp_path, prog):
    """/root/reference/example.cpp and example_srt.cpp compiled UNMODIFIED twice (oracle/Makefile): against the oracle library
    and against libaprilasr.so. Same PCM, same model -> byte-identical stdout (partials, finals, silences as the CLI prints
    them; subtitle blocks with their time_ms-derived timestamps for the SRT tool)."""
    import subprocess
    ref = os.path.join(C.ROOT, "oracle", "_ref", prog + "_ref")
    mine = os.path.join(C.ROOT, "oracle", "_ref", prog + "_b200")
    if not (os.path.exists(ref) and os.path.exists(mine)):
        pytest.skip("example binaries not built (need /root/reference at build time)")
    pcm = C.gen_stream(0, 96000)
    raw = tmp_path / "a.raw"
    raw.write_bytes(pcm.tobytes())
    a = subprocess.run([ref, str(raw), C.TINY_MODEL], capture_output=True, timeout=300)
    b = subprocess.run([mine, str(raw), C.TINY_MODEL], capture_output=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-300:], b.stderr[-300:])
    assert a.stdout == b.stdout
    assert len(b.stdout) > 200
    if prog == "example":
        assert b"Model name: synthetic-tiny" in b.stdout


def _oracle_many(model, pcms, threads=8):
    import threading
    out = [None] * len(pcms)

    def work(k):
        for i in range(k, len(pcms), threads):
            out[i] = C.run_oracle(model, pcms[i])
    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    for t in ts: t.start()
    for t in ts: t.join()
    return out


def test_several_sessions_per_search_cluster():
    """More sessions than search clusters (148 SMs / 8 CTAs = 18): every cluster walks 3 sessions in lock step, so a session
    that ends its step on a blank can sit through extra joiner passes of its cluster mates. Regression: a context reset on
    that blank (2.2 s without emission, april_session.c:411-416) must still reload the decoder row at the next step."""
    pcms = [C.gen_stream(200 + i, 16000 * 10 + 777 * i) for i in range(40)]
    got = C.run_product_batch(C.TINY_MODEL, pcms)
    want = _oracle_many(C.TINY_MODEL, pcms)
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if g.keys() != w.keys()]
    assert not bad, "sessions with differing callbacks: %s" % bad


def test_full_length_batch_matches_reference():
    """BASELINE configs[1] at full size: 64 sessions x 30 s through the batched path; the first 12 sessions are also run
    through the oracle (30 s each, one thread per session). The whole callback sequence (partials, finals, silences, token
    ids, flags, times) must be identical; one near-tied decision in 12 long streams is tolerated and reported, because two
    fp32 implementations that differ by 1e-5 in a logit cannot agree on a margin smaller than that."""
    path = C.model_path("full")
    pcms = [C.gen_stream(i, 16000 * 30) for i in range(64)]
    got = C.run_product_batch(path, pcms)
    want = _oracle_many(path, pcms[:12], threads=12)
    diffs = [C.callback_diff(got[i], want[i]) for i in range(12)]
    bad = [i for i in range(12) if got[i].keys() != want[i].keys()]
    print("full-length parity: %d of 12 sessions identical; differing: %s; (callbacks differing, finals identical): %s" % (12 - len(bad), bad, diffs))
    assert all(f for _, f in diffs), "final results differ"
    assert len(bad) <= 1
    assert sum(len(t) for r, t in want[0].log if r == 2) > 100      # the streams really produce text


def test_long_form_carry_over_in_many_small_feeds():
    """BASELINE configs[4] in miniature: long streams fed in many calls, so LSTM state, feature-row carry, greedy context
    and the active-token window survive hundreds of passes. 4 sessions x 120 s (mid-size model with a contractive recurrence,
    generated at test time), 0.5 s per feed call, ragged lengths; the oracle gets the same chunking."""
    import threading
    path = C.model_path("mid")
    pcms = [C.gen_stream(400 + i, 16000 * 120 + 4321 * i) for i in range(4)]
    got = C.run_product_batch(path, pcms, chunk=8000)
    want = [None] * 4

    def work(i):
        want[i] = C.run_oracle(path, pcms[i], chunk=8000)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts: t.start()
    for t in ts: t.join()
    for i in range(4):
        nd, finals_ok = C.callback_diff(got[i], want[i])
        print("long-form session %d: %d callbacks, %d differ, finals identical: %s" % (i, len(want[i].log), nd, finals_ok))
        assert finals_ok, "session %d: final results differ" % i
        assert nd <= 3, "session %d: %d callbacks differ" % (i, nd)      # tentative-token near-ties only (see common.callback_diff)
        assert len(want[i].log) > 500


def test_fbank_v2_variant_is_bit_compatible():
    """Same golden rows through the opt-in front-end variant (APRIL_B200_FBANK=v2: eight DC-sum chains per warp; correct on
    hardware, speed not yet measured, hence not the default). Runs in a subprocess because the switch is read once per process."""
    import subprocess, sys
    code = (
        "import os, sys, ctypes, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import april_asr_b200 as A, common as C\n"
        "g = np.load(os.path.join(C.GOLDEN, 'fbank_stream0_3s.npz')); pcm = C.gen_stream(0, 48000)\n"
        "lib = C.product_lib(); m = A.Model(C.TINY_MODEL, lib); out = np.zeros((400, 80), np.float32)\n"
        "n = lib.c.aab_debug_fbank(m.handle, pcm.ctypes.data_as(ctypes.POINTER(ctypes.c_short)), pcm.size, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 400)\n"
        "ref = g['rows']; assert n == ref.shape[0]\n"
        "mism = int((out[:n].view(np.uint32) != ref.view(np.uint32)).sum()); print('mismatching', mism)\n"
        "assert np.abs(out[:n] - ref).max() <= 2e-6 and mism <= ref.size // 1000\n"
        "cols = C.run_product_batch(C.TINY_MODEL, [C.gen_stream(2, 64000)]); want = C.load_golden('tiny_callbacks.json')['stream2']\n"
    ) % (C.ROOT, os.path.join(C.ROOT, "tests"), os.path.join(C.ROOT, "tools"))
    env = dict(os.environ, APRIL_B200_FBANK="v2")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]


@pytest.mark.xfail(strict=False, reason="written after this round's GPU budget was spent; not yet run on hardware")
def test_async_ring_overflow_reports_cant_keep_up():
    """audio_provider.c:59-64 / april_session.c:482-492: an async session whose 48 000-sample ring would overflow drops the
    chunk and calls the handler with ERROR_CANT_KEEP_UP (count 0) from the feeding thread."""
    import time
    lib = C.product_lib()
    m = A.Model(C.TINY_MODEL, lib)
    col = C.Collected()
    s = A.Session(m, col.cb, asynchronous=True, no_rt=True)
    s.feed_pcm16(C.gen_stream(50, 100000))                # > 48 000 in one call: cannot fit whatever the batcher has drained
    assert (3, []) in col.keys()
    s.feed_pcm16(C.gen_stream(51, 16000))                 # a chunk that fits is still accepted afterwards
    time.sleep(0.3)
    s.flush(); time.sleep(0.5)
    assert any(r in (1, 2) for r, _ in col.keys()) or len(col.keys()) >= 1
    s.close(); m.close()
