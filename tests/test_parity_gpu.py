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
    """Any chunking of the same stream must give the ORACLE's callbacks for it (the oracle itself is chunking-invariant:
    tests/test_oracle.py::test_oracle_chunk_size_invariance)."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    pcm = C.gen_stream(9, 64000)
    want = C.run_oracle(C.TINY_MODEL, pcm).keys()
    for chunk in (None, 640, 3200, 7001):
        assert C.run_product_batch(C.TINY_MODEL, [pcm], chunk=chunk)[0].keys() == want, "chunk %s" % chunk


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


def _oracle_many(model, pcms, threads=8, chunk=None):
    import threading
    out = [None] * len(pcms)

    def work(k):
        for i in range(k, len(pcms), threads):
            out[i] = C.run_oracle(model, pcms[i], chunk=chunk)
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


_ORACLE_30S = {}


def _oracle_30s(path, ids):
    """Oracle callback logs of the 30 s benchmark streams, computed once per test session and shared between the tests below."""
    todo = [i for i in ids if i not in _ORACLE_30S]
    if todo:
        res = _oracle_many(path, [C.gen_stream(i, 16000 * 30) for i in todo], threads=min(len(todo), 24))
        for i, r in zip(todo, res):
            _ORACLE_30S[i] = r
    return [_ORACLE_30S[i] for i in ids]


CHECKED_STREAMS = 24


def test_full_length_batch_matches_reference():
    """BASELINE configs[1] at full size: 64 sessions x 30 s through the batched path; 24 of them are also run through the
    oracle (30 s each, one thread per session). The whole callback sequence (partials, finals, silences, token ids, flags,
    times) is compared."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    path = C.model_path("full")
    pcms = [C.gen_stream(i, 16000 * 30) for i in range(64)]
    got = C.run_product_batch(path, pcms)
    ids = list(range(CHECKED_STREAMS))
    want = _oracle_30s(path, ids)
    C.check_streams(path, pcms, got, want, ids, "64 x 30 s")
    assert sum(len(t) for r, t in want[0].log if r == 2) > 100      # the streams really produce text


def test_512_sessions_on_one_gpu_match_reference():
    """BASELINE configs[2]'s session count on ONE GPU (north_star's operating point): 512 sessions x 30 s in one batched pass
    (four 128-session groups per scan launch); 16 sessions spread over the groups are checked against the oracle."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    from gen_pcm import gen_streams
    path = C.model_path("full")
    order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]          # oracle streams (shared with the test above) ...
    where = [0, 37, 127, 128, 129, 200, 255, 256, 300, 383, 384, 400, 450, 500, 510, 511]   # ... placed at these batch positions
    ids = list(range(1000, 1512))
    for pos, sid in zip(where, order):
        ids[pos] = sid
    pcms = gen_streams(ids, 16000 * 30)
    got = C.run_product_batch(path, pcms)
    want = _oracle_30s(path, order)
    C.check_streams(path, pcms, got, want, where, "512 x 30 s")


def test_256_sessions_streaming_40ms_chunks_match_reference():
    """BASELINE configs[3]: 256 sessions fed 640-sample (40 ms) chunks, one batched pass per tick (T = 1 encoder step per pass:
    every per-slot state carried across hundreds of passes); 16 sessions are checked against the oracle fed the same chunks."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    path = C.model_path("full")
    n = 16000 * 6
    pcms = [C.gen_stream(2000 + i, n) for i in range(256)]
    got = C.run_product_batch(path, pcms, chunk=640)
    where = [0, 1, 63, 64, 100, 127, 128, 129, 170, 191, 192, 200, 230, 250, 254, 255]
    want = _oracle_many(path, [pcms[i] for i in where], threads=16, chunk=640)
    C.check_streams(path, pcms, got, want, where, "256 x 40 ms chunks", chunk=640)


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
        assert len(want[i].log) > 500
    C.check_streams(path, pcms, got, want, [0, 1, 2, 3], "4 x 120 s in 0.5 s feeds", chunk=8000)


def test_generic_window_front_end_matches_the_golden_rows_too():
    """The default front-end for the 512-sample window is the radix-8 kernel (test_fbank_rows_match_reference_golden); windows of
    another size take the generic radix-2 kernel. Force it (APRIL_B200_FBANK=generic, read once per process, hence the subprocess)
    and hold it to the same golden rows and the same golden callback log."""
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
        "gl = C.load_golden('tiny_callbacks.json')['stream2']\n"
        "col = C.run_product_batch(C.TINY_MODEL, [C.gen_stream(gl['stream'], gl['samples'])], chunk=gl['chunk'])[0]\n"
        "want = [(r, [(t, f, ms) for t, f, ms, lp in toks]) for r, toks in gl['log']]\n"
        "assert col.keys() == want, 'callbacks differ from the golden log'\n"
    ) % (C.ROOT, os.path.join(C.ROOT, "tests"), os.path.join(C.ROOT, "tools"))
    env = dict(os.environ, APRIL_B200_FBANK="generic")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]


def test_fp32_simt_tier_is_a_tight_device_side_reference():
    """APRIL_B200_PRECISION=fp32 runs the whole encoder in plain fp32 SIMT kernels (no tensor cores, per-step recurrence): slow, but
    the device-side ground truth the tensor-core tier is developed against. It must sit an order of magnitude closer to the oracle
    than the 2e-3 bar: 5e-5 on encoder_out and raw logits of the full model (measured 8e-6 / 1.5e-5)."""
    import json, subprocess, sys
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    code = (
        "import sys, json, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import common as C\n"
        "path = C.model_path('full'); pcm = C.gen_stream(123, 16000 * 4)\n"
        "p = C.run_product_batch(path, [pcm], tap=True)[0]; o = C.run_oracle(path, pcm, tap=True)\n"
        "e = max(float(np.abs(a - b).max()) for a, b in zip(o.enc, p.enc)); l = max(float(np.abs(a - b).max()) for a, b in zip(o.logits, p.logits))\n"
        "print(json.dumps({'enc': e, 'logits': l, 'same': o.keys() == p.keys(), 'steps': len(o.enc), 'passes': len(o.logits) == len(p.logits)}))\n"
    ) % (C.ROOT, os.path.join(C.ROOT, "tests"), os.path.join(C.ROOT, "tools"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, APRIL_B200_PRECISION="fp32"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-600:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    print("fp32 SIMT tier vs oracle:", res)
    assert res["same"] and res["passes"] and res["steps"] > 50
    assert res["enc"] < 5e-5 and res["logits"] < 5e-5


@pytest.mark.parametrize("env", [{"APRIL_B200_GREEDY": "pair"}, {"APRIL_B200_GREEDY": "cluster", "APRIL_B200_OVERLAP": "0"},
                                 {"APRIL_B200_OVERLAP": "4"}, {"APRIL_B200_OVERLAP": "7", "APRIL_B200_GREEDY": "pair"}],
                         ids=["search-on-cta-pairs", "search-on-8-cta-clusters", "search-under-the-encoder-4-chunks", "search-under-the-encoder-7-chunks"])
def test_every_search_schedule_matches_the_oracle(env):
    """The engine has two search kernels (8-CTA clusters with the joiner in registers; CTA pairs streaming it from L2) and two
    schedules (search after the encoder; search of time chunk c under the encoder of chunk c + 1). Whatever the defaults pick for a
    batch size, each combination is held to the oracle's callbacks here: 28 ragged sessions x ~20 s in one feed + flush (mid model),
    and a second feed afterwards so that the state a schedule leaves behind is exercised too. The switches are read at model load."""
    import os
    path = C.model_path("mid")
    pcms = [C.gen_stream(2100 + i, 16000 * 20 + 3111 * i) for i in range(28)]          # the overlapped schedule serves 24..64 sessions
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        got = C.run_product_batch(path, pcms, chunk=16000 * 14)          # two feeds: 14 s, then the rest
    finally:
        for k, v in saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
    want = _oracle_many(path, pcms, chunk=16000 * 14)
    C.check_streams(path, pcms, got, want, list(range(28)), "search schedule %s" % env, chunk=16000 * 14)
