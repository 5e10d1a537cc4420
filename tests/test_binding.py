"""The reference's own Python binding (bindings/python/april_asr: _april_c_ffi.py:118-136 loads `libaprilasr.so` from its package
directory; _april.py:99-179 wraps Model / Session) run UNMODIFIED — as the sourceless bytecode package oracle/Makefile compiles
from /root/reference — against this repo's libaprilasr.so. Callbacks must equal the oracle's (§8f-3 binding conformance)."""
import json
import os
import shutil
import subprocess
import sys

import pytest

import common as C

PYB = os.path.join(C.ROOT, "oracle", "_ref", "pybinding", "april_asr")

CHILD = r'''
import json, sys
sys.path.insert(0, sys.argv[1])
import april_asr                                   # the reference's package, bytecode only
model = april_asr.Model(sys.argv[2])
log = []
def handler(result, tokens):
    log.append([int(result), [[t.token, bool(t.word_boundary), bool(t.sentence_end), int(round(t.time * 1000))] for t in tokens]])
session = april_asr.Session(model, handler)
raw = open(sys.argv[3], "rb").read()
chunk = int(sys.argv[4]) * 2
for i in range(0, len(raw), chunk):
    session.feed_pcm16(raw[i:i + chunk])
session.flush()
info = [model.get_name(), model.get_language(), model.get_sample_rate(), session.get_rt_speedup()]
del session
print(json.dumps({"info": info, "log": log}))
'''


def _run_binding(tmp_path, so_path, pcm, chunk):
    if not os.path.exists(os.path.join(PYB, "__init__.pyc")):
        pytest.skip("reference binding not compiled (oracle/Makefile needs /root/reference at build time)")
    pkg_root = tmp_path / "site"
    shutil.copytree(PYB, pkg_root / "april_asr")
    shutil.copy(so_path, pkg_root / "april_asr" / "libaprilasr.so")     # where _load_library() looks (linux)
    raw = tmp_path / "a.raw"
    raw.write_bytes(pcm.tobytes())
    r = subprocess.run([sys.executable, "-c", CHILD, str(pkg_root), C.TINY_MODEL, str(raw), str(chunk)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def _oracle_log(pcm, chunk):
    o = C.run_oracle(C.TINY_MODEL, pcm, chunk=chunk)
    return [[r, [[t, bool(f & 1), bool(f & 2), ms] for t, f, ms in toks]] for r, toks in o.keys()]


def test_reference_python_binding_harness_against_the_oracle_library(tmp_path):
    """CPU: the same harness over the oracle library — proves the bytecode package and the comparison are sound."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    pcm = C.gen_stream(2, 48000)
    got = _run_binding(tmp_path, C.ORACLE_SO, pcm, 3200)
    assert got["info"][:3] == ["synthetic-tiny-seed1234", "en-us", 16000]
    assert got["log"] == _oracle_log(pcm, 3200)
    assert len(got["log"]) > 10


@pytest.mark.gpu
def test_reference_python_binding_unmodified_against_this_library(tmp_path):
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    pcm = C.gen_stream(0, 96000)
    got = _run_binding(tmp_path, C.PRODUCT_SO, pcm, 4000)
    assert got["info"] == ["synthetic-tiny-seed1234", "en-us", 16000, 1.0]
    assert got["log"] == _oracle_log(pcm, 4000)
    assert len(got["log"]) > 20
