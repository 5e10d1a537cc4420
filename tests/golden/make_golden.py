#!/usr/bin/env python3
"""Generates the committed golden fixtures by RUNNING THE REFERENCE in this container:
  * oracle/_ref/libaprilasr_ref.so = /root/reference/src/*.c compiled unmodified (oracle/Makefile)
    over oracle/fake_ort (CPU fp32 restatement of the ONNX graphs);
  * fbank rows come straight from the reference's fbank.c/pocketfft.c symbols in that library.
Run from the repo root after `make -C oracle`:  python tests/golden/make_golden.py
The reference publishes no golden vectors of its own, so these are the pins.
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import common as C  # noqa: E402


class FBankOptions(ctypes.Structure):          # /root/reference/src/fbank.h:26-66
    _fields_ = [("sample_freq", ctypes.c_int), ("frame_shift_ms", ctypes.c_int), ("frame_length_ms", ctypes.c_int),
                ("num_bins", ctypes.c_int), ("round_pow2", ctypes.c_bool), ("mel_low", ctypes.c_int), ("mel_high", ctypes.c_int),
                ("snip_edges", ctypes.c_bool), ("pull_segment_count", ctypes.c_int), ("pull_segment_step", ctypes.c_int),
                ("use_sonic", ctypes.c_bool), ("remove_dc_offset", ctypes.c_bool), ("preemph_coeff", ctypes.c_float)]


def reference_fbank_rows(pcm):
    """Feature rows exactly as the reference session sees them (april_session.c:501-538 slicing)."""
    lib = C.oracle_lib().c
    lib.make_fbank.argtypes, lib.make_fbank.restype = [FBankOptions], ctypes.c_void_p
    lib.fbank_accept_waveform.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t]
    lib.fbank_pull_segments.argtypes, lib.fbank_pull_segments.restype = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t], ctypes.c_bool
    lib.free_fbank.argtypes = [ctypes.c_void_p]
    o = FBankOptions(16000, 10, 25, 80, True, 20, 0, True, 9, 4, False, True, 0.97)
    fb = lib.make_fbank(o)
    rows = {}
    step = 0
    seg = np.zeros((9, 80), np.float32)
    for i in range(0, len(pcm), 3200):
        wave = (pcm[i:i + 3200].astype(np.float32) / np.float32(32768.0)).astype(np.float32)
        lib.fbank_accept_waveform(fb, wave.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(wave))
        while lib.fbank_pull_segments(fb, seg.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), seg.nbytes):
            for r in range(9):
                rows[4 * step + r] = seg[r].copy()
            step += 1
    lib.free_fbank(fb)
    n = max(rows) + 1 if rows else 0
    return np.stack([rows[i] for i in range(n)]) if n else np.zeros((0, 80), np.float32)


def main():
    assert C.oracle_lib() is not None, "build the oracle first: make -C oracle"
    out = {}
    for sid, nsamp, chunk in ((0, 96000, None), (1, 96000, 640), (2, 64000, 3777), (3, 4000, None), (4, 0, None)):
        pcm = C.gen_stream(sid, nsamp) if nsamp else np.zeros(0, np.int16)
        col = C.run_oracle(C.TINY_MODEL, pcm, chunk=chunk)
        out["stream%d" % sid] = dict(stream=sid, samples=nsamp, chunk=chunk, pcm_sha=C.pcm_digest(pcm),
                                     log=[[r, [[t, f, ms, lp] for t, f, ms, lp in toks]] for r, toks in col.log])
    with open(os.path.join(HERE, "tiny_callbacks.json"), "w") as f:
        json.dump(out, f, indent=0)
    pcm = C.gen_stream(0, 48000)
    rows = reference_fbank_rows(pcm)
    np.savez_compressed(os.path.join(HERE, "fbank_stream0_3s.npz"), rows=rows, pcm_sha=C.pcm_digest(pcm))
    print("golden written:", {k: len(v["log"]) for k, v in out.items()}, rows.shape)


if __name__ == "__main__":
    main()
