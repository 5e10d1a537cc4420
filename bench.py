#!/usr/bin/env python3
"""bench.py — audio frames/s of the batched RNN-T hot path (fbank -> encoder -> decoder/joiner ->
greedy search) behind the april C ABI, on N B200s (one process per GPU, sessions sharded, no collective
on the data path), next to the reference's CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch: `sessions` fresh streams of `seconds` of synthetic
16 kHz s16 PCM are fed whole (aab_feed_pcm16_batch) and flushed (aab_flush_batch) — BASELINE.json
configs[1] at N=1 (64 x 30 s), configs[2] sharded 64 per GPU at N>1.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

FRAMES_PER_STREAM = lambda n: (n - 512) // 160 + 1          # frames = floor((N-512)/160)+1
FLOP_PER_STEP_ENCODER = 2.0 * 84_196_240


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", str(gpu_index), "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=3)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], None, set()
        for r in rows:
            if len(r) < 9 or r[0] != str(self.idx):
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def ensure_model(rank, dist):
    path = os.path.join(ROOT, "build", "full.april")
    if rank == 0 and not os.path.exists(path):
        import make_synthetic_april
        os.makedirs(os.path.dirname(path), exist_ok=True)
        make_synthetic_april.generate(path + ".tmp", "full", 1234)
        os.replace(path + ".tmp", path)
    if dist is not None:
        dist.barrier()
    return path


def time_cpu_reference(model_path, seconds_per_session, n_threads, first_id=0):
    """The reference's own host code (oracle/_ref: /root/reference/src/*.c compiled unmodified over the CPU
    fp32 ONNX interpreter), one OS thread per session like the reference (1 intra-op thread,
    april_model.c:54-55). Returns (frames/s, wall s, frames)."""
    import april_asr_b200 as A
    from gen_pcm import gen_stream
    so = os.path.join(ROOT, "oracle", "_ref", "libaprilasr_ref.so")
    lib = A.Lib(so)
    m = A.Model(model_path, lib)
    n = int(16000 * seconds_per_session)
    pcms = [gen_stream(first_id + i, n) for i in range(n_threads)]
    counters = [(ctypes.c_uint64 * 4)() for _ in range(n_threads)]

    def _noop(ud, r, cnt, toks):
        return None
    handler = A.HANDLER(_noop)
    sessions = []
    for i in range(n_threads):
        cfg = A.AprilConfig(); cfg.handler = handler; cfg.userdata = None; cfg.flags = 0
        sessions.append(lib.c.aas_create_session(m.handle, cfg))

    def work(i):
        lib.c.aas_feed_pcm16(sessions[i], pcms[i].ctypes.data_as(ctypes.POINTER(ctypes.c_short)), pcms[i].size)
        lib.c.aas_flush(sessions[i])
    ts = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    wall = time.perf_counter() - t0
    for s in sessions:
        lib.c.aas_free(s)
    m.close()
    frames = n_threads * FRAMES_PER_STREAM(n)
    return frames / wall, wall, frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sessions", type=int, default=64, help="sessions per GPU")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--cpu-seconds", type=float, default=2.0, help="audio seconds per session in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="bulk", choices=["bulk", "stream"], help="stream = BASELINE configs[3]: 40 ms chunks through the batched feed, per-chunk latency")
    ap.add_argument("--ticks", type=int, default=250)
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if a.impl == "reference":
        if rank != 0:
            return
        path = ensure_model(0, None)
        nthr = max(1, min(ncores, 32))
        vals = []
        for i in range(a.warmup + a.steps):
            fps, wall, frames = time_cpu_reference(path, a.cpu_seconds, nthr, first_id=1000 * i)
            if i >= a.warmup:
                vals.append((fps, wall, frames))
        fps = sum(v[2] for v in vals) / sum(v[1] for v in vals)
        sample = "%d sessions x %.1f s of the same synthetic 16 kHz PCM, one thread per session, feed whole clip + flush" % (nthr, a.cpu_seconds)
        print(json.dumps({
            "impl": "reference", "metric": "audio_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * sum(v[1] for v in vals) / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "64 concurrent sessions x 30 s synthetic 16 kHz PCM (BASELINE configs[1]); CPU arm runs a bounded sample of it", "sample": sample},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": nthr, "kind": "reference", "sample": sample,
                             "note": "reference host sources compiled unmodified over a CPU fp32 restatement of the ONNX graphs (ONNXRuntime itself is not in this image)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "x_realtime_per_stream": a.cpu_seconds / (sum(v[1] for v in vals) / len(vals)),
        }))
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"          # NCCL_DEBUG=VERSION/INFO prints to stdout, which must carry exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    os.environ["APRIL_B200_DEVICE"] = str(local)
    import april_asr_b200 as A
    from april_asr_b200 import _ffi, dist_util
    from gen_pcm import gen_stream

    path = ensure_model(rank, dist)
    lib = A.Lib()
    m = A.Model(path, lib)
    S, n = a.sessions, int(16000 * a.seconds)
    pcms = [gen_stream(i, n) for i in dist_util.shard_stream_ids(rank, world, S)]
    handler = ctypes.cast(lib.c.aab_counting_handler, A.HANDLER)
    counters = (ctypes.c_uint64 * 4)()

    def make_sessions():
        out = []
        for _ in range(S):
            cfg = A.AprilConfig(); cfg.handler = handler; cfg.userdata = ctypes.cast(counters, ctypes.c_void_p); cfg.flags = 0
            h = lib.c.aas_create_session(m.handle, cfg)
            assert h
            out.append(h)
        return out

    hs_t = ctypes.c_void_p * S
    ptrs = (ctypes.POINTER(ctypes.c_short) * S)(*[p.ctypes.data_as(ctypes.POINTER(ctypes.c_short)) for p in pcms])
    cnts = (ctypes.c_size_t * S)(*[p.size for p in pcms])

    def one_step():
        ss = make_sessions()
        arr = hs_t(*ss)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.c.aab_feed_pcm16_batch(arr, ptrs, cnts, S)
        rc |= lib.c.aab_flush_batch(arr, S)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert rc == 0
        for h in ss:
            lib.c.aas_free(h)
        return dt

    if a.mode == "stream":
        # BASELINE configs[3]: S sessions fed 40 ms (640-sample) chunks; latency = time inside one batched feed call
        # (handlers run before it returns), measured on the host around the C-ABI call.
        ss = make_sessions()
        arr = hs_t(*ss)
        chunk = 640
        lat = []
        for k in range(a.ticks + 20):
            off = (k * chunk) % (n - chunk)
            p2 = (ctypes.POINTER(ctypes.c_short) * S)(*[ctypes.cast(p.ctypes.data + 2 * off, ctypes.POINTER(ctypes.c_short)) for p in pcms])
            c2 = (ctypes.c_size_t * S)(*([chunk] * S))
            t0 = time.perf_counter()
            rc = lib.c.aab_feed_pcm16_batch(arr, p2, c2, S)
            dt = time.perf_counter() - t0
            assert rc == 0
            if k >= 20:
                lat.append(dt * 1e3)
        lat.sort()
        p50, p99 = lat[len(lat) // 2], lat[min(len(lat) - 1, int(len(lat) * 0.99))]
        if rank == 0:
            print(json.dumps({"metric": "chunk_latency_ms", "mode": "stream", "sessions": S, "chunk_ms": 40, "ticks": a.ticks, "p50_ms": p50, "p99_ms": p99,
                              "mean_ms": sum(lat) / len(lat), "frames_per_sec": S * 4 / (sum(lat) / len(lat) / 1e3), "x_realtime_per_stream": 40.0 / (sum(lat) / len(lat)),
                              "config": {"workload": "%d sessions streaming 40 ms chunks through aab_feed_pcm16_batch (BASELINE configs[3]), 1 GPU" % S}}))
        for h in ss:
            lib.c.aas_free(h)
        m.close()
        return
    for _ in range(a.warmup):
        one_step()
    st = _ffi.AprilB200Stats()
    lib.c.aab_set_timing(m.handle, 1)
    lib.c.aab_get_stats(m.handle, ctypes.byref(st), 1)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    walls = [one_step() for _ in range(a.steps)]
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    lib.c.aab_get_stats(m.handle, ctypes.byref(st), 1)
    lib.c.aab_set_timing(m.handle, 0)

    (wall_s, compute_s), (frames_step,) = dist_util.combine(dist, "cuda", [sum(walls), st.compute_ms / 1e3], [S * FRAMES_PER_STREAM(n)])
    frames_total = frames_step * a.steps

    if rank == 0:
        pk, pk_src = peaks()
        steps_per_pass = st.encoder_steps / a.steps
        gemm_tflops = st.gemm_flops / (st.gemm_ms / 1e3) / 1e12 if st.gemm_ms > 0 else None
        peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        scan_tflops = (st.encoder_steps * 2.0 * 31_457_280) / max(1e-9, st.lstm_scan_ms / 1e3) / 1e12
        out = {
            "metric": "audio_frames_per_sec", "value": frames_total / compute_s, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 * compute_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d concurrent sessions x %.0f s synthetic 16 kHz s16 PCM per GPU, feed whole clip + flush (BASELINE configs[%d]); random-init aprilv0-shaped LSTM-transducer (12 x LSTMP 1024/512 + FF 2048, vocab 500)" % (S, a.seconds, 1 if world == 1 else 2),
                       "sessions_per_gpu": S, "seconds_per_session": a.seconds, "parallelism": "sessions sharded, %d per GPU, no collective on the data path" % S,
                       "l2": "working set (337 MB fp32 weights + GBs of activations per pass) exceeds the 126 MB L2; no explicit flush",
                       "timing": "value: CUDA events on the engine stream around the device-resident part of each pass (after the PCM upload), max over ranks; e2e: wall clock around the C-ABI calls with host buffers"},
            "x_realtime_per_stream": a.seconds / (compute_s / a.steps),
            "e2e": {"value": frames_total / wall_s, "unit": "frames/s", "ms_per_step": 1000.0 * wall_s / a.steps,
                    "h2d_bytes_per_step": st.h2d_bytes / a.steps, "d2h_bytes_per_step": st.d2h_bytes / a.steps,
                    "x_realtime_per_stream": a.seconds / (wall_s / a.steps)},
            "gpu_launches": int(st.kernel_launches),
            "clocks": clocks,
            # dominant kernel of the step: the persistent recurrent scan (one launch = one LSTM layer over all steps of a pass).
            # algorithmic FLOPs = the reference's W_hh (4H x D) and W_hr (D x H) products per session-step, counted once.
            "roofline": {"bound": "tensor", "kernel": "scan3::lstm_scan3_kernel (recurrent W_hh / W_hr part of one LSTM layer, all steps of a pass; %.0f %% of the step)" % (100.0 * st.lstm_scan_ms / max(1e-9, st.compute_ms)),
                         "achieved": scan_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": scan_tflops / peak_tf,
                         "peak_source": pk_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                         "traffic": 1173.2e6, "traffic_note": "dram bytes read + written by one launch (748 steps x 64 sessions), ncu --set full, profiles/r01b_ncu_scan3.raw.csv; algorithmic bytes of that launch: 784e6 (gate pre-activations G, fp32) + 196e6 (m out, fp16 hi/lo) + 16.8e6 (W_eff slice set) = 997e6",
                         "why_far_from_peak": "latency-bound dependent chain: %d sequential layer-steps per pass, each a grid-wide exchange of m(t) through L2 (TMA landing ~0.7 us + MMA + TMEM drain + pair exchange + cell + release ~1 us + acquire ~1.5 us); N = 64 sessions per MMA and a measured >= 84 clk issue cost per tcgen05.mma whatever its N (profiles/r01b_summary.md)" % int(12 * steps_per_pass / S),
                         "us_per_layer_step": 1e3 * st.lstm_scan_ms / max(1.0, (st.encoder_steps / S) * 12),
                         "hbm_view": {"achieved_GBps": (st.encoder_steps * 12 * (4 * 1024 * 4 + 2 * 1024 * 2)) / max(1e-9, st.lstm_scan_ms / 1e3) / 1e9, "peak_GBps": pk.get("hbm_gbs"), "note": "algorithmic bytes: fp32 gate rows in, fp16 (hi, lo) m rows out, per session-step and layer"},
                         "precision": "fp16x3 split (3-4 tensor-core products per algorithmic product; callbacks identical to the fp32 reference on 23 of 24 checked 30 s streams); algorithmic FLOPs counted once",
                         "gemm": {"kernel": "tc::gemm_tc_kernel<3,*,2> time-batched projections (W_ih, W_hr out, FF1, FF2, conv3, conv-out, enc_proj), CTA pairs 256x256x64",
                                  "achieved": gemm_tflops, "frac": (gemm_tflops / peak_tf) if gemm_tflops else None, "ms_per_step": st.gemm_ms / a.steps,
                                  "note": "3 MMAs per algorithmic product: the tensor pipe does 3x this figure"},
                         "gemm_ms_per_step": st.gemm_ms / a.steps, "greedy_ms_per_step": st.greedy_ms / a.steps, "lstm_scan_ms_per_step": st.lstm_scan_ms / a.steps,
                         "whole_pass_tflops": steps_per_pass * FLOP_PER_STEP_ENCODER / (st.compute_ms / a.steps / 1e3) / 1e12,
                         "whole_pass_frac": steps_per_pass * FLOP_PER_STEP_ENCODER / (st.compute_ms / a.steps / 1e3) / 1e12 / peak_tf},
            "results": {"partial_callbacks": int(counters[0]), "final_callbacks": int(counters[1]), "silence_callbacks": int(counters[2]), "final_tokens": int(counters[3]),
                        "encoder_steps_per_pass": steps_per_pass, "joiner_passes_per_pass": st.joiner_passes / a.steps},
        }
        if not a.no_cpu_baseline and world == 1:
            nthr = max(1, min(ncores, 32))
            fps, wall, frames = time_cpu_reference(path, a.cpu_seconds, nthr, first_id=5000)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": nthr, "kind": "reference", "per_session_frames_per_sec": fps / nthr, "per_session_x_realtime": fps / nthr / 100.0,
                                   "sample": "%d sessions x %.1f s of the same synthetic PCM, one thread per session (reference uses 1 ORT thread per session), feed + flush; %.1f s wall" % (nthr, a.cpu_seconds, wall),
                                   "note": "reference host sources compiled unmodified + CPU fp32 restatement of the ONNX graphs (ONNXRuntime unavailable in this image)"}
        print(json.dumps(out))
    m.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
