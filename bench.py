#!/usr/bin/env python3
"""bench.py — audio frames/s of the batched RNN-T hot path (fbank -> encoder -> decoder/joiner -> greedy search) behind the
april C ABI, on N B200s (one process per GPU, sessions sharded, no collective on the data path), next to the reference's
CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--sessions S] [--seconds T]

One "step" = one pass of the hot path over one batch: `sessions` fresh streams of `seconds` of synthetic 16 kHz s16 PCM are
fed whole (aab_feed_pcm16_batch) and flushed (aab_flush_batch) — BASELINE.json configs[1] at N=1 (64 x 30 s), configs[2]
sharded 64 per GPU at N>1. The default line also carries, as extra blocks measured in the same process after the timed steps:
  "b512":     512 sessions x 30 s on this GPU (north_star's operating point),
  "stream":   configs[3] — 256 synchronous sessions, one native thread each, 40 ms chunks through plain aas_feed_pcm16
              (coalesced into batched passes inside the library), p50/p99 of the per-chunk latency,
  "cpu_baseline": the reference's own host code on this box's cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

FRAMES_PER_STREAM = lambda n: (n - 512) // 160 + 1          # frames = floor((N-512)/160)+1
FLOP_PER_STEP_ENCODER = 2.0 * 84_196_240                    # BASELINE.md §2 / SURVEY.md §8d
FLOP_PER_STEP_LSTM_GEMM = 2.0 * 12 * 4_718_592              # W_ih + W_hh + W_hr of the 12 layers: the subset north_star's >= 50 % refers to
FLOP_PER_STEP_RECURRENT = 2.0 * 12 * (4096 * 512 + 512 * 1024)   # W_hh + W_hr: what the scan kernel replaces, counted as the reference computes it
ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "libaprilasr_ref.so")


def steps_of_clip(n_samples):
    """Encoder steps the reference runs for one clip fed whole and flushed (plan.h restates it)."""
    frames = FRAMES_PER_STREAM(n_samples)
    feed = max(0, (frames - 9) // 4 + 1)
    return feed + 28


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(kernel_substr):
    """dram bytes read + written by one launch of the dominant kernel, from the newest committed `ncu --set full` raw page of it
    under profiles/ (per launch); None when there is no capture of the current kernel."""
    import csv, glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*.raw.csv"))):
        try:
            rows = list(csv.reader(open(f)))
        except Exception:
            continue
        if len(rows) < 3:
            continue
        hdr = rows[0]
        try:
            ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except ValueError:
            continue
        units = rows[1]
        scale = lambda u: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        for r in rows[2:]:
            if len(r) > max(ik, ir, iw) and kernel_substr in r[ik]:
                try:
                    best = {"bytes": float(r[ir].replace(",", "")) * scale(units[ir]) + float(r[iw].replace(",", "")) * scale(units[iw]), "file": os.path.relpath(f, ROOT)}
                except ValueError:
                    pass
    return best


class ClockSampler:
    """`nvidia-smi -lms 50` running in the background from before the warm-up (its start-up alone can take longer than a short timed
    region); every row carries nvidia-smi's own timestamp and only the rows inside [begin(), end()] — the timed region — are used."""
    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        q = "timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", str(gpu_index), "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def begin(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.end()
        time.sleep(0.12)                                   # let the sample that straddles end() be written
        self.p.terminate()
        try:
            self.p.wait(timeout=3)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, total = [], None, set(), 0
        slack = datetime.timedelta(milliseconds=30)
        for r in rows:
            if len(r) < 10 or r[1] != str(self.idx):
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f")
                clk, cmax = float(r[2]), float(r[3])
            except ValueError:
                continue
            total += 1
            if self.t0 is not None and not (self.t0 - slack <= ts <= self.t1 + slack):
                continue
            sm.append(clk); mx = cmax
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "sampling": "nvidia-smi -lms 50, rows time-stamped inside the timed region (%d rows over the whole run)" % total}


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


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own host code (oracle/_ref: /root/reference/src/*.c compiled unmodified over the CPU fp32 ONNX
# interpreter), one native thread per session like the reference (1 intra-op thread, april_model.c:54-55), driven through
# aas_feed_pcm16 + aas_flush by tools/stream_driver.cpp. A step of the reference costs the same whatever the clip (fixed-size
# graphs, batch 1), so a bounded sample is scaled to the benchmarked clip BY ENCODER STEPS: the flush tail (28 steps) weighs
# in the reported value exactly as it does in a 30 s clip, whatever the sample's own length.
# ------------------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, model_path):
        import april_asr_b200 as A
        import stream_driver as SD
        from gen_pcm import gen_stream
        self.A, self.SD, self.gen = A, SD, gen_stream
        self.lib = A.Lib(ORACLE_SO)
        self.lib.c.fake_ort_get_run_counts.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
        self.m = A.Model(model_path, self.lib)

    def enc_runs(self):
        c = (ctypes.c_uint64 * 3)()
        self.lib.c.fake_ort_get_run_counts(c)
        return int(c[0])

    def sample(self, n_threads, seconds, first_id=0):
        """n_threads sessions x `seconds` fed whole + flushed. Returns (wall s, encoder steps run)."""
        n = int(16000 * seconds)
        pcms = [self.gen(first_id + i, n) for i in range(n_threads)]
        r0 = self.enc_runs()
        wall_ms, _ = self.SD.bulk(self.lib, self.m.handle, pcms, 0)
        return wall_ms / 1e3, self.enc_runs() - r0

    def measure(self, n_threads, budget_s, clip_seconds, first_id=0):
        """Calibrates on a 0.5 s clip, then times ONE sample sized to about budget_s. Value = frames/s the box would reach on
        n_threads clips of clip_seconds (feed + flush), i.e. steps/s of the sample x frames-per-step of the benchmarked clip."""
        w0, s0 = self.sample(n_threads, 0.5, first_id)
        per_step = w0 / max(1, s0 / n_threads)                          # s per encoder step and thread (threads run side by side)
        steps_budget = max(40.0, budget_s / per_step)
        secs = min(clip_seconds, max(1.0, (steps_budget - 28) * 0.04))
        wall, steps = self.sample(n_threads, secs, first_id + 100)
        n_clip = int(16000 * clip_seconds)
        clip_frames, clip_steps = FRAMES_PER_STREAM(n_clip), steps_of_clip(n_clip)
        value = (steps / wall) * clip_frames / clip_steps
        return {"value": value, "unit": "frames/s", "cores": n_threads, "kind": "reference",
                "sample": "%d sessions x %.1f s of the same synthetic PCM, one native thread per session through aas_feed_pcm16 + aas_flush: %d encoder steps in %.1f s wall; scaled by encoder steps to %.0f s clips (%d frames per %d steps)" % (n_threads, secs, steps, wall, clip_seconds, clip_frames, clip_steps),
                "encoder_steps_per_sec": steps / wall, "per_session_x_realtime": (steps / wall / n_threads) * clip_frames / clip_steps / 100.0,
                "wall_s": wall + w0,
                "note": "reference host sources compiled unmodified (oracle/_ref) over a CPU fp32 restatement of the ONNX graphs; ONNXRuntime itself is not in this image"}

    def close(self):
        self.m.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sessions", type=int, default=64, help="sessions per GPU")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--cpu-budget", type=float, default=None, help="wall seconds of one CPU sample (default: 20 s inside the b200 arm; 150 s / (steps + warmup) in the reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the b512 and stream blocks")
    ap.add_argument("--mode", default="bulk", choices=["bulk", "stream", "soak"], help="stream = only BASELINE configs[3] (256 sessions x 40 ms chunks via aas_feed_pcm16); soak = BASELINE configs[4] per GPU: long-form streams fed in many calls (state carried across all of them)")
    ap.add_argument("--soak-minutes", type=float, default=60.0, help="audio minutes per session in --mode soak")
    ap.add_argument("--soak-chunk", type=float, default=10.0, help="audio seconds per feed call in --mode soak")
    ap.add_argument("--ticks", type=int, default=250)
    ap.add_argument("--no-pin", action="store_true", help="leave the input PCM buffers pageable (the library stages them into pinned memory itself)")
    ap.add_argument("--lanes", type=int, default=0, help="ONE process driving this many GPUs through the library's own placement (SURVEY.md par. 8e): sessions x lanes streams, no torchrun; 0 = one process per GPU")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workload = "%d concurrent sessions x %.0f s synthetic 16 kHz s16 PCM per GPU, feed whole clip + flush (BASELINE configs[%d]); random-init aprilv0-shaped LSTM-transducer (12 x LSTMP 1024/512 + FF 2048, vocab 500)" % (
        a.sessions, a.seconds, 1 if world == 1 else 2)

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if a.impl == "reference":
        if rank != 0:
            return
        path = ensure_model(0, None)
        nthr = max(1, min(ncores, 32))
        arm = CpuArm(path)
        budget = a.cpu_budget or max(4.0, 150.0 / max(1, a.steps + a.warmup))   # the whole run ends within a few minutes whatever K and W
        vals = []
        for i in range(a.warmup + a.steps):
            r = arm.measure(nthr, budget, a.seconds, first_id=1000 * i)
            if i >= a.warmup:
                vals.append(r)
        arm.close()
        fps = statistics.mean(v["value"] for v in vals)
        cb = dict(vals[-1]); cb["value"] = fps
        n_clip = int(16000 * a.seconds)
        print(json.dumps({
            "impl": "reference", "metric": "audio_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * a.sessions * FRAMES_PER_STREAM(n_clip) / fps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sessions_per_gpu": a.sessions, "seconds_per_session": a.seconds,
                       "sample": cb["sample"], "ms_per_step_note": "time this box's %d host threads would need for one step of the workload (sessions x clip), from the sample" % nthr},
            "cpu_baseline": cb,
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "x_realtime_per_stream": cb["per_session_x_realtime"],
        }))
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if a.lanes > 0:
        assert world == 1, "--lanes is the single-process layout"
        os.environ.pop("APRIL_B200_DEVICE", None); os.environ.pop("LOCAL_RANK", None)
        os.environ["APRIL_B200_MAX_DEVICES"] = str(a.lanes)      # aam_create_model builds one engine per device; aas_create_session picks the least-loaded
    else:
        os.environ["APRIL_B200_DEVICE"] = str(local)   # one process per GPU: pin this process's models to its device
    import april_asr_b200 as A
    import stream_driver as SD
    from april_asr_b200 import _ffi, dist_util
    from gen_pcm import gen_stream, gen_streams

    path = ensure_model(rank, dist)
    lib = A.Lib()
    t_load = time.perf_counter()
    m = A.Model(path, lib)
    t_load = time.perf_counter() - t_load      # aam_create_model: parse, upload, fp16 splits, W_eff fold, decoder table
    handler = ctypes.cast(lib.c.aab_counting_handler, A.HANDLER)
    pk, pk_src = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])

    def run_bulk(S, seconds, warmup, steps, stream_ids):
        """warmup + steps passes of S fresh sessions x seconds; returns the per-config measurements (this rank's)."""
        n = int(16000 * seconds)
        pcms = gen_streams(stream_ids, n)
        # the step's inputs live in PINNED host memory (page-locked once, outside the timed region, through the library's own call):
        # the H2D copies inside the timed region are DMA'd straight from these buffers. --no-pin leaves them pageable: the library then
        # copies every feed into its pinned staging buffer first.
        pinned = [p for p in pcms if not a.no_pin and lib.c.aab_pin_host_memory(ctypes.c_void_p(p.ctypes.data), p.nbytes) == 0]
        counters = (ctypes.c_uint64 * 4)()
        hs_t = ctypes.c_void_p * S
        ptrs = (ctypes.POINTER(ctypes.c_short) * S)(*[p.ctypes.data_as(ctypes.POINTER(ctypes.c_short)) for p in pcms])
        cnts = (ctypes.c_size_t * S)(*[p.size for p in pcms])

        def one_step():
            ss = []
            for _ in range(S):
                cfg = A.AprilConfig(); cfg.handler = handler; cfg.userdata = ctypes.cast(counters, ctypes.c_void_p); cfg.flags = 0
                h = lib.c.aas_create_session(m.handle, cfg)
                assert h
                ss.append(h)
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
        sampler = ClockSampler(local) if rank == 0 else None
        for _ in range(warmup):
            one_step()
        st = _ffi.AprilB200Stats()
        lib.c.aab_set_timing(m.handle, 1)
        lib.c.aab_get_stats(m.handle, ctypes.byref(st), 1)
        for i in range(4):
            counters[i] = 0
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler: sampler.begin()
        walls = [one_step() for _ in range(steps)]
        torch.cuda.synchronize()
        if sampler: sampler.end()
        if dist is not None:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        for p in pinned:
            lib.c.aab_unpin_host_memory(ctypes.c_void_p(p.ctypes.data))
        lib.c.aab_get_stats(m.handle, ctypes.byref(st), 1)
        lib.c.aab_set_timing(m.handle, 0)
        return {"S": S, "n": n, "walls": walls, "st": st, "clocks": clocks, "counters": [int(x) for x in counters], "pinned_inputs": len(pinned) == len(pcms)}

    def roofline_of(r, steps):
        st = r["st"]
        sess_steps = st.encoder_steps                                       # session-steps over the timed steps
        scan_s, lgemm_s = st.lstm_scan_ms / 1e3, st.lstm_gemm_ms / 1e3
        scan_tf = sess_steps * FLOP_PER_STEP_RECURRENT / max(1e-9, scan_s) / 1e12
        lstm_tf = sess_steps * FLOP_PER_STEP_LSTM_GEMM / max(1e-9, scan_s + lgemm_s) / 1e12
        gemm_tf = st.gemm_flops / max(1e-9, st.gemm_ms / 1e3) / 1e12
        whole_tf = sess_steps * FLOP_PER_STEP_ENCODER / max(1e-9, st.compute_ms / 1e3) / 1e12
        layer_steps = max(1.0, (sess_steps / r["S"]) * 12)
        # the engine picks the scan kernel by the pass's session count (engine.cu: scan4_layer)
        if r["S"] <= 64: kname, kdesc = "lstm_scan3_kernel", "scan3::lstm_scan3_kernel (<= 64 sessions: latency-optimised CTA pairs, W_eff stacked on the UMMA M dimension)"
        elif r["S"] <= 256: kname, kdesc = "lstm_scan4_kernel", "scan4::lstm_scan4_kernel (65-256 sessions: sessions on the UMMA M dimension, K split across a 2-CTA cluster)"
        else: kname, kdesc = "lstm_scan5_kernel", "scan5::lstm_scan5_kernel (> 256 sessions: cta_group::2 pair MMAs, items of 256 sessions, cluster of 4)"
        tr = ncu_traffic(kname)
        return {"bound": "tensor", "kernel": kdesc + "; the recurrent W_hh / W_hr part of one LSTM layer, all steps of a pass per launch; %.0f %% of the step" % (100.0 * st.lstm_scan_ms / max(1e-9, st.compute_ms)),
                "achieved": scan_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": scan_tf / peak_tf,
                "peak_source": pk_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                "traffic": tr["bytes"] if tr else None,
                "traffic_note": ("dram bytes read + written by one launch, from the committed ncu page %s" % tr["file"]) if tr else "no ncu --set full capture of this kernel committed under profiles/ yet",
                "algorithmic_note": "FLOPs per launch = steps x sessions x 2 x (4H x D + D x H) (the reference's W_hh and W_hr products, counted once; the kernel issues 3 fp16 products per algorithmic product over W_eff = W_hh W_hr); duration = CUDA events on the engine stream around every launch of the timed steps",
                "us_per_layer_step": 1e6 * scan_s / layer_steps,
                "lstm_gemm_subset": {"what": "W_ih + W_hh + W_hr of all layers (113.25 MFLOP per session-step): the scan launches plus the W_ih and W_hr GEMM launches",
                                     "achieved": lstm_tf, "frac": lstm_tf / peak_tf, "ms_per_step": 1e3 * (scan_s + lgemm_s) / steps},
                "gemm": {"kernel": "tc::gemm_tc_kernel<3,*,2> time-batched projections (W_ih, W_hr out, FF1, FF2, conv3, conv-out, enc_proj), CTA pairs 256x256x64",
                         "achieved": gemm_tf, "frac": gemm_tf / peak_tf, "ms_per_step": st.gemm_ms / steps, "note": "3 MMAs per algorithmic product: the tensor pipe does 3x this figure"},
                "ms_per_step_by_class": {"lstm_scan": st.lstm_scan_ms / steps, "gemm": st.gemm_ms / steps, "of_which_lstm_gemm": st.lstm_gemm_ms / steps,
                                         "joiner_greedy": st.greedy_ms / steps, "fbank": st.frontend_ms / steps},
                "whole_pass": {"achieved": whole_tf, "frac": whole_tf / peak_tf, "what": "168.39 MFLOP of encoder per session-step over the whole device-resident pass"},
                "precision": "fp16x3 split (3 tensor-core products per algorithmic product, fp32 accumulation in TMEM); algorithmic FLOPs counted once"}

    def stream_block(S, ticks):
        n = 16000 * 12
        pcms = gen_streams(range(10_000, 10_000 + S), n)
        st = _ffi.AprilB200Stats()
        lib.c.aab_get_stats(m.handle, ctypes.byref(st), 1)
        lat, cnt = SD.stream(lib, m.handle, pcms, 640, 20, ticks, 0)
        lib.c.aab_get_stats(m.handle, ctypes.byref(st), 1)
        per_tick = np.sort(lat.max(axis=1))                              # a tick is done when its slowest stream has its callbacks
        flat = np.sort(lat.reshape(-1))
        q = lambda v, p: float(v[min(len(v) - 1, int(len(v) * p))])
        return {"workload": "%d synchronous sessions, one native thread each, 40 ms (640-sample) chunks through plain aas_feed_pcm16, all streams' chunks arriving together (BASELINE configs[3])" % S,
                "sessions": S, "ticks": ticks, "latency_ms": {"p50": q(flat, 0.5), "p99": q(flat, 0.99), "max": float(flat[-1])},
                "latency_note": "per chunk and stream: aas_feed_pcm16 returned - entered; a synchronous session's handler runs before the call returns",
                "tick_ms": {"p50": q(per_tick, 0.5), "p99": q(per_tick, 0.99)},
                "x_realtime_per_stream": 40.0 / q(per_tick, 0.5), "frames_per_sec": S * 4 / (float(per_tick.mean()) / 1e3),
                "device_passes": int(st.passes), "mean_sessions_per_pass": st.pass_sessions / max(1, st.passes),
                "device_ms_per_pass": st.device_ms / max(1, st.passes), "of_which_after_upload_ms": st.compute_ms / max(1, st.passes),   # CUDA events on the engine stream: the rest of a chunk's latency is host (coalescing window, planning, staging, wake-ups, callbacks)
                "callbacks": cnt}

    if a.mode == "stream":
        sb = stream_block(a.sessions if a.sessions != 64 else 256, a.ticks)
        if rank == 0:
            print(json.dumps({"metric": "chunk_latency_ms", "value": sb["latency_ms"]["p50"], "unit": "ms", "n_gpus": world, "steps": a.ticks, "warmup": 20, "higher_is_better": False,
                              "ms_per_step": sb["tick_ms"]["p50"], "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": sb["workload"]}, "stream": sb}))
        m.close()
        return

    if a.mode == "soak":
        # BASELINE configs[4]: 128 x 1 h over 8 GPUs = 16 long-form sessions per GPU. Each session loops a 60 s synthetic clip; every
        # feed call advances all sessions by --soak-chunk seconds, so LSTM state, feature carry and greedy context survive
        # minutes * 60 / chunk passes. Throughput = frames over the device time of all passes (max over ranks).
        S = a.sessions if a.sessions != 64 else 16
        clip = gen_streams(dist_util.shard_stream_ids(rank, world, S), 16000 * 60)
        nchunk = int(16000 * a.soak_chunk)
        calls = int(a.soak_minutes * 60 / a.soak_chunk)
        counters = (ctypes.c_uint64 * 4)()
        ss = []
        for _ in range(S):
            cfg = A.AprilConfig(); cfg.handler = handler; cfg.userdata = ctypes.cast(counters, ctypes.c_void_p); cfg.flags = 0
            ss.append(lib.c.aas_create_session(m.handle, cfg))
        arr = (ctypes.c_void_p * S)(*ss)
        cnts = (ctypes.c_size_t * S)(*([nchunk] * S))
        st = _ffi.AprilB200Stats()
        lib.c.aab_get_stats(m.handle, ctypes.byref(st), 1)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(calls):
            off = (k * nchunk) % (16000 * 60 - nchunk)
            ptrs = (ctypes.POINTER(ctypes.c_short) * S)(*[ctypes.cast(p.ctypes.data + 2 * off, ctypes.POINTER(ctypes.c_short)) for p in clip])
            assert lib.c.aab_feed_pcm16_batch(arr, ptrs, cnts, S) == 0
        assert lib.c.aab_flush_batch(arr, S) == 0
        wall = time.perf_counter() - t0
        lib.c.aab_get_stats(m.handle, ctypes.byref(st), 1)
        for h in ss:
            lib.c.aas_free(h)
        (wall_s, dev_s), (frames,) = dist_util.combine(dist, "cuda", [wall, st.device_ms / 1e3], [st.frames])
        if rank == 0:
            print(json.dumps({"metric": "audio_frames_per_sec", "mode": "soak", "value": frames / dev_s, "unit": "frames/s", "n_gpus": world, "steps": calls, "warmup": 0,
                              "ms_per_step": 1e3 * dev_s / calls, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": "%d long-form sessions per GPU x %.0f min, %.0f s per feed call (%d calls + flush), state carried across all calls (BASELINE configs[4] per GPU)" % (S, a.soak_minutes, a.soak_chunk, calls)},
                              "x_realtime_per_stream": a.soak_minutes * 60 / dev_s, "e2e": {"value": frames / wall_s, "unit": "frames/s", "x_realtime_per_stream": a.soak_minutes * 60 / wall_s},
                              "results": {"encoder_steps": int(st.encoder_steps), "final_tokens": int(counters[3]), "final_callbacks": int(counters[1]), "silence_callbacks": int(counters[2])}}))
        m.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    if a.lanes > 0:
        # single process, all GPUs: the batched calls span the lanes, whose passes run concurrently; times are the max over lanes (aab_get_stats)
        S = a.sessions * a.lanes
        r = run_bulk(S, a.seconds, a.warmup, a.steps, list(range(S)))
        st = r["st"]
        frames = S * FRAMES_PER_STREAM(r["n"]) * a.steps
        print(json.dumps({"metric": "audio_frames_per_sec", "mode": "lanes", "value": frames / (st.compute_ms / 1e3), "unit": "frames/s", "n_gpus": a.lanes, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": st.compute_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16x3-split tensor-core products)", "data": "synthetic",
                          "config": {"workload": workload, "sessions_per_gpu": a.sessions, "layout": "ONE process, %d GPUs, sessions placed by the library (least-loaded lane), per-lane passes of one batched call run concurrently; no collective" % a.lanes},
                          "e2e": {"value": frames / sum(r["walls"]), "unit": "frames/s", "ms_per_step": 1e3 * sum(r["walls"]) / a.steps, "h2d_bytes_per_step": st.h2d_bytes / a.steps, "d2h_bytes_per_step": st.d2h_bytes / a.steps},
                          "gpu_launches": int(st.kernel_launches), "model_load_s": t_load, "clocks": r["clocks"], "final_tokens": r["counters"][3]}))
        m.close()
        return
    r = run_bulk(a.sessions, a.seconds, a.warmup, a.steps, dist_util.shard_stream_ids(rank, world, a.sessions))
    st = r["st"]
    (wall_s, compute_s), (frames_step,) = dist_util.combine(dist, "cuda", [sum(r["walls"]), st.compute_ms / 1e3], [a.sessions * FRAMES_PER_STREAM(r["n"])])
    frames_total = frames_step * a.steps
    out = None
    if rank == 0:
        out = {
            "metric": "audio_frames_per_sec", "value": frames_total / compute_s, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 * compute_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16x3-split tensor-core products, fp32 accumulate and fp32 everywhere else)", "data": "synthetic",
            "config": {"workload": workload, "sessions_per_gpu": a.sessions, "seconds_per_session": a.seconds,
                       "parallelism": "sessions sharded, %d per GPU, no collective on the data path" % a.sessions,
                       "l2": "working set (337 MB fp32 weights + GBs of activations per pass) exceeds the 126 MB L2; no explicit flush",
                       "timing": "value: device-resident part of each pass (CUDA events on the engine stream, after the PCM upload), max over ranks; e2e: wall clock around the C-ABI calls with host buffers (H2D of the inputs, compute, D2H of the decisions, result replay, callbacks)",
                       "inputs": "pinned host memory (aab_pin_host_memory once per buffer, outside the timed region): DMA straight from the caller's buffers" if r["pinned_inputs"] else "pageable host memory: the library copies each feed into its pinned staging buffer, then H2D"},
            "x_realtime_per_stream": a.seconds / (compute_s / a.steps),
            "e2e": {"value": frames_total / wall_s, "unit": "frames/s", "ms_per_step": 1000.0 * wall_s / a.steps,
                    "h2d_bytes_per_step": st.h2d_bytes / a.steps, "d2h_bytes_per_step": st.d2h_bytes / a.steps,
                    "of_which_upload_ms": (st.device_ms - st.compute_ms) / a.steps,      # pinned staging + H2D before the first kernel (this rank)
                    "of_which_host_ms": 1000.0 * sum(r["walls"]) / a.steps - st.device_ms / a.steps,   # planning, result replay, callbacks (this rank)
                    "x_realtime_per_stream": a.seconds / (wall_s / a.steps)},
            "gpu_launches": int(st.kernel_launches), "model_load_s": t_load,
            "clocks": r["clocks"],
            "comm": {"backend": "nccl" if world > 1 else None, "nranks": world, "use": "start/stop barrier and max-over-ranks of the timings only"},
            "roofline": roofline_of(r, a.steps),
            "results": {"partial_callbacks": r["counters"][0], "final_callbacks": r["counters"][1], "silence_callbacks": r["counters"][2], "final_tokens": r["counters"][3],
                        "encoder_steps_per_pass": st.encoder_steps / a.steps, "joiner_passes_per_pass": st.joiner_passes / a.steps},
        }
    if world == 1 and not a.no_extra:
        # north_star's operating point on one GPU, and configs[3], in the same process (weights already resident)
        if a.sessions != 512:
            rb = run_bulk(512, a.seconds, 1, 2, range(512))
            sb_st, steps_b = rb["st"], 2
            comp = sb_st.compute_ms / 1e3
            fr = 512 * FRAMES_PER_STREAM(rb["n"]) * steps_b
            out["b512"] = {"workload": "512 concurrent sessions x %.0f s on this one GPU (north_star operating point; configs[2] puts 64 on each of 8 GPUs instead)" % a.seconds,
                           "value": fr / comp, "unit": "frames/s", "steps": steps_b, "warmup": 1, "ms_per_step": 1e3 * comp / steps_b, "x_realtime_per_stream": a.seconds / (comp / steps_b),
                           "e2e": {"value": fr / sum(rb["walls"]), "unit": "frames/s", "ms_per_step": 1e3 * sum(rb["walls"]) / steps_b,
                                   "h2d_bytes_per_step": sb_st.h2d_bytes / steps_b, "d2h_bytes_per_step": sb_st.d2h_bytes / steps_b},
                           "clocks": rb["clocks"], "roofline": roofline_of(rb, steps_b), "final_tokens": rb["counters"][3]}
        out["stream"] = stream_block(256, a.ticks)
    if rank == 0:
        if not a.no_cpu_baseline and world == 1 and os.path.exists(ORACLE_SO):
            arm = CpuArm(path)
            nthr = max(1, min(ncores, 32))
            cb = arm.measure(nthr, a.cpu_budget or 20.0, a.seconds, first_id=5000)
            one = arm.measure(1, (a.cpu_budget or 20.0) / 2, 10.0, first_id=7000)           # configs[0]: one session, one core, 10 s clip
            cb["single_session"] = {"value": one["value"], "unit": "frames/s", "x_realtime": one["per_session_x_realtime"], "sample": one["sample"],
                                    "workload": "BASELINE configs[0]: 1 session, 10 s clip, 1 core"}
            arm.close()
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    m.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
