"""Seeded synthetic 16 kHz s16 PCM: band-limited noise bursts, silences
long enough for the 2200 ms silence path (src/april_session.c:411), a few sine sweeps.

Deterministic: numpy legacy RandomState(seed) only. Test/bench infrastructure.
"""
import numpy as np

BASE_SEED = 0x5EED0000


def gen_stream(stream_id, n_samples, sr=16000):
    rs = np.random.RandomState((BASE_SEED + int(stream_id)) & 0x7FFFFFFF)
    out = np.zeros(n_samples, dtype=np.float64)
    pos = int(rs.randint(0, sr // 4))
    kind_cycle = 0
    while pos < n_samples:
        r = rs.rand()
        if r < 0.55:                                    # noise burst 0.3 .. 1.6 s
            n = int(sr * (0.3 + 1.3 * rs.rand()))
            w = rs.randn(n) * 3000.0
            # cheap band-limiting: one-pole low-pass with random corner + slow AM
            a = 0.25 + 0.7 * rs.rand()
            y = np.empty(n)
            acc = 0.0
            # vectorised one-pole via cumulative trick is unstable; use lfilter-free loop in chunks
            from_scipy = _onepole(w, a)
            y = from_scipy
            am = 0.6 + 0.4 * np.sin(2 * np.pi * (2 + 6 * rs.rand()) * np.arange(n) / sr + rs.rand() * 6.28)
            seg = y * am * (1.5 / max(a, 0.3))
        elif r < 0.75:                                  # sine sweep 0.2 .. 0.8 s
            n = int(sr * (0.2 + 0.6 * rs.rand()))
            f0, f1 = 150 + 800 * rs.rand(), 500 + 3000 * rs.rand()
            t = np.arange(n) / sr
            ph = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / (n / sr))
            seg = 6000.0 * np.sin(ph) * np.hanning(n)
        else:                                           # silence; every 3rd one is > 2.5 s
            kind_cycle += 1
            dur = 2.6 + 0.8 * rs.rand() if kind_cycle % 3 == 0 else 0.1 + 0.6 * rs.rand()
            n = int(sr * dur)
            seg = rs.randn(n) * (0.0 if rs.rand() < 0.5 else 2.0)
        m = min(n, n_samples - pos)
        out[pos:pos + m] += seg[:m]
        pos += n
    return np.clip(np.round(out), -32768, 32767).astype(np.int16)


def _onepole(x, a):
    try:
        from scipy.signal import lfilter
        return lfilter([a], [1.0, -(1.0 - a)], x)
    except Exception:                                   # pragma: no cover
        y = np.empty_like(x)
        acc = 0.0
        for i, v in enumerate(x):
            acc = a * v + (1.0 - a) * acc
            y[i] = acc
        return y


def _gen_job(args):
    return gen_stream(*args)


def gen_streams(stream_ids, n_samples, workers=None):
    """Many streams: worker processes (spawned, so it is safe after CUDA has been initialised) when the job is big."""
    ids = list(stream_ids)
    if len(ids) * n_samples < 64 * 480000 + 1:
        return [gen_stream(i, n_samples) for i in ids]
    import multiprocessing as mp
    import os
    workers = workers or max(2, min(16, (os.cpu_count() or 4)))
    with mp.get_context("spawn").Pool(workers) as pool:
        return pool.map(_gen_job, [(i, n_samples) for i in ids], chunksize=8)


def gen_batch(n_streams, n_samples, first_id=0):
    return np.stack([gen_stream(first_id + i, n_samples) for i in range(n_streams)])


if __name__ == "__main__":
    import sys
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
    x = gen_stream(0, n)
    print(x[:16], float(np.abs(x).mean()))
