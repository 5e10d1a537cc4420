#!/usr/bin/env python
"""Full-length parity check at BASELINE configs[1]'s size: S sessions x T seconds through the batched product path
(GPU), the first K of them through the oracle (oracle/_ref: the reference's host code over the CPU fp32 interpreter,
one thread per session). Prints one JSON line: per checked session whether the whole callback sequence (result type,
token, flags, time_ms) is identical, and where it first differs if not.  TEST INFRASTRUCTURE: uses oracle/."""
import argparse, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import common as C  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sessions", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--check", type=int, default=8)
    ap.add_argument("--model", default=None)
    a = ap.parse_args()
    path = a.model or C.model_path("full")
    n = int(16000 * a.seconds)
    pcms = [C.gen_stream(i, n) for i in range(a.sessions)]
    t0 = time.perf_counter()
    got = C.run_product_batch(path, pcms)
    t_gpu = time.perf_counter() - t0
    want = [None] * a.check

    def work(i):
        want[i] = C.run_oracle(path, pcms[i])
    ts = [threading.Thread(target=work, args=(i,)) for i in range(a.check)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    t_cpu = time.perf_counter() - t0
    rows = []
    for i in range(a.check):
        g, w = got[i].keys(), want[i].keys()
        k = next((j for j, (x, y) in enumerate(zip(g, w)) if x != y), None)
        if k is None and len(g) != len(w): k = min(len(g), len(w))
        nd, fin_ok = C.callback_diff(got[i], want[i])
        rows.append({"session": i, "callbacks_gpu": len(g), "callbacks_oracle": len(w), "identical": k is None, "first_diff_callback": k, "callbacks_differing": nd, "finals_identical": fin_ok,
                     "final_tokens": sum(len(t) for r, t in w if r == 2)})
    print(json.dumps({"check": "full-length parity", "sessions": a.sessions, "seconds": a.seconds, "checked": a.check,
                      "identical": sum(r["identical"] for r in rows), "finals_identical": sum(r["finals_identical"] for r in rows), "gpu_wall_s": round(t_gpu, 2), "oracle_wall_s": round(t_cpu, 2), "rows": rows}))


if __name__ == "__main__":
    main()
