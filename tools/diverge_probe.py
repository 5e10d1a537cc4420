#!/usr/bin/env python
"""Dev probe (TEST INFRASTRUCTURE, uses oracle/): per-step error of the encoder output and joiner logits between the GPU
path and the oracle over one long stream, and the logit margins at the first differing greedy decision."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import common as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--stream", type=int, default=3)
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--model", default="full")
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--samples", type=int, default=0)
ap.add_argument("--batch", type=int, default=1, help="run the stream as member of a batch of this many sessions (stream ids 0..batch-1)")
a = ap.parse_args()
path = C.model_path(a.model)
nsamp = a.samples or int(16000 * a.seconds)
pcm = C.gen_stream(a.stream, nsamp)
if a.batch > 1:
    pcms = [C.gen_stream(i, nsamp) for i in range(a.batch)]
    g = C.run_product_batch(path, pcms, tap=True, chunk=a.chunk or None)[a.stream]
else:
    g = C.run_product_batch(path, [pcm], tap=True, chunk=a.chunk or None)[0]
w = C.run_oracle(path, pcm, tap=True, chunk=a.chunk or None)
n = min(len(g.enc), len(w.enc))
err = np.array([np.abs(g.enc[t] - w.enc[t]).max() for t in range(n)])
mag = np.array([np.abs(w.enc[t]).max() for t in range(n)])
print("enc steps", len(g.enc), len(w.enc), "max abs err", err.max(), "at", int(err.argmax()), "typ |enc|", float(np.median(mag)))
print("enc err by 100-step block:", [float("%.2e" % err[i:i + 100].max()) for i in range(0, n, 100)])
m = min(len(g.logits), len(w.logits))
first = None
for i in range(m):
    lg, lw = g.logits[i], w.logits[i]
    e = np.abs(lg - lw).max()
    if e > 1e-2:
        first = i; break
print("joiner passes", len(g.logits), len(w.logits), "first pass with |dlogit| > 1e-2:", first)
upto = first if first is not None else m
le = np.array([np.abs(g.logits[i] - w.logits[i]).max() for i in range(upto)])
print("logit err before that: max %.3e median %.3e" % (le.max() if len(le) else 0, np.median(le) if len(le) else 0))
if first is not None and first > 0:
    for i in range(max(0, first - 3), first):
        lw = w.logits[i]; lg = g.logits[i]
        o = np.argsort(lw)[::-1]
        print("pass", i, "oracle top3", [(int(k), float(lw[k])) for k in o[:3]], "gpu same idx", [float(lg[k]) for k in o[:3]], "blank(0)", float(lw[0]), float(lg[0]), "err", float(np.abs(lg - lw).max()))
gk, wk = g.keys(), w.keys()
k = next((j for j, (x, y) in enumerate(zip(gk, wk)) if x != y), None)
print("callbacks", len(gk), len(wk), "first differing callback:", k)
if k is not None:
    for j in range(max(0, k - 1), min(len(gk), len(wk), k + 2)):
        print(" gpu   ", j, gk[j][0], len(gk[j][1]), gk[j][1][:3], "...", gk[j][1][-4:])
        print(" oracle", j, wk[j][0], len(wk[j][1]), wk[j][1][:3], "...", wk[j][1][-4:])
