#!/usr/bin/env python
"""Dev: where do the pair scan kernel's m rows differ from scan4's? Runs one pass of B sessions x `seconds` under both and compares the
dumped layer-0 m planes (APRIL_B200_DUMP_M) per (step, session, unit)."""
import argparse, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="mid"); ap.add_argument("--sessions", type=int, default=150); ap.add_argument("--seconds", type=float, default=2.0); ap.add_argument("--ragged", type=int, default=0); ap.add_argument("--tap", type=int, default=0); ap.add_argument("--flush", type=int, default=0); ap.add_argument("--layer", type=int, default=0); ap.add_argument("--a", default=""); ap.add_argument("--b", default="APRIL_B200_SCAN=pair"); ap.add_argument("--child", default=None)
a = ap.parse_args()
if a.child:
    import common as C
    C.run_product_batch(C.model_path(a.model), [C.gen_stream(700 + i, int(16000 * a.seconds) + (977 * i if a.ragged else 0)) for i in range(a.sessions)], flush=bool(a.flush), tap=bool(a.tap))
    sys.exit(0)
out = {}
for arm, env in (("scan4", dict(kv.split("=", 1) for kv in a.a.split())), ("pair", dict(kv.split("=", 1) for kv in a.b.split()))):
    f = tempfile.mktemp(suffix=".bin")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--model", a.model, "--sessions", str(a.sessions), "--seconds", str(a.seconds), "--ragged", str(a.ragged), "--tap", str(a.tap), "--flush", str(a.flush), "--child", "1"],
                       env=dict(os.environ, APRIL_B200_DUMP_M=f, APRIL_B200_DUMP_LAYER=str(a.layer), **env), capture_output=True, text=True, timeout=600)
    if r.returncode != 0: print(arm, "failed", r.stderr[-500:]); sys.exit(1)
    raw = open(f, "rb").read(); hdr = np.frombuffer(raw[:16], np.int32); T1, B, H = int(hdr[0]), int(hdr[1]), int(hdr[2])
    out[arm] = np.frombuffer(raw[16:], np.float16).reshape(2, T1, B, H).astype(np.float32)
dl = np.abs(out["scan4"][1] - out["pair"][1]); print("LO plane: max diff", dl.max(), "first bad row block (> 1e-6)", next((t for t in range(dl.shape[0]) if dl[t].max() > 1e-6), None), "max |lo|", np.abs(out["pair"][1]).max(), np.abs(out["scan4"][1]).max())
d = np.abs(out["scan4"][0] - out["pair"][0])
print("config", vars(a)); print("planes", out["scan4"][0].shape, "max diff", d.max(), "first bad row block", next((t for t in range(d.shape[0]) if d[t].max() > 1e-3), None))
for t in range(min(4, d.shape[0])):
    dt = d[t]
    bad_units = np.where(dt.max(axis=0) > 1e-3)[0]; bad_sess = np.where(dt.max(axis=1) > 1e-3)[0]
    print("row block", t, "max", dt.max(), "bad units", len(bad_units), bad_units[:48].tolist(), "bad sessions", len(bad_sess), bad_sess[:20].tolist(), "...", bad_sess[-5:].tolist())
    if len(bad_units): print("   per-unit max diff (first 64 units):", np.round(dt.max(axis=0)[:64], 3).tolist())
