#!/usr/bin/env python
"""Dev: where do the pair scan kernel's m rows differ from scan4's? Runs one pass of B sessions x `seconds` under both and compares the
dumped layer-0 m planes (APRIL_B200_DUMP_M) per (step, session, unit)."""
import argparse, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="mid"); ap.add_argument("--sessions", type=int, default=150); ap.add_argument("--seconds", type=float, default=2.0); ap.add_argument("--ragged", type=int, default=0); ap.add_argument("--tap", type=int, default=0); ap.add_argument("--flush", type=int, default=0); ap.add_argument("--layer", type=int, default=0); ap.add_argument("--a", default=""); ap.add_argument("--b", default="APRIL_B200_SCAN=pair"); ap.add_argument("--emulate", type=int, default=-1, help="layer whose W_eff the dumped layer used: check both arms against an fp64 emulation"); ap.add_argument("--child", default=None)
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
    npl = 2 * T1 * B * H
    out[arm] = np.frombuffer(raw[16:16 + 2 * npl], np.float16).reshape(2, T1, B, H).astype(np.float32)
    out[arm + "_G"] = np.frombuffer(raw[16 + 2 * npl:], np.float32).reshape(T1 - 1, B, 4, H)
if a.emulate >= 0:
    import make_synthetic_april as MS
    Wt = MS.make_weights(MS.PRESETS[a.model], 1234)
    weff = Wt["l%d_w_hh" % a.emulate].astype(np.float64) @ Wt["l%d_w_hr" % a.emulate].astype(np.float64)      # [4H, H]
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    for arm in ("scan4", "pair"):
        G = out[arm + "_G"].astype(np.float64); P = out[arm].astype(np.float64)
        c = np.zeros((B, H)); m = P[0][0] + P[1][0]
        print("emulation of arm %s (fp64, W_eff of layer %d): G identical in both arms: %s" % (arm, a.emulate, np.array_equal(out["scan4_G"], out["pair_G"])))
        for t in range(min(4, T1 - 1)):
            pre = G[t] + np.einsum("gjk,bk->bgj", weff.reshape(4, H, H), m)
            c = sig(pre[:, 1]) * c + sig(pre[:, 0]) * np.tanh(pre[:, 2]); mm = sig(pre[:, 3]) * np.tanh(c)
            got = P[0][t + 1] + P[1][t + 1]
            err = np.abs(mm - got)
            print("   step %d: max |emulated - arm| = %.3g (units above 1e-3: %d of %d)" % (t, err.max(), int((err.max(axis=0) > 1e-3).sum()), H))
            if err.max() > 1e-3:
                # which part of the product is missing? compare the pre-activation error with partial products
                full = np.einsum("gjk,bk->bgj", weff.reshape(4, H, H), m); lo_half = np.einsum("gjk,bk->bgj", weff.reshape(4, H, H)[:, :, :H // 2], m[:, :H // 2])
                print("      per-unit max err (first 32):", np.round(err.max(axis=0)[:32], 3).tolist())
                print("      |W m| max %.3g; |W m (K half 0 only)| max %.3g" % (np.abs(full).max(), np.abs(lo_half).max()))
            m = got; # follow the arm's own state so that each step is checked on its own
            # c follows the emulation (the arm's c is not dumped): after a bad step the later steps are not meaningful
for k in list(out):
    if k.endswith("_G"): del out[k]
print("nan counts (hi plane) a/b:", int(np.isnan(out["scan4"][0]).sum()), int(np.isnan(out["pair"][0]).sum()))
for k in out: out[k] = np.nan_to_num(out[k], nan=9.0)
dl = np.abs(out["scan4"][1] - out["pair"][1]); print("LO plane: max diff", dl.max(), "first bad row block (> 1e-6)", next((t for t in range(dl.shape[0]) if dl[t].max() > 1e-6), None), "max |lo|", np.abs(out["pair"][1]).max(), np.abs(out["scan4"][1]).max())
d = np.abs(out["scan4"][0] - out["pair"][0])
print("config", vars(a)); print("planes", out["scan4"][0].shape, "max diff", d.max(), "first bad row block", next((t for t in range(d.shape[0]) if d[t].max() > 1e-3), None))
for t in range(min(4, d.shape[0])):
    dt = d[t]
    bad_units = np.where(dt.max(axis=0) > 1e-3)[0]; bad_sess = np.where(dt.max(axis=1) > 1e-3)[0]
    print("row block", t, "max", dt.max(), "bad units", len(bad_units), bad_units[:48].tolist(), "bad sessions", len(bad_sess), bad_sess[:20].tolist(), "...", bad_sess[-5:].tolist())
    if len(bad_units): print("   per-unit max diff (first 64 units):", np.round(dt.max(axis=0)[:64], 3).tolist())
