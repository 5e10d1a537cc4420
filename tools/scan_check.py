#!/usr/bin/env python
"""Development check for a new recurrent-scan kernel: encoder outputs (tapped through aab_set_tap) of the same streams under two
engine configurations selected by environment variables (each run in its own process: the switches are read once at model load),
and optionally against the oracle.  TEST INFRASTRUCTURE: the oracle arm uses oracle/_ref.

  python tools/scan_check.py --model tiny --sessions 3 --seconds 5 --a APRIL_B200_PRECISION=fp32 --b "" --oracle 1
"""
import argparse, json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))


def child(a):
    import numpy as np
    import common as C
    path = C.model_path(a.model)
    pcms = [C.gen_stream(700 + i, int(16000 * a.seconds) + (977 * i if a.ragged else 0)) for i in range(a.sessions)]
    cols = C.run_product_batch(path, pcms, tap=bool(a.tap), chunk=a.chunk or None)
    out = {}
    for i, c in enumerate(cols):
        out["enc%d" % i] = np.stack(c.enc) if c.enc else np.zeros((0, 1), np.float32)
        out["lg%d" % i] = np.stack(c.logits) if c.logits else np.zeros((0, 1), np.float32)
        out["keys%d" % i] = np.array(json.dumps(c.keys()))
    np.savez(a.child, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--sessions", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--ragged", type=int, default=1)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--a", default="APRIL_B200_SCAN=simt", help="space-separated VAR=VALUE list of arm A")
    ap.add_argument("--b", default="", help="arm B")
    ap.add_argument("--oracle", type=int, default=0, help="also run the first N sessions through the oracle")
    ap.add_argument("--tap", type=int, default=1, help="0: compare callback logs only (the taps switch some engine paths off)")
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child:
        return child(a)
    import numpy as np
    res = {}
    for arm, envs in (("a", a.a), ("b", a.b)):
        env = dict(os.environ)
        for kv in envs.split():
            k, v = kv.split("=", 1); env[k] = v
        f = tempfile.mktemp(suffix=".npz")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--model", a.model, "--sessions", str(a.sessions), "--seconds", str(a.seconds),
                            "--ragged", str(a.ragged), "--chunk", str(a.chunk), "--tap", str(a.tap), "--child", f], env=env, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            print("arm %s (%s) FAILED rc=%d\n%s" % (arm, envs, r.returncode, (r.stdout + r.stderr)[-1500:]))
            res[arm] = None
            continue
        err = [l for l in r.stderr.splitlines() if "scan trace" in l or l.startswith("  ") or "libapril" in l]
        if err: print("\n".join(err[:12]))
        res[arm] = dict(np.load(f)); os.unlink(f)
    if res["a"] is not None and res["b"] is not None:
        worst = 0.0; nkeys = 0; worst_lg = 0.0
        for i in range(a.sessions):
            la, lb = res["a"]["lg%d" % i], res["b"]["lg%d" % i]
            if la.shape == lb.shape and la.size:
                worst_lg = max(worst_lg, float(np.abs(la - lb).max()))
            elif la.shape != lb.shape:
                print("session %d: logit shapes differ %s %s" % (i, la.shape, lb.shape))
            ea, eb = res["a"]["enc%d" % i], res["b"]["enc%d" % i]
            if ea.shape != eb.shape:
                print("session %d: shapes differ %s %s" % (i, ea.shape, eb.shape)); continue
            d = np.abs(ea - eb)
            if d.size:
                per_step = d.max(axis=1)
                worst = max(worst, float(d.max()))
                if d.max() > 1e-3:
                    first = int(np.argmax(per_step > 1e-3))
                    print("session %d: max |a-b| %.3g, first step above 1e-3: %d of %d, nan a/b: %d/%d" % (i, d.max(), first, len(per_step), int(np.isnan(ea).sum()), int(np.isnan(eb).sum())))
            nkeys += int(str(res["a"]["keys%d" % i]) == str(res["b"]["keys%d" % i]))
        print("A=[%s] vs B=[%s] model=%s sessions=%d: max |enc_a - enc_b| = %.3g, max |logits_a - logits_b| = %.3g; identical callback logs: %d of %d" % (a.a, a.b, a.model, a.sessions, worst, worst_lg, nkeys, a.sessions))
    if a.oracle and res["b"] is not None:
        import common as C
        path = C.model_path(a.model)
        for i in range(min(a.oracle, a.sessions)):
            pcm = C.gen_stream(700 + i, int(16000 * a.seconds) + (977 * i if a.ragged else 0))
            o = C.run_oracle(path, pcm, tap=True, chunk=a.chunk or None)
            eo, eb = np.stack(o.enc), res["b"]["enc%d" % i]
            same = json.dumps(o.keys()) == str(res["b"]["keys%d" % i])
            print("oracle vs B session %d: max |enc| err %.3g over %d steps; callbacks identical: %s" % (i, float(np.abs(eo - eb).max()) if eo.shape == eb.shape else -1, len(eo), same))


if __name__ == "__main__":
    main()
