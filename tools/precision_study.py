#!/usr/bin/env python
"""Where does the tensor-core tier's logit error come from, and how often does it flip a greedy decision?
Runs N streams of the full model once through the oracle (taps: encoder output per step, raw joiner logits per pass) and
through the product under several engine configurations (one process each: the switches are read at model load), then reports
per configuration the max / median error of encoder_out and logits, and the number of greedy decisions whose margin
(|best non-blank - blank threshold| as april_session.c:329-358 computes it) is smaller than twice that configuration's
logit error: the decisions that CAN flip.  TEST INFRASTRUCTURE: uses oracle/_ref.

  python tools/precision_study.py --streams 2 --seconds 10 --out gpurun_out/precision.json
"""
import argparse, json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))

CONFIGS = [
    ("default (fp16x3 tcgen05 GEMMs + scan)", ""),
    ("tcgen05 GEMMs, fp32 SIMT recurrence", "APRIL_B200_SCAN=simt"),
    ("fp32 SIMT tier", "APRIL_B200_PRECISION=fp32"),
]


def child(a):
    import numpy as np
    import common as C
    path = C.model_path(a.model)
    pcms = [C.gen_stream(900 + i, int(16000 * a.seconds)) for i in range(a.streams)]
    cols = C.run_product_batch(path, pcms, tap=True)
    out = {}
    for i, c in enumerate(cols):
        out["enc%d" % i] = np.stack(c.enc); out["lg%d" % i] = np.stack(c.logits); out["keys%d" % i] = np.array(json.dumps(c.keys()))
    np.savez(a.child, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="full")
    ap.add_argument("--streams", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child:
        return child(a)
    import numpy as np
    import common as C
    path = C.model_path(a.model)
    pcms = [C.gen_stream(900 + i, int(16000 * a.seconds)) for i in range(a.streams)]
    report = {"model": a.model, "streams": a.streams, "seconds": a.seconds, "configs": []}
    runs = []
    for name, envs in CONFIGS:
        env = dict(os.environ)
        for kv in envs.split():
            k, v = kv.split("=", 1); env[k] = v
        f = tempfile.mktemp(suffix=".npz")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--model", a.model, "--streams", str(a.streams), "--seconds", str(a.seconds), "--child", f],
                           env=env, capture_output=True, text=True, timeout=1200)
        runs.append((name, envs, dict(np.load(f)) if r.returncode == 0 else None, r.stderr[-400:]))
        if r.returncode == 0: os.unlink(f)
    want = [C.run_oracle(path, p, tap=True) for p in pcms]       # one after the other: the oracle's tap is a process-wide hook
    for name, envs, got, err in runs:
        if got is None:
            report["configs"].append({"name": name, "env": envs, "failed": err}); continue
        e_enc, e_lg, can_flip, decisions, same = [], [], 0, 0, 0
        for i in range(a.streams):
            eo, lo = np.stack(want[i].enc), np.stack(want[i].logits)
            eg, lg = got["enc%d" % i], got["lg%d" % i]
            same += int(json.dumps(want[i].keys()) == str(got["keys%d" % i]))
            if eo.shape != eg.shape or lo.shape != lg.shape:
                continue
            e_enc.append(np.abs(eo - eg).max(axis=1)); e_lg.append(np.abs(lo - lg).max(axis=1))
            # margins of the decisions the reference takes on these logits (blank id 0): blank vs best non-blank, with the
            # early-emit (1.0), punctuation (3.5) and tentative-token (4.0) offsets it can apply: the closest one counts
            best = lo[:, 1:].max(axis=1); d = lo[:, 0] - best
            margin = np.min(np.abs(np.stack([d, d - 1.0, d - 3.5, d - 4.0])), axis=0)
            top2 = np.sort(lo[:, 1:], axis=1)[:, -2:]
            margin = np.minimum(margin, top2[:, 1] - top2[:, 0])
            decisions += len(margin)
            can_flip += int((margin < 2.0 * float(np.abs(lo - lg).max())).sum())
        ee, el = np.concatenate(e_enc), np.concatenate(e_lg)
        report["configs"].append({"name": name, "env": envs, "encoder_out_max_abs_err": float(ee.max()), "encoder_out_median_err": float(np.median(ee)),
                                  "logits_max_abs_err": float(el.max()), "logits_median_err": float(np.median(el)), "joiner_passes": int(decisions),
                                  "decisions_within_2x_max_err_of_a_threshold": can_flip, "streams_with_identical_callbacks": same})
    print(json.dumps(report, indent=1))
    if a.out:
        json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
