#!/bin/bash
set -u
O=gpurun_out/r2t; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1
one() {  # sessions, env...
  B=$1; shift
  env "$@" timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra --sessions $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=r['ms_per_step_by_class']; print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'scan_ms', round(c['lstm_scan'],2), 'gemm_ms', round(c['gemm'],2), 'greedy_ms', round(c['joiner_greedy'],2), 'tokens', d['results']['final_tokens'])
    elif 'rror' in l or 'timeout' in l: print(l.rstrip()[:200])
" | tail -2; }
for V in 0 2 3 4 5; do echo "== B=64 overlap=$V (halving chunks)"; one 64 APRIL_B200_OVERLAP=$V APRIL_B200_TIMING=1; done > $O/sweep.log 2>&1
for V in 0 4; do echo "== B=32 overlap=$V"; one 32 APRIL_B200_OVERLAP=$V; echo "== B=8 overlap=$V"; one 8 APRIL_B200_OVERLAP=$V; done >> $O/sweep.log 2>&1
cat $O/sweep.log
timeout 1200 python -m pytest tests -m gpu -q -x -k "search_schedule or full_length or long_form or golden" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log; tail -6 $O/pytest_sel.log
