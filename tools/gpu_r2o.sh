#!/bin/bash
set -u
O=gpurun_out/r2o; mkdir -p $O
export APRIL_B200_TIMING=1
timeout 600 python tools/scan_check.py --model full --sessions 300 --seconds 3 --a "" --b "APRIL_B200_SCAN=pair" --oracle 1 > $O/pair_full300.log 2>&1; tail -3 $O/pair_full300.log
timeout 600 python tools/scan_check.py --model mid --sessions 520 --seconds 3 --a "" --b "APRIL_B200_SCAN=pair" --oracle 1 > $O/pair_mid520.log 2>&1; tail -3 $O/pair_mid520.log
one() {  # sessions, env...
  B=$1; shift
  env "$@" timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --sessions $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=r['ms_per_step_by_class']; print('ms/step', round(d['ms_per_step'],2), 'scan_ms', round(c['lstm_scan'],2), 'us/layer-step', round(r['us_per_layer_step'],3), 'gemm_ms', round(c['gemm'],2), 'greedy_ms', round(c['joiner_greedy'],2), 'fbank', round(c['fbank'],2), 'tokens', d['results']['final_tokens'])
    elif 'rror' in l: print(l.rstrip()[:200])
" | tail -2; }
for B in 160 256 384 512; do
  echo "== B=$B scan4"; one $B X=1
  echo "== B=$B pair";  one $B APRIL_B200_SCAN=pair
done > $O/sweep.log 2>&1
cat $O/sweep.log
