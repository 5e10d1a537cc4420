#!/bin/bash
# Dev: one bench line per configuration, condensed. Usage on the GPU box:  . tools/sweep.sh; one 64 APRIL_B200_OVERLAP=0; one 512 APRIL_B200_GREEDY=cluster
one() {  # sessions, VAR=VALUE...
  B=$1; shift
  env APRIL_B200_TIMING=1 "$@" timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra --sessions $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('scan trace') or l.startswith('  '): print(l.rstrip())
    elif l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=r['ms_per_step_by_class']
        print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'scan_ms', round(c['lstm_scan'],2), 'us/layer-step', round(r['us_per_layer_step'],3), 'gemm_ms', round(c['gemm'],2), 'search_ms', round(c['joiner_greedy'],2), 'fbank', round(c['fbank'],2), 'tokens', d['results']['final_tokens'])
    elif 'rror' in l or 'timeout' in l: print(l.rstrip()[:200])
" | tail -3; }
