run() { env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('scan trace') or l.startswith('  '): print(l.rstrip())
    elif l.startswith('{'):
        d=json.loads(l); print('ms/step', round(d['ms_per_step'],2), 'scan_ms', round(d['roofline']['lstm_scan_ms_per_step'],2), 'us/layer-step', round(d['roofline']['us_per_layer_step'],3), 'gemm_ms', round(d['roofline']['gemm_ms_per_step'],2), 'greedy_ms', round(d['roofline'].get('greedy_ms_per_step',0),2), 'tokens', d['results']['final_tokens'])
    elif 'timeout' in l or 'rror' in l: print(l.rstrip())
" | tail -5; }
