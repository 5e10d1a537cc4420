#!/bin/bash
# Round 2, GPU call D: scan3 (<= 64 sessions) + scan4 dispatch on the shared m planes; greedy_tc phase trace; CUDA graphs + coalescing
# for the streaming block; whole -m gpu suite; default bench line.
set -u
O=gpurun_out/r2d; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1

timeout 300 python tools/scan_check.py --model tiny --sessions 3 --seconds 5 --oracle 1 > $O/check_tiny.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 48 --seconds 5 --oracle 1 > $O/check_full48.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 200 --seconds 5 > $O/check_full200.log 2>&1
tail -n 3 $O/check_*.log

run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra $BARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('scan trace') or l.startswith('  ') or l.startswith('greedy trace'): print(l.rstrip())
    elif l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=r['ms_per_step_by_class']; print('$tag', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'scan_ms', round(c['lstm_scan'],2), 'us/layer-step', round(r['us_per_layer_step'],3), 'gemm_ms', round(c['gemm'],2), 'lstm_gemm_ms', round(c['of_which_lstm_gemm'],2), 'greedy_ms', round(c['joiner_greedy'],2), 'fbank_ms', round(c['fbank'],2), 'tokens', d['results']['final_tokens'])
    elif 'timeout' in l or 'rror' in l or 'libapril' in l: print(l.rstrip())
" | tail -6; }
{
BARGS="--sessions 64"
run d_b64 APRIL_B200_SCAN_TRACE=1
run d_gt_b64 APRIL_B200_GREEDY=tc APRIL_B200_GREEDY_TRACE=1
BARGS="--sessions 32"
run d_b32 A=1
BARGS="--sessions 96"
run d_b96 A=1
BARGS="--sessions 512"
run d_b512 A=1
run d_gt_b512 APRIL_B200_GREEDY=tc APRIL_B200_GREEDY_TRACE=1
} > $O/sweep.log 2>&1
cat $O/sweep.log

timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
tail -n 30 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json')); print('default', d['ms_per_step'], d['e2e']['ms_per_step']); print('b512', d['b512']['ms_per_step'], d['b512']['e2e']['ms_per_step']); print('stream', d['stream']['latency_ms'], d['stream']['tick_ms'], d['stream']['device_passes'], d['stream']['mean_sessions_per_pass'])"
tail -n 5 $O/bench_default.err
timeout 300 env APRIL_B200_GRAPHS=0 python bench.py --mode stream --sessions 256 > $O/stream_nograph.json 2> $O/stream_nograph.err
python -c "
import json; d=json.load(open('$O/stream_nograph.json')); s=d['stream']; print('stream no graphs', s['latency_ms'], s['tick_ms'], s['device_passes'], s['mean_sessions_per_pass'])"
