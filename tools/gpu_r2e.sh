#!/bin/bash
# Round 2, GPU call E: pair-MMA scan (scan5) validation and timing; per-request coalescer + graphs on the streaming block; near-tie
# policy in the whole -m gpu suite; default bench line.
set -u
O=gpurun_out/r2e; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1

timeout 600 python tools/scan_check.py --model full --sessions 300 --seconds 4 --a "" --b "APRIL_B200_SCAN=pair" --oracle 1 > $O/pair_full300.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 520 --seconds 3 --a "" --b "APRIL_B200_SCAN=pair" > $O/pair_full520.log 2>&1
timeout 300 python tools/scan_check.py --model mid --sessions 150 --seconds 6 --a "" --b "APRIL_B200_SCAN=pair" --oracle 1 > $O/pair_mid150.log 2>&1
tail -n 4 $O/pair_*.log

run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra $BARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('scan trace') or l.startswith('  ') or l.startswith('greedy trace'): print(l.rstrip())
    elif l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=r['ms_per_step_by_class']; print('$tag', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'scan_ms', round(c['lstm_scan'],2), 'us/layer-step', round(r['us_per_layer_step'],3), 'gemm_ms', round(c['gemm'],2), 'lstm_gemm_ms', round(c['of_which_lstm_gemm'],2), 'greedy_ms', round(c['joiner_greedy'],2), 'fbank_ms', round(c['fbank'],2), 'tokens', d['results']['final_tokens'])
    elif 'timeout' in l or 'rror' in l or 'libapril' in l: print(l.rstrip())
" | tail -6; }
{
BARGS="--sessions 512"
run e_pair_b512 APRIL_B200_SCAN=pair APRIL_B200_SCAN_TRACE=1
run e_b512 A=1
BARGS="--sessions 256"
run e_pair_b256 APRIL_B200_SCAN=pair
BARGS="--sessions 160"
run e_pair_b160 APRIL_B200_SCAN=pair
run e_b160 A=1
} > $O/sweep.log 2>&1
cat $O/sweep.log

timeout 1500 python -m pytest tests -m gpu -q -rP > $O/pytest_gpu.log 2>&1
tail -n 30 $O/pytest_gpu.log
grep -h "near-ties\|sync coalescing\|native sync\|single sync" $O/pytest_gpu.log | head
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json')); print('default', d['ms_per_step'], d['e2e']['ms_per_step']); print('b512', d['b512']['ms_per_step'], d['b512']['e2e']['ms_per_step']); print('stream', d['stream']['latency_ms'], d['stream']['tick_ms'], d['stream']['device_passes'], d['stream']['mean_sessions_per_pass']); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
tail -n 5 $O/bench_default.err
