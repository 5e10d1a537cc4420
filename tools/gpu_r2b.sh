#!/bin/bash
# Round 2, GPU call B: two-team scan4 (correctness + timing), fbank512, mid-model divergence probe, precision study, the whole
# -m gpu suite, and the new bench.py default line.
set -u
O=gpurun_out/r2b; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1

timeout 300 python tools/scan_check.py --model tiny --sessions 3 --seconds 5 --oracle 1 > $O/check_tiny.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 200 --seconds 6 > $O/check_full200.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 300 --seconds 3 --a "APRIL_B200_SCAN=v3" --b "APRIL_B200_SCAN_CS=4" > $O/check_cs4.log 2>&1
tail -n 2 $O/check_*.log

run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra $BARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('scan trace') or l.startswith('  '): print(l.rstrip())
    elif l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=r['ms_per_step_by_class']; print('$tag', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'scan_ms', round(c['lstm_scan'],2), 'us/layer-step', round(r['us_per_layer_step'],3), 'gemm_ms', round(c['gemm'],2), 'lstm_gemm_ms', round(c['of_which_lstm_gemm'],2), 'greedy_ms', round(c['joiner_greedy'],2), 'fbank_ms', round(c['fbank'],2), 'tokens', d['results']['final_tokens'])
    elif 'timeout' in l or 'rror' in l or 'libapril' in l: print(l.rstrip())
" | tail -6; }
{
BARGS="--sessions 64"
run v4t2_b64 APRIL_B200_SCAN_TRACE=1
run v4t2_g4_b64 APRIL_B200_SCAN_GROUPS=4
run v4t2_g1_b64 APRIL_B200_SCAN_GROUPS=1
run v4t2_m4_b64 APRIL_B200_SCAN_MMAW=4
run v4t2_cs4_b64 APRIL_B200_SCAN_CS=4
BARGS="--sessions 512"
run v4t2_b512 APRIL_B200_SCAN_TRACE=1
run v4t2_cs4_b512 APRIL_B200_SCAN_CS=4 APRIL_B200_SCAN_TRACE=1
BARGS="--sessions 256"
run v4t2_b256 A=1
BARGS="--sessions 128"
run v4t2_b128 A=1
} > $O/sweep.log 2>&1
cat $O/sweep.log

timeout 600 python tools/diverge_probe.py --model mid --stream 400 --samples 1920000 --chunk 8000 > $O/diverge_mid.log 2>&1
tail -n 25 $O/diverge_mid.log
timeout 900 python tools/precision_study.py --streams 2 --seconds 10 --out $O/precision.json > $O/precision.log 2>&1
tail -n 40 $O/precision.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
tail -n 40 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
tail -c 3000 $O/bench_default.json; tail -n 5 $O/bench_default.err
