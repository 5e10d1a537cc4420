#!/bin/bash
# Round 2, GPU call A: first hardware contact of the wide-batch scan (kernels_scan4.cuh): parity against the round-1 kernel and
# the oracle, then timing of its variants at 64 and 512 sessions, then the whole -m gpu suite.
set -u
O=gpurun_out/r2a; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full')" > $O/gen.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.log 2>&1

timeout 300 python tools/scan_check.py --model tiny --sessions 3 --seconds 5 --oracle 1 > $O/check_tiny.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 2 --seconds 4 --oracle 1 > $O/check_full.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 200 --seconds 6 > $O/check_full200.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 130 --seconds 3 --a "APRIL_B200_SCAN=v3" --b "APRIL_B200_SCAN_CS=4 APRIL_B200_SCAN_MMAW=2" > $O/check_cs4m2.log 2>&1
timeout 600 python tools/scan_check.py --model full --sessions 130 --seconds 3 --a "APRIL_B200_SCAN_CS=4 APRIL_B200_SCAN_MMAW=1" --b "APRIL_B200_SCAN_MMAW=2" > $O/check_cs4m1_cs2m2.log 2>&1
tail -n 3 $O/check_*.log

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
run v3_b64 APRIL_B200_SCAN=v3
run v4_cs2m4_b64 APRIL_B200_SCAN_TRACE=1
run v4_cs2m4_g2_b64 APRIL_B200_SCAN_GROUPS=2
run v4_cs2m2_b64 APRIL_B200_SCAN_MMAW=2
run v4_cs4m2_b64 APRIL_B200_SCAN_CS=4 APRIL_B200_SCAN_MMAW=2 APRIL_B200_SCAN_TRACE=1
run v4_cs4m1_b64 APRIL_B200_SCAN_CS=4 APRIL_B200_SCAN_MMAW=1
run v4_fbankv2_b64 APRIL_B200_FBANK=v2
BARGS="--sessions 512"
run v4_cs2m4_b512 APRIL_B200_SCAN_TRACE=1
run v4_cs2m2_b512 APRIL_B200_SCAN_MMAW=2
run v4_cs4m2_b512 APRIL_B200_SCAN_CS=4 APRIL_B200_SCAN_MMAW=2 APRIL_B200_SCAN_TRACE=1
run v4_cs4m1_b512 APRIL_B200_SCAN_CS=4 APRIL_B200_SCAN_MMAW=1
BARGS="--sessions 256"
run v4_cs2m4_b256 A=1
run v4_cs4m2_b256 APRIL_B200_SCAN_CS=4 APRIL_B200_SCAN_MMAW=2
} > $O/sweep.log 2>&1
cat $O/sweep.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
tail -n 15 $O/pytest_gpu.log
