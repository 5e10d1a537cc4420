#!/bin/bash
set -u
O=gpurun_out/r2s; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1
{
timeout 600 python tools/scan_check.py --model full --sessions 64 --seconds 30 --tap 0 --a "" --b "APRIL_B200_OVERLAP=4" --oracle 1
timeout 600 python tools/scan_check.py --model full --sessions 37 --seconds 12 --tap 0 --a "" --b "APRIL_B200_OVERLAP=6" --oracle 1
timeout 300 python tools/scan_check.py --model mid --sessions 20 --seconds 20 --tap 0 --a "" --b "APRIL_B200_OVERLAP=4" --oracle 2
} > $O/check.log 2>&1
grep -E "A=\[|oracle vs|FAILED|rror" $O/check.log | cut -c1-300
export APRIL_B200_TIMING=1
one() {  # sessions, env...
  B=$1; shift
  env "$@" timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra --sessions $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=r['ms_per_step_by_class']; print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'scan_ms', round(c['lstm_scan'],2), 'gemm_ms', round(c['gemm'],2), 'greedy_ms', round(c['joiner_greedy'],2), 'tokens', d['results']['final_tokens'])
    elif 'rror' in l or 'timeout' in l: print(l.rstrip()[:200])
" | tail -2; }
for V in 0 3 4 6 8; do echo "== B=64 overlap=$V"; one 64 APRIL_B200_OVERLAP=$V; done > $O/sweep.log 2>&1
for V in 0 4; do echo "== B=32 overlap=$V"; one 32 APRIL_B200_OVERLAP=$V; done >> $O/sweep.log 2>&1
unset APRIL_B200_TIMING
for V in 0 4 6; do echo "== B=64 overlap=$V, no per-class timing"; one 64 APRIL_B200_OVERLAP=$V; done >> $O/sweep.log 2>&1
cat $O/sweep.log
BENCH="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra"
cap() {   # name, kernel regex, env, bench args
  env $3 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -c 1 -o $O/$1 -f $BENCH $4 > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  grep -E "ERROR|passes" $O/$1.log | head -3
}
cap scan4_b200 lstm_scan4_kernel X=1 "--sessions 200"
cap scan5_b512 lstm_scan5_kernel X=1 "--sessions 512"
cap greedy_pair_b512 greedy_pair_kernel X=1 "--sessions 512"
rm -f $O/*.ncu-rep
APRIL_B200_OVERLAP=4 timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_overlap.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_overlap.log; tail -5 $O/pytest_gpu_overlap.log
