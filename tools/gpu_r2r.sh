#!/bin/bash
set -u
O=gpurun_out/r2r; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1
{
timeout 300 python tools/scan_check.py --model tiny --sessions 5 --seconds 5 --a "APRIL_B200_GREEDY=cluster" --b "APRIL_B200_GREEDY=pair" --oracle 2
timeout 300 python tools/scan_check.py --model mid --sessions 70 --seconds 6 --a "APRIL_B200_GREEDY=cluster" --b "APRIL_B200_GREEDY=pair" --oracle 2
timeout 600 python tools/scan_check.py --model full --sessions 64 --seconds 8 --a "APRIL_B200_GREEDY=cluster" --b "APRIL_B200_GREEDY=pair" --oracle 1
timeout 600 python tools/scan_check.py --model full --sessions 300 --seconds 3 --a "APRIL_B200_GREEDY=cluster" --b "APRIL_B200_GREEDY=pair"
} > $O/check.log 2>&1
grep -E "A=\[|oracle vs|FAILED|rror" $O/check.log | cut -c1-300
export APRIL_B200_TIMING=1
one() {  # sessions, env...
  B=$1; shift
  env "$@" timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --sessions $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; c=r['ms_per_step_by_class']; print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'scan_ms', round(c['lstm_scan'],2), 'gemm_ms', round(c['gemm'],2), 'greedy_ms', round(c['joiner_greedy'],2), 'tokens', d['results']['final_tokens'])
    elif 'rror' in l or 'timeout' in l: print(l.rstrip()[:200])
" | tail -2; }
for B in 16 64 160 512; do
  echo "== B=$B cluster"; one $B APRIL_B200_GREEDY=cluster
  echo "== B=$B pair";  one $B APRIL_B200_GREEDY=pair
done > $O/sweep.log 2>&1
cat $O/sweep.log
unset APRIL_B200_TIMING
BENCH="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra"
cap() {   # name, kernel regex, env, bench args
  env $3 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -c 1 -o $O/$1 -f $BENCH $4 > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  grep -E "ERROR|passes" $O/$1.log | head -3
}
cap scan4_b96 lstm_scan4_kernel X=1 "--sessions 96"
cap scan5_b160 lstm_scan5_kernel APRIL_B200_SCAN=pair "--sessions 160"
cap greedy_pair_b64 greedy_pair_kernel APRIL_B200_GREEDY=pair ""
rm -f $O/*.ncu-rep
APRIL_B200_GREEDY=pair timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_pair.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_pair.log; tail -5 $O/pytest_gpu_pair.log
