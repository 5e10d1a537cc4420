#!/bin/bash
# Profiling pass on the GPU box (see /opt/skills/guides/B200_PROFILING.md): launch list of the default bench command + one `--set full`
# capture per hot kernel. Numbers printed by bench.py under ncu are NOT bench values. (The scan kernels carry their cluster dims at compile
# time: ncu cannot replay a launch that has BOTH the cluster-dimension and the cooperative launch attribute — LaunchFailed on pass 2.)
set -u
OUT=gpurun_out/prof; mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra"
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full')"   # generate the model once
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches_b64.csv $BENCH > $OUT/launches_b64.log 2>&1
cap() {   # name, kernel regex (demangled name), launch-skip, extra bench args
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -o $OUT/$1 -f $BENCH $4 > $OUT/$1.log 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
}
cap scan3_b64 lstm_scan3_kernel 0 ""
cap scan4_b200 lstm_scan4_kernel 0 "--sessions 200"
cap scan5_b512 lstm_scan5_kernel 0 "--sessions 512"
cap gemm_wih "gemm_tc_kernel<.int.3, .int.0, .int.2>" 1 ""
cap greedy greedy_cluster_kernel 0 ""
cap greedy_pair_b512 greedy_pair_kernel 0 "--sessions 512"
cap fbank512 fbank512_kernel 0 ""
cap conv12 conv12_im2col_kernel 0 ""
ls -la $OUT | grep -v ncu-rep
