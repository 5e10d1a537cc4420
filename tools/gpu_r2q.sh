#!/bin/bash
# Round 2 final validation: the two scan captures ncu could not replay under --set full (minimal metric set, one pass), the GPU test suite,
# the default bench line, the reference arm, the streaming and soak modes.
set -u
O=gpurun_out/r2q; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full')" > $O/gen.log 2>&1
BENCH="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
cap1() {  # name, kernel regex, bench args
  timeout 600 ncu --metrics $M --clock-control none --kernel-name-base demangled -k "regex:$2" -c 1 -o $O/$1 -f $BENCH $3 > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  tail -2 $O/$1.raw.csv | cut -c1-600
}
cap1 scan4_b200 lstm_scan4_kernel "--sessions 200"
cap1 scan5_b512 lstm_scan5_kernel "--sessions 512"
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -14 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; cut -c1-600 $O/bench_default.json
timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-400 $O/bench_reference.json
timeout 600 python bench.py --mode stream > $O/bench_stream.json 2> $O/bench_stream.err; echo "stream rc=$?"; cut -c1-700 $O/bench_stream.json
timeout 900 python bench.py --mode soak > $O/bench_soak16.json 2> $O/bench_soak16.err; echo "soak16 rc=$?"; cut -c1-700 $O/bench_soak16.json
timeout 900 python bench.py --mode soak --sessions 128 > $O/bench_soak128.json 2> $O/bench_soak128.err; echo "soak128 rc=$?"; cut -c1-700 $O/bench_soak128.json
