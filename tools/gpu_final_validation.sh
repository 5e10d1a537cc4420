#!/bin/bash
# Round 2, final validation on one GPU: smoke, the GPU suite, the default bench line (+ b512, stream, cpu_baseline blocks), the reference
# arm, the streaming mode, and the ncu launch list of the default command.
set -u
O=gpurun_out/r3c; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -14 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; cut -c1-300 $O/bench_default.json
timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
timeout 600 python bench.py --mode stream > $O/bench_stream.json 2> $O/bench_stream.err; echo "stream rc=$?"
timeout 600 python bench.py --sessions 512 --no-extra --no-cpu-baseline > $O/bench_b512.json 2> $O/bench_b512.err; echo "b512 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_b64.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra > $O/launches_b64.log 2>&1; echo "launch list rc=$?"
