#!/bin/bash
# ONE process driving all 8 GPUs of a box through the library's own placement (SURVEY.md par. 8e)
set -u
O=gpurun_out/r3e; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 85 python bench.py --lanes 8 --steps 3 --warmup 2 --no-cpu-baseline --no-extra > $O/bench_lanes8.json 2> $O/bench_lanes8.err; echo "lanes8 rc=$?"; cut -c1-900 $O/bench_lanes8.json; tail -3 $O/bench_lanes8.err
