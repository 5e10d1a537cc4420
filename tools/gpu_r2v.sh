#!/bin/bash
# Round 2, two GPUs of one box: the default line once more (clock sampling fixed), one process per GPU under torchrun (the driver's
# layout), ONE process driving both GPUs through the library's lanes, and the placement test.
set -u
O=gpurun_out/r2v; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full')" > $O/gen.log 2>&1
timeout 900 python bench.py > $O/bench_default_n1.json 2> $O/bench_default_n1.err; echo "n1 rc=$?"; cut -c1-200 $O/bench_default_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"; cut -c1-300 $O/bench_n2.json
timeout 600 python bench.py --lanes 2 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_lanes2.json 2> $O/bench_lanes2.err; echo "lanes rc=$?"; cut -c1-700 $O/bench_lanes2.json
timeout 600 python -m pytest tests -m gpu -q -k "placed_on_every_gpu or coalesced" > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log
