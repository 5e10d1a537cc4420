#!/bin/bash
set -u
O=gpurun_out/r2i; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('mid')" > $O/gen.log 2>&1
{
for l in 0 1 2; do
  timeout 200 python tools/pair_debug.py --model mid --sessions 1 --seconds 2 --a "APRIL_B200_SCAN=nosmall" --b "" --layer $l 2>&1 | grep -E "config|LO plane|planes|row block [0-3] " | cut -c1-260
done
for l in 0 1; do
  timeout 200 python tools/pair_debug.py --model mid --sessions 150 --seconds 2 --layer $l 2>&1 | grep -E "config|LO plane|planes|row block [0-3] " | cut -c1-260
done
} > $O/bisect2.log 2>&1
cat $O/bisect2.log
