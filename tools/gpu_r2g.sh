#!/bin/bash
set -u
O=gpurun_out/r2g; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('mid')" > $O/gen.log 2>&1
{
for cfg in "--layer 0" "--layer 2" "--layer 0 --ragged 1" "--layer 2 --ragged 1" "--layer 2 --tap 1" "--layer 2 --ragged 1 --tap 1 --flush 1"; do
  timeout 200 python tools/pair_debug.py --model mid --sessions 150 --seconds 3 $cfg 2>&1 | grep -E "config|planes|row block [0-3] "
done
for cfg in "--layer 2" "--layer 2 --tap 1"; do
  timeout 200 python tools/pair_debug.py --model mid --sessions 1 --seconds 3 --a "APRIL_B200_SCAN=nosmall" --b "" $cfg 2>&1 | grep -E "config|planes|row block [0-3] "
done
} > $O/bisect.log 2>&1
cat $O/bisect.log | cut -c1-400
