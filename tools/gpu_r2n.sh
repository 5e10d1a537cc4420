#!/bin/bash
set -u
O=gpurun_out/r2n; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
F='config|emulat|step|per-unit|W m|planes|Traceback|Error'
{
for i in 1 2 3; do
  echo "=== pair, layer 1 (run $i)"
  timeout 300 python tools/pair_debug.py --model mid --sessions 150 --seconds 2 --layer 1 --emulate 1 2>&1 | grep -E "$F" | cut -c1-300
  echo "=== scan3, layer 1 (run $i)"
  timeout 300 python tools/pair_debug.py --model mid --sessions 1 --seconds 2 --layer 1 --emulate 1 --a "APRIL_B200_SCAN=nosmall" --b "" 2>&1 | grep -E "$F" | cut -c1-300
done
timeout 600 python tools/scan_check.py --model mid --sessions 150 --seconds 6 --a "" --b "APRIL_B200_SCAN=pair" --oracle 2
} > $O/diag.log 2>&1
cat $O/diag.log
