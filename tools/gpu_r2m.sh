#!/bin/bash
set -u
O=gpurun_out/r2m; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('mid')" > $O/gen.log 2>&1
F='config|emulat|step|per-unit|W m|planes|Traceback|Error'
{
for L in 0 1; do
  echo "=== pair, layer $L"
  timeout 300 python tools/pair_debug.py --model mid --sessions 150 --seconds 2 --layer $L --emulate $L 2>&1 | grep -E "$F" | cut -c1-400
  echo "=== scan3, layer $L"
  timeout 300 python tools/pair_debug.py --model mid --sessions 1 --seconds 2 --layer $L --emulate $L --a "APRIL_B200_SCAN=nosmall" --b "" 2>&1 | grep -E "$F" | cut -c1-400
done
} > $O/diag.log 2>&1
cat $O/diag.log
