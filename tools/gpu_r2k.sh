#!/bin/bash
set -u
O=gpurun_out/r2k; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('mid')" > $O/gen.log 2>&1
F='config|nan counts|LO plane|planes|row block [0-2] '
{
for X in APRIL_B200_DBG_W0=1 APRIL_B200_DBG_POISON=1 APRIL_B200_DBG_SYNC=1; do
  echo "=== pair, layer 1, $X"
  timeout 200 python tools/pair_debug.py --model mid --sessions 150 --seconds 2 --layer 1 --a "$X" --b "APRIL_B200_SCAN=pair $X" 2>&1 | grep -E "$F" | cut -c1-200
  echo "=== scan3, layer 1, $X"
  timeout 200 python tools/pair_debug.py --model mid --sessions 1 --seconds 2 --layer 1 --a "APRIL_B200_SCAN=nosmall $X" --b "$X" 2>&1 | grep -E "$F" | cut -c1-200
done
echo "=== pair, layer 0, poison"
timeout 200 python tools/pair_debug.py --model mid --sessions 150 --seconds 2 --layer 0 --a "APRIL_B200_DBG_POISON=1" --b "APRIL_B200_SCAN=pair APRIL_B200_DBG_POISON=1" 2>&1 | grep -E "$F" | cut -c1-200
} > $O/diag.log 2>&1
cat $O/diag.log
