#!/bin/bash
set -u
O=gpurun_out/r2f; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('mid')" > $O/gen.log 2>&1
timeout 300 python tools/pair_debug.py --model mid --sessions 1 --seconds 2 --a "APRIL_B200_SCAN=nosmall" --b "" > $O/dbg_scan3_mid_b1.log 2>&1
timeout 300 python tools/pair_debug.py --model mid --sessions 40 --seconds 2 --a "APRIL_B200_SCAN=nosmall" --b "" > $O/dbg_scan3_mid_b40.log 2>&1
timeout 300 python tools/pair_debug.py --model tiny --sessions 40 --seconds 2 --a "APRIL_B200_SCAN=nosmall" --b "" > $O/dbg_scan3_tiny_b40.log 2>&1
timeout 300 python tools/pair_debug.py --model mid --sessions 150 --seconds 2 > $O/dbg_pair_mid_b150.log 2>&1
timeout 300 python tools/pair_debug.py --model tiny --sessions 150 --seconds 2 > $O/dbg_pair_tiny_b150.log 2>&1
tail -n 12 $O/dbg_*.log
