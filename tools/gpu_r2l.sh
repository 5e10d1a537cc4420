#!/bin/bash
set -u
O=gpurun_out/r2l; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('mid')" > $O/gen.log 2>&1
cat > /tmp/vw.py <<'PY'
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'tools')
import common as C
model = C.model_path(sys.argv[1]); n = int(sys.argv[2])
C.run_product_batch(model, [C.gen_stream(700 + i, 32000) for i in range(n)], flush=False)
PY
APRIL_B200_VERIFY_W=1 APRIL_B200_SCAN=pair timeout 200 python /tmp/vw.py mid 150 2>&1 | grep "verify W" > $O/verify_mid.log
F='config|nan counts|LO plane|planes|row block [0-2] '
{
cat $O/verify_mid.log
for W in 1 2; do for L in 0 1; do
  echo "=== pair, layer $L, all layers use the scan weights of layer $W"
  timeout 200 python tools/pair_debug.py --model mid --sessions 150 --seconds 2 --layer $L --a "APRIL_B200_DBG_WL=$W" --b "APRIL_B200_SCAN=pair APRIL_B200_DBG_WL=$W" 2>&1 | grep -E "$F" | cut -c1-200
done; done
echo "=== scan3, layer 0, all layers use the scan weights of layer 1"
timeout 200 python tools/pair_debug.py --model mid --sessions 1 --seconds 2 --layer 0 --a "APRIL_B200_SCAN=nosmall APRIL_B200_DBG_WL=1" --b "APRIL_B200_DBG_WL=1" 2>&1 | grep -E "$F" | cut -c1-200
echo "=== scan3, layer 1, plain (repeat x2)"
for i in 1 2; do timeout 200 python tools/pair_debug.py --model mid --sessions 1 --seconds 2 --layer 1 --a "APRIL_B200_SCAN=nosmall" --b "" 2>&1 | grep -E "$F" | cut -c1-200; done
} > $O/diag.log 2>&1
cat $O/diag.log
