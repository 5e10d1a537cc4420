#!/bin/bash
set -u
O=gpurun_out/r2j; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('mid')" > $O/gen.log 2>&1
cat > /tmp/vw.py <<'PY'
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'tools')
import common as C
model = C.model_path(sys.argv[1]); n = int(sys.argv[2])
C.run_product_batch(model, [C.gen_stream(700 + i, 32000) for i in range(n)], flush=False)
PY
APRIL_B200_VERIFY_W=1 APRIL_B200_SCAN=pair timeout 200 python /tmp/vw.py mid 150 > $O/verify_mid.log 2>&1
APRIL_B200_VERIFY_W=1 APRIL_B200_SCAN=pair timeout 200 python /tmp/vw.py tiny 150 > $O/verify_tiny.log 2>&1
grep -h "verify W" $O/verify_mid.log $O/verify_tiny.log
