#!/bin/bash
set -u
O=gpurun_out/r2h; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('mid')" > $O/gen.log 2>&1
cat > /tmp/mc.py <<'PY'
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'tools')
import common as C
n = int(sys.argv[2]); model = C.model_path(sys.argv[1])
cols = C.run_product_batch(model, [C.gen_stream(700 + i, 16000 + 977 * i) for i in range(n)], chunk=8000)
print("callbacks", [len(c.log) for c in cols][:4])
PY
timeout 800 compute-sanitizer --tool memcheck --print-limit 40 --error-exitcode 0 python /tmp/mc.py mid 1 > $O/memcheck_mid_b1.log 2>&1
grep -E "Invalid|out of bounds|Address|at .* in|by thread|ERROR SUMMARY|callbacks" $O/memcheck_mid_b1.log | head -60
timeout 800 compute-sanitizer --tool memcheck --print-limit 40 --error-exitcode 0 python /tmp/mc.py tiny 70 > $O/memcheck_tiny_b70.log 2>&1
grep -E "Invalid|out of bounds|Address|at .* in|by thread|ERROR SUMMARY|callbacks" $O/memcheck_tiny_b70.log | head -40
