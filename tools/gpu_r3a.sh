#!/bin/bash
set -u
O=gpurun_out/r3a; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "gemm or golden or streaming_40ms or full_size_model or chunk" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log; tail -6 $O/pytest_sel.log
for V in 1 0; do
  APRIL_B200_SPLITK=$V timeout 600 python bench.py --mode stream > $O/stream_splitk$V.json 2> $O/stream_splitk$V.err
  APRIL_B200_SPLITK=$V timeout 600 python bench.py --sessions 8 --no-extra --no-cpu-baseline > $O/b8_splitk$V.json 2> $O/b8_splitk$V.err
done
APRIL_B200_SPLITK=1 timeout 600 python bench.py --no-extra --no-cpu-baseline > $O/b64_splitk1.json 2> $O/b64_splitk1.err
python - <<'PY'
import json
def last(f): return json.loads([l for l in open(f) if l.startswith('{')][-1])
for V in (1,0):
    s=last('gpurun_out/r3a/stream_splitk%d.json'%V)['stream']; print('splitk',V,'stream', s['latency_ms'], 'device/pass', round(s['device_ms_per_pass'],3))
    d=last('gpurun_out/r3a/b8_splitk%d.json'%V); print('   b8', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['roofline']['ms_per_step_by_class'])
d=last('gpurun_out/r3a/b64_splitk1.json'); print('b64', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['roofline']['ms_per_step_by_class'])
PY
