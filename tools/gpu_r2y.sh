#!/bin/bash
set -u
O=gpurun_out/r2y; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "512_sessions or pinned or search_schedule" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log; tail -4 $O/pytest_sel.log
timeout 600 python bench.py --sessions 512 --no-extra --no-cpu-baseline > $O/bench_b512.json 2> $O/bench_b512.err
timeout 600 python bench.py --sessions 160 --no-extra --no-cpu-baseline > $O/bench_b160.json 2> $O/bench_b160.err
python - <<'PY'
import json
for f in ['bench_b512','bench_b160']:
    d=json.loads([l for l in open('gpurun_out/r2y/%s.json'%f) if l.startswith('{')][-1])
    print(f, round(d['ms_per_step'],2), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items()})
PY
