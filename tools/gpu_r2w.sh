#!/bin/bash
set -u
O=gpurun_out/r2w; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full'); common.model_path('mid')" > $O/gen.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "pinned or damaged or coalesced or async" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log; tail -5 $O/pytest_sel.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_pinned.json 2> $O/bench_pinned.err; echo "rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-extra --no-pin > $O/bench_nopin.json 2> $O/bench_nopin.err; echo "rc=$?"
python - <<'PY'
import json
for f in ['bench_pinned','bench_nopin']:
    d=json.loads([l for l in open('gpurun_out/r2w/%s.json'%f) if l.startswith('{')][-1])
    print(f, round(d['ms_per_step'],2), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items()}, d['config'].get('inputs','')[:40], d['clocks'])
    if 'b512' in d: print('   b512', round(d['b512']['ms_per_step'],2), d['b512'].get('e2e'))
PY
