#!/bin/bash
set -u
O=gpurun_out/r2x; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -10 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2x/bench_default.json') if l.startswith('{')][-1])
print('default', round(d['ms_per_step'],2), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items()}, d['clocks']['samples'], d['clocks']['sm_mhz'])
print('b512', round(d['b512']['ms_per_step'],2), d['b512'].get('e2e',{}).get('ms_per_step'))
print('stream', d['stream']['latency_ms'])
PY
