#!/bin/bash
set -u
O=gpurun_out/r2z; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); import common; common.model_path('full')" > $O/gen.log 2>&1
timeout 600 python bench.py --sessions 160 --no-extra --no-cpu-baseline > $O/bench_b160.json 2> $O/bench_b160.err
timeout 600 python bench.py --sessions 256 --no-extra --no-cpu-baseline > $O/bench_b256.json 2> $O/bench_b256.err
python - <<'PY'
import json
for f in ['bench_b160','bench_b256']:
    d=json.loads([l for l in open('gpurun_out/r2z/%s.json'%f) if l.startswith('{')][-1])
    print(f, round(d['ms_per_step'],2), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items()})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_stream.csv python bench.py --mode stream --ticks 4 > $O/launches_stream.log 2>&1; echo "tick list rc=$?"
python - <<'PY'
import csv,collections,re
rows=list(csv.reader(open('gpurun_out/r2z/launches_stream.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
names=[]; 
for r in rows[hi+1:]:
    if len(r)<=im: continue
    name=re.sub(r'\(.*','',r[ik]).replace('void ','').replace('april::','')
    v=float(r[im].replace(',',''))*{'ns':1e-3,'us':1.0,'ms':1e3}.get(r[iu],1e-3)
    names.append((name,v))
# the last tick: from the last fbank512 to the end
idx=[i for i,(n,v) in enumerate(names) if 'carry_in' in n]
print('kernels', len(names), 'ticks seen', len(idx))
if len(idx)>=2:
    a,b=idx[-2],idx[-1]
    tot=collections.OrderedDict(); cnt=collections.Counter()
    for n,v in names[a:b]: tot[n]=tot.get(n,0)+v; cnt[n]+=1
    T=sum(tot.values()); print('one tick: %d kernels, %.1f us of kernel time'%(b-a,T))
    for k,v in sorted(tot.items(), key=lambda x:-x[1]): print('%8.1f us %5.1f %% %3d  %s'%(v,100*v/T,cnt[k],k[:80]))
PY
