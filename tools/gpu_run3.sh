#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -12 gpurun_out/r2_pytest.log
timeout 600 python tools/time_fk_variants.py 'build/variants/libirt_[ab]_*.so' > gpurun_out/r2_fkvar.log 2>&1
tail -3 gpurun_out/r2_fkvar.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench.json').read().strip().split('\n')[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'fkonly',d['roofline']['fk_only'])
    print('e2e',d['e2e']['value'],d['e2e']['roofline'],d['e2e']['small'])
    ec=d['edge_check']; print('k3',ec['ms_per_sweep'],ec['roofline']['frac'],ec['build_s']); print('k2',ec['k2']); print(ec.get('replanning_tick'),ec.get('replanning_tick_with_path')); print(ec.get('low_collision_env'))
    print('cpu',d['cpu_baseline'])
except Exception as e: print('parse error',e)
PY
timeout 600 python tools/time_k2.py 300000 10 > gpurun_out/r2_k2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_k2_launches.csv python tools/time_k2.py 300000 10 > gpurun_out/r2_k2_ncu.log 2>&1
tail -4 gpurun_out/r2_k2_plain.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_k2_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except: continue
    u=r[iu]; ms=v/1e6 if u.startswith('ns') else (v/1e3 if u.startswith('us') else v)
    k=r[ik][:60]; tot[k]+=ms; cnt[k]+=1
for k,v in tot.most_common(25): print('%-60s %6d %10.3f ms'%(k,cnt[k],v))
PY
