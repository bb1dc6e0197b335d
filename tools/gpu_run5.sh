#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -6 gpurun_out/r2_pytest.log
timeout 600 python tools/ncu_fk.py k2 60000 > gpurun_out/r2_k2small_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_k2_launches.csv python tools/ncu_fk.py k2 60000 > gpurun_out/r2_k2_ncu.log 2>&1
cat gpurun_out/r2_k2small_plain.log | tail -3
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_k2_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
mine=[r for r in rows[1:]]
idx=[i for i,r in enumerate(mine) if 'edge_init' in r[ik]]
seg=mine[idx[len(idx)//2]-8:]
tot=collections.Counter(); cnt=collections.Counter()
for r in seg:
    try: v=float(r[iv].replace(',',''))
    except: continue
    u=r[iu]; ms=v/1e6 if u.startswith('ns') else (v/1e3 if u.startswith('us') else v)
    k=r[ik].replace('<unnamed>::','').split('(')[0][:50]; tot[k]+=ms; cnt[k]+=1
print('second call total %.2f ms'%sum(tot.values()))
for k,v in tot.most_common(14): print('%-50s %5d %9.3f ms'%(k,cnt[k],v))
PY
ncu --set full --clock-control none --import-source on -k regex:swept_voxel_raster_kernel -s 2 -c 1 -o gpurun_out/r2_raster python tools/ncu_fk.py k2 60000 > gpurun_out/r2_raster_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:edge_subdivide -s 12 -c 1 -o gpurun_out/r2_subdivide python tools/ncu_fk.py k2 60000 > gpurun_out/r2_subdiv_ncu.log 2>&1
timeout 300 python tools/ncu_fk.py fkv > gpurun_out/r2_fkv_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fk_rk4 -s 2 -c 1 -o gpurun_out/r2_k1 python tools/ncu_fk.py fkv > gpurun_out/r2_k1_ncu.log 2>&1
ncu --set full --clock-control none -k regex:self_collision -s 4 -c 2 -o gpurun_out/r2_selfcol python tools/ncu_fk.py fkv > gpurun_out/r2_selfcol_ncu.log 2>&1
ncu --set full --clock-control none -k regex:dfma_peak -s 2 -c 1 -o gpurun_out/r2_dfma python tools/ncu_fk.py peak > gpurun_out/r2_dfma_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
