#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
rm -f gpurun_out/r2_fk_noflags.log
IRT_B200_LIB=$PWD/build/variants/libirt_c_main.so timeout 300 python tools/time_fk.py >> gpurun_out/r2_fk_noflags.log 2>&1
IRT_FK_SMEM=1 IRT_B200_LIB=$PWD/build/variants/libirt_f_smem255.so timeout 300 python tools/time_fk.py >> gpurun_out/r2_fk_noflags.log 2>&1
cat gpurun_out/r2_fk_noflags.log
timeout 900 python tools/time_k2.py 1000000 17 > gpurun_out/r2_k2.log 2>&1
tail -4 gpurun_out/r2_k2.log
timeout 600 python tools/ncu_fk.py k2 60000 > gpurun_out/r2_k2small_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_k2_launches.csv python tools/ncu_fk.py k2 60000 > gpurun_out/r2_k2_ncu.log 2>&1
tail -2 gpurun_out/r2_k2small_plain.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_k2_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
mine=rows[1:]
idx=[i for i,r in enumerate(mine) if 'edge_init' in r[ik]]
seg=mine[idx[len(idx)//2]-8:]
tot=collections.Counter(); cnt=collections.Counter()
for r in seg:
    try: v=float(r[iv].replace(',',''))
    except: continue
    u=r[iu]; ms=v/1e6 if u.startswith('ns') else (v/1e3 if u.startswith('us') else v)
    k=r[ik].replace('<unnamed>::','').split('(')[0][:50]; tot[k]+=ms; cnt[k]+=1
print('second call total %.2f ms'%sum(tot.values()))
for k,v in tot.most_common(8): print('%-50s %5d %9.3f ms'%(k,cnt[k],v))
PY
