#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log
rm -f gpurun_out/r2_fk_noflags.log
for v in c_main; do
  IRT_B200_LIB=$PWD/build/variants/libirt_$v.so timeout 300 python tools/time_fk.py >> gpurun_out/r2_fk_noflags.log 2>&1
done
cat gpurun_out/r2_fk_noflags.log
timeout 600 python tools/time_fk_variants.py 'build/variants/libirt_c_*.so' > gpurun_out/r2_fkvar.log 2>&1
tail -3 gpurun_out/r2_fkvar.log
timeout 900 python tools/time_k2.py 1000000 17 > gpurun_out/r2_k2.log 2>&1
tail -4 gpurun_out/r2_k2.log
timeout 1500 python bench.py --steps 5 --warmup 3 --cpu-seconds 4 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2_bench.err
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
