#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -4 gpurun_out/r2_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_N1.json 2> gpurun_out/r2_bench_N1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_N1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_N1.json').read().strip().split('\n')[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'fkonly',d['roofline']['fk_only'], d['clocks'])
    print('e2e',d['e2e']['value'],d['e2e']['roofline']['frac'],d['e2e']['small']['value'])
    ec=d['edge_check']; print('k3',ec['ms_per_sweep'],ec['roofline']['frac'],ec['build_s']); print('k2',ec['k2']['seconds'],ec['k2']['frac_of_fp64_peak']); print(ec.get('replanning_tick')['ms_per_tick'],ec.get('replanning_tick_with_path')['ms_words_on_host']); print(ec.get('low_collision_env'))
    print('cpu',d['cpu_baseline']['value'], ec['cpu_baseline']['value'], ec['cpu_baseline']['verdicts_equal_gpu'], ec['cpu_baseline']['k2_vs_oracle'])
except Exception as e: print('parse error',e)
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fk_rk4 -s 2 -c 1 -o gpurun_out/r2_k1_final python tools/ncu_fk.py fkv > gpurun_out/r2_k1_ncu.log 2>&1
K='regex:fk_|self_coll|voxel_and|swept|edge_|raster_|scan_|round_|vertex_heads|env_|dfma|copy_i64|row_offsets|xchg'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out/r2_k1_final.ncu-rep gpurun_out/r2_launches.csv
