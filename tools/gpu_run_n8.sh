#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_N8.json 2> gpurun_out/r2_bench_N8.err; echo "bench n8 rc=$?"
tail -c 1200 gpurun_out/r2_bench_N8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_N8.json').read().strip().split('\n')[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
    print('e2e',d['e2e']['value'],d['e2e']['roofline'], d['e2e']['small']['value'])
    ec=d['edge_check']; print('k3',ec['ms_per_sweep'],ec['exchange']); print('build',ec['build_s']); print(ec.get('replanning_tick')); print(ec.get('low_collision_env'))
    print(d.get('host_affinity_rank0'))
except Exception as e: print('parse error',e)
PY
