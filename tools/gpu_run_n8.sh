#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_N8.json 2> gpurun_out/r2_bench_N8.err; echo "bench n8 rc=$?"
tail -c 600 gpurun_out/r2_bench_N8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_N8.json').read().strip().split('\n')[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
    print('e2e',d['e2e']['value'],d['e2e']['roofline']['frac'], d['e2e']['small']['value'])
    ec=d['edge_check']; x=ec['exchange']; print('k3',ec['ms_per_sweep'],x['nccl_all_gather_ms_per_sweep'],x['mismatches_vs_nccl'],x['mismatches_vs_unsharded']); print('build',ec['build_s']['edge_voxel_cache']); print(ec.get('replanning_tick')['ms_per_tick']); print(ec.get('low_collision_env')['ms_per_sweep'])
except Exception as e: print('parse error',e)
PY
