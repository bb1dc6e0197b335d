#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_n2.log
tail -15 gpurun_out/r2_pytest_n2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --roadmap-vertices 300000 > gpurun_out/r2_bench_n2_small.json 2> gpurun_out/r2_bench_n2_small.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r2_bench_n2_small.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2_small.json').read().strip().split('\n')[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
    print('e2e',d['e2e']['value'],d['e2e']['roofline']['frac'])
    ec=d['edge_check']; print('k3',ec['ms_per_sweep'],ec['exchange']); print('build',ec['build_s']); print(ec.get('replanning_tick'))
    print(d.get('host_affinity_rank0'))
except Exception as e: print('parse error',e)
PY
