#!/bin/bash
# round-2 GPU check: parity tests, K1 variants, K2 timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
timeout 600 python tools/time_fk_variants.py 'build/variants/libirt_*.so' > gpurun_out/r2_fkvar.log 2>&1
cat gpurun_out/r2_fkvar.log | tail -5
timeout 900 python tools/time_k2.py 1000000 17 > gpurun_out/r2_k2.log 2>&1
tail -6 gpurun_out/r2_k2.log
IRT_B200_TRACE=1 timeout 600 python tools/time_k2.py 300000 10 > gpurun_out/r2_k2_trace.log 2>&1
tail -40 gpurun_out/r2_k2_trace.log
