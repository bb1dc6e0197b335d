#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -12 gpurun_out/r2_pytest.log
rm -f gpurun_out/r2_fk_noflags.log
for v in a_r1 b_main; do
  IRT_B200_LIB=$PWD/build/variants/libirt_$v.so timeout 300 python tools/time_fk.py >> gpurun_out/r2_fk_noflags.log 2>&1
done
IRT_FK_SMEM=1 IRT_B200_LIB=$PWD/build/variants/libirt_e_smem.so timeout 300 python tools/time_fk.py >> gpurun_out/r2_fk_noflags.log 2>&1
cat gpurun_out/r2_fk_noflags.log
timeout 600 python tools/time_fk_variants.py 'build/variants/libirt_[ab]_*.so' > gpurun_out/r2_fkvar.log 2>&1
tail -3 gpurun_out/r2_fkvar.log
IRT_B200_TRACE=1 timeout 900 python tools/time_k2.py 1000000 17 > gpurun_out/r2_k2_trace.log 2>&1
tail -40 gpurun_out/r2_k2_trace.log
