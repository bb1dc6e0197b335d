#!/bin/bash
mkdir -p gpurun_out
for v in a_r1 c_static d_persist; do
  IRT_B200_LIB=$PWD/build/variants/libirt_$v.so timeout 300 python tools/time_fk.py >> gpurun_out/r2_fk_noflags.log 2>&1
done
cat gpurun_out/r2_fk_noflags.log
PKG=interactive-rate-tendons_b200
g++ -std=c++17 -O1 tests/cpp/test_host_mirror.cpp -o /tmp/thm -L$PKG -lirt_b200 -Loracle -loracle -Wl,-rpath,$PWD/$PKG -Wl,-rpath,$PWD/oracle -fopenmp || exit 1
IRT_B200_DEBUG_SYNC=1 timeout 600 stdbuf -o0 -e0 /tmp/thm > gpurun_out/r2_debugsync.log 2>&1
echo "thm rc=$?" >> gpurun_out/r2_debugsync.log
tail -c 2500 gpurun_out/r2_debugsync.log
