for v in rs_minb7 rs_minb8 rs_minb6s1 rs_minb7s1; do
  if [ -z "$v" ]; then unset IRT_B200_LIB; else export IRT_B200_LIB=build/variants/libirt_$v.so; fi
  IRT_B200_TRACE=2 python tools/time_k2.py 1000000 17 > gpurun_out/r2_k2var_$v.log 2>&1
  echo "== ${v:-default}"; grep "indexed edges" gpurun_out/r2_k2var_$v.log | tail -1
  grep "timeline\]" gpurun_out/r2_k2var_$v.log | tail -99 | grep "raster" | awk '{print $3, $6, $7, $8}' | sort -n | head -2 | tr '\n' ' '; echo
done
