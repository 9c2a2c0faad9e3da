out=gpurun_out/r2_tune_k1.log
: > $out
for lib in tune/lib_k1r*.so; do
  echo "=== $lib" >> $out
  TKB_LIB=$PWD/$lib timeout 180 python tools/knob_case.py >> $out 2>&1 || { echo "PARITY FAILED" >> $out; continue; }
  TKB_K1_SERIAL=1 TKB_LIB=$PWD/$lib timeout 300 python tools/profile_case.py 256 160 2 2>&1 | grep -E "stage ms" >> $out
  TKB_LIB=$PWD/$lib timeout 300 python tools/profile_case.py 256 160 3 2>&1 | grep -E "stage ms" >> $out
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 >> $out
