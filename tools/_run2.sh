out=gpurun_out/r2_tune_k2h.log
: > $out
timeout 180 python tools/knob_case.py >> $out 2>&1 || echo "PARITY FAILED" >> $out
TKB_K1_SERIAL=1 timeout 300 python tools/profile_case.py 256 160 2 2>&1 | grep -E "stage ms" >> $out
timeout 300 python tools/profile_case.py 256 160 3 2>&1 | grep -E "stage ms" >> $out
TKB_LIB=$PWD/tune/lib_defer_stats.so timeout 300 python tools/profile_case.py 256 160 2>&1 | grep -E "K2 pipe|stage ms" >> $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3 >> $out
