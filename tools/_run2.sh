for i in 1 2; do
for l in "" tune/lib_old.so; do
echo "lib=$l" >> gpurun_out/r2_tune_log4.log
if [ -n "$l" ]; then export TKB_LIB=$PWD/$l; else unset TKB_LIB; fi
TKB_K1_SERIAL=1 timeout 300 python tools/profile_case.py 256 160 2>&1 | grep -E "stage ms" >> gpurun_out/r2_tune_log4.log
done
done
unset TKB_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2 >> gpurun_out/r2_tune_log4.log
