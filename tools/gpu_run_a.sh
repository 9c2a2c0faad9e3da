#!/bin/bash
# 1-GPU check: parity suite, bench line (both arms), calibration counters of the algorithmic flops
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests1.log
python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench1_ref.json 2> gpurun_out/r2_bench1_ref.err
python tools/calibrate_k2_flops.py > gpurun_out/r2_cal_plain.log 2>&1 &&
ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum \
    -k regex:eval_integrand_kernel --csv --log-file gpurun_out/r2_k2_calibration.csv python tools/calibrate_k2_flops.py > gpurun_out/r2_cal_ncu.log 2>&1
