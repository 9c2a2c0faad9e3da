#!/bin/bash
# 1-GPU evidence run: parity suite, both bench arms, launch list, ncu --set full of the six K2 launches (KV1_NSp/NSm paired) + K1 at C3 size
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests6.log
python bench.py > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench6_ref.json 2> gpurun_out/r2_bench6_ref.err
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_ncu_launches.log 2>&1
python tools/profile_case.py 256 160 > gpurun_out/r2_plain_p.log 2>&1 &&
ncu --set full --clock-control none -k regex:'assemble_cst_pipe_kernel|assemble_cst_pair_kernel|assemble_fixed_kernel' -c 7 -o /tmp/r2_full256e \
    python tools/profile_case.py 256 160 > gpurun_out/r2_ncu_b.log 2>&1
ncu -i /tmp/r2_full256e.ncu-rep --page raw --csv > gpurun_out/r2_full256e_raw.csv 2>> gpurun_out/r2_ncu_b.log
ncu --set full --clock-control none --import-source on -k regex:assemble_cst_pair_kernel -c 1 -o gpurun_out/r2_k2_pair \
    python tools/profile_case.py 256 160 > gpurun_out/r2_ncu_pair.log 2>&1
