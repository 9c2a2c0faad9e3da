#!/bin/bash
# Round-2 first GPU call: toolchain probe, parity suite, bench, ncu --set full of the seven K2 launches + K1 at 256 xi.
mkdir -p gpurun_out
{
  echo "== which gfortran / f951 / flang / nvfortran / ifx"; which gfortran f951 flang flang-new nvfortran ifx ifort 2>&1
  echo "== gcc -print-prog-name=f951"; gcc -print-prog-name=f951
  echo "== import tiktaalik"; python -c "import tiktaalik; print(tiktaalik.__file__)" 2>&1 | tail -1
  echo "== baseline/_ref oracle/_ref"; ls -la baseline/_ref oracle/_ref 2>&1
  echo "== nproc"; nproc; lscpu | grep -E 'Model name|Thread|Core|Socket'
  echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
} > gpurun_out/r2_probe.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests0.log
python bench.py > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
python tools/profile_case.py 256 160 > gpurun_out/r2_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'assemble_cst_pipe_kernel|assemble_fixed_kernel' -c 8 -o /tmp/r2_full256 \
    python tools/profile_case.py 256 160 > gpurun_out/r2_ncu.log 2>&1
ncu -i /tmp/r2_full256.ncu-rep --page raw --csv > gpurun_out/r2_full256_raw.csv 2>> gpurun_out/r2_ncu.log
ls -la /tmp/r2_full256.ncu-rep >> gpurun_out/r2_ncu.log
sz=$(stat -c %s /tmp/r2_full256.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 45000000 ]; then cp /tmp/r2_full256.ncu-rep gpurun_out/r2_full256.ncu-rep; fi
