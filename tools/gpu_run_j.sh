#!/bin/bash
# last check of the round: full GPU suite and the default bench line with the final library
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests7.log
python bench.py > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
