#!/bin/bash
# 8-GPU state at the end of the round (paired K2 launch): C3 with the copy-engine exchange and with NCCL, C5 sweep
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $T bench.py --gpus 8 --no-extras > gpurun_out/final2_bench_n8_c3.json 2> gpurun_out/final2_bench_n8_c3.err
timeout 600 $T bench.py --gpus 8 --no-extras --gather nccl > gpurun_out/final2_bench_n8_c3_nccl.json 2> gpurun_out/final2_bench_n8_c3_nccl.err
timeout 900 $T bench.py --gpus 8 --steps 2 --warmup 1 --no-extras --workload c5 > gpurun_out/final2_bench_n8_c5.json 2> gpurun_out/final2_bench_n8_c5.err
