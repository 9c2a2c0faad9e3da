#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $T tools/check_shard_nccl.py > gpurun_out/r2_shard${N}b.log 2>&1; echo "rc=$?" >> gpurun_out/r2_shard${N}b.log
timeout 600 $T bench.py --gpus $N --steps 3 --warmup 3 --no-extras > gpurun_out/r2_bench_n${N}_c3_kern.json 2> gpurun_out/r2_bench_n${N}_c3_kern.err
TKB_P2P_COPY=memcpy timeout 600 $T bench.py --gpus $N --steps 3 --warmup 3 --no-extras > gpurun_out/r2_bench_n${N}_c3_memcpy.json 2> gpurun_out/r2_bench_n${N}_c3_memcpy.err
timeout 900 $T bench.py --gpus $N --steps 2 --warmup 1 --no-extras --workload c5 > gpurun_out/r2_bench_n${N}_c5_kern.json 2> gpurun_out/r2_bench_n${N}_c5_kern.err
