#!/bin/bash
# N-GPU run (N = $1): exchange checks, then the bench with the collective inside the timed step, C3 and C5
N=${1:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $T tools/check_shard_nccl.py > gpurun_out/r2_shard$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_shard$N.log
timeout 600 $T bench.py --gpus $N --steps 3 --warmup 3 --no-extras > gpurun_out/r2_bench_n${N}_c3.json 2> gpurun_out/r2_bench_n${N}_c3.err
timeout 600 $T bench.py --gpus $N --steps 3 --warmup 3 --no-extras --gather nccl > gpurun_out/r2_bench_n${N}_c3_nccl.json 2> gpurun_out/r2_bench_n${N}_c3_nccl.err
timeout 900 $T bench.py --gpus $N --steps 2 --warmup 1 --no-extras --workload c5 > gpurun_out/r2_bench_n${N}_c5.json 2> gpurun_out/r2_bench_n${N}_c5.err
nvidia-smi topo -m > gpurun_out/r2_topo$N.txt 2>&1
