#!/bin/bash
# 2-GPU check: sharded build + every form of the exchange against the unsharded result, then the bench with the collective in the timed step
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $T tools/check_shard_nccl.py --bw > gpurun_out/r2_shard2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_shard2.log
timeout 600 $T bench.py --gpus 2 --steps 3 --warmup 2 --no-extras > gpurun_out/r2_bench_n2_p2p.json 2> gpurun_out/r2_bench_n2_p2p.err
timeout 600 $T bench.py --gpus 2 --steps 3 --warmup 2 --no-extras --gather nccl > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err
timeout 900 $T bench.py --gpus 2 --steps 2 --warmup 1 --no-extras --workload c5 > gpurun_out/r2_bench_n2_c5.json 2> gpurun_out/r2_bench_n2_c5.err
