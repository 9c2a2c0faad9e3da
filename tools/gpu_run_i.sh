#!/bin/bash
# 2-GPU regression of the multi-GPU path with the final library: both exchange forms verified against the per-rank matrices, C5 chunks, smoke
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
python __graft_entry__.py smoke > gpurun_out/r2_smoke2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke2.log
timeout 600 $T tools/check_shard_nccl.py > gpurun_out/r2_shard2d.log 2>&1; echo "rc=$?" >> gpurun_out/r2_shard2d.log
timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 1 --no-extras --verify > gpurun_out/r2_v4_p2p.json 2> gpurun_out/r2_v4_p2p.err
timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 1 --no-extras --verify --gather nccl > gpurun_out/r2_v4_nccl.json 2> gpurun_out/r2_v4_nccl.err
timeout 900 $T bench.py --gpus 2 --steps 1 --warmup 1 --no-extras --verify --workload c5 > gpurun_out/r2_v4_c5.json 2> gpurun_out/r2_v4_c5.err
