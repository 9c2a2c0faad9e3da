#!/bin/bash
# 2-GPU regression of the multi-GPU path after the last changes: verify flag, both exchange forms, C5 chunks; plus smoke and the default bench
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke.log
timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 1 --no-extras --verify > gpurun_out/r2_v2_p2p.json 2> gpurun_out/r2_v2_p2p.err
timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 1 --no-extras --verify --gather nccl > gpurun_out/r2_v2_nccl.json 2> gpurun_out/r2_v2_nccl.err
timeout 900 $T bench.py --gpus 2 --steps 1 --warmup 1 --no-extras --verify --workload c5 > gpurun_out/r2_v2_c5.json 2> gpurun_out/r2_v2_c5.err
timeout 900 $T bench.py --gpus 2 > gpurun_out/r2_v2_full.json 2> gpurun_out/r2_v2_full.err
timeout 300 $T bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r2_v2_ref.json 2> gpurun_out/r2_v2_ref.err
