#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523"
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests3.log
python examples/walkthrough.py > gpurun_out/r2_walkthrough.log 2>&1; echo "rc=$?" >> gpurun_out/r2_walkthrough.log
timeout 600 $T tools/check_shard_nccl.py > gpurun_out/r2_shard2c.log 2>&1; echo "rc=$?" >> gpurun_out/r2_shard2c.log
timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 1 --no-extras --verify > gpurun_out/r2_v3_p2p.json 2> gpurun_out/r2_v3_p2p.err
timeout 900 $T bench.py --gpus 2 > gpurun_out/r2_v3_full.json 2> gpurun_out/r2_v3_full.err
