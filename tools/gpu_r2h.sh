#!/bin/bash
# round-2 GPU session H (2 GPUs): the benchmark under torchrun as the driver launches it, weak-scaling headline + config-5 strong line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2h_gpus.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > $O/r2h_bench_2gpu.log 2>&1
tail -c 2500 $O/r2h_bench_2gpu.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > $O/r2h_ref_2gpu.log 2>&1
tail -c 600 $O/r2h_ref_2gpu.log
