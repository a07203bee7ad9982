#!/bin/bash
# round-2 GPU session Z (2 GPUs): the driver's launch of both arms at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2z_bench_n2.json 2> $O/r2z_bench_n2.err; echo "n2 rc=$?"; tail -1 $O/r2z_bench_n2.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2z_ref_n2.json 2> $O/r2z_ref_n2.err; echo "ref n2 rc=$?"; tail -1 $O/r2z_ref_n2.json | cut -c1-200
