#!/bin/bash
# round-2 GPU session Z4 (4 GPUs): the driver's launch of both arms at N = 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 > $O/r2z4_bench_n4.json 2> $O/r2z4_bench_n4.err; echo "n4 rc=$?"; tail -1 $O/r2z4_bench_n4.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > $O/r2z4_ref_n4.json 2> $O/r2z4_ref_n4.err; echo "ref n4 rc=$?"; tail -1 $O/r2z4_ref_n4.json | cut -c1-200
