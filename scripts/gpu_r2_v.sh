#!/bin/bash
# round 2, call V (8 GPUs): bench at N=8 with the shipped kernels; the multi-device group tests on real devices
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2v; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench rc=$?"
tail -c 300 $O/bench_8gpu.err
timeout 300 python -m pytest tests/test_gpu_mc_group.py -m gpu -x -q --timeout=200 > $O/pytest_mc_group.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mc_group.log
tail -3 $O/pytest_mc_group.log
