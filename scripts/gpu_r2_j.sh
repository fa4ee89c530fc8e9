#!/bin/bash
# round 2, call J (2 GPUs): the multi-GPU entry of the C ABI on real devices, the console programs on 2 GPUs, bench at N=2
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2j; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 600 python -m pytest tests/test_gpu_mc_group.py tests/test_gpu_config3.py tests/test_gpu_facade.py -m gpu -x -q --timeout=400 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?"
tail -c 400 $O/bench_2gpu.err
