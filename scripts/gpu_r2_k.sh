#!/bin/bash
# round 2, call K (8 GPUs): copy-in ceiling of the box, bench at N=8, the console programs and the waterfall on 8 GPUs
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2k; mkdir -p $O
nvidia-smi -L > $O/smi.txt; nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | head -25 > $O/lscpu.txt
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 8 --master-port 29521 scripts/h2d_ceiling.py > $O/h2d_8.json 2> $O/h2d_8.err
timeout 300 $T --nproc-per-node 4 --master-port 29522 scripts/h2d_ceiling.py > $O/h2d_4.json 2> $O/h2d_4.err
timeout 900 $T --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench rc=$?"
tail -c 300 $O/bench_8gpu.err
# the reference's console program on 8 GPUs through the C++ drop-in (ldpc_mc_run_multi: threads + NCCL inside the library)
W=$PWD/fixedpointldpc_b200
TMP=$(mktemp -d); python - "$TMP" <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import fixedpointldpc_b200 as fp
g = np.load("tests/golden/reference_vectors.npz")
tmp = sys.argv[1]
code = fp.codes.wifi_1944_r12(); code.save(os.path.join(tmp, "H_802.11_IndZero.txt"))
parity = np.setdiff1d(np.arange(code.n), g["wifi_info_index"].astype(np.int64)).astype(np.int32)
fp.Generator(code=code, parity_cols=parity).save(os.path.join(tmp, "H_802.11_IndZerog.txt"))
a5 = fp.codes.array_p47_r5()
parity = np.setdiff1d(np.arange(a5.n), g["a5_info_index"].astype(np.int64)).astype(np.int32)
fp.Generator(code=a5, parity_cols=parity).save(os.path.join(tmp, "G_array_forward.txt"))
PY
( cd $TMP && for n in 1 8; do /usr/bin/time -f "wall %e s" -o time_$n.txt env LDPC_GPUS=$n $W/ldpc_wrapper_wifi <<< 2 > wifi_$n.txt 2>&1; cat wifi_$n.txt time_$n.txt; done
  /usr/bin/time -f "wall %e s" env LDPC_STREAM=philox LDPC_MC_ROUND=1048576 timeout 600 $W/ldpc_wrapper_a5 sweep 4.5 5.5 0.5 waterfall.csv 100 > sweep.txt 2>&1; cat sweep.txt waterfall.csv waterfall.csv_log.txt ) > $O/console_8gpu.txt 2>&1
cp $TMP/waterfall.csv $TMP/waterfall.csv_log.txt $O/ 2>/dev/null
tail -20 $O/console_8gpu.txt
