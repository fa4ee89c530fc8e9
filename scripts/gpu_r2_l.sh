#!/bin/bash
# round 2, call L (8 GPUs): bench at N=8 with the fixed copy-in chunk schedule, the console programs and the waterfall on 8 GPUs
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2l; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench rc=$?"
tail -c 300 $O/bench_8gpu.err
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
( cd $TMP
  for n in 1 8; do
    t0=$(date +%s.%N); echo 2 | LDPC_GPUS=$n timeout 300 $W/ldpc_wrapper_wifi; t1=$(date +%s.%N)
    echo "ldpc_wrapper_wifi on $n GPU(s): wall $(echo "$t1 - $t0" | bc) s"
  done
  t0=$(date +%s.%N); LDPC_STREAM=philox LDPC_MC_ROUND=1048576 timeout 600 $W/ldpc_wrapper_a5 sweep 4.5 5.5 0.5 waterfall.csv 100; t1=$(date +%s.%N)
  echo "waterfall (philox stream, 8 GPUs): wall $(echo "$t1 - $t0" | bc) s"
  cat waterfall.csv waterfall.csv_log.txt ) > $O/console_8gpu.txt 2>&1
tail -25 $O/console_8gpu.txt
