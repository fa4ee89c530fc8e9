#!/bin/bash
# round 2, call F: A/B against the previous kernel on the same box, occupancy, per-CTA trip counts, then parity
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2f; mkdir -p $O
B="timeout 120 python bench.py --only --no-cpu --steps 5 --warmup 3"
for c in wifi a5 c79; do
  LDPC_B200_LIB=$PWD/scratch/libldpc_prev.so $B --code $c > $O/${c}_prev.json 2>&1
  $B --code $c > $O/${c}_new.json 2>&1
done
LDPC_NO_STAGE=1 $B --code wifi > $O/wifi_new_nostage.json 2>&1
for v in prev new; do
  for c in wifi a5; do
    LDPC_B200_LIB=$PWD/scratch/libldpc_${v}_timing.so timeout 120 python bench.py --only --code $c --steps 1 --warmup 3 --no-cpu --frames 65536 --e2e-frames 1024 2>&1 | grep "phase cycles" > $O/phase_${v}_${c}.txt
  done
done
for f in $O/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["frames_per_s"]), round(d["operating_point"]["frames_per_s"]), round(d["operating_point"]["frac_of_30it_frame_iteration_rate"],4), d["run"])
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done
tail -n 30 $O/phase_*.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 --deselect tests/test_gpu_parity_at_scale.py > $O/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/pytest.log
tail -15 $O/pytest.log
