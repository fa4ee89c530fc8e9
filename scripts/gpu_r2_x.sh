#!/bin/bash
# round 2, call X: 802.11 variable phase fed by one table (order + degree + edge offsets): parity + same-box A/B
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2x; mkdir -p $O
P=$PWD/fixedpointldpc_b200
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
LDPC_B200_LIB=$P/libldpc_b200_l1.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=300 -k wifi > $O/pytest_l1.log 2>&1; echo "pytest rc=$?" >> $O/pytest_l1.log
tail -3 $O/pytest_l1.log
B="timeout 120 python bench.py --only --no-cpu --steps 8 --warmup 3 --code wifi"
for v in old base l1 old base l1; do
  L=$P/libldpc_b200_$v.so; [ $v = base ] && L=$P/libldpc_b200.so
  LDPC_B200_LIB=$L $B > $O/wifi_${v}_$RANDOM.json 2>&1
done
for f in $O/*.json; do echo -n "$f "; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["frames_per_s"]), round(d["operating_point"]["frames_per_s"]), round(d["operating_point"]["frac_of_30it_frame_iteration_rate"],4))
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done | tee $O/summary.txt
