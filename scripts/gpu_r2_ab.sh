#!/bin/bash
# same-box A/B of a variant library (scripts/build_variant.py NAME ...) against the shipped one: parity file on the variant,
# then bench --only per code.   usage: gpu_r2_ab.sh NAME [codes...]
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
V=$1; shift; CODES=${*:-wifi a5 c79 a24}
O=gpurun_out/ab_$V; mkdir -p $O
P=$PWD/fixedpointldpc_b200
LDPC_B200_LIB=$P/libldpc_b200_$V.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
B="timeout 120 python bench.py --only --no-cpu --steps 8 --warmup 3"
for c in $CODES; do
  $B --code $c > $O/${c}_base.json 2>&1
  LDPC_B200_LIB=$P/libldpc_b200_$V.so $B --code $c > $O/${c}_$V.json 2>&1
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
