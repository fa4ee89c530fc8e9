#!/bin/bash
# round 2, call Q: word-set count as a compile-time constant for the named codes: parity subset + same-box A/B
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2q; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_montecarlo.py -m gpu -x -q --timeout=300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
B="timeout 120 python bench.py --only --no-cpu --steps 8 --warmup 3"
for c in wifi a5 c79 a24; do
  $B --code $c > $O/${c}_ws.json 2>&1
  LDPC_RUNTIME_W=1 $B --code $c > $O/${c}_rt.json 2>&1
done
for f in $O/*.json; do echo -n "$f "; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["frames_per_s"]), round(d["operating_point"]["frames_per_s"]), round(d["operating_point"]["frac_of_30it_frame_iteration_rate"],4))
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done
