#!/bin/bash
# round 2, call T: array p47 r24 with two / three checks interleaved per thread; parity of the new variable-phase sums
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_montecarlo.py -m gpu -x -q --timeout=300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for sp in 2 3; do
  LDPC_A24_SPLIT=$sp timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=300 -k a24 > $O/pytest_split$sp.log 2>&1; echo "pytest rc=$?" >> $O/pytest_split$sp.log
  tail -2 $O/pytest_split$sp.log
done
B="timeout 120 python bench.py --only --no-cpu --steps 8 --warmup 3"
$B --code a24 > $O/a24_base.json 2>&1
LDPC_A24_SPLIT=2 $B --code a24 > $O/a24_split2.json 2>&1
LDPC_A24_SPLIT=3 $B --code a24 > $O/a24_split3.json 2>&1
for c in wifi a5 c79; do $B --code $c > $O/${c}_base.json 2>&1; done
for f in $O/*.json; do echo -n "$f "; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["frames_per_s"]), round(d["operating_point"]["frames_per_s"]), round(d["operating_point"]["frac_of_30it_frame_iteration_rate"],4),
          "e2e", round(d.get("e2e",{}).get("frac_of_device_rate",0),4), round(d.get("e2e_i16",{}).get("frac_of_device_rate",0),4), d["run"]["threads"])
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done | tee $O/summary.txt
