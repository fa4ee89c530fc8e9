#!/bin/bash
# round 2, call D: refill rework with the shared address space kept, f64 decoder, Monte-Carlo group
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q --timeout=200 > $O/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/pytest.log
tail -15 $O/pytest.log
B="timeout 120 python bench.py --only --no-cpu --steps 5 --warmup 3"
for c in wifi a5 c79 a24; do
  $B --code $c > $O/${c}_stage.json 2>&1 || echo "bench $c failed/timeout"
done
LDPC_NO_STAGE=1 $B --code wifi > $O/wifi_nostage.json 2>&1
LDPC_NO_STAGE=1 $B --code c79 > $O/c79_nostage.json 2>&1
LDPC_A5_CLOSED=1 $B --code a5 > $O/a5_closed.json 2>&1
LDPC_A24_512=1 $B --code a24 > $O/a24_r128_auto.json 2>&1
LDPC_A24_512=1 $B --code a24 --threads 512 > $O/a24_r128_t512.json 2>&1
for f in $O/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["frames_per_s"]), round(d["operating_point"]["frames_per_s"]), round(d["operating_point"]["frac_of_30it_frame_iteration_rate"],4), d["run"], 'e2e %.3f %.3f'%(d['e2e']['frac_of_device_rate'], d['e2e_i16']['frac_of_device_rate']))
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done
