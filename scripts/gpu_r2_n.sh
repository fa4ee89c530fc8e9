#!/bin/bash
# round 2, call N: refill without the barrier behind the header update (lean mode): parity subset + throughput
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2n; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_montecarlo.py -m gpu -x -q --timeout=300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
B="timeout 120 python bench.py --only --no-cpu --steps 8 --warmup 3"
for c in wifi a5 c79 a24; do $B --code $c > $O/${c}.json 2>&1; done
for f in $O/*.json; do echo -n "$f "; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["frames_per_s"]), round(d["operating_point"]["frames_per_s"]), round(d["operating_point"]["frac_of_30it_frame_iteration_rate"],4), 'e2e %.3f %.3f'%(d['e2e']['frac_of_device_rate'], d['e2e_i16']['frac_of_device_rate']))
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done
