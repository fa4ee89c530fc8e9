#!/bin/bash
# round 2, call S: instruction-mix experiments (variant libraries, same box) + the copy-in ramp of the fed pipeline
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2s; mkdir -p $O
P=$PWD/fixedpointldpc_b200
LDPC_B200_LIB=$P/libldpc_b200_x123.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=300 > $O/pytest_x123.log 2>&1; echo "pytest rc=$?" >> $O/pytest_x123.log
tail -3 $O/pytest_x123.log
B="timeout 120 python bench.py --only --no-cpu --steps 8 --warmup 3"
for c in wifi a5 c79 a24; do
  for v in base x1 x2 x3 x123; do
    case "$c-$v" in c79-x2|c79-x3|a24-x2|a24-x3) continue;; esac
    L=$P/libldpc_b200_$v.so; [ $v = base ] && L=$P/libldpc_b200.so
    LDPC_B200_LIB=$L $B --code $c > $O/${c}_$v.json 2>&1
  done
done
LDPC_FEED_NO_RAMP=1 $B --code wifi > $O/wifi_noramp.json 2>&1
$B --code wifi --e2e-frames 65536 > $O/wifi_e2e64k.json 2>&1
for f in $O/*.json; do echo -n "$f "; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["frames_per_s"]), round(d["operating_point"]["frames_per_s"]), round(d["operating_point"]["frac_of_30it_frame_iteration_rate"],4),
          "e2e", round(d.get("e2e",{}).get("frac_of_device_rate",0),4), round(d.get("e2e_i16",{}).get("frac_of_device_rate",0),4))
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done | tee $O/summary.txt
