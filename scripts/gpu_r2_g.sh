#!/bin/bash
# round 2, call G: degree masks (small hot code) on both refill structures, same box
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2g; mkdir -p $O
B="timeout 120 python bench.py --only --no-cpu --steps 8 --warmup 3"
LDPC_B200_LIB=$PWD/scratch/libldpc_prev.so $B --code wifi > $O/wifi_prev.json 2>&1
LDPC_B200_LIB=$PWD/scratch/libldpc_prev_mask.so $B --code wifi > $O/wifi_prevmask.json 2>&1
$B --code wifi > $O/wifi_new.json 2>&1
LDPC_NO_STAGE=1 $B --code wifi > $O/wifi_new_nostage.json 2>&1
LDPC_WIFI_ALL_DEGREES=1 $B --code wifi > $O/wifi_new_alldeg.json 2>&1
for c in a5 c79; do
  LDPC_B200_LIB=$PWD/scratch/libldpc_prev.so $B --code $c > $O/${c}_prev.json 2>&1
  $B --code $c > $O/${c}_new.json 2>&1
done
LDPC_A24_512=1 $B --code a24 > $O/a24_new_r128.json 2>&1
for f in $O/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["frames_per_s"]), round(d["operating_point"]["frames_per_s"]), round(d["operating_point"]["frac_of_30it_frame_iteration_rate"],4), 'e2e %.3f %.3f'%(d['e2e']['frac_of_device_rate'], d['e2e_i16']['frac_of_device_rate']), d["run"].get("stage_rows"))
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-300:])
PY
done
