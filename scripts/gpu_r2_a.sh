#!/bin/bash
# round 2, call A: parity tests, full bench line, a24 / a5 launch-shape experiments, a24 ncu capture
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
tail -c 600 $O/bench_full.err
B="timeout 300 python bench.py --only --no-cpu --steps 5 --warmup 3"
$B --code a24 > $O/a24_default.json 2>&1
$B --code a24 --threads 576 > $O/a24_t576.json 2>&1
$B --code a24 --threads 640 > $O/a24_t640.json 2>&1
LDPC_A24_512=1 $B --code a24 > $O/a24_512regs128.json 2>&1
$B --code a5 > $O/a5_table.json 2>&1
LDPC_A5_CLOSED=1 $B --code a5 > $O/a5_closed.json 2>&1
for f in $O/a24_*.json $O/a5_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["frames_per_s"], d["operating_point"]["frames_per_s"], d["run"])
except Exception as e:
    print("ERR", e, open(sys.argv[1]).read()[-400:])
PY
done
# a24 ncu capture (one launch of 16384 frames x 30 iterations)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 3 -c 1 -o $O/prof_a24_r2a -f \
  python bench.py --code a24 --only --precision 16 --steps 1 --warmup 3 --no-cpu --frames 16384 > $O/ncu_a24.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summarise.py $O/prof_a24_r2a.ncu-rep a24 r2a --frames 16384 --iters 30 --outdir $O/summaries > $O/summarise.log 2>&1
tail -3 $O/summarise.log
