#!/bin/bash
# round 2, call M: whole gpu suite + smoke + whole bench line (both arms) with the kernel of the round
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2m; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > $O/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/pytest.log
tail -4 $O/pytest.log
cp gpurun_out/parity_at_scale_last.txt gpurun_out/facade_latency.txt $O/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
tail -c 300 $O/bench_full.err
