#!/bin/bash
# round 2, call I: parity at scale + ncu captures of every kernel + launch list
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity_at_scale.py -m gpu -x -q --timeout=800 > $O/pytest_scale.log 2>&1; echo "pytest rc=$?" >> $O/pytest_scale.log
tail -4 $O/pytest_scale.log
cp gpurun_out/parity_at_scale_last.txt $O/ 2>/dev/null
for c in wifi a5 c79 a24; do
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 3 -c 1 -o $O/prof_${c} -f \
    python bench.py --code $c --only --precision 16 --steps 1 --warmup 3 --no-cpu --frames 16384 > $O/ncu_${c}.log 2>&1
  python scripts/ncu_summarise.py $O/prof_${c}.ncu-rep $c r2 --frames 16384 --iters 30 --outdir $O/summaries >> $O/summarise.log 2>&1
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 5 -c 1 -o $O/prof_wifi_op -f \
    python bench.py --code wifi --only --precision 16 --steps 1 --warmup 3 --no-cpu --frames 32768 > $O/ncu_wifi_op.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"encode_kernel|channel_kernel|decode_kernel" -c 4 -o $O/prof_aux -f \
    python scripts/ncu_aux.py > $O/ncu_aux.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_wifi.csv \
    python bench.py --only --steps 2 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
ls -la $O
