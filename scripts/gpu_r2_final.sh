#!/bin/bash
# round 2, call R: final ncu captures of the shipped kernels (summaries only), then whole suite + smoke + both bench arms
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2u; mkdir -p $O
for c in wifi a5 c79 a24; do
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 3 -c 1 -o $O/prof_${c} -f \
    python bench.py --code $c --only --precision 16 --steps 1 --warmup 3 --no-cpu --frames 16384 > $O/ncu_${c}.log 2>&1
  python scripts/ncu_summarise.py $O/prof_${c}.ncu-rep $c r2 --frames 16384 --iters 30 --outdir $O/summaries >> $O/summarise.log 2>&1
  rm -f $O/prof_${c}.ncu-rep
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 5 -c 1 -o $O/prof_wifi_op -f \
    python bench.py --code wifi --only --precision 16 --steps 1 --warmup 3 --no-cpu --frames 32768 > $O/ncu_wifi_op.log 2>&1
python scripts/ncu_summarise.py $O/prof_wifi_op.ncu-rep wifi_op r2 --frames 32768 --iters 9.85 --outdir $O/summaries >> $O/summarise.log 2>&1
rm -f $O/prof_wifi_op.ncu-rep
timeout 400 ncu --set full --clock-control none -k regex:"encode_kernel|channel_kernel|decode_kernel" -c 4 -o $O/prof_aux -f \
    python scripts/ncu_aux.py > $O/ncu_aux.log 2>&1
python scripts/ncu_raw_text.py $O/prof_aux.ncu-rep > $O/summaries/ncu_raw_aux_r2.txt 2>> $O/summarise.log
rm -f $O/prof_aux.ncu-rep
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > $O/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/pytest.log
tail -4 $O/pytest.log
cp gpurun_out/parity_at_scale_last.txt gpurun_out/facade_latency.txt $O/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_wifi_r2.csv \
    python bench.py --only --no-cpu --steps 5 --warmup 3 > $O/ncu_launches.log 2>&1
du -sh $O
