#!/bin/bash
# round 2, call I: ncu captures of every kernel (summarised on the box; the .ncu-rep files stay there) + launch list
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2i; mkdir -p $O
for c in wifi a5 c79 a24; do
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 3 -c 1 -o $O/prof_${c} -f \
    python bench.py --code $c --only --precision 16 --steps 1 --warmup 3 --no-cpu --frames 16384 > $O/ncu_${c}.log 2>&1
  python scripts/ncu_summarise.py $O/prof_${c}.ncu-rep $c r2 --frames 16384 --iters 30 --outdir $O/summaries >> $O/summarise.log 2>&1
  rm -f $O/prof_${c}.ncu-rep
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 5 -c 1 -o $O/prof_wifi_op -f \
    python bench.py --code wifi --only --precision 16 --steps 1 --warmup 3 --no-cpu --frames 32768 > $O/ncu_wifi_op.log 2>&1
grep -o '"operating_point": {[^}]*}' $O/ncu_wifi_op.log > $O/wifi_op_point.txt
python scripts/ncu_summarise.py $O/prof_wifi_op.ncu-rep wifi_op r2 --frames 32768 --iters 9.85 --outdir $O/summaries >> $O/summarise.log 2>&1
rm -f $O/prof_wifi_op.ncu-rep
timeout 400 ncu --set full --clock-control none -k regex:"encode_kernel|channel_kernel|decode_kernel" -c 4 -o $O/prof_aux -f \
    python scripts/ncu_aux.py > $O/ncu_aux.log 2>&1
python scripts/ncu_raw_text.py $O/prof_aux.ncu-rep > $O/summaries/ncu_raw_aux_r2.txt 2>> $O/summarise.log
rm -f $O/prof_aux.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bench_wifi.csv \
    python bench.py --only --steps 2 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
du -sh $O; ls -la $O $O/summaries
