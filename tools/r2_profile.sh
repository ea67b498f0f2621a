#!/bin/bash
# round-2 evidence (run under gpurun, one ncu family per call):
#   tools/r2_profile.sh <tag> launches c3|c5    ncu per-launch durations of one timed step
#   tools/r2_profile.sh <tag> full c3|c5        ncu --set full of the k_trace (+ k_shade on c3) launches of one timed step
TAG=$1; MODE=$2; WL=$3
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline"
case $MODE-$WL in
launches-*)
  $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}_$WL.csv $CMD > gpurun_out/ncu_launch.log 2>&1
  tail -1 gpurun_out/ncu_launch.log | cut -c1-200 ;;
full-c3)   # per step: raygen, 10 x k_trace, 9 x k_shade, ... ; counting pass (1 subframe) + 3 warm-up steps come first = 4 x 19 trace/shade launches
  $CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -s 76 -c 19 -o gpurun_out/prof_${TAG}_c3 $CMD > gpurun_out/ncu_full.log 2>&1
  tail -1 gpurun_out/ncu_full.log | cut -c1-200
  python tools/ncu_summary.py gpurun_out/prof_${TAG}_c3.ncu-rep gpurun_out/ncu_${TAG}_c3.json "tools/r2_profile.sh $TAG full c3: ncu --set full --clock-control none -k regex:k_trace|k_shade -s 76 -c 19; $CMD (the 10 k_trace + 9 k_shade launches of the first timed step = 8 subframes in one wavefront)" > /dev/null
  ncu -i gpurun_out/prof_${TAG}_c3.ncu-rep --page source --csv --kernel-name regex:k_trace --launch-skip 2 --launch-count 1 2>/dev/null | gzip > gpurun_out/ncu_${TAG}_c3_trace_source.csv.gz
  ncu -i gpurun_out/prof_${TAG}_c3.ncu-rep --page source --csv --kernel-name regex:k_shade --launch-skip 1 --launch-count 1 2>/dev/null | gzip > gpurun_out/ncu_${TAG}_c3_shade_source.csv.gz
  rm -f gpurun_out/prof_${TAG}_c3.ncu-rep ;;
full-c5)   # per step (one subframe = 265 M samples): 6 row-chunk groups x 6 k_trace launches = 36, all captured; counting pass + 3 warm-up steps = 144 launches first
  $CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_trace" -s 144 -c 36 -o gpurun_out/prof_${TAG}_c5 $CMD > gpurun_out/ncu_full.log 2>&1
  tail -1 gpurun_out/ncu_full.log | cut -c1-200
  python tools/ncu_summary.py gpurun_out/prof_${TAG}_c5.ncu-rep gpurun_out/ncu_${TAG}_c5.json "tools/r2_profile.sh $TAG full c5: ncu --set full --clock-control none -k regex:k_trace -s 144 -c 36; $CMD (all 36 k_trace launches of the timed step = 6 row-chunk groups of at most 48 Mi paths, 10 M-triangle scene, tree larger than L2)" > /dev/null
  rm -f gpurun_out/prof_${TAG}_c5.ncu-rep ;;
esac
cp gpurun_out/plain.log gpurun_out/plain_${TAG}_${MODE}_$WL.log 2>/dev/null
