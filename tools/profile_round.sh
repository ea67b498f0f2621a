#!/bin/bash
# refreshes the judged evidence for the current build (run under gpurun, ONE ncu pass per call):
#   tools/profile_round.sh <tag> bench     smoke, GPU tests, the C3 bench line and the other workloads (no profiler)
#   tools/profile_round.sh <tag> launches  ncu per-launch durations of two C3 frames (after the same command ran plain)
#   tools/profile_round.sh <tag> full      one ncu --set full capture of k_trace / k_shade (after the same command ran plain)
TAG=${1:-r01}; MODE=${2:-bench}
case $MODE in
bench)
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
  python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
  python bench.py --steps 8 --warmup 3 > gpurun_out/bench_${TAG}_c3.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
  for wl in c1 c2 c4 c5; do
    python bench.py --workload $wl --steps 4 --warmup 3 > gpurun_out/bench_${TAG}_$wl.json 2>> gpurun_out/bench_$TAG.err
  done
  python - "$TAG" <<'PY'
import json, sys
for wl in ("c3", "c1", "c2", "c4", "c5"):
    try:
        d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}_{wl}.json"))
        print(wl, round(d["value"], 1), "Msamples/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), "rays/s", f'{d["rays_per_s"]:.3g}',
              "frac", round(d.get("roofline", {}).get("frac", 0), 3), "cpu", round(d.get("cpu_baseline", {}).get("value", 0), 2))
    except Exception as e:
        print(wl, "failed", e)
PY
  ;;
launches)
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  tail -2 gpurun_out/ncu_launch.log | cut -c1-160
  ;;
full)
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -s 6 -c 4 -o gpurun_out/prof_${TAG}_trace_shade python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log | cut -c1-160
  ;;
esac
