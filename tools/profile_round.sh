#!/bin/bash
# refreshes the judged evidence for the current build (run under gpurun): smoke, GPU tests, bench line, ncu launch list,
# one ncu --set full capture of the dominant kernels.  usage: tools/profile_round.sh <tag>
TAG=${1:-r01}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_${TAG}_c3.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -s 6 -c 4 -o gpurun_out/prof_${TAG}_trace_shade python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-160
