#!/bin/bash
# multi-GPU evidence (run under gpurun --gpus N): tools/r2_multi.sh <tag> <N> <workloads...>
TAG=$1; N=$2; shift 2
for wl in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wl --steps 6 --warmup 3 \
    > gpurun_out/bench_${TAG}_${wl}_n$N.json 2> gpurun_out/bench_${TAG}_${wl}_n$N.err || tail -5 gpurun_out/bench_${TAG}_${wl}_n$N.err
  python - "$TAG" "$wl" "$N" <<'PY'
import json, sys
t, wl, n = sys.argv[1:4]
try:
    d = json.load(open(f"gpurun_out/bench_{t}_{wl}_n{n}.json"))
    print(wl, "N", n, round(d["value"], 1), "Msamples/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), "parity", d.get("multi_gpu_parity", {}).get("identical"))
except Exception as e:
    print(wl, n, "failed", e)
PY
done
