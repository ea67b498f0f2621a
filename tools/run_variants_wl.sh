#!/bin/bash
# quick bench of every variants/lib_*.so on the given workloads (under gpurun): tools/run_variants_wl.sh c5 c2
for wl in "$@"; do
for f in variants/lib_*.so; do
  FOVPT_LIB=$PWD/$f python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - "$f" "$wl" <<'PY'
import json, sys
d=json.load(open("gpurun_out/v.json"))
k=d["kernel_ms_per_step"]
print(f"{sys.argv[2]} {sys.argv[1]:36s} {d['value']:7.1f} Msamples/s  step {d['ms_per_step']:.3f} ms  trace {k['k_trace']:.3f} shade {k['k_shade']:.3f}")
PY
done
done
