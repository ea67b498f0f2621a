#!/bin/bash
# runs the quick C3 bench for every variants/lib_*.so (under gpurun)
for f in variants/lib_*.so; do
  FOVPT_LIB=$PWD/$f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - "$f" <<'PY'
import json, sys
d=json.load(open("gpurun_out/v.json"))
k=d["kernel_ms_per_step"]
print(f"{sys.argv[1]:40s} {d['value']:7.1f} Msamples/s  frame {d['ms_per_step']:.3f} ms  trace {k['k_trace']:.3f} shade {k['k_shade']:.3f}")
PY
done
