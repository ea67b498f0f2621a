#!/bin/bash
# what the per-kernel CUDA events cost a frame (run under gpurun)
for e in 1 0; do
  FOVPT_KERNEL_TIMING=$e python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ev_$e.json 2>/dev/null
  FOVPT_KERNEL_TIMING=$e python bench.py --steps 8 --warmup 3 --no-cpu-baseline --emulate-world 8 > gpurun_out/ev8_$e.json 2>/dev/null
done
python - <<'PY'
import json
for e in (1, 0):
    d = json.load(open(f"gpurun_out/ev_{e}.json")); d8 = json.load(open(f"gpurun_out/ev8_{e}.json"))
    print("events", e, "frame ms", round(d["ms_per_step"], 4), "Msamples/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "| 1/8 share ms", round(d8["ms_per_step"], 4))
PY
