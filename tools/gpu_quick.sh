#!/bin/bash
# quick GPU loop used during kernel work: parity suite, then the C3 bench line (no CPU baseline)
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print({k:d[k] for k in ("value","ms_per_step","rays_per_s","kernel_ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"])
PY
