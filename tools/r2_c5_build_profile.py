"""C5 scene ingest + BVH build (+ one frame) as a stand-alone program, for an ncu launch list of the builder kernels and
for time-to-first-frame: prints host-side wall times of scene generation, upload + build, first frame."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fovpathtracing_optixcodelatest_b200 import native, workloads
from fovpathtracing_optixcodelatest_b200.renderer import ProbeData, SampleRenderer

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = workloads.BY_NAME[name]()
t0 = time.perf_counter(); sc = wl.scene(); t1 = time.perf_counter()
r = SampleRenderer(sc, 0, wl.apply(native.default_config())); t2 = time.perf_counter()
r.resize((wl.width, wl.height)); r.setCamera(wl.camera); r.setProbe(ProbeData(wl.probe()).BuildCDF())
r.launchParams.frame.c_x, r.launchParams.frame.c_y = wl.gaze_xy()
t3 = time.perf_counter()
for _ in range(frames):
    r.render()
t4 = time.perf_counter()
st = r.stats()
print(f"{name}: {sc.n_tris} tris | scene generation {t1 - t0:.3f} s | upload + BVH build {t2 - t1:.3f} s (device build {st['bvh_build_ms']:.1f} ms, "
      f"{st['bvh_nodes']} nodes) | probe + buffers {t3 - t2:.3f} s | {frames} frame(s) {t4 - t3:.3f} s | time to first frame (ingest + build + frame) {t4 - t1:.3f} s")
r.close()
