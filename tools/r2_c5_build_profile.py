"""C5 scene ingest + BVH build (+ one frame) as a stand-alone program, for an ncu launch list of the builder kernels and
for time-to-first-frame: prints host-side wall times of scene generation, context creation, mesh hand-over,
commit (flatten + upload + device build), first frame.  FOVPT_BUILD_TRACE=1 adds the library's own phase times on stderr."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fovpathtracing_optixcodelatest_b200 import native, workloads
from fovpathtracing_optixcodelatest_b200.renderer import ProbeData, SampleRenderer

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
repeat = int(sys.argv[3]) if len(sys.argv) > 3 else 2          # the second ingest runs on a warm context / warm allocator
wl = workloads.BY_NAME[name]()
t0 = time.perf_counter(); sc = wl.scene(); t1 = time.perf_counter()
ctx = native.Context(0); ctx.set_config(wl.apply(native.default_config())); t2 = time.perf_counter()
print(f"{name}: {sc.n_tris} tris | scene generation {t1 - t0:.3f} s | CUDA context + library {t2 - t1:.3f} s")
for k in range(repeat):
    ta = time.perf_counter()
    ctx.scene_begin()
    for tex in sc.textures: ctx.add_texture(tex)
    for m in sc.meshes: ctx.add_mesh(m.vtx, m.idx, m.material, m.texcoord, m.texture_id)
    tb = time.perf_counter()
    ctx.commit(); tc = time.perf_counter()
    st = ctx.stats()
    print(f"  ingest {k}: hand-over (fovpt_scene_add_*) {tb - ta:.3f} s | commit {tc - tb:.3f} s (device build {st['bvh_build_ms']:.1f} ms, {st['bvh_nodes']} nodes) "
          f"| total {tc - ta:.3f} s = {sc.n_tris / (tc - ta) / 1e6:.1f} Mtris/s")
ctx.close()
t3 = time.perf_counter()
r = SampleRenderer(sc, 0, wl.apply(native.default_config())); t4 = time.perf_counter()
r.resize((wl.width, wl.height)); r.setCamera(wl.camera); r.setProbe(ProbeData(wl.probe()).BuildCDF())
r.launchParams.frame.c_x, r.launchParams.frame.c_y = wl.gaze_xy()
t5 = time.perf_counter()
for _ in range(frames):
    r.render()
t6 = time.perf_counter()
print(f"  renderer: ctor (ingest + build) {t4 - t3:.3f} s | probe + buffers {t5 - t4:.3f} s | {frames} frame(s) {t6 - t5:.3f} s | "
      f"time to first frame on a warm process {t6 - t3:.3f} s")
r.close()
