"""Where do GPU and oracle frames differ?  (run under gpurun)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orc
from fovpathtracing_optixcodelatest_b200 import native, scenes, workloads
from fovpathtracing_optixcodelatest_b200.renderer import SampleRenderer, ProbeData

def run(depth, frames, acc, w=384, h=216, sub0=0):
    wl = workloads.scaled(workloads.c2(), w, h)
    ri, ro = wl.inner_radius, wl.outer_radius
    sc, cam, probe = wl.scene(), wl.camera, wl.probe()
    o = orc.Oracle("port"); o.add_scene(sc); o.commit(); o.set_probe(probe)
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, w / h)
    lp, accum, frame = o.make_params(w, h, cam.eye, U, V, W)
    cfg = native.default_config(); cfg.inner_radius, cfg.outer_radius = ri, ro
    cfg.max_depth = depth; cfg.accumulate = acc; cfg.seed_all_passes_with_subframe = acc
    r = SampleRenderer(sc, 0, cfg); r.resize((w, h)); r.setCamera(cam); r.setProbe(ProbeData(probe).BuildCDF())
    r.launchParams.frame.c_x, r.launchParams.frame.c_y = w // 2, h // 2
    lp.frame.subframe_index = sub0; r.launchParams.frame.subframe_index = sub0
    fc = orc.FovCfg.sv5(inner_radius=ri, outer_radius=ro); op = orc.Opts.sv5(max_depth=depth, accumulate=acc, seed_all_passes_with_subframe=acc)
    for _ in range(frames):
        o.render(lp, fc, op); r.render()
    g = r.downloadAccum()[..., :3].astype(np.float64); a = accum[..., :3].astype(np.float64)
    rel = np.abs(a - g).max(axis=2) / (1 + np.abs(a).max(axis=2))
    cnt = o.counters(); st = r.stats()
    print(f"depth {depth} frames {frames} acc {acc}: close(1e-4) {(rel<=1e-4).mean():.4f} close(1e-6) {(rel<=1e-6).mean():.4f} max rel {rel.max():.3g} rmse {((a-g)**2).mean()/a.mean()**2:.3g}"
          f" rays o {cnt['radiance_rays']} g {st['radiance_rays']} shadow o {cnt['shadow_rays']} g {st['shadow_rays']}")
    # by ring
    y, x = np.mgrid[0:h, 0:w]; rr = np.hypot(x - w // 2, y - h // 2)
    for name, m in (("fovea", rr <= ri), ("mid", (rr > ri + 1) & (rr <= ro)), ("periph", rr > ro + 2)):
        print(f"    {name:6s} close(1e-4) {(rel[m]<=1e-4).mean():.4f}  n={m.sum()}")
    bad = np.argwhere(rel > 1e-4)[:6]
    for (yy, xx) in bad: print('      px', xx, yy, a[yy, xx], g[yy, xx])
    r.close(); o.close()

run(4, 1, 1)
run(4, 2, 1)
run(4, 1, 1, sub0=1)
run(4, 1, 1, sub0=2)
run(4, 1, 0, sub0=1)
