"""Quick GPU-vs-oracle check used during bring-up (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.orc import Oracle, Opts, FovCfg
from fovpathtracing_optixcodelatest_b200 import scenes, native
from fovpathtracing_optixcodelatest_b200.renderer import SampleRenderer, ProbeData

def run(scene, cam, w, h, probe_img, ri, ro, name):
    o = Oracle("port"); o.add_scene(scene); o.commit(); o.set_probe(probe_img)
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, w / h)
    lp, accum, frame = o.make_params(w, h, cam.eye, U, V, W)
    t = time.time()
    o.render(lp, FovCfg.sv5(inner_radius=ri, outer_radius=ro), Opts.sv5())
    t_cpu = time.time() - t
    cnt = o.counters()

    cfg = native.default_config(); cfg.inner_radius = ri; cfg.outer_radius = ro
    r = SampleRenderer(scene, 0, cfg)
    r.resize((w, h)); r.setCamera(cam)
    pd = ProbeData(probe_img).BuildCDF(); r.setProbe(pd)
    r.launchParams.frame.c_x, r.launchParams.frame.c_y = w // 2, h // 2
    gU = np.array(r.launchParams.camera.U.tolist(), np.float32)
    print(name, "camera U equal:", (gU.view(np.uint32) == U.view(np.uint32)).all())
    px, cx, py, cy = o.get_probe()
    print(name, "cdf equal:", all((a.view(np.uint32) == b.view(np.uint32)).all() for a, b in
          ((px, pd.pdfValuesX), (cx, pd.cdfValuesX), (py, pd.pdfValuesY), (cy, pd.cdfValuesY))))
    r.render()
    st0 = r.stats()
    r.launchParams.frame.subframe_index = 0
    r.ctx.reset_stats()
    t = time.time(); r.render(); t_gpu = time.time() - t
    st = r.stats()
    g_accum = r.downloadAccum(); g_frame = r.downloadPixels()
    print(name, "cpu s", round(t_cpu, 3), "gpu wall s", round(t_gpu, 4), "gpu frame ms", st["last_frame_ms"],
          {k: round(v, 3) for k, v in st.items() if k.startswith("last_")})
    print(name, "oracle", cnt)
    print(name, "gpu   ", {k: st[k] for k in ("samples", "radiance_rays", "shadow_rays", "kernel_launches", "bvh_nodes", "bvh_tris", "bvh_build_ms")})
    a = accum[..., :3].astype(np.float64); g = g_accum[..., :3].astype(np.float64)
    print(name, "mean oracle", a.mean(), "gpu", g.mean(), "nan", np.isnan(g).sum())
    diff = np.abs(a - g)
    close = (diff <= 1e-4 * (1 + np.abs(a))).all(axis=2)
    print(name, "pixels matching to 1e-4:", close.mean(), "rMSE", ((a - g) ** 2).mean() / (a.mean() ** 2 + 1e-12))
    print(name, "frame equal frac", (frame == g_frame).mean())
    # primary hits per pass
    from oracle.orc import Opts as O
    from fovpathtracing_optixcodelatest_b200.structs import LaunchParams
    import ctypes as C
    for (factor, fill, rin, rout, spp, lw, lh, offx, offy) in (
            (4, 4, float(ro), 1e9, 8, w // 4, h // 4, 0, 0),
            (2, 2, float(ri), float(ro + 2), 16, ro + 2, ro + 2, (w // 2 - (ro + 2)) & 0xffffffff, (h // 2 - (ro + 2)) & 0xffffffff),
            (1, 1, 0.0, float(ri + 1), 32, 2 * (ri + 1), 2 * (ri + 1), (w // 2 - (ri + 1)) & 0xffffffff, (h // 2 - (ri + 1)) & 0xffffffff)):
        lp2, acc2, fr2 = o.make_params(w, h, cam.eye, U, V, W)
        lp2.frame.factor_x = lp2.frame.factor_y = factor; lp2.frame.factor_z = 1; lp2.frame.fillSize = fill
        lp2.frame.r_inner = rin; lp2.frame.r_outer = rout; lp2.samples_per_launch = 1
        lp2.frame.offset_x = offx; lp2.frame.offset_y = offy
        tri = np.zeros(lw * lh, np.int32); tuv = np.zeros((lw * lh, 3), np.float32); dr = np.zeros((lw * lh, 3), np.float32)
        op = O.sv5(max_depth=0); op.primary_tri = tri.ctypes.data; op.primary_tuv = tuv.ctypes.data; op.primary_dir = dr.ctypes.data
        o.launch(lp2, lw, lh, op)
        lp3 = LaunchParams(); C.memmove(C.byref(lp3), C.byref(lp2), C.sizeof(lp3))
        lp3.probe = r.launchParams.probe
        gtri, gtuv, gdir = r.ctx.primary_hits(lp3, lw, lh)
        print(name, f"pass f={factor}: tri equal", (tri == gtri).all(), "mismatch", int((tri != gtri).sum()), "of", tri.size,
              "tuv bits equal", (tuv.view(np.uint32) == gtuv.view(np.uint32)).all(),
              "dir bits equal", (dr.view(np.uint32)[tri > -2] == gdir.view(np.uint32)[tri > -2]).all(),
              "active", int((tri > -2).sum()), "hits", int((tri >= 0).sum()))
    # random rays
    rng = np.random.default_rng(5)
    n = 200000
    lo = np.array([m.vtx.min(0) for m in scene.meshes]).min(0); hi = np.array([m.vtx.max(0) for m in scene.meshes]).max(0)
    org = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, np.full((n, 1), 0.01, np.float32), d.astype(np.float32), np.full((n, 1), 1e16, np.float32)], 1).astype(np.float32)
    ot, otuv = o.trace(rays, 0); gt, gtuv2 = r.ctx.trace(rays, 0)
    print(name, "random closest: equal", (ot == gt).all(), int((ot != gt).sum()), "tuv bits", (otuv.view(np.uint32) == gtuv2.view(np.uint32)).all(), "hit frac", (ot >= 0).mean())
    oa, _ = o.trace(rays, 1); ga, _ = r.ctx.trace(rays, 1)
    print(name, "random any: equal", (oa == ga).all(), int((oa != ga).sum()), "occluded frac", oa.mean())
    r.close()

if __name__ == "__main__":
    sc = scenes.cornell_box()
    run(sc, scenes.CORNELL_CAMERA, 960, 540, scenes.constant_probe(64, 32, 2.5), 74, 241, "cornell")
