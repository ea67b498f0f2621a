"""CPU, build container only (skipped where oracle/_ref is absent): the restatement against the
reference's own code compiled unmodified, live, at larger sizes than the committed fixtures."""
import numpy as np

from fovpathtracing_optixcodelatest_b200 import scenes
from fovpathtracing_optixcodelatest_b200.structs import Material


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_bsdf_random_bitexact(ref):
    R, P = ref.Oracle("reference"), ref.Oracle("port")
    rng = np.random.default_rng(123)
    for i in range(3000):
        m = Material.default()
        m.color.set(rng.random(3)); m.metallic = rng.random(); m.subsurface = rng.random() if i % 2 else 0.0
        m.specular = rng.random(); m.roughness = rng.random(); m.specularTint = rng.random(); m.clearcoat = rng.random()
        m.clearcoatGloss = rng.random(); m.transmission = float(rng.choice([0.0, 0.4, 1.0, rng.random()]))
        if i % 7 == 0:
            m.eta = 0.0                       # IOR inferred from specular (Material.h:40-46)
        N = rng.normal(size=3); N /= np.linalg.norm(N); V = rng.normal(size=3); V /= np.linalg.norm(V)
        if V @ N < 0:
            V = -V
        L = rng.normal(size=3); L /= np.linalg.norm(L)
        alb = rng.random(3); seed = int(rng.integers(0, 2 ** 32)); e = (1.0, 1.4) if i % 2 else (1.4, 1.0)
        assert (bits(R.bsdf_sample(m, alb, e[0], e[1], N, V, seed)) == bits(P.bsdf_sample(m, alb, e[0], e[1], N, V, seed))).all()
        assert (bits(R.bsdf_eval(m, alb, e[0], e[1], N, V, L)) == bits(P.bsdf_eval(m, alb, e[0], e[1], N, V, L))).all()
    R.close(); P.close()


def test_atrium_frame_bitexact_and_tonemap_closure(ref):
    """one foveated 320x180 frame of the Sponza-class scene: accumulation buffer, RGBA8 frame, ray and
    BVH-visit counters identical; and the re-typed tonemap of orc_tonemap closes against the frame
    buffer the unmodified raygen program wrote (oracle/ref_driver.cpp note)."""
    sc = scenes.atrium(detail=0.25)
    probe = scenes.sky_probe(256, 128, 2)
    out = {}
    for kind in ("reference", "port"):
        o = ref.Oracle(kind)
        o.add_scene(sc); o.commit(); o.set_probe(probe)
        cam = scenes.SPONZA_CAMERA
        U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, 320 / 180)
        lp, accum, frame = o.make_params(320, 180, cam.eye, U, V, W)
        o.render(lp, ref.FovCfg.sv5(inner_radius=24, outer_radius=80), ref.Opts.sv5(jitter_yx=1))
        out[kind] = (accum.copy(), frame.copy(), o.counters(), o)
    a, b = out["reference"], out["port"]
    assert (bits(a[0]) == bits(b[0])).all() and (a[1] == b[1]).all() and a[2] == b[2]
    R = a[3]
    acc, frm = a[0].reshape(-1, 4), a[1].reshape(-1)
    for i in range(0, len(frm), 97):
        assert R.tonemap(acc[i, :3]) == int(frm[i])
    a[3].close(); b[3].close()


def test_primary_hits_bitexact(ref):
    sc = scenes.cornell_box()
    got = {}
    for kind in ("reference", "port"):
        o = ref.Oracle(kind)
        o.add_scene(sc); o.commit(); o.set_probe(scenes.constant_probe(8, 4, 2.5))
        cam = scenes.CORNELL_CAMERA
        U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, 1.0)
        lp, accum, frame = o.make_params(128, 128, cam.eye, U, V, W)
        lp.frame.factor_x = lp.frame.factor_y = lp.frame.factor_z = 1; lp.frame.fillSize = 1
        lp.frame.r_inner, lp.frame.r_outer = 0.0, 1e9; lp.samples_per_launch = 1
        tri = np.zeros(128 * 128, np.int32); tuv = np.zeros((128 * 128, 3), np.float32); d = np.zeros((128 * 128, 3), np.float32)
        op = ref.Opts.sv5(jitter_yx=1); op.primary_tri, op.primary_tuv, op.primary_dir = tri.ctypes.data, tuv.ctypes.data, d.ctypes.data
        o.launch(lp, 128, 128, op)
        got[kind] = (tri, tuv, d); o.close()
    for x, y in zip(got["reference"], got["port"]):
        assert (x.view(np.uint32) == y.view(np.uint32)).all()
    assert (got["port"][0] >= 0).mean() > 0.3


def test_shadow_catcher_frame_bitexact(ref):
    """MATERIAL_FLAG_SHADOW_CATCHER (Material.h:9): secondary rays pass through the catcher (deviceProgram.cu:646-651),
    primary hits accumulate SampleShadow into alpha (:691-694, :347-387) and composite the backplate (:558)"""
    sc = scenes.cornell_with_shadow_catcher()
    out = {}
    for kind in ("reference", "port"):
        o = ref.Oracle(kind)
        o.add_scene(sc); o.commit(); o.set_probe(scenes.sky_probe(64, 32, seed=4))
        cam = scenes.CORNELL_CAMERA
        U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, 160 / 90)
        lp, accum, frame = o.make_params(160, 90, cam.eye, U, V, W)
        o.render(lp, ref.FovCfg.sv5(inner_radius=12, outer_radius=40), ref.Opts.sv5(jitter_yx=1))
        out[kind] = (accum.copy(), frame.copy(), o.counters()); o.close()
    a, b = out["reference"], out["port"]
    assert (bits(a[0]) == bits(b[0])).all() and (a[1] == b[1]).all() and a[2] == b[2]
