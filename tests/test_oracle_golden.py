"""CPU: the oracle restatement (oracle/liboracle.so, kind "port") against the golden fixtures that
tests/golden/make_golden.py generated from the REFERENCE's own sources compiled for the host
(oracle/_ref), and against the known-answer vectors of SURVEY.md Appendix B.  Bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from fovpathtracing_optixcodelatest_b200 import scenes
from fovpathtracing_optixcodelatest_b200.structs import Material


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture()
def P(port):
    o = port.Oracle("port")
    yield o
    o.close()


def test_appendix_b_known_answers(P):
    # cuda/random.h:34-49, maths.h:176-210
    assert [P.tea4(0, 0), P.tea4(1, 0), P.tea4(0, 1), P.tea4(5, 0), P.tea4(518400 + 960, 7)] == \
        [1576399551, 3231205618, 1569133783, 1535087949, 693573382]
    r, s = P.rnd_seq(P.tea4(0, 0), 2)
    assert abs(r[0] - 0.294449925) < 1e-9 and abs(r[1] - 0.695515215) < 1e-9 and s == 4105309513
    u, f = P.random_seq(P.tea4(0, 0), 3)
    assert list(u) == [211779248, 4095045287, 3168626450]
    np.testing.assert_allclose(f, [0.0493086986, 0.953452051, 0.737753332], rtol=0, atol=1e-9)
    U, V, W = P.camera_uvw((-1293.07, 154.681, 0), (200, 0, 0), (0, 1, 0), 45.0, 1920 / 1080)
    np.testing.assert_allclose(U, [0, 0, 1105.35083], atol=1e-3); np.testing.assert_allclose(V, [64.0709686, 618.449829, 0], atol=1e-3)
    np.testing.assert_allclose(W, [1493.06995, -154.681, 0], atol=1e-3)


def test_appendix_b_bsdf_and_probe(P):
    m = Material.default()
    o = P.bsdf_sample(m, (.5, .5, .5), 1.0, 1.4, (0, 1, 0), np.array([0.3, 0.8, 0.1]) / np.linalg.norm([0.3, 0.8, 0.1]), P.tea4(5, 0))
    np.testing.assert_allclose(o[:3], [-0.249102, -0.964911, -0.083034], atol=2e-6)
    assert abs(o[3] - 0.388768) < 2e-6 and abs(o[5] - 0.080581) < 2e-6
    img = np.zeros((4, 8, 4), np.float32)
    img[..., :3] = (1 + np.arange(32, dtype=np.float32)).reshape(4, 8, 1); img[..., 3] = 1
    P.set_probe(img)
    pdfX, cdfX, pdfY, cdfY = P.get_probe()
    np.testing.assert_allclose(cdfY, [0.0681818202, 0.25757575, 0.568181813, 1], rtol=1e-7)
    assert abs(pdfX[0, 0] - 0.027777778) < 1e-8 and cdfX[0, 7] == 1.0
    s = P.probe_sample(1234)
    np.testing.assert_allclose(s[:3], [-0.707107008, -4.37113883e-08, 0.707106531], atol=1e-6)
    assert s[3] == 24.0 and abs(s[6] - 0.0736881271) < 1e-7
    uv = P.probe_eval(np.array([0.3, 0.8, 0.1]) / np.linalg.norm([0.3, 0.8, 0.1]))[:2]
    np.testing.assert_allclose(uv, [0.551208, 0.119823], atol=1e-6)


def test_rng_golden(P):
    g = np.load(os.path.join(GOLDEN, "rng.npz"))
    assert [P.tea4(int(a), int(b)) for a, b in g["tea_in"]] == list(g["tea_out"])
    r, s = P.rnd_seq(int(g["tea_out"][0]), 64)
    assert (bits(r) == bits(g["rnd"])).all() and s == int(g["rnd_seed"])
    u, f = P.random_seq(int(g["tea_out"][0]), 256)
    assert (u == g["rand_u32"]).all() and (bits(f) == bits(g["rand_f"])).all()
    uvw = np.concatenate(P.camera_uvw((-1293.07, 154.681, 0), (200, 0, 0), (0, 1, 0), 45.0, 1920 / 1080))
    assert (bits(uvw) == bits(g["sponza_uvw"])).all()


def test_bsdf_golden(P):
    g = np.load(os.path.join(GOLDEN, "bsdf.npz"))
    n = len(g["seeds"])
    branches = set()
    for i in range(n):
        m = Material.from_buffer_copy(g["mats"][i].tobytes())
        s = P.bsdf_sample(m, g["alb"][i], float(g["eta"][i, 0]), float(g["eta"][i, 1]), g["N"][i], g["V"][i], int(g["seeds"][i]))
        e = P.bsdf_eval(m, g["alb"][i], float(g["eta"][i, 0]), float(g["eta"][i, 1]), g["N"][i], g["V"][i], g["L"][i])
        assert (bits(s) == bits(g["sample"][i])).all(), i
        assert (bits(e) == bits(g["eval"][i])).all(), i
        branches.add(int(s[9]))
    assert branches == {2, 4, 5, 7}          # every RNG-consumption branch of BSDFSample is exercised


def test_probe_and_tonemap_golden(P):
    g = np.load(os.path.join(GOLDEN, "probe.npz"))
    P.set_probe(scenes.sky_probe(64, 32, seed=4))
    for a, k in zip(P.get_probe(), ("pdfX", "cdfX", "pdfY", "cdfY")):
        assert (bits(a) == bits(g[k])).all(), k
    for s, exp in zip(g["seeds"], g["sample"]):
        assert (bits(P.probe_sample(int(s))) == bits(exp)).all()
    for d, exp in zip(g["dirs"], g["eval"]):
        assert (bits(P.probe_eval(d)) == bits(exp)).all()
    assert [P.tonemap(a) for a in g["tm_in"]] == list(g["tm_out"])


def _render(port, probe, gaze, frames):
    sc = scenes.cornell_box()
    sc.textures.append(scenes.checker_texture(32)); sc.meshes[0].texcoord = np.array([[0, 0], [0, 3], [3, 3], [3, 0]], np.float32); sc.meshes[0].texture_id = 0
    o = port.Oracle("port")
    o.add_scene(sc); o.commit(); o.set_probe(probe)
    cam = scenes.CORNELL_CAMERA
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, 160 / 90)
    lp, accum, frame = o.make_params(160, 90, cam.eye, U, V, W, gaze=gaze)
    for _ in range(frames):
        # jitter_yx=1: the fixture came from g++'s right-to-left evaluation of deviceProgram.cu:479
        o.render(lp, port.FovCfg.sv5(inner_radius=12, outer_radius=40), port.Opts.sv5(threads=1, jitter_yx=1))   # threads=1: out-of-frame splats clamp onto border pixels, sequential last writer wins
    c = o.counters(); o.close()
    return accum, frame, np.array([c[k] for k in sorted(c)], np.uint64)


def test_foveated_frames_golden(port):
    """whole foveated frames (three passes, textures, both materials, two subframes; and a gaze close
    to the frame corner so that the unsigned offsets wrap) match the reference build."""
    g = np.load(os.path.join(GOLDEN, "frames.npz"))
    a, f, c = _render(port, scenes.sky_probe(64, 32, seed=4), None, 2)
    assert (bits(a) == bits(g["accum_sky"])).all() and (f == g["frame_sky"]).all() and (c == g["counts_sky"]).all()
    a, f, c = _render(port, scenes.constant_probe(16, 8, 2.5), (20, 70), 1)
    assert (bits(a) == bits(g["accum_wrap"])).all() and (f == g["frame_wrap"]).all() and (c == g["counts_wrap"]).all()


def test_shadow_catcher_frame_golden(port):
    """a frame with a shadow-catcher floor (Material.h:9, deviceProgram.cu:646-651,691-694,347-387) matches the reference build"""
    g = np.load(os.path.join(GOLDEN, "frames.npz"))
    o = port.Oracle("port")
    o.add_scene(scenes.cornell_with_shadow_catcher()); o.commit(); o.set_probe(scenes.sky_probe(64, 32, seed=4))
    cam = scenes.CORNELL_CAMERA
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, 160 / 90)
    lp, accum, frame = o.make_params(160, 90, cam.eye, U, V, W)
    o.render(lp, port.FovCfg.sv5(inner_radius=12, outer_radius=40), port.Opts.sv5(threads=1, jitter_yx=1))
    c = o.counters(); o.close()
    assert (bits(accum) == bits(g["accum_catcher"])).all() and (frame == g["frame_catcher"]).all()
    assert (np.array([c[k] for k in sorted(c)], np.uint64) == g["counts_catcher"]).all()


def test_bvh_never_changes_a_hit(P):
    """closest / any-hit through the canonical BVH == brute force over all triangles"""
    sc = scenes.atrium(detail=0.12)
    P.add_scene(sc); P.commit()
    rng = np.random.default_rng(9)
    n = 4000
    lo = np.array([m.vtx.min(0) for m in sc.meshes]).min(0); hi = np.array([m.vtx.max(0) for m in sc.meshes]).max(0)
    org = rng.uniform(lo, hi, (n, 3)); d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, np.full((n, 1), 0.01), d, np.full((n, 1), 1e16)], 1).astype(np.float32)
    t0, uv0 = P.trace(rays, 0); t2, uv2 = P.trace(rays, 2)
    assert (t0 == t2).all() and (bits(uv0) == bits(uv2)).all() and (t0 >= 0).mean() > 0.5
    a1, _ = P.trace(rays, 1); a3, _ = P.trace(rays, 3)
    assert (a1 == a3).all() and 0.05 < a1.mean() < 1.0


def test_empty_and_tiny_scenes(P):
    """edge cases: no geometry (every ray misses -> backplate only), one triangle (single-leaf tree)"""
    P.commit(); P.set_probe(scenes.constant_probe(8, 4, 2.5))
    U, V, W = P.camera_uvw((0, 0, -5), (0, 0, 0), (0, 1, 0), 45.0, 1.0)
    lp, accum, frame = P.make_params(16, 16, (0, 0, -5), U, V, W)
    cfg = __import__("oracle.orc", fromlist=["x"]).FovCfg.sv5(foveated=0, spp_uniform=2)
    P.render(lp, cfg, __import__("oracle.orc", fromlist=["x"]).Opts.sv5())
    assert np.allclose(accum[..., :3], 2.5) and P.counters()["shadow_rays"] == 0
    P.add_mesh([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], [[0, 1, 2]], Material.diffuse((.5, .5, .5)))
    P.commit()
    tri, _ = P.trace(np.array([[0, 0, -5, 0.01, 0, 0, 1, 1e16], [5, 5, -5, 0.01, 0, 0, 1, 1e16]], np.float32), 0)
    assert list(tri) == [0, -1]


@pytest.mark.parametrize("mesh", ["sphere", "cube"])
def test_ray_triangle_test_is_watertight(P, mesh):
    """R-TRACE property (OptiX's triangle test is watertight; the reference has no source for it): rays fired from inside a
    CLOSED mesh exactly at its shared edges and vertices must hit something -- none may slip between two triangles -- both
    by brute force over all triangles and through the BVH; any-hit rays from outside aimed the same way are all occluded.
    (Round 1's Moeller-Trumbore test lost ~20 % of such rays on the sphere.)"""
    sc = scenes.closed_sphere(4) if mesh == "sphere" else scenes.closed_cube(12)
    c = np.asarray([3.0, -7.0, 11.0] if mesh == "sphere" else [1.5, 2.5, -3.5])
    P.add_scene(sc); P.commit()
    rays, _ = scenes.edge_and_vertex_rays(sc, 100000, 1, [c, c + [17.0, 5.0, -11.0], c + [-30.0, 20.0, 9.0]])
    t, _ = P.trace(rays, 0)
    assert (t >= 0).all(), f"{(t < 0).sum()} of {len(t)} edge/vertex rays leaked through the closed mesh (BVH)"
    tb, _ = P.trace(rays[:20000], 2)
    assert (tb >= 0).all(), f"{(tb < 0).sum()} rays leaked (brute force)"
    if mesh == "sphere":      # from outside, at edges that face the origin (on the silhouette a grazing ray may legitimately miss)
        eye = c + [400.0, 130.0, -220.0]
        outside, tgt = scenes.edge_and_vertex_rays(sc, 50000, 2, [eye])
        facing = ((tgt - c) * (eye - tgt)).sum(1) / (np.linalg.norm(tgt - c, axis=1) * np.linalg.norm(eye - tgt, axis=1)) > 0.1
        to, _ = P.trace(outside[facing], 0)
        assert facing.sum() > 10000 and (to >= 0).all(), f"{(to < 0).sum()} rays from outside leaked"
    # occlusion rays cull back faces (deviceProgram.cu:242): seen from inside, the mesh must be wound inwards to be front-facing
    P2 = type(P)("port")
    m = sc.meshes[0]
    P2.add_mesh(m.vtx, m.idx[:, ::-1].copy(), m.material); P2.commit()
    occ, _ = P2.trace(rays, 1)
    P2.close()
    assert (occ == 1).all(), f"{(occ == 0).sum()} shadow rays leaked"


def test_oracle_filter_and_features_properties(port):
    """the oracle's reconstruction filter and denoiser features (orc_opts.filter / features; port only -- the reference ships an
    empty Filter.h and commented-out feature stores): filter 0 through the same code path is the box splat bit for bit, filter 1
    leaves the fovea alone and keeps a constant image constant; normals are unit length, albedo is the material colour"""
    from fovpathtracing_optixcodelatest_b200 import scenes
    w, h, ri, ro = 160, 90, 12, 40
    sc, cam = scenes.cornell_box(), scenes.CORNELL_CAMERA
    out = {}
    for f in (0, 1):
        o = port.Oracle("port"); o.add_scene(sc); o.commit(); o.set_probe(scenes.sky_probe(32, 16, 1))
        U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, w / h)
        lp, accum, frame = o.make_params(w, h, cam.eye, U, V, W)
        nb, ab = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
        lp.frame.normal_buffer, lp.frame.albedo_buffer = nb.ctypes.data, ab.ctypes.data
        o.render(lp, port.FovCfg.sv5(inner_radius=ri, outer_radius=ro), port.Opts.sv5(filter=f, features=1, threads=1))
        out[f] = (accum.copy(), nb, ab); o.close()
    o = port.Oracle("port"); o.add_scene(sc); o.commit(); o.set_probe(scenes.sky_probe(32, 16, 1))
    lp, accum, frame = o.make_params(w, h, cam.eye, U, V, W)
    o.render(lp, port.FovCfg.sv5(inner_radius=ri, outer_radius=ro), port.Opts.sv5(threads=1)); o.close()
    assert (accum.view(np.uint32) == out[0][0].view(np.uint32)).all()
    y, x = np.mgrid[0:h, 0:w]; rr = np.hypot(x - w // 2, y - h // 2)
    assert (out[0][0][rr < ri - 2] == out[1][0][rr < ri - 2]).all()
    assert (out[0][0][rr > ro + 8] != out[1][0][rr > ro + 8]).any()
    nb, ab = out[0][1], out[0][2]
    hit = np.linalg.norm(nb[..., :3], axis=2) > 0
    fov = hit & (rr < ri - 2)                                      # fovea: 32 samples of the same wall -> mean of equal unit vectors
    ln = np.linalg.norm(nb[..., :3], axis=2)
    assert fov.any() and (ln[hit] <= 1.0 + 1e-5).all() and (ln[fov] > 0.5).all()
    colours = {tuple(np.float32(m.material.color.tolist())) for m in sc.meshes}
    centre = ab[h // 2, w // 2, :3]
    assert min(np.abs(np.array(c) - centre).max() for c in colours) < 1e-6
