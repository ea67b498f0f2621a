"""Generates tests/golden/*.npz from oracle/_ref/libfovref.so -- i.e. from the REFERENCE's own
deviceProgram.cu / Disney.cuh / Probe.cuh / maths.h / sample.h / random.h / helpers.h / Probe.h /
Model.cpp / Camera.cpp compiled unmodified for the host (oracle/Makefile target `ref`).
Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures pin the CPU oracle (and through it the CUDA path) where the reference tree is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fovpathtracing_optixcodelatest_b200 import scenes  # noqa: E402
from fovpathtracing_optixcodelatest_b200.structs import Material  # noqa: E402
from oracle import orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def rand_material(rng):
    m = Material.default()
    m.color.set(rng.random(3)); m.metallic = rng.random(); m.subsurface = rng.random() if rng.random() < 0.5 else 0.0
    m.specular = rng.random(); m.roughness = rng.random(); m.specularTint = rng.random(); m.clearcoat = rng.random()
    m.clearcoatGloss = rng.random(); m.transmission = float(rng.choice([0.0, 0.4, 1.0, rng.random()]))
    return m


def bsdf_cases(n=600, seed=11):
    rng = np.random.default_rng(seed)
    mats = np.zeros((n, 104), np.uint8); alb = np.zeros((n, 3), np.float32); eta = np.zeros((n, 2), np.float32)
    N = np.zeros((n, 3), np.float32); V = np.zeros((n, 3), np.float32); L = np.zeros((n, 3), np.float32); seeds = np.zeros(n, np.uint32)
    for i in range(n):
        m = rand_material(rng) if i % 3 else (Material.default() if i % 2 else Material.diffuse((.7, .7, .7)))
        mats[i] = np.frombuffer(bytes(m), np.uint8)
        nn = rng.normal(size=3); nn /= np.linalg.norm(nn); vv = rng.normal(size=3); vv /= np.linalg.norm(vv)
        if vv @ nn < 0:
            vv = -vv
        ll = rng.normal(size=3); ll /= np.linalg.norm(ll)
        N[i], V[i], L[i] = nn, vv, ll
        alb[i] = rng.random(3); eta[i] = (1.0, 1.4) if i % 2 else (1.4, 1.0); seeds[i] = rng.integers(0, 2 ** 32)
    return mats, alb, eta, N, V, L, seeds


def eval_bsdf(o, mats, alb, eta, N, V, L, seeds):
    n = len(seeds)
    s_out = np.zeros((n, 10), np.float32); e_out = np.zeros((n, 4), np.float32)
    for i in range(n):
        m = Material.from_buffer_copy(mats[i].tobytes())
        s_out[i] = o.bsdf_sample(m, alb[i], float(eta[i, 0]), float(eta[i, 1]), N[i], V[i], int(seeds[i]))
        e_out[i] = o.bsdf_eval(m, alb[i], float(eta[i, 0]), float(eta[i, 1]), N[i], V[i], L[i])
    return s_out, e_out


def small_frame(o, scene, cam, w, h, probe, ri, ro, gaze=None, frames=1):
    o.add_scene(scene); o.commit(); o.set_probe(probe)
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, w / h)
    lp, accum, frame = o.make_params(w, h, cam.eye, U, V, W, gaze=gaze)
    for _ in range(frames):
        o.render(lp, orc.FovCfg.sv5(inner_radius=ri, outer_radius=ro), orc.Opts.sv5(threads=1, jitter_yx=1 if o.kind == "reference" else 0))   # threads=1: clamped border splats are last-writer-wins
    return accum.copy(), frame.copy(), o.counters()


def main():
    R = orc.Oracle("reference")
    # --- integer RNG / KATs
    tea_in = np.array([(0, 0), (1, 0), (0, 1), (5, 0), (518400 + 960, 7), (123456789, 42), (0xffffffff, 0xffffffff)], np.uint32)
    tea_out = np.array([R.tea4(int(a), int(b)) for a, b in tea_in], np.uint32)
    rnd, rnd_seed = R.rnd_seq(int(tea_out[0]), 64)
    ru, rf = R.random_seq(int(tea_out[0]), 256)
    uvw = np.concatenate(R.camera_uvw((-1293.07, 154.681, 0), (200, 0, 0), (0, 1, 0), 45.0, 1920 / 1080))
    np.savez_compressed(os.path.join(OUT, "rng.npz"), tea_in=tea_in, tea_out=tea_out, rnd=rnd, rnd_seed=np.uint32(rnd_seed), rand_u32=ru, rand_f=rf, sponza_uvw=uvw)
    # --- BSDF
    mats, alb, eta, N, V, L, seeds = bsdf_cases()
    s_out, e_out = eval_bsdf(R, mats, alb, eta, N, V, L, seeds)
    np.savez_compressed(os.path.join(OUT, "bsdf.npz"), mats=mats, alb=alb, eta=eta, N=N, V=V, L=L, seeds=seeds, sample=s_out, eval=e_out)
    # --- probe
    img = scenes.sky_probe(64, 32, seed=4)
    R.set_probe(img)
    pdfX, cdfX, pdfY, cdfY = R.get_probe()
    pseeds = (np.arange(400, dtype=np.uint64) * 7919 + 1).astype(np.uint32)
    ps = np.stack([R.probe_sample(int(s)) for s in pseeds])
    rng = np.random.default_rng(3); dirs = rng.normal(size=(200, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True); dirs = dirs.astype(np.float32)
    pe = np.stack([R.probe_eval(d) for d in dirs])
    tm_in = (rng.random((300, 3)) ** 3 * 4).astype(np.float32)
    tm_out = np.array([R.tonemap(a) for a in tm_in], np.uint32)
    np.savez_compressed(os.path.join(OUT, "probe.npz"), pdfX=pdfX, cdfX=cdfX, pdfY=pdfY, cdfY=cdfY, seeds=pseeds, sample=ps, dirs=dirs, eval=pe, tm_in=tm_in, tm_out=tm_out)
    R.close()
    # --- small foveated frames through the unmodified raygen / closest-hit programs
    sc = scenes.cornell_box()
    sc.textures.append(scenes.checker_texture(32)); sc.meshes[0].texcoord = np.array([[0, 0], [0, 3], [3, 3], [3, 0]], np.float32); sc.meshes[0].texture_id = 0
    R = orc.Oracle("reference")
    a1, f1, c1 = small_frame(R, sc, scenes.CORNELL_CAMERA, 160, 90, scenes.sky_probe(64, 32, seed=4), 12, 40, frames=2)
    R.close(); R = orc.Oracle("reference")
    a2, f2, c2 = small_frame(R, sc, scenes.CORNELL_CAMERA, 160, 90, scenes.constant_probe(16, 8, 2.5), 12, 40, gaze=(20, 70))   # offset wraps (gaze < R)
    R.close()
    R = orc.Oracle("reference")
    a3, f3, c3 = small_frame(R, scenes.cornell_with_shadow_catcher(), scenes.CORNELL_CAMERA, 160, 90, scenes.sky_probe(64, 32, seed=4), 12, 40)
    R.close()
    np.savez_compressed(os.path.join(OUT, "frames.npz"), accum_sky=a1, frame_sky=f1, counts_sky=np.array([c1[k] for k in sorted(c1)], np.uint64),
                        accum_wrap=a2, frame_wrap=f2, counts_wrap=np.array([c2[k] for k in sorted(c2)], np.uint64),
                        accum_catcher=a3, frame_catcher=f3, counts_catcher=np.array([c3[k] for k in sorted(c3)], np.uint64))
    # --- loader: the reference's loadOBJ on the committed Cornell OBJ
    obj = os.path.join(ROOT, "assets", "cornell.obj")
    scenes.write_obj(scenes.cornell_box(), obj)
    R = orc.Oracle("reference"); R.load_obj(obj)
    ms = R.meshes()
    tris = [m["vtx"][m["idx"].astype(np.int64)].reshape(-1, 9) for m in ms]      # resolved triangles (vertex order is unspecified, A.9)
    np.savez_compressed(os.path.join(OUT, "loader.npz"), n_meshes=len(ms), counts=np.array([len(t) for t in tris]), tris=np.concatenate(tris),
                        colors=np.array([m["material"].color.tolist() for m in ms], np.float32),
                        emission=np.array([m["material"].emission.tolist() for m in ms], np.float32),
                        materials=np.stack([np.frombuffer(bytes(m["material"]), np.uint8) for m in ms]),
                        texture_ids=np.array([m["texture_id"] for m in ms]))
    R.close()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
