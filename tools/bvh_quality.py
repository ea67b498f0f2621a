"""BVH quality probe (run under gpurun): mean inner-node / leaf / triangle visits per ray of the device
BVH for a workload's primary-like and diffuse-bounce-like rays, next to the oracle's canonical SAH BVH."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fovpathtracing_optixcodelatest_b200 import native, workloads, scenes
from oracle import orc

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
wl = workloads.BY_NAME[name]()
sc = wl.scene()
ctx = native.Context(0); t = time.time(); ctx.upload_scene(sc); print("upload+build s", round(time.time() - t, 3), ctx.stats()["bvh_build_ms"], "ms build;", ctx.stats()["bvh_nodes"], "nodes")
rng = np.random.default_rng(1)
n = 400000
cam = wl.camera
U, V, W = scenes.camera_uvw(cam, wl.width / wl.height)
px = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
d = px[:, :1] * U[None] + px[:, 1:] * V[None] + W[None]; d /= np.linalg.norm(d, axis=1, keepdims=True)
prim = np.concatenate([np.tile(np.asarray(cam.eye, np.float32), (n, 1)), np.full((n, 1), 0.01, np.float32), d.astype(np.float32), np.full((n, 1), 1e16, np.float32)], 1).astype(np.float32)
tri, tuv = ctx.trace(prim, 0)
hit = tri >= 0
P = prim[hit, :3] + tuv[hit, :1] * prim[hit, 4:7]
dd = rng.normal(size=(hit.sum(), 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
sec = np.concatenate([P, np.full((len(P), 1), 0.01), dd, np.full((len(P), 1), 1e16)], 1).astype(np.float32)
o = orc.Oracle("port"); o.add_scene(sc); o.commit()
for label, rays in (("primary", prim), ("diffuse", sec)):
    for mode, mname in ((2, "closest"), (3, "any")):
        _, c = ctx.trace(rays, mode)
        o.reset_counters(); o.trace(rays[:50000], mode - 2); k = o.counters()
        nn = (k["closest_nodes"] if mode == 2 else k["any_nodes"]) / 50000; nt = (k["closest_tris"] if mode == 2 else k["any_tris"]) / 50000
        print(f"{label:8s} {mname:8s} gpu nodes {c[:,0].mean():7.2f} leaves {c[:,1].mean():6.2f} tris {c[:,2].mean():6.2f} | oracle SAH nodes {nn:7.2f} tris {nt:6.2f}")
