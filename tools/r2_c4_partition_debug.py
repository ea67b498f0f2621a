"""debug: C4 view 0, 8-rank interleaved partition emulated in one process (ranks render one after the other into the same
buffers) vs the unpartitioned render; prints the differing pixels"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fovpathtracing_optixcodelatest_b200 import native, workloads, partition
from fovpathtracing_optixcodelatest_b200.renderer import ProbeData, SampleRenderer
wl = workloads.c4(); w, h = wl.width, wl.height
cam, gaze = wl.view_at(0)
def mk():
    r = SampleRenderer(wl.scene(), 0, wl.apply(native.default_config()))
    r.resize((w, h)); r.setCamera(cam); r.setProbe(ProbeData(wl.probe()).BuildCDF())
    r.launchParams.frame.c_x, r.launchParams.frame.c_y = gaze
    return r
r = mk(); r.render(); A = r.downloadAccum().copy(); F = r.downloadPixels().copy(); r.close()
world, tile = 8, 16
owner = partition.interleaved(w, h, world, tile)
r = mk()
for rk in range(world):
    r.ctx.set_tile_owner(tile, owner, rk)
    r.launchParams.frame.subframe_index = 0
    r.render()
B = r.downloadAccum().copy(); G = r.downloadPixels().copy(); r.close()
d = (A.view(np.uint32) != B.view(np.uint32)).any(axis=2)
ys, xs = np.nonzero(d)
print("differing", d.sum(), "frame differing", (F != G).sum())
for y, x in list(zip(ys, xs))[:40]:
    print((x, y), "full", A[y, x], "part", B[y, x], "dist to gaze %.1f" % np.hypot(x - gaze[0], y - gaze[1]))
print("x range", xs.min() if len(xs) else None, xs.max() if len(xs) else None, "y range", ys.min() if len(ys) else None, ys.max() if len(ys) else None)
