"""GPU experiment (round 2): recover the texture unit's filter weight as a function of the coordinate.
A W x 2 texture whose texels alternate 0 / 255 along x (constant along y), sampled at v = 0.25 (row centre, beta = 0):
hw = alpha or 1 - alpha.  Dumps (u, hw) of fine sweeps to gpurun_out/r2_texsweep.npz for offline analysis."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fovpathtracing_optixcodelatest_b200 import native, scenes

out = {}
for W in (4, 37, 1024):
    sc = scenes.cornell_box()
    row = ((np.arange(W) % 2) * 255).astype(np.uint32)
    tex = np.stack([row * 0x010101 | 0xff000000] * 2).astype(np.uint32)
    sc.textures.append(tex)
    ctx = native.Context(0); ctx.upload_scene(sc)
    if W == 4:
        u = (np.arange(0, 1 << 18, dtype=np.float64) / (1 << 18) * 1.5 - 0.25).astype(np.float32)
    elif W == 37:
        u = (0.3 + np.arange(0, 1 << 18, dtype=np.float64) / (1 << 18) * (3.0 / 37)).astype(np.float32)
    else:
        u = (0.7 + np.arange(0, 1 << 18, dtype=np.float64) / (1 << 18) * (3.0 / 1024)).astype(np.float32)
    uv = np.stack([u, np.full_like(u, 0.25)], 1)
    sw, hw = ctx.texture_check(0, uv, 0)
    out[f"u{W}"] = u; out[f"hw{W}"] = hw[:, 0].copy(); out[f"sw{W}"] = sw[:, 0].copy()
    ctx.close()
np.savez_compressed(os.path.join("gpurun_out", "r2_texsweep.npz"), **out)
print("ok")
