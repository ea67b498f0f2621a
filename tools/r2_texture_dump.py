"""GPU experiment (round 2): dump texture-unit results for a random 64x64 RGBA8 texture at random coordinates
(gpurun_out/r2_texdump.npz) so the blend arithmetic can be fitted offline."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fovpathtracing_optixcodelatest_b200 import native, scenes
rng = np.random.default_rng(5)
out = {}
for name, shape in (("p2", (64, 64)), ("odd", (53, 37))):
    sc = scenes.cornell_box()
    tex = rng.integers(0, 2 ** 32, shape, dtype=np.uint64).astype(np.uint32)
    sc.textures.append(tex)
    ctx = native.Context(0); ctx.upload_scene(sc)
    uv = rng.uniform(-1.0, 2.0, (300000, 2)).astype(np.float32)
    sw, hw = ctx.texture_check(0, uv, 0)
    out["tex_" + name] = tex; out["uv_" + name] = uv; out["hw_" + name] = hw
    ctx.close()
np.savez_compressed(os.path.join("gpurun_out", "r2_texdump.npz"), **out)
print("ok")
