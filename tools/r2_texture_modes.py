"""GPU experiment (round 2): which software definition of the bilinear weights reproduces the texture unit?
Prints, per candidate mode of fovpt_texture_check, the max / mean abs error against tex2D<float4> and the share of
bit-identical channel values.  modes: 0 float weights, 1 frac rounded to 8 bits, 2 coordinate rounded to 1/256 texel,
3 frac truncated, 4 coordinate floored."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fovpathtracing_optixcodelatest_b200 import native, scenes

rng = np.random.default_rng(11)
for shape in ((53, 37), (64, 64), (256, 1024)):
    sc = scenes.cornell_box()
    sc.textures.append(rng.integers(0, 2 ** 32, shape, dtype=np.uint64).astype(np.uint32))
    ctx = native.Context(0); ctx.upload_scene(sc)
    uv = np.concatenate([rng.uniform(-2.0, 3.0, (400000, 2)), rng.uniform(0.0, 1.0, (400000, 2))]).astype(np.float32)
    for mode in (-1, 0, 1, 2, 3, 4):
        sw, hw = ctx.texture_check(0, uv, mode)
        e = np.abs(sw[:, :3] - hw[:, :3])
        print(f"tex {shape} mode {mode:2d}: max {e.max():.3e} mean {e.mean():.3e} identical {np.mean(e == 0):.4f}  hw*255 granularity sample {hw[:3,:3].ravel()}")
    ctx.close()
