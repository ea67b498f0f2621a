"""The BASELINE.json configurations (C1..C5) as concrete, seeded inputs.

A workload is pure data: scene arrays, camera, probe image, frame size, gaze, and the
``fovpt_config`` values.  bench.py and the tests build the product renderer and the CPU checkers
from the same object, so both sides see identical inputs.
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import scenes


@dataclass
class Workload:
    name: str
    description: str
    scene_fn: object
    camera: scenes.CameraSpec
    probe_fn: object
    width: int
    height: int
    gaze: Optional[Tuple[int, int]] = None        # None = frame centre
    foveated: int = 1
    inner_radius: int = 74
    outer_radius: int = 241
    spp: Tuple[int, int, int] = (8, 16, 32)       # periphery, mid, fovea  (SimplePathtracer.cpp:142,170,193)
    spp_uniform: int = 4
    max_depth: int = 4
    frames: int = 1                               # accumulated subframes that make up the quoted spp
    accumulate: int = 0
    seed_all_passes_with_subframe: int = 0
    tmin: float = 0.01
    view_at: object = None                        # optional: step index -> (camera, gaze) for moving-gaze / stereo workloads
    _scene: object = field(default=None, repr=False)
    _probe: object = field(default=None, repr=False)

    def scene(self):
        if self._scene is None:
            self._scene = self.scene_fn()
        return self._scene

    def probe(self):
        if self._probe is None:
            self._probe = self.probe_fn()
        return self._probe

    def gaze_xy(self):
        return self.gaze if self.gaze is not None else (self.width // 2, self.height // 2)

    def apply(self, cfg):
        """fill a native.Config / oracle FovCfg-like object with this workload's values"""
        for k, v in (("foveated", self.foveated), ("inner_radius", self.inner_radius), ("outer_radius", self.outer_radius),
                     ("spp_periphery", self.spp[0]), ("spp_mid", self.spp[1]), ("spp_fovea", self.spp[2]),
                     ("spp_uniform", self.spp_uniform), ("max_depth", self.max_depth), ("accumulate", self.accumulate),
                     ("seed_all_passes_with_subframe", self.seed_all_passes_with_subframe), ("tmin", self.tmin)):
            if hasattr(cfg, k):
                setattr(cfg, k, v)
        return cfg


def c1():
    return Workload("C1", "Cornell box OBJ, 512x512, 1 spp, max depth 1, no foveation",
                    scenes.cornell_box, scenes.CORNELL_CAMERA, lambda: scenes.constant_probe(64, 32, 2.5),
                    512, 512, foveated=0, spp_uniform=1, max_depth=1)


def c2():
    # 64 foveal spp = 2 accumulated frames of (8,16,32); radii are the reference's own commented 2x values
    return Workload("C2", "Cornell box 1920x1080 foveated (gaze at centre), 64 spp, depth 8, Lambertian + Disney",
                    scenes.cornell_box, scenes.CORNELL_CAMERA, lambda: scenes.constant_probe(64, 32, 2.5),
                    1920, 1080, inner_radius=148, outer_radius=482, max_depth=8, frames=2, accumulate=1,
                    seed_all_passes_with_subframe=1)


def c3():
    # 256 foveal spp = 8 accumulated frames of (8,16,32); reference radii 74/241 (SimplePathtracer.cpp:20-21)
    return Workload("C3", "Sponza-class atrium (~262k tris) 1920x1080 foveated, 256 spp, depth 8, probe env light",
                    scenes.atrium, scenes.SPONZA_CAMERA, lambda: scenes.sky_probe(2048, 1024, 2),
                    1920, 1080, max_depth=8, frames=8, accumulate=1, seed_all_passes_with_subframe=1)


def c4_eye(eye_index=0, frame_k=0, ipd=6.5):
    # stereo pair: cameras +-IPD/2 along U; gaze path mirrors the counter*0.05 motion of main.cpp:387
    import math
    cam = scenes.SPONZA_CAMERA
    U, _, _ = scenes.camera_uvw(cam, 1.0)
    u = U / np.linalg.norm(U)
    s = (-0.5 if eye_index == 0 else 0.5) * ipd
    eye = tuple(float(cam.eye[i] + s * u[i]) for i in range(3))
    look = tuple(float(cam.lookat[i] + s * u[i]) for i in range(3))
    gaze = (int(1080 + 600 * math.cos(0.05 * frame_k)), int(1080 + 600 * math.sin(0.05 * frame_k)))
    return Workload(f"C4.{eye_index}", "Stereo VR eye 2160x2160 foveated, moving gaze",
                    scenes.atrium, scenes.CameraSpec(eye, look, cam.up, cam.fovy), lambda: scenes.sky_probe(2048, 1024, 2),
                    2160, 2160, gaze=gaze, inner_radius=148, outer_radius=482, max_depth=4)


def c5(n_tris=10_000_000):
    return Workload("C5", "Synthetic 10M-triangle procedural mesh, 3840x2160, uniform 32 spp x 32 frames",
                    lambda: scenes.displaced_grid(n_tris), scenes.CameraSpec((0.0, 900.0, 2600.0), (0.0, 0.0, 0.0), (0, 1, 0), 45.0),
                    lambda: scenes.sky_probe(2048, 1024, 2), 3840, 2160, foveated=0, spp_uniform=32, max_depth=4,
                    frames=32, accumulate=1, seed_all_passes_with_subframe=1)


def c4():
    """BASELINE configs[3] as a bench workload: one step = one eye frame of the stereo pair; step k renders eye k % 2
    of stereo frame k // 2, whose gaze point follows the circular path of c4_eye (per-frame gaze update)."""
    wl = c4_eye(0, 0)
    wl.name = "C4"
    wl.description = "Stereo VR pair 2x2160x2160 foveated, moving gaze (one step = one eye frame, gaze updated every stereo frame)"
    wl.view_at = lambda k: (c4_eye(k % 2, k // 2).camera, c4_eye(k % 2, k // 2).gaze)
    return wl


BY_NAME = {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}


def scaled(wl: Workload, width: int, height: int, radius_scale: float = None):
    """the same workload on a smaller frame (parity tests at sizes the CPU oracle finishes in seconds)"""
    import copy
    w2 = copy.copy(wl)
    s = radius_scale if radius_scale is not None else width / wl.width
    w2.width, w2.height = width, height
    w2.inner_radius = max(2, int(round(wl.inner_radius * s))); w2.outer_radius = max(w2.inner_radius + 2, int(round(wl.outer_radius * s)))
    if wl.gaze is not None:
        w2.gaze = (int(wl.gaze[0] * width / wl.width), int(wl.gaze[1] * height / wl.height))
    return w2
