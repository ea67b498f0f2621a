"""Deterministic synthetic scenes, probes and cameras for the BASELINE.json configurations.

The reference ships no assets (no .obj/.mtl/.hdr anywhere in the tree; ``PT_sv5_/main.cpp:196-218``
points at absolute Windows paths), so the workloads are authored here:

* ``cornell_box()``      -- C1/C2: a 555-unit Cornell box (open front), two blocks, an emissive quad.
                            Also written as OBJ+MTL (``write_obj``) in the form ``loadOBJ``
                            (``PT_sv5_/Model.cpp:138-217``) accepts: one ``usemtl`` per ``o`` group and a
                            material on every face.
* ``atrium()``           -- C3/C4: a procedural "Sponza-class" colonnaded atrium, ~262k triangles, sized
                            like Crytek Sponza so the reference's camera preset
                            (``main.cpp:240-254``) sits inside it.
* ``displaced_grid()``   -- C5: procedural height-field mesh of any triangle count.
* ``sky_probe()`` / ``constant_probe()`` -- lat-long HDR probes (``main.cpp:160-187`` loadProbe/loadColor).

Everything is numpy with fixed seeds: the same arrays are produced on every machine.
"""
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .structs import Material


@dataclass
class Mesh:
    name: str
    vtx: np.ndarray                  # (n,3) float32
    idx: np.ndarray                  # (m,3) uint32
    material: Material
    texcoord: Optional[np.ndarray] = None   # (n,2) float32
    texture_id: int = -1


@dataclass
class Scene:
    meshes: List[Mesh] = field(default_factory=list)
    textures: List[np.ndarray] = field(default_factory=list)   # (h,w) uint32 RGBA8, row 0 = v 0
    name: str = ""

    @property
    def n_tris(self):
        return int(sum(len(m.idx) for m in self.meshes))


@dataclass
class CameraSpec:
    eye: tuple
    lookat: tuple
    up: tuple = (0.0, 1.0, 0.0)
    fovy: float = 45.0


def camera_uvw(cam: CameraSpec, aspect: float):
    """float32 re-statement of sutil::Camera::UVWFrame (reference sutil/Camera.cpp:34-46).

    Used by the product host layer; the checkers use their own (orc_camera_uvw)."""
    f = np.float32
    eye = np.asarray(cam.eye, f); lookat = np.asarray(cam.lookat, f); up = np.asarray(cam.up, f)

    def cross(a, b):
        return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], f)

    def length(a):
        return f(np.sqrt(f(f(a[0] * a[0]) + f(a[1] * a[1])) + f(a[2] * a[2])))

    def normalize(a):
        return (a * f(f(1.0) / length(a))).astype(f)

    W = (lookat - eye).astype(f)
    wlen = length(W)
    U = normalize(cross(W, up))
    V = normalize(cross(U, W))
    vlen = f(wlen * f(math.tan(float(f(f(0.5) * f(cam.fovy)) * f(math.pi) / f(180.0)))))
    V = (V * vlen).astype(f)
    U = (U * f(vlen * f(aspect))).astype(f)
    return U, V, W


# ----------------------------------------------------------------------------- primitives

def _quad(a, b, c, d):
    """two triangles a-b-c, a-c-d (counter-clockwise seen from the side the normal points to)."""
    v = np.array([a, b, c, d], np.float32)
    i = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    return v, i


def _merge(parts):
    vs, is_, base = [], [], 0
    for v, i in parts:
        vs.append(np.asarray(v, np.float32).reshape(-1, 3)); is_.append(np.asarray(i, np.uint32).reshape(-1, 3) + base)
        base += len(vs[-1])
    return np.concatenate(vs), np.concatenate(is_)


def _block(p0, p1, p2, p3, height):
    """a box standing on the floor: footprint p0..p3 (ccw seen from above), given height."""
    b = [np.array([p[0], 0.0, p[1]], np.float32) for p in (p0, p1, p2, p3)]
    t = [x + np.array([0, height, 0], np.float32) for x in b]
    parts = [_quad(t[0], t[1], t[2], t[3])]   # top, normal +y when footprint is cw in (x,z) seen from +y
    for k in range(4):
        n = (k + 1) % 4
        parts.append(_quad(b[k], b[n], t[n], t[k]))
    return _merge(parts)


def cornell_box(disney_half=True) -> Scene:
    """Cornell box in the classic 555-unit frame.  9 meshes, 36 triangles.

    ``disney_half``: odd meshes keep the reference's default Disney ``Material()`` parameters
    (only Kd/Ke overwritten, exactly what ``loadOBJ`` does, Model.cpp:190-191); even meshes are the
    diffuse parameterisation (C2: "Lambertian + Disney materials").
    """
    white, red, green = (0.73, 0.73, 0.73), (0.65, 0.05, 0.05), (0.12, 0.45, 0.15)
    L = 555.0
    parts = [
        ("floor", _quad((0, 0, 0), (0, 0, L), (L, 0, L), (L, 0, 0)), white, (0, 0, 0)),
        ("ceiling", _quad((0, L, 0), (L, L, 0), (L, L, L), (0, L, L)), white, (0, 0, 0)),
        ("back", _quad((0, 0, L), (0, L, L), (L, L, L), (L, 0, L)), white, (0, 0, 0)),
        ("left", _quad((L, 0, 0), (L, 0, L), (L, L, L), (L, L, 0)), red, (0, 0, 0)),
        ("right", _quad((0, 0, 0), (0, L, 0), (0, L, L), (0, 0, L)), green, (0, 0, 0)),
        ("light", _quad((213, L - 1.0, 227), (343, L - 1.0, 227), (343, L - 1.0, 332), (213, L - 1.0, 332)),
         (0.78, 0.78, 0.78), (17.0, 12.0, 4.0)),
        ("short_block", _block((130, 65), (82, 225), (240, 272), (290, 114), 165.0), white, (0, 0, 0)),
        ("tall_block", _block((423, 247), (265, 296), (314, 456), (472, 406), 330.0), white, (0, 0, 0)),
        ("front_lip", _quad((0, 0, 0), (L, 0, 0), (L, -20.0, 0), (0, -20.0, 0)), white, (0, 0, 0)),
    ]
    sc = Scene(name="cornell")
    for k, (name, (v, i), kd, ke) in enumerate(parts):
        if disney_half and (k % 2 == 1):
            m = Material.default(); m.color.set(kd); m.emission.set(ke)
        else:
            m = Material.diffuse(kd, ke)
        sc.meshes.append(Mesh(name, np.asarray(v, np.float32), np.asarray(i, np.uint32), m))
    return sc


CORNELL_CAMERA = CameraSpec(eye=(278.0, 273.0, -800.0), lookat=(278.0, 273.0, 0.0), fovy=40.0)
# reference preset for Crytek Sponza, PT_sv5_/main.cpp:240-254
SPONZA_CAMERA = CameraSpec(eye=(-1293.07, 154.681, 0.0), lookat=(200.0, 0.0, 0.0), fovy=45.0)


def write_obj(scene: Scene, path: str):
    """OBJ + MTL in the form the reference loader handles: one ``o`` group / one ``usemtl`` each,
    Kd -> Material.color, Ke -> Material.emission (Model.cpp:190-191)."""
    base = os.path.splitext(os.path.basename(path))[0]
    mtl_path = os.path.join(os.path.dirname(path), base + ".mtl")
    with open(mtl_path, "w") as f:
        for m in scene.meshes:
            f.write(f"newmtl {m.name}\nKd {m.material.color.x:.9g} {m.material.color.y:.9g} {m.material.color.z:.9g}\n"
                    f"Ke {m.material.emission.x:.9g} {m.material.emission.y:.9g} {m.material.emission.z:.9g}\n\n")
    with open(path, "w") as f:
        f.write(f"mtllib {base}.mtl\n")
        base_v = 1
        for m in scene.meshes:
            f.write(f"o {m.name}\n")
            for v in m.vtx:
                f.write(f"v {v[0]:.9g} {v[1]:.9g} {v[2]:.9g}\n")
            f.write(f"usemtl {m.name}\n")
            for t in m.idx:
                f.write(f"f {t[0] + base_v} {t[1] + base_v} {t[2] + base_v}\n")
            base_v += len(m.vtx)


# ----------------------------------------------------------------------------- probes

def constant_probe(w, h, value=2.5):
    """loadColor (main.cpp:175-187): a frame-sized constant ambient probe; sv5 uses 5.0*0.5."""
    a = np.empty((h, w, 4), np.float32)
    a[..., :3] = np.float32(value)
    a[..., 3] = 1.0
    return a


def sky_probe(w=2048, h=1024, seed=2, sun_dir=(0.35, 0.8, 0.25), sun_radiance=600.0, sun_cos=0.9985):
    """Procedural lat-long HDR: gradient sky + ground + a small sun disc + faint seeded cloud noise.
    Row r is v=(r+0.5)/h with v=0 at +y (ProbeUVToDir, Probe.cuh:48-58)."""
    rng = np.random.default_rng(seed)
    v = (np.arange(h, dtype=np.float32) + 0.5) / h
    u = (np.arange(w, dtype=np.float32) + 0.5) / w
    theta = v * np.float32(math.pi)
    phi = u * np.float32(2 * math.pi)
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    x = -st * np.cos(phi)[None, :]
    y = np.broadcast_to(ct, (h, w))
    z = -st * np.sin(phi)[None, :]
    up = np.clip(y, 0.0, 1.0)
    sky = np.stack([0.35 + 0.25 * (1 - up), 0.55 + 0.25 * (1 - up), 1.0 + 0.0 * up], -1) * (1.2 + 1.5 * (1 - up)[..., None] ** 3)
    ground = np.array([0.22, 0.2, 0.17], np.float32) * 0.6
    rgb = np.where((y >= 0)[..., None], sky, ground[None, None, :]).astype(np.float32)
    # low-frequency clouds
    coarse = rng.random((h // 32 + 2, w // 32 + 2)).astype(np.float32)
    yy = np.linspace(0, coarse.shape[0] - 1.001, h, dtype=np.float32); xx = np.linspace(0, coarse.shape[1] - 1.001, w, dtype=np.float32)
    y0 = yy.astype(int); x0 = xx.astype(int); fy = (yy - y0)[:, None]; fx = (xx - x0)[None, :]
    n = (coarse[y0][:, x0] * (1 - fy) * (1 - fx) + coarse[y0 + 1][:, x0] * fy * (1 - fx)
         + coarse[y0][:, x0 + 1] * (1 - fy) * fx + coarse[y0 + 1][:, x0 + 1] * fy * fx)
    rgb *= (0.85 + 0.3 * n)[..., None]
    s = np.asarray(sun_dir, np.float32); s = s / np.linalg.norm(s)
    cosang = x * s[0] + y * s[1] + z * s[2]
    sun = (cosang > sun_cos)
    rgb[sun] += np.array([1.0, 0.93, 0.8], np.float32) * np.float32(sun_radiance)
    out = np.empty((h, w, 4), np.float32)
    out[..., :3] = rgb
    out[..., 3] = 1.0
    return out


# ----------------------------------------------------------------------------- textures

def checker_texture(size=256, a=(200, 190, 170), b=(120, 95, 80), cells=8, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    sel = ((yy * cells // size) + (xx * cells // size)) % 2
    noise = rng.integers(-12, 13, (size, size))
    out = np.zeros((size, size), np.uint32)
    for c in range(3):
        ch = np.clip(np.where(sel, a[c], b[c]) + noise, 0, 255).astype(np.uint32)
        out |= ch << (8 * c)
    out |= np.uint32(255) << 24
    return out
