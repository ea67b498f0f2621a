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

    def add(self, vtx, idx, material, texcoord=None, texture_id=-1, name=""):
        self.meshes.append(Mesh(name or f"mesh{len(self.meshes)}", np.ascontiguousarray(vtx, np.float32).reshape(-1, 3),
                                np.ascontiguousarray(idx, np.uint32).reshape(-1, 3), material, texcoord, texture_id))
        return len(self.meshes) - 1


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
    """Cornell box in the classic 555-unit frame.  9 meshes, 34 triangles.

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


def cornell_with_shadow_catcher() -> Scene:
    """Cornell box whose floor is a shadow catcher (reference MATERIAL_FLAG_SHADOW_CATCHER, Material.h:9:
    deviceProgram.cu:646-651 pass-through for secondary rays, :691-694 + SampleShadow :347-387 alpha accumulation)"""
    sc = cornell_box()
    m = sc.meshes[0].material.copy()
    m.flags = 1
    sc.meshes[0].material = m
    return sc


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


# ----------------------------------------------------------------------------- procedural meshes

def _grid(origin, du, dv, nu, nv, disp=None, uv_scale=(1.0, 1.0)):
    """(nu x nv) quads spanning origin + s*du + t*dv, s,t in [0,1]; optional displacement along the
    normal disp(s,t).  Triangles are counter-clockwise seen from the side cross(du,dv) points to."""
    o = np.asarray(origin, np.float64); du = np.asarray(du, np.float64); dv = np.asarray(dv, np.float64)
    s = np.linspace(0.0, 1.0, nu + 1); t = np.linspace(0.0, 1.0, nv + 1)
    S, T = np.meshgrid(s, t, indexing="xy")
    P = o[None, None, :] + S[..., None] * du + T[..., None] * dv
    if disp is not None:
        n = np.cross(du, dv); n = n / np.linalg.norm(n)
        P = P + disp(S, T)[..., None] * n
    vtx = P.reshape(-1, 3).astype(np.float32)
    tc = np.stack([S * uv_scale[0], T * uv_scale[1]], -1).reshape(-1, 2).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
    a = (jj * (nu + 1) + ii).reshape(-1); b = a + 1; c = a + (nu + 1) + 1; d = a + (nu + 1)
    idx = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.uint32)
    return vtx, idx, tc


def _lathe(profile_r, profile_y, center, nseg, uv_scale=(4.0, 4.0)):
    """surface of revolution around +y through `center`; profile from bottom to top; outward normals."""
    r = np.asarray(profile_r, np.float64); y = np.asarray(profile_y, np.float64)
    ang = np.linspace(0.0, 2 * math.pi, nseg + 1)
    A, R = np.meshgrid(ang, r, indexing="xy")
    Y = np.broadcast_to(y[:, None], A.shape)
    P = np.stack([center[0] + R * np.cos(A), center[1] + Y, center[2] - R * np.sin(A)], -1)
    vtx = P.reshape(-1, 3).astype(np.float32)
    nr = len(r) - 1
    S = A / (2 * math.pi); T = np.broadcast_to(np.linspace(0, 1, len(r))[:, None], A.shape)
    tc = np.stack([S * uv_scale[0], T * uv_scale[1]], -1).reshape(-1, 2).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(nseg), np.arange(nr), indexing="xy")
    a = (jj * (nseg + 1) + ii).reshape(-1); b = a + 1; c = a + (nseg + 1) + 1; d = a + (nseg + 1)
    idx = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.uint32)
    return vtx, idx, tc


def _icosphere(level):
    t = (1.0 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = np.array(v, np.float64); v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array(f, np.int64)
    for _ in range(level):
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        key = np.sort(edges, axis=1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]; mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v); v = np.concatenate([v, mid])
        n = len(f); m01 = base + inv[:n]; m12 = base + inv[n:2 * n]; m20 = base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)])
    return v, f


def closed_sphere(level=5, radius=100.0, center=(3.0, -7.0, 11.0)) -> Scene:
    """a closed, consistently wound icosphere (20 * 4^level triangles, every edge shared by exactly two of them)"""
    v, f = _icosphere(level)
    sc = Scene()
    sc.add((v * radius + np.asarray(center)).astype(np.float32), f.astype(np.uint32), _disney((.7, .7, .7)))
    return sc


def closed_cube(n=24, half=50.0, center=(1.5, 2.5, -3.5)) -> Scene:
    """a closed axis-aligned cube whose faces are n x n grids with shared vertices: axis-parallel edges and rays are where
    edge functions become exactly zero"""
    g = np.linspace(-half, half, n + 1)
    pts = {}
    verts, tris = [], []

    def vid(p):
        k = tuple(np.round(p, 6))
        if k not in pts:
            pts[k] = len(verts); verts.append(p)
        return pts[k]
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            a1, a2 = (axis + 1) % 3, (axis + 2) % 3
            for i in range(n):
                for j in range(n):
                    q = []
                    for (u, w) in ((g[i], g[j]), (g[i + 1], g[j]), (g[i + 1], g[j + 1]), (g[i], g[j + 1])):
                        p = np.zeros(3); p[axis] = sgn * half; p[a1] = u; p[a2] = w
                        q.append(vid(p))
                    if sgn > 0:
                        tris += [(q[0], q[1], q[2]), (q[0], q[2], q[3])]
                    else:
                        tris += [(q[0], q[2], q[1]), (q[0], q[3], q[2])]
    sc = Scene()
    sc.add((np.array(verts) + np.asarray(center)).astype(np.float32), np.array(tris, np.uint32), _disney((.7, .7, .7)))
    return sc


def edge_and_vertex_rays(scene: Scene, n, seed, origins):
    """rays from the given origins aimed EXACTLY (to float32 rounding) at points on the shared edges (90 %) and at the
    vertices (10 %) of the scene's first mesh; returns (rays[n*len(origins), 8], per-ray target)"""
    rng = np.random.default_rng(seed)
    m = scene.meshes[0]
    v = m.vtx.astype(np.float64); f = m.idx.astype(np.int64)
    tri = f[rng.integers(0, len(f), n)]
    e = rng.integers(0, 3, n)
    a = v[tri[np.arange(n), e]]; b = v[tri[np.arange(n), (e + 1) % 3]]
    s = rng.random(n)[:, None]
    s[: n // 10] = 0.0
    tgt = a + (b - a) * s
    out = []
    for o in origins:
        o = np.asarray(o, np.float64)
        d = tgt - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = np.zeros((n, 8), np.float32); r[:, 0:3] = o; r[:, 3] = 0.01; r[:, 4:7] = d; r[:, 7] = 1e16
        out.append(r)
    return np.concatenate(out), np.concatenate([tgt] * len(origins))


def _value_noise3(p, seed, freq):
    """cheap deterministic lattice noise in [-1,1] evaluated at points p (n,3)."""
    q = p * freq
    i = np.floor(q).astype(np.int64); fr = q - i
    fr = fr * fr * (3 - 2 * fr)

    def h(ix, iy, iz):
        n = (ix * 73856093) ^ (iy * 19349663) ^ (iz * 83492791) ^ (seed * 2654435761)
        n = (n ^ (n >> 13)) * 1274126177
        return ((n ^ (n >> 16)) & 0xffff) / 32767.5 - 1.0

    out = 0.0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (fr[:, 0] if dx else 1 - fr[:, 0]) * (fr[:, 1] if dy else 1 - fr[:, 1]) * (fr[:, 2] if dz else 1 - fr[:, 2])
                out = out + w * h(i[:, 0] + dx, i[:, 1] + dy, i[:, 2] + dz)
    return out


def _disney(color, **kw):
    m = Material.default()        # what loadOBJ leaves: reference defaults + Kd, Ke = 0
    m.color.set(color); m.emission.set((0.0, 0.0, 0.0))
    for k, v in kw.items():
        setattr(m, k, v)
    return m


def atrium(seed=1, detail=1.0) -> Scene:
    """Procedural "Sponza-class" atrium (C3/C4): a two-storey colonnaded hall with an open roof,
    arches, galleries, hanging drapes, vases and a displaced sculpture; ~262k triangles at detail=1.
    Extents follow Crytek Sponza (x in [-1900,1800], y in [0,1500], z in [-1100,1100]) so that the
    reference camera preset (main.cpp:240-254) stands inside the hall looking down its long axis."""
    rng = np.random.default_rng(seed)
    sc = Scene(name="atrium")
    sc.textures.append(checker_texture(256, (205, 190, 165), (150, 128, 105), 16, seed))      # 0 stone tiles
    sc.textures.append(checker_texture(256, (190, 170, 150), (170, 150, 128), 32, seed + 1))  # 1 brick
    x0, x1, z0, z1, H = -1900.0, 1800.0, -1100.0, 1100.0, 1500.0
    D = lambda n: max(2, int(round(n * detail)))

    def add(name, v, i, mat, tc=None, tex=-1):
        sc.meshes.append(Mesh(name, np.asarray(v, np.float32), np.asarray(i, np.uint32), mat,
                              None if tc is None else np.asarray(tc, np.float32), tex))

    stone = _disney((0.78, 0.74, 0.66))
    brick = _disney((0.72, 0.66, 0.58), roughness=0.9)
    # floor: tiles with shallow grout relief, normal +y
    v, i, tc = _grid((x0, 0, z1), (x1 - x0, 0, 0), (0, 0, z0 - z1), D(180), D(108),
                     lambda S, T: 2.0 * np.cos(S * 2 * math.pi * 42) * np.cos(T * 2 * math.pi * 25), (21.0, 12.5))
    add("floor", v, i, stone, tc, 0)
    # four walls, normals pointing inward, brick relief
    relief = lambda S, T: 3.0 * np.sin(S * 2 * math.pi * 60) * np.sin(T * 2 * math.pi * 30)
    v, i, tc = _grid((x0, 0, z0), (x1 - x0, 0, 0), (0, H, 0), D(112), D(44), relief, (24.0, 10.0)); add("wall_north", v, i, brick, tc, 1)
    v, i, tc = _grid((x1, 0, z1), (x0 - x1, 0, 0), (0, H, 0), D(112), D(44), relief, (24.0, 10.0)); add("wall_south", v, i, brick, tc, 1)
    v, i, tc = _grid((x0, 0, z1), (0, 0, z0 - z1), (0, H, 0), D(72), D(44), relief, (14.0, 10.0)); add("wall_west", v, i, brick, tc, 1)
    v, i, tc = _grid((x1, 0, z0), (0, 0, z1 - z0), (0, H, 0), D(72), D(44), relief, (14.0, 10.0)); add("wall_east", v, i, brick, tc, 1)
    # colonnades: two rows, two storeys
    col_x = np.linspace(x0 + 250.0, x1 - 250.0, 10)
    storeys = ((0.0, 620.0), (680.0, 1250.0))
    for si, (yb, yt) in enumerate(storeys):
        hgt = yt - yb
        prof_t = np.linspace(0.0, 1.0, D(26))
        base_r = 58.0 - 6.0 * si
        # entasis + base + capital flare + fluting handled by radial ripple
        prof_r = base_r * (1.0 - 0.12 * prof_t) + 22.0 * np.exp(-((prof_t - 0.0) / 0.05) ** 2) + 26.0 * np.exp(-((prof_t - 1.0) / 0.06) ** 2)
        for zi, zc in enumerate((-430.0, 430.0)):
            for ci, xc in enumerate(col_x):
                v, i, tc = _lathe(prof_r, prof_t * hgt, (xc, yb, zc), D(30))
                ang = np.arctan2(-(v[:, 2] - zc), v[:, 0] - xc)
                ripple = 1.0 + 0.03 * np.cos(ang * 12)
                v[:, 0] = xc + (v[:, 0] - xc) * ripple; v[:, 2] = zc + (v[:, 2] - zc) * ripple
                hue = 0.70 + 0.08 * rng.random()
                add(f"column_{si}_{zi}_{ci}", v, i, _disney((hue, hue * 0.96, hue * 0.88)), tc, -1)
        # arches between neighbouring columns: half tori in the x-y plane
        for zi, zc in enumerate((-430.0, 430.0)):
            parts = []
            for ci in range(len(col_x) - 1):
                xa, xb = col_x[ci], col_x[ci + 1]
                R = 0.5 * (xb - xa) - 40.0; rr = 34.0
                nu, nv = D(24), D(14)
                u = np.linspace(0.0, math.pi, nu + 1); w = np.linspace(0.0, 2 * math.pi, nv + 1)
                Uu, Ww = np.meshgrid(u, w, indexing="xy")
                cx_ = 0.5 * (xa + xb); cy_ = yt - R * 0.55
                X = cx_ - (R + rr * np.cos(Ww)) * np.cos(Uu)
                Y = cy_ + (R + rr * np.cos(Ww)) * np.sin(Uu) * 0.55
                Z = zc + rr * np.sin(Ww)
                P = np.stack([X, Y, Z], -1).reshape(-1, 3)
                ii, jj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
                a = (jj * (nu + 1) + ii).reshape(-1); b = a + 1; c = a + (nu + 1) + 1; d = a + (nu + 1)
                parts.append((P, np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])))
            v, i = _merge(parts)
            add(f"arches_{si}_{zi}", v, i, _disney((0.80, 0.76, 0.68)))
    # gallery slabs (upper floor) between colonnade and wall, top and underside
    for zi, (za, zb) in enumerate(((z0, -380.0), (380.0, z1))):
        v, i, tc = _grid((x0, 650.0, zb), (x1 - x0, 0, 0), (0, 0, za - zb), D(128), D(16), None, (20.0, 4.0)); add(f"gallery_top_{zi}", v, i, stone, tc, 0)
        v, i, tc = _grid((x0, 620.0, za), (x1 - x0, 0, 0), (0, 0, zb - za), D(128), D(16), None, (20.0, 4.0)); add(f"gallery_bottom_{zi}", v, i, stone, tc, 0)
        # roof over the gallery (normal down), the nave stays open to the sky
        v, i, tc = _grid((x0, 1290.0, za), (x1 - x0, 0, 0), (0, 0, zb - za), D(96), D(14), None, (20.0, 4.0)); add(f"roof_{zi}", v, i, _disney((0.55, 0.42, 0.33)), tc, -1)
    # hanging drapes across the nave, folded cloth; two-sided by construction (single sheet)
    cloth = ((0.75, 0.08, 0.06), (0.10, 0.45, 0.14), (0.12, 0.16, 0.62), (0.78, 0.62, 0.10), (0.55, 0.10, 0.45), (0.85, 0.85, 0.80))
    for k, xc in enumerate(np.linspace(x0 + 600.0, x1 - 500.0, 6)):
        ph = rng.random() * 6.28
        fold = lambda S, T, ph=ph: 38.0 * np.sin(S * 2 * math.pi * 9 + ph) * (0.35 + 0.65 * T) + 25.0 * np.sin(T * 7.0 + ph)
        v, i, tc = _grid((xc, 1180.0, -330.0), (0, 0, 660.0), (0, -520.0, 0), D(66), D(62), fold, (6.0, 6.0))
        add(f"drape_{k}", v, i, _disney(cloth[k % len(cloth)], roughness=0.8, metallic=0.1, transmission=0.0), tc, -1)
    # vases along the nave (smooth metallic / clearcoat materials)
    vt = np.linspace(0.0, 1.0, D(34))
    vr = 30.0 + 55.0 * np.sin(vt * math.pi) ** 1.5 * (1.0 - 0.45 * vt) + 18.0 * np.exp(-((vt - 1.0) / 0.08) ** 2)
    for k, xc in enumerate(np.linspace(x0 + 450.0, x1 - 450.0, 8)):
        zc = -250.0 if k % 2 == 0 else 250.0
        v, i, tc = _lathe(vr, vt * 190.0, (xc, 0.0, zc), D(30))
        m = _disney((0.25 + 0.5 * rng.random(), 0.3 + 0.4 * rng.random(), 0.3 + 0.5 * rng.random()),
                    roughness=0.25 + 0.3 * rng.random(), metallic=0.8 if k % 3 == 0 else 0.2, clearcoat=1.0 if k % 3 == 1 else 0.0,
                    transmission=0.0 if k % 4 else 0.9)
        add(f"vase_{k}", v, i, m, tc, -1)
    # sculpture at the far end: displaced icosphere on a plinth
    lvl = 5 if detail >= 0.99 else max(2, int(round(5 + math.log2(max(detail, 1e-3)))))
    sv, sf = _icosphere(lvl)
    disp = 1.0 + 0.22 * _value_noise3(sv, seed, 2.3) + 0.08 * _value_noise3(sv, seed + 7, 6.1)
    sv = sv * disp[:, None] * 170.0 + np.array([1250.0, 330.0, 0.0])
    add("sculpture", sv, sf, _disney((0.82, 0.78, 0.70), roughness=0.55))
    parts = [_quad((1130, 160, 120), (1370, 160, 120), (1370, 160, -120), (1130, 160, -120))]
    for (ax, az), (bx, bz) in (((1130, 120), (1370, 120)), ((1370, 120), (1370, -120)), ((1370, -120), (1130, -120)), ((1130, -120), (1130, 120))):
        parts.append(_quad((ax, 0, az), (bx, 0, bz), (bx, 160, bz), (ax, 160, az)))
    v, i = _merge(parts)
    add("plinth", v, i, stone)
    return sc


def displaced_grid(n_tris=10_000_000, seed=3, extent=4000.0) -> Scene:
    """C5: one height-field mesh with ~n_tris triangles (fractal sine terrain), split into bands so
    that no single mesh is absurdly large; deterministic."""
    n = int(math.ceil(math.sqrt(n_tris / 2.0)))
    rng = np.random.default_rng(seed)
    ph = rng.random(8) * 6.28
    sc = Scene(name=f"grid{n_tris}")
    bands = max(1, n // 512)
    rows = np.linspace(0, n, bands + 1).astype(int)
    for b in range(bands):
        r0, r1 = rows[b], rows[b + 1]
        s = np.linspace(0.0, 1.0, n + 1, dtype=np.float64); t = np.linspace(r0 / n, r1 / n, r1 - r0 + 1, dtype=np.float64)
        S, T = np.meshgrid(s, t, indexing="xy")
        Hh = sum((180.0 / (k + 1)) * np.sin(S * (k + 1) * 9.0 + ph[k]) * np.cos(T * (k + 1) * 7.0 + ph[(k + 3) % 8]) for k in range(5))
        P = np.stack([(S - 0.5) * extent, Hh, (0.5 - T) * extent], -1).reshape(-1, 3).astype(np.float32)
        nu, nv = n, r1 - r0
        ii, jj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
        a = (jj * (nu + 1) + ii).reshape(-1); bb = a + 1; c = a + (nu + 1) + 1; d = a + (nu + 1)
        idx = np.concatenate([np.stack([a, bb, c], 1), np.stack([a, c, d], 1)]).astype(np.uint32)
        col = (0.35 + 0.4 * rng.random(), 0.45 + 0.3 * rng.random(), 0.3 + 0.3 * rng.random())
        sc.meshes.append(Mesh(f"band_{b}", P, idx, _disney(col)))
    return sc
