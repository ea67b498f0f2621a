"""ctypes driver for the two CPU checkers (TEST INFRASTRUCTURE -- see oracle/orc_api.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  ``load("port")`` opens oracle/liboracle.so (the restatement), ``load("reference")``
opens oracle/_ref/libfovref.so (the reference's sources compiled unmodified for the host).
"""
import ctypes as C
import os

import numpy as np

from fovpathtracing_optixcodelatest_b200.structs import LaunchParams, Material

HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {"port": os.path.join(HERE, "liboracle.so"), "reference": os.path.join(HERE, "_ref", "libfovref.so")}


class Opts(C.Structure):
    _fields_ = [
        ("max_depth", C.c_int32), ("tmin", C.c_float), ("tmax", C.c_float), ("jitter_yx", C.c_int32),
        ("threads", C.c_int32), ("accumulate", C.c_int32), ("exposure", C.c_float),
        ("seed_all_passes_with_subframe", C.c_int32),
        ("primary_tri", C.c_void_p), ("primary_tuv", C.c_void_p), ("primary_dir", C.c_void_p),
        ("features", C.c_int32), ("filter", C.c_int32),
    ]

    @classmethod
    def sv5(cls, **kw):
        o = cls(max_depth=4, tmin=0.01, tmax=1e16, jitter_yx=0, threads=0, accumulate=0, exposure=16.0,
                seed_all_passes_with_subframe=0, features=0, filter=0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class FovCfg(C.Structure):
    _fields_ = [("foveated", C.c_int32), ("inner_radius", C.c_int32), ("outer_radius", C.c_int32),
                ("spp_periphery", C.c_int32), ("spp_mid", C.c_int32), ("spp_fovea", C.c_int32),
                ("spp_uniform", C.c_int32)]

    @classmethod
    def sv5(cls, **kw):
        o = cls(foveated=1, inner_radius=74, outer_radius=241, spp_periphery=8, spp_mid=16, spp_fovea=32,
                spp_uniform=4)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


def available(kind):
    return os.path.exists(PATHS[kind])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, kind="port"):
        path = PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (make -C oracle {'port' if kind == 'port' else 'ref'})")
        L = C.CDLL(path)
        self.L = L
        L.orc_kind.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_tea4.restype = C.c_uint32
        L.orc_tea4.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_tonemap.restype = C.c_uint32
        L.orc_tonemap.argtypes = [C.c_void_p, C.c_float]
        for name in ("orc_add_texture", "orc_add_mesh", "orc_load_obj", "orc_add_box", "orc_mesh_count",
                     "orc_mesh_info", "orc_mesh_get", "orc_texture_count", "orc_texture_info", "orc_texture_get",
                     "orc_commit", "orc_set_probe", "orc_get_probe", "orc_camera_uvw", "orc_launch", "orc_render",
                     "orc_trace", "orc_probe_sample", "orc_probe_eval"):
            getattr(L, name).restype = C.c_int
        L.orc_add_texture.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_add_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_load_obj.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_add_box.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mesh_count.argtypes = [C.c_void_p]
        L.orc_mesh_info.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.orc_mesh_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.orc_texture_count.argtypes = [C.c_void_p]
        L.orc_texture_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_texture_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_commit.argtypes = [C.c_void_p]
        L.orc_set_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_get_probe.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orc_camera_uvw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.orc_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_reset_counters.argtypes = [C.c_void_p]
        L.orc_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_rnd_seq.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_random_seq.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_bsdf_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_bsdf_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_probe_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_probe_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.kind = L.orc_kind().decode()
        self.ctx = C.c_void_p(L.orc_create())
        self._keep = []

    def close(self):
        if self.ctx:
            self.L.orc_destroy(self.ctx)
            self.ctx = None

    def _chk(self, rc):
        if rc < 0:
            raise RuntimeError(self.L.orc_last_error(self.ctx).decode())
        return rc

    # ---- scene
    def add_texture(self, rgba8):
        a = np.ascontiguousarray(rgba8, dtype=np.uint32)
        return self._chk(self.L.orc_add_texture(self.ctx, _ptr(a), a.shape[1], a.shape[0]))

    def add_mesh(self, vtx, idx, material, texcoord=None, texture_id=-1):
        v = np.ascontiguousarray(vtx, dtype=np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        t = None if texcoord is None else np.ascontiguousarray(texcoord, dtype=np.float32).reshape(-1, 2)
        return self._chk(self.L.orc_add_mesh(self.ctx, _ptr(v), len(v), _ptr(i), len(i), _ptr(t), C.byref(material), texture_id))

    def add_scene(self, scene):
        for tex in scene.textures:
            self.add_texture(tex)
        for m in scene.meshes:
            self.add_mesh(m.vtx, m.idx, m.material, m.texcoord, m.texture_id)

    def load_obj(self, path):
        return self._chk(self.L.orc_load_obj(self.ctx, path.encode()))

    def add_box(self, material, pos, ext):
        p = np.asarray(pos, np.float32); e = np.asarray(ext, np.float32)
        return self._chk(self.L.orc_add_box(self.ctx, C.byref(material), _ptr(p), _ptr(e)))

    def meshes(self):
        out = []
        for m in range(self.L.orc_mesh_count(self.ctx)):
            nv, nt, ntc, tex = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            self._chk(self.L.orc_mesh_info(self.ctx, m, C.byref(nv), C.byref(nt), C.byref(ntc), C.byref(tex)))
            v = np.zeros((nv.value, 3), np.float32); i = np.zeros((nt.value, 3), np.uint32)
            tc = np.zeros((ntc.value, 2), np.float32); mat = Material()
            self._chk(self.L.orc_mesh_get(self.ctx, m, _ptr(v), _ptr(i), _ptr(tc) if ntc.value else None, C.byref(mat)))
            out.append(dict(vtx=v, idx=i, texcoord=tc if ntc.value else None, material=mat, texture_id=tex.value))
        return out

    def textures(self):
        out = []
        for t in range(self.L.orc_texture_count(self.ctx)):
            w, h = C.c_int(), C.c_int()
            self._chk(self.L.orc_texture_info(self.ctx, t, C.byref(w), C.byref(h)))
            a = np.zeros((h.value, w.value), np.uint32)
            self._chk(self.L.orc_texture_get(self.ctx, t, _ptr(a)))
            out.append(a)
        return out

    def commit(self):
        self._chk(self.L.orc_commit(self.ctx))

    # ---- probe / camera
    def set_probe(self, rgba, offset=(0.0, 0.0, 0.0)):
        a = np.ascontiguousarray(rgba, dtype=np.float32)
        assert a.ndim == 3 and a.shape[2] == 4
        o = np.asarray(offset, np.float32)
        self._probe_shape = a.shape[:2]
        self._chk(self.L.orc_set_probe(self.ctx, a.shape[1], a.shape[0], _ptr(a), _ptr(o)))

    def get_probe(self):
        h, w = self._probe_shape
        pdfX = np.zeros((h, w), np.float32); cdfX = np.zeros((h, w), np.float32)
        pdfY = np.zeros(h, np.float32); cdfY = np.zeros(h, np.float32)
        self._chk(self.L.orc_get_probe(self.ctx, _ptr(pdfX), _ptr(cdfX), _ptr(pdfY), _ptr(cdfY)))
        return pdfX, cdfX, pdfY, cdfY

    def camera_uvw(self, eye, lookat, up, fovy, aspect):
        e = np.asarray(eye, np.float32); l = np.asarray(lookat, np.float32); u = np.asarray(up, np.float32)
        o = np.zeros(9, np.float32)
        self.L.orc_camera_uvw(_ptr(e), _ptr(l), _ptr(u), fovy, aspect, _ptr(o))
        return o[0:3].copy(), o[3:6].copy(), o[6:9].copy()

    # ---- rendering
    def make_params(self, w, h, eye, U, V, W, gaze=None, subframe=0):
        """LaunchParams with freshly allocated host accum/frame buffers (returned alongside)."""
        lp = LaunchParams()
        accum = np.zeros((h, w, 4), np.float32)
        frame = np.zeros((h, w), np.uint32)
        lp.frame.accum_buffer = accum.ctypes.data
        lp.frame.frame_buffer = frame.ctypes.data
        lp.frame.size_x, lp.frame.size_y = w, h
        lp.frame.subframe_index = subframe
        gx, gy = gaze if gaze is not None else (w // 2, h // 2)
        lp.frame.c_x, lp.frame.c_y = gx, gy
        lp.camera.eye.set(eye); lp.camera.U.set(U); lp.camera.V.set(V); lp.camera.W.set(W)
        lp.viewport_x, lp.viewport_y = w, h
        lp.white = 1.0
        lp._keep = (accum, frame)          # the blob holds raw pointers into these arrays
        return lp, accum, frame

    def launch(self, lp, lw, lh, opts):
        self._chk(self.L.orc_launch(self.ctx, C.byref(lp), lw, lh, C.byref(opts)))

    def render(self, lp, cfg, opts):
        self._chk(self.L.orc_render(self.ctx, C.byref(lp), C.byref(cfg), C.byref(opts)))

    def counters(self):
        a = np.zeros(8, np.uint64)
        self.L.orc_get_counters(self.ctx, _ptr(a))
        return dict(radiance_rays=int(a[0]), shadow_rays=int(a[1]), closest_nodes=int(a[2]), closest_tris=int(a[3]),
                    any_nodes=int(a[4]), any_tris=int(a[5]), samples=int(a[6]))

    def reset_counters(self):
        self.L.orc_reset_counters(self.ctx)

    def trace(self, rays, mode):
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        tri = np.zeros(len(r), np.int32); tuv = np.zeros((len(r), 3), np.float32)
        self._chk(self.L.orc_trace(self.ctx, _ptr(r), len(r), mode, _ptr(tri), _ptr(tuv)))
        return tri, tuv

    # ---- unit hooks
    def tea4(self, a, b):
        return int(self.L.orc_tea4(a, b))

    def rnd_seq(self, seed, n):
        out = np.zeros(n, np.float32); s = C.c_uint32()
        self.L.orc_rnd_seq(seed, n, _ptr(out), C.byref(s))
        return out, s.value

    def random_seq(self, seed, n):
        u = np.zeros(n, np.uint32); f = np.zeros(n, np.float32)
        self.L.orc_random_seq(seed, n, _ptr(u), _ptr(f))
        return u, f

    def bsdf_sample(self, material, albedo, etaI, etaO, N, V, seed):
        a = np.asarray(albedo, np.float32); n = np.asarray(N, np.float32); v = np.asarray(V, np.float32)
        o = np.zeros(10, np.float32)
        self.L.orc_bsdf_sample(C.byref(material), _ptr(a), etaI, etaO, _ptr(n), _ptr(v), seed, _ptr(o))
        return o

    def bsdf_eval(self, material, albedo, etaI, etaO, N, V, Ldir):
        a = np.asarray(albedo, np.float32); n = np.asarray(N, np.float32); v = np.asarray(V, np.float32)
        l = np.asarray(Ldir, np.float32); o = np.zeros(4, np.float32)
        self.L.orc_bsdf_eval(C.byref(material), _ptr(a), etaI, etaO, _ptr(n), _ptr(v), _ptr(l), _ptr(o))
        return o

    def probe_sample(self, seed):
        o = np.zeros(7, np.float32)
        self._chk(self.L.orc_probe_sample(self.ctx, seed, _ptr(o)))
        return o

    def probe_eval(self, d):
        dd = np.asarray(d, np.float32); o = np.zeros(6, np.float32)
        self._chk(self.L.orc_probe_eval(self.ctx, _ptr(dd), _ptr(o)))
        return o

    def tonemap(self, accum, exposure=16.0):
        a = np.asarray(accum, np.float32)
        return int(self.L.orc_tonemap(_ptr(a), exposure))
