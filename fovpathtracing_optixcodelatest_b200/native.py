"""ctypes binding of ``libfovpt.so`` (C ABI: ``include/fovpt.h``) -- the product path.

There is no CPU fallback: if the shared library is missing or no CUDA device is present this
module raises.  Nothing here imports ``oracle/``.
"""
import ctypes as C
import os

import numpy as np

from .structs import LaunchParams, Material, Probe

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfovpt.so")


class Config(C.Structure):
    """``fovpt_config``; ``Config.default()`` = sv5's literals."""
    _fields_ = [
        ("max_depth", C.c_int32), ("tmin", C.c_float), ("tmax", C.c_float), ("exposure", C.c_float),
        ("accumulate", C.c_int32), ("foveated", C.c_int32), ("inner_radius", C.c_int32), ("outer_radius", C.c_int32),
        ("spp_periphery", C.c_int32), ("spp_mid", C.c_int32), ("spp_fovea", C.c_int32), ("spp_uniform", C.c_int32),
        ("seed_all_passes_with_subframe", C.c_int32), ("skip_wasted_trace", C.c_int32), ("fast_math", C.c_int32),
        ("filter", C.c_int32), ("features", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64), ("radiance_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("frames", C.c_uint64), ("bvh_nodes", C.c_uint64), ("bvh_tris", C.c_uint64),
        ("bvh_build_ms", C.c_float), ("last_frame_ms", C.c_float), ("last_extend_ms", C.c_float),
        ("last_shade_ms", C.c_float), ("last_shadow_ms", C.c_float), ("last_raygen_ms", C.c_float),
        ("last_resolve_ms", C.c_float), ("trace_launches", C.c_uint64),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "fovpt_create", "fovpt_destroy", "fovpt_last_error", "fovpt_stream", "fovpt_default_config", "fovpt_set_config",
    "fovpt_get_config", "fovpt_scene_begin", "fovpt_scene_add_texture", "fovpt_scene_add_mesh", "fovpt_scene_commit",
    "fovpt_probe_upload", "fovpt_resize", "fovpt_launch", "fovpt_render_frame", "fovpt_sync", "fovpt_download",
    "fovpt_download_accum", "fovpt_trace", "fovpt_primary_hits", "fovpt_set_tile_owner", "fovpt_get_stats",
    "fovpt_reset_stats", "fovpt_bvh_info", "fovpt_bvh_layout", "fovpt_bvh_read",
    "fovpt_peer_frame_create", "fovpt_peer_frame_open", "fovpt_peer_frame_signal", "fovpt_peer_frame_wait",
    "fovpt_peer_frame_release", "fovpt_peer_frame_close", "fovpt_set_kernel_timing",
    "fovpt_render_frames", "fovpt_peer_frame_accum", "fovpt_forget_frame", "fovpt_texture_check",
    "fovpt_feature_buffers", "fovpt_download_features", "fovpt_upload_accum",
]

_lib = None


def load_library():
    """dlopen libfovpt.so and declare signatures.  Raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FOVPT_LIB", LIB_PATH)        # developer knob: A/B-test another build of the same library
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with __graft_entry__.build() "
                           f"(make -C fovpathtracing_optixcodelatest_b200/csrc); there is no CPU fallback")
    L = C.CDLL(path)
    vp, i32, u64p = C.c_void_p, C.c_int, C.POINTER(C.c_uint64)
    L.fovpt_create.argtypes = [C.POINTER(vp), i32, vp]
    L.fovpt_destroy.argtypes = [vp]
    L.fovpt_last_error.argtypes = [vp]; L.fovpt_last_error.restype = C.c_char_p
    L.fovpt_stream.argtypes = [vp]; L.fovpt_stream.restype = vp
    L.fovpt_default_config.argtypes = [C.POINTER(Config)]; L.fovpt_default_config.restype = None
    L.fovpt_set_config.argtypes = [vp, C.POINTER(Config)]
    L.fovpt_get_config.argtypes = [vp, C.POINTER(Config)]
    L.fovpt_scene_begin.argtypes = [vp]
    L.fovpt_scene_add_texture.argtypes = [vp, vp, i32, i32]
    L.fovpt_scene_add_mesh.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, vp, vp, i32]
    L.fovpt_scene_commit.argtypes = [vp, u64p]
    L.fovpt_probe_upload.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.fovpt_resize.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(vp)]
    L.fovpt_launch.argtypes = [vp, vp, C.c_size_t, C.c_uint, C.c_uint]
    L.fovpt_render_frame.argtypes = [vp, vp]
    L.fovpt_render_frames.argtypes = [vp, vp, i32]
    L.fovpt_forget_frame.argtypes = [vp]
    L.fovpt_feature_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.fovpt_download_features.argtypes = [vp, vp, vp, vp, C.c_size_t]
    L.fovpt_texture_check.argtypes = [vp, i32, vp, C.c_size_t, i32, vp, vp]
    L.fovpt_peer_frame_accum.argtypes = [vp, i32, C.POINTER(vp)]
    L.fovpt_sync.argtypes = [vp]
    L.fovpt_download.argtypes = [vp, vp, C.c_size_t]
    L.fovpt_download_accum.argtypes = [vp, vp, C.c_size_t]
    L.fovpt_upload_accum.argtypes = [vp, vp, C.c_size_t]
    L.fovpt_trace.argtypes = [vp, vp, C.c_size_t, i32, vp, vp]
    L.fovpt_primary_hits.argtypes = [vp, vp, C.c_uint, C.c_uint, vp, vp, vp]
    L.fovpt_set_tile_owner.argtypes = [vp, i32, i32, i32, vp, i32]
    L.fovpt_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.fovpt_reset_stats.argtypes = [vp]
    L.fovpt_bvh_info.argtypes = [vp, u64p, u64p]
    L.fovpt_bvh_layout.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fovpt_bvh_read.argtypes = [vp, vp, vp]
    L.fovpt_set_kernel_timing.argtypes = [vp, i32]
    L.fovpt_peer_frame_create.argtypes = [vp, C.c_size_t, i32, i32, vp, C.POINTER(vp)]
    L.fovpt_peer_frame_open.argtypes = [vp, vp, i32, C.POINTER(vp)]
    for n in ("signal", "wait", "release", "close"):
        getattr(L, "fovpt_peer_frame_" + n).argtypes = [vp]
    for n in EXPORTS:
        if n not in ("fovpt_last_error", "fovpt_stream", "fovpt_default_config"):
            getattr(L, n).restype = C.c_int
    _lib = L
    return L


def default_config() -> Config:
    k = Config()
    load_library().fovpt_default_config(C.byref(k))
    return k


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """Thin owner of one ``fovpt_ctx`` (one GPU, one stream)."""

    def __init__(self, device=0, stream=None):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.fovpt_create(C.byref(self.h), device, stream)
        if rc != 0:
            raise RuntimeError(f"fovpt_create failed (code {rc}): no usable CUDA device {device}; "
                               f"libfovpt has no CPU fallback")
        self.device = device

    def close(self):
        if self.h:
            self.L.fovpt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.fovpt_last_error(self.h).decode())

    # ---- config
    def set_config(self, cfg: Config):
        self._chk(self.L.fovpt_set_config(self.h, C.byref(cfg)))

    def get_config(self) -> Config:
        k = Config()
        self._chk(self.L.fovpt_get_config(self.h, C.byref(k)))
        return k

    # ---- scene
    def scene_begin(self):
        self._chk(self.L.fovpt_scene_begin(self.h))

    def add_texture(self, rgba8):
        a = np.ascontiguousarray(rgba8, dtype=np.uint32)
        r = self.L.fovpt_scene_add_texture(self.h, _p(a), a.shape[1], a.shape[0])
        if r < 0:
            raise RuntimeError(self.L.fovpt_last_error(self.h).decode())
        return r

    def add_mesh(self, vtx, idx, material: Material, texcoord=None, texture_id=-1):
        v = np.ascontiguousarray(vtx, dtype=np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        t = None if texcoord is None else np.ascontiguousarray(texcoord, dtype=np.float32).reshape(-1, 2)
        r = self.L.fovpt_scene_add_mesh(self.h, _p(v), len(v), _p(i), len(i), _p(t), C.byref(material), texture_id)
        if r < 0:
            raise RuntimeError(self.L.fovpt_last_error(self.h).decode())
        return r

    def upload_scene(self, scene):
        self.scene_begin()
        for tex in scene.textures:
            self.add_texture(tex)
        for m in scene.meshes:
            self.add_mesh(m.vtx, m.idx, m.material, m.texcoord, m.texture_id)
        return self.commit()

    def commit(self):
        t = C.c_uint64()
        self._chk(self.L.fovpt_scene_commit(self.h, C.byref(t)))
        return t.value

    # ---- probe / buffers
    def probe_upload(self, rgba, pdfX, cdfX, pdfY, cdfY, offset=(0.0, 0.0, 0.0)) -> Probe:
        a = np.ascontiguousarray(rgba, dtype=np.float32)
        h, w = a.shape[:2]
        arrs = [np.ascontiguousarray(x, dtype=np.float32) for x in (pdfX, cdfX, pdfY, cdfY)]
        o = np.asarray(offset, np.float32)
        pb = Probe()
        self._chk(self.L.fovpt_probe_upload(self.h, w, h, _p(a), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(o), C.byref(pb)))
        return pb

    def resize(self, w, h):
        a, f = C.c_void_p(), C.c_void_p()
        self._chk(self.L.fovpt_resize(self.h, w, h, C.byref(a), C.byref(f)))
        return a.value, f.value

    # ---- rendering
    def launch(self, lp: LaunchParams, lw, lh):
        self._chk(self.L.fovpt_launch(self.h, C.byref(lp), C.sizeof(lp), lw, lh))

    def render_frame(self, lp: LaunchParams):
        self._chk(self.L.fovpt_render_frame(self.h, C.byref(lp)))

    def render_frames(self, lp: LaunchParams, n_subframes: int):
        """n accumulated subframes traced as one wavefront (bit-identical to n render_frame calls)"""
        self._chk(self.L.fovpt_render_frames(self.h, C.byref(lp), n_subframes))

    def sync(self):
        self._chk(self.L.fovpt_sync(self.h))

    def download(self, w, h):
        out = np.empty((h, w), np.uint32)
        self._chk(self.L.fovpt_download(self.h, _p(out), w * h))
        return out

    def download_into(self, out):
        self._chk(self.L.fovpt_download(self.h, _p(out), out.size))

    def download_accum(self, w, h):
        out = np.empty((h, w, 4), np.float32)
        self._chk(self.L.fovpt_download_accum(self.h, _p(out), w * h))
        return out

    def upload_accum(self, accum):
        a = np.ascontiguousarray(accum, dtype=np.float32)
        self._chk(self.L.fovpt_upload_accum(self.h, _p(a), a.size // 4))

    def feature_buffers(self):
        """device pointers (color, normal, albedo) of the denoiser buffers at the size of the last resize"""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self.L.fovpt_feature_buffers(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def download_features(self, w, h):
        out = [np.empty((h, w, 4), np.float32) for _ in range(3)]
        self._chk(self.L.fovpt_download_features(self.h, _p(out[0]), _p(out[1]), _p(out[2]), w * h))
        return dict(color=out[0], normal=out[1], albedo=out[2])

    def trace(self, rays, mode):
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        tri = np.zeros(len(r), np.int32); tuv = np.zeros((len(r), 3), np.float32)
        self._chk(self.L.fovpt_trace(self.h, _p(r), len(r), mode, _p(tri), _p(tuv)))
        return tri, tuv

    def texture_check(self, texture_id, uv, mode=-1):
        """(software fetch, hardware texture-unit fetch) of texture `texture_id` at uv[n,2]; both n x 4 float32"""
        u = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        sw = np.zeros((len(u), 4), np.float32); hw = np.zeros((len(u), 4), np.float32)
        self._chk(self.L.fovpt_texture_check(self.h, texture_id, _p(u), len(u), mode, _p(sw), _p(hw)))
        return sw, hw

    def primary_hits(self, lp: LaunchParams, lw, lh):
        n = lw * lh
        tri = np.zeros(n, np.int32); tuv = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32)
        self._chk(self.L.fovpt_primary_hits(self.h, C.byref(lp), lw, lh, _p(tri), _p(tuv), _p(d)))
        return tri, tuv, d

    def set_tile_owner(self, tile_size, owner, rank):
        if owner is None:
            self._chk(self.L.fovpt_set_tile_owner(self.h, 0, 0, 0, None, 0))
            return
        o = np.ascontiguousarray(owner, dtype=np.uint8)
        self._chk(self.L.fovpt_set_tile_owner(self.h, tile_size, o.shape[1], o.shape[0], _p(o), rank))

    # ---- shared frame of a multi-GPU job (include/fovpt.h, fovpt_peer_frame_*)
    def peer_frame_create(self, npix, world, owner_rank=0):
        """owner: allocate the shared RGBA8 frame; returns (128-byte handle for the other ranks, device pointer)"""
        handle = (C.c_ubyte * 128)(); ptr = C.c_void_p()
        self._chk(self.L.fovpt_peer_frame_create(self.h, npix, world, owner_rank, handle, C.byref(ptr)))
        return bytes(handle), ptr.value

    def peer_frame_open(self, handle: bytes, rank):
        """contributor: map the owner's frame; returns the device pointer to put in LaunchParams.frame.frame_buffer"""
        buf = (C.c_ubyte * 128).from_buffer_copy(handle); ptr = C.c_void_p()
        self._chk(self.L.fovpt_peer_frame_open(self.h, buf, rank, C.byref(ptr)))
        return ptr.value

    def peer_frame_accum(self, mirror=True):
        """device pointer of the shared float4 accumulation buffer; mirror: every rank stores its pixels' accumulation there"""
        ptr = C.c_void_p()
        self._chk(self.L.fovpt_peer_frame_accum(self.h, 1 if mirror else 0, C.byref(ptr)))
        return ptr.value

    def peer_frame_signal(self):
        self._chk(self.L.fovpt_peer_frame_signal(self.h))

    def peer_frame_wait(self):
        self._chk(self.L.fovpt_peer_frame_wait(self.h))

    def peer_frame_release(self):
        self._chk(self.L.fovpt_peer_frame_release(self.h))

    def peer_frame_close(self):
        self._chk(self.L.fovpt_peer_frame_close(self.h))

    def set_kernel_timing(self, on: bool):
        self._chk(self.L.fovpt_set_kernel_timing(self.h, 1 if on else 0))

    def stats(self) -> dict:
        s = Stats()
        self._chk(self.L.fovpt_get_stats(self.h, C.byref(s)))
        return s.asdict()

    def reset_stats(self):
        self._chk(self.L.fovpt_reset_stats(self.h))

    def bvh(self):
        nn, nt = C.c_uint64(), C.c_uint64()
        self._chk(self.L.fovpt_bvh_info(self.h, C.byref(nn), C.byref(nt)))
        wide, nbytes = C.c_int(), C.c_int()
        self._chk(self.L.fovpt_bvh_layout(self.h, C.byref(wide), C.byref(nbytes)))
        self.bvh_wide = wide.value
        nodes = np.zeros((nn.value, nbytes.value // 4), np.uint32); ids = np.zeros(nt.value, np.uint32)
        self._chk(self.L.fovpt_bvh_read(self.h, _p(nodes), _p(ids)))
        return nodes, ids
