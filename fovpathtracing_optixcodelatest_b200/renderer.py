"""Python mirror of the reference host API for the foveated path-tracing path.

Same names, argument meaning, call-order contract and error behaviour as the reference's C++
``class SampleRenderer`` (``PT_sv5_/SimplePathtracer.h:45-189``) and ``struct ProbeData``
(``PT_sv5_/Probe.h:5-100``); the C++ drop-in itself lives in ``host/SimplePathtracer.{h,cpp}``.
Everything below the method bodies is the C ABI of ``libfovpt`` -- no OptiX, no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import native
from .scenes import CameraSpec, Scene, camera_uvw
from .structs import LaunchParams


class ProbeData:
    """Host probe + luminance CDFs (reference ``Probe.h:5-100``).  ``BuildCDF`` follows
    ``Probe.h:29-77`` operation for operation in float32 (sequential running sums)."""

    def __init__(self, data=None, offset=(0.0, 0.0, 0.0)):
        self.valid = False
        self.data = None
        self.width = self.height = 0
        self.offset = tuple(offset)
        if data is not None:
            self.data = np.ascontiguousarray(data, dtype=np.float32)
            self.height, self.width = self.data.shape[:2]

    def BuildCDF(self):
        f = np.float32
        d = self.data
        weight = (d[..., 0] * f(0.3) + d[..., 1] * f(0.6) + d[..., 2] * f(0.1)).astype(f)   # Luminance, maths.h:163-166
        cdfX = np.cumsum(weight, axis=1, dtype=f)          # float32 running sum, left to right
        totalX = cdfX[:, -1].copy()
        inv = (f(1.0) / totalX).astype(f)
        self.pdfValuesX = (weight * inv[:, None]).astype(f)
        self.cdfValuesX = (cdfX * inv[:, None]).astype(f)
        cumY = np.cumsum(totalX, dtype=f)
        totalY = cumY[-1]
        self.cdfValuesY = (cumY / totalY).astype(f)
        self.pdfValuesY = (totalX / totalY).astype(f)
        self.valid = True
        return self


class SampleRenderer:
    """Drop-in for the reference renderer on this path.

    Call order contract (``PT_sv5_/main.cpp:297-306``): ctor -> ``resize`` -> ``setCamera`` ->
    ``setProbe`` -> per frame: set ``launchParams.frame.c_x/c_y`` (gaze) -> ``render`` ->
    ``downloadPixels``.  ``launchParams`` is public and writable exactly as in the reference.
    """

    def __init__(self, model: Scene, device: int = 0, config: native.Config = None):
        self.ctx = native.Context(device)                 # raises without a GPU / built extension
        self.launchParams = LaunchParams()
        self.model = model
        self.config = config if config is not None else native.default_config()
        self.ctx.set_config(self.config)
        self.launchParams.traversable = self.ctx.upload_scene(model)     # buildAccel, SimplePathtracer.cpp:58
        self.lastSetCamera = None

    # -- SampleRenderer::resize, SimplePathtracer.cpp:228-274
    def resize(self, newSize):
        w, h = int(newSize[0]), int(newSize[1])
        if w == 0 or h == 0:
            return
        accum, frame = self.ctx.resize(w, h)
        lp = self.launchParams
        lp.frame.size_x, lp.frame.size_y = w, h
        lp.frame.accum_buffer, lp.frame.frame_buffer = accum, frame
        lp.viewport_x, lp.viewport_y = w, h
        self._feature_buffers()

    # the denoiser buffers resize() allocates in the reference (SimplePathtracer.cpp:245-260); here only when they are written
    def _feature_buffers(self):
        lp = self.launchParams
        if self.config.features and lp.frame.size_x:
            lp.frame.color_buffer, lp.frame.normal_buffer, lp.frame.albedo_buffer = self.ctx.feature_buffers()
        else:
            lp.frame.color_buffer = lp.frame.normal_buffer = lp.frame.albedo_buffer = 0

    # -- SampleRenderer::setCamera, SimplePathtracer.cpp:282-289 (aspect from frame.size)
    def setCamera(self, camera: CameraSpec):
        lp = self.launchParams
        if lp.frame.size_y == 0:
            raise RuntimeError("setCamera before resize: frame.size is 0 (reference divides by it, SimplePathtracer.cpp:286)")
        self.lastSetCamera = camera
        U, V, W = camera_uvw(camera, np.float32(lp.frame.size_x) / np.float32(lp.frame.size_y))
        lp.camera.eye.set(camera.eye); lp.camera.U.set(U); lp.camera.V.set(V); lp.camera.W.set(W)

    # -- SampleRenderer::setProbe, SimplePathtracer.cpp:292-308
    def setProbe(self, probe: ProbeData):
        if not probe.valid:
            raise RuntimeError("Probe Data is not valid")          # Probe.h:104-105
        self.launchParams.probe = self.ctx.probe_upload(probe.data, probe.pdfValuesX, probe.cdfValuesX,
                                                        probe.pdfValuesY, probe.cdfValuesY, probe.offset)

    def setConfig(self, config: native.Config):
        self.config = config
        self.ctx.set_config(config)
        self._feature_buffers()

    # -- SampleRenderer::render, SimplePathtracer.cpp:77-214 (three passes, one submission, then sync)
    def render(self):
        if self.launchParams.frame.size_x == 0:
            return
        self.ctx.render_frame(self.launchParams)
        self.ctx.sync()                                    # CUDA_SYNC_CHECK, :212

    def renderSubframes(self, n, sync=True):
        """n successive render() calls of a progressive loop as one submission (include/fovpt.h fovpt_render_frames)"""
        if self.launchParams.frame.size_x == 0:
            return
        self.ctx.render_frames(self.launchParams, int(n))
        if sync:
            self.ctx.sync()

    def render_async(self):
        if self.launchParams.frame.size_x == 0:
            return
        self.ctx.render_frame(self.launchParams)

    # -- SampleRenderer::downloadPixels, SimplePathtracer.cpp:276-280
    def downloadPixels(self, h_pixels=None):
        lp = self.launchParams
        if h_pixels is None:
            return self.ctx.download(lp.frame.size_x, lp.frame.size_y)
        self.ctx.download_into(h_pixels)
        return h_pixels

    def downloadAccum(self):
        lp = self.launchParams
        return self.ctx.download_accum(lp.frame.size_x, lp.frame.size_y)

    # -- checkpoint / resume of a progressive render (the reference has none): the accumulation buffer and
    #    frame.subframe_index are the only persistent state of the path
    def saveCheckpoint(self, path):
        lp = self.launchParams
        np.savez(path, accum=self.downloadAccum(), subframe_index=np.uint32(lp.frame.subframe_index), size=np.array([lp.frame.size_x, lp.frame.size_y]))

    def loadCheckpoint(self, path):
        lp = self.launchParams
        d = np.load(path)
        if tuple(int(v) for v in d["size"]) != (lp.frame.size_x, lp.frame.size_y):
            raise RuntimeError("checkpoint was written for another frame size")
        self.ctx.upload_accum(d["accum"])
        lp.frame.subframe_index = int(d["subframe_index"])

    def downloadFeatures(self):
        """dict(color, normal, albedo) of (h, w, 4) float32: LaunchParams.frame.{color,normal,albedo}_buffer (config.features = 1)"""
        lp = self.launchParams
        return self.ctx.download_features(lp.frame.size_x, lp.frame.size_y)

    def stats(self):
        return self.ctx.stats()

    def close(self):
        self.ctx.close()
