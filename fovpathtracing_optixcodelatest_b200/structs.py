"""ctypes mirrors of the reference's launch-contract structs.

These are the byte layouts that cross the C ABI (``include/fovpt.h``) as opaque blobs:

* ``LaunchParams``  248 B  -- reference ``PT_sv5_/LaunchParams.h:49-91``
* ``Material``      104 B  -- reference ``PT_sv5_/Material.h:11-70`` (defaults: ctor ``:13-38``)
* ``Probe``          64 B  -- reference ``PT_sv5_/Probe.cuh:6-21``

Offsets are the ones g++ measured on the unmodified reference headers;
they are asserted at import time and again, on the C side, by static_asserts in
``csrc/fovpt_layout.h``.
"""
import ctypes as C


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def set(self, v):
        self.x, self.y, self.z = float(v[0]), float(v[1]), float(v[2])

    def tolist(self):
        return [self.x, self.y, self.z]


class Probe(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("data", C.c_uint64),
        ("offset", Float3), ("_pad0", C.c_uint32),
        ("pdfValuesX", C.c_uint64), ("cdfValuesX", C.c_uint64),
        ("pdfValuesY", C.c_uint64), ("cdfValuesY", C.c_uint64),
    ]


class Frame(C.Structure):
    _fields_ = [
        ("accum_buffer", C.c_uint64), ("frame_buffer", C.c_uint64),
        ("color_buffer", C.c_uint64), ("normal_buffer", C.c_uint64), ("albedo_buffer", C.c_uint64),
        ("size_x", C.c_int32), ("size_y", C.c_int32),
        ("subframe_index", C.c_uint32),
        ("factor_x", C.c_uint32), ("factor_y", C.c_uint32), ("factor_z", C.c_uint32),
        ("fillSize", C.c_int32), ("_pad0", C.c_uint32),
        ("c_x", C.c_uint32), ("c_y", C.c_uint32),
        ("r_inner", C.c_float), ("r_outer", C.c_float),
        ("offset_x", C.c_uint32), ("offset_y", C.c_uint32),
        ("redraw", C.c_uint32), ("_pad1", C.c_uint32),
    ]


class Camera(C.Structure):
    _fields_ = [("eye", Float3), ("U", Float3), ("V", Float3), ("W", Float3)]


class LaunchParams(C.Structure):
    _fields_ = [
        ("frame", Frame),
        ("camera", Camera),
        ("samples_per_launch", C.c_uint32), ("_pad0", C.c_uint32),
        ("traversable", C.c_uint64),
        ("probe", Probe),
        ("viewport_x", C.c_int32), ("viewport_y", C.c_int32),
        ("white", C.c_float), ("_pad1", C.c_uint32),
    ]


class Material(C.Structure):
    _fields_ = [
        ("emission", Float3), ("color", Float3), ("absorption", Float3),
        ("eta", C.c_float), ("metallic", C.c_float), ("subsurface", C.c_float),
        ("specular", C.c_float), ("roughness", C.c_float), ("specularTint", C.c_float),
        ("anisotropic", C.c_float), ("sheen", C.c_float), ("sheenTint", C.c_float),
        ("clearcoat", C.c_float), ("clearcoatGloss", C.c_float), ("transmission", C.c_float),
        ("bump", C.c_float), ("bumpTile", Float3), ("flags", C.c_int32),
    ]

    @classmethod
    def default(cls):
        """The reference constructor's values (Material.h:13-38)."""
        m = cls()
        m.color.set((1.0, 0.0, 0.0)); m.emission.set((1.0, 1.0, 1.0)); m.absorption.set((1.0, 1.0, 1.0))
        m.eta = 1.4; m.metallic = 0.5; m.subsurface = 0.0; m.specular = 1.0; m.roughness = 1.0
        m.specularTint = 1.0; m.anisotropic = 0.0; m.sheen = 0.0; m.sheenTint = 0.0
        m.clearcoat = 0.0; m.clearcoatGloss = 1.0; m.transmission = 0.4; m.bump = 0.0
        m.bumpTile.set((1.0, 1.0, 1.0)); m.flags = 0
        return m

    @classmethod
    def diffuse(cls, color, emission=(0.0, 0.0, 0.0)):
        """'Lambertian' expressed through Disney parameters."""
        m = cls.default()
        m.color.set(color); m.emission.set(emission)
        m.transmission = 0.0; m.metallic = 0.0; m.specular = 0.0; m.subsurface = 0.0
        m.clearcoat = 0.0; m.roughness = 1.0
        return m

    def copy(self):
        m = Material()
        C.memmove(C.byref(m), C.byref(self), C.sizeof(Material))
        return m


MATERIAL_FLAG_SHADOW_CATCHER = 1

_EXPECT = {
    (LaunchParams, None): 248, (Material, None): 104, (Probe, None): 64, (Frame, None): 104,
}
assert C.sizeof(LaunchParams) == 248, C.sizeof(LaunchParams)
assert C.sizeof(Material) == 104
assert C.sizeof(Probe) == 64
assert Frame.size_x.offset == 40 and Frame.subframe_index.offset == 48 and Frame.factor_x.offset == 52
assert Frame.fillSize.offset == 64 and Frame.c_x.offset == 72 and Frame.r_inner.offset == 80
assert Frame.r_outer.offset == 84 and Frame.offset_x.offset == 88 and Frame.redraw.offset == 96
assert LaunchParams.camera.offset == 104 and LaunchParams.samples_per_launch.offset == 152
assert LaunchParams.traversable.offset == 160 and LaunchParams.probe.offset == 168
assert LaunchParams.viewport_x.offset == 232 and LaunchParams.white.offset == 240
assert Probe.data.offset == 8 and Probe.offset.offset == 16 and Probe.pdfValuesX.offset == 32
assert Probe.cdfValuesY.offset == 56
assert Material.color.offset == 12 and Material.eta.offset == 36 and Material.transmission.offset == 80
assert Material.bumpTile.offset == 88 and Material.flags.offset == 100
