"""The C++ host layer (fovpathtracing_optixcodelatest_b200/host: SampleRenderer, loadOBJ / addBox, ProbeData,
sutil::Camera -- the drop-in for the reference's C++ classes) driven through its flat test hooks (host/capi.cpp).

CPU part: OBJ ingest against the fixture the REFERENCE's own Model.cpp produced (tests/golden/loader.npz),
BuildCDF / Camera against the reference-generated fixtures, struct sizes.  GPU part: the C++ SampleRenderer
call sequence of the reference's main.cpp, compared with the CPU oracle and with the python binding."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from fovpathtracing_optixcodelatest_b200 import scenes
from fovpathtracing_optixcodelatest_b200.structs import LaunchParams, Material

HOST_LIB = os.path.join(ROOT, "fovpathtracing_optixcodelatest_b200", "libfovpt_host.so")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def H():
    assert os.path.exists(HOST_LIB), "libfovpt_host.so missing: run __graft_entry__.build()"
    L = C.CDLL(HOST_LIB)
    vp = C.c_void_p
    L.fovpth_last_error.restype = C.c_char_p
    for n in ("fovpth_model_new", "fovpth_load_obj", "fovpth_renderer_new", "fovpth_launch_params"):
        getattr(L, n).restype = vp
    L.fovpth_load_obj.argtypes = [C.c_char_p]
    L.fovpth_model_free.argtypes = [vp]
    L.fovpth_add_box.argtypes = [vp, vp, vp, vp]
    L.fovpth_add_mesh.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int]
    L.fovpth_add_texture.argtypes = [vp, vp, C.c_int, C.c_int]
    L.fovpth_mesh_count.argtypes = [vp]
    L.fovpth_mesh_info.argtypes = [vp, C.c_int] + [vp] * 4
    L.fovpth_mesh_get.argtypes = [vp, C.c_int] + [vp] * 4
    L.fovpth_texture_count.argtypes = [vp]
    L.fovpth_texture_get.argtypes = [vp, C.c_int, vp, vp, vp]
    L.fovpth_decode_image.argtypes = [C.c_char_p, vp, vp, vp, C.c_size_t]
    L.fovpth_partition.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.fovpth_build_cdf.argtypes = [C.c_int, C.c_int] + [vp] * 5
    L.fovpth_camera_uvw.argtypes = [vp, vp, vp, C.c_float, C.c_float, vp]
    L.fovpth_renderer_new.argtypes = [vp, C.c_int, vp]
    L.fovpth_renderer_free.argtypes = [vp]
    L.fovpth_resize.argtypes = [vp, C.c_int, C.c_int]
    L.fovpth_set_camera.argtypes = [vp, vp, vp, vp, C.c_float]
    L.fovpth_set_probe.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
    L.fovpth_launch_params.argtypes = [vp]
    L.fovpth_render.argtypes = [vp]
    L.fovpth_render_to_output_buffer.argtypes = [vp, C.c_int, vp]
    L.fovpth_download_pixels.argtypes = [vp, vp]
    L.fovpth_download_accum.argtypes = [vp, vp]
    L.fovpth_save_checkpoint.argtypes = [vp, C.c_char_p]
    L.fovpth_load_checkpoint.argtypes = [vp, C.c_char_p]
    L.fovpth_render_subframes.argtypes = [vp, C.c_int]
    return L


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def model_meshes(H, m):
    out = []
    for i in range(H.fovpth_mesh_count(m)):
        nv, nt, ntc, tex = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        H.fovpth_mesh_info(m, i, C.byref(nv), C.byref(nt), C.byref(ntc), C.byref(tex))
        v = np.zeros((nv.value, 3), np.float32); idx = np.zeros((nt.value, 3), np.uint32); mat = Material()
        tc = np.zeros((ntc.value, 2), np.float32)
        H.fovpth_mesh_get(m, i, p(v), p(idx), p(tc) if ntc.value else None, C.byref(mat))
        out.append(dict(vtx=v, idx=idx, texcoord=tc if ntc.value else None, material=mat, texture_id=tex.value))
    return out


# ------------------------------------------------------------------------------------------- CPU
def test_struct_sizes_match_reference(H):
    assert [H.fovpth_sizeof(i) for i in range(3)] == [248, 104, 64]        # LaunchParams, Material, Probe


def test_loadobj_matches_reference_loader(H):
    """assets/cornell.obj through the C++ loadOBJ == what the reference's Model.cpp (tinyobjloader) made of it"""
    g = np.load(os.path.join(GOLDEN, "loader.npz"))
    m = H.fovpth_load_obj(os.path.join(ROOT, "assets", "cornell.obj").encode())
    assert m, H.fovpth_last_error()
    ms = model_meshes(H, m)
    assert len(ms) == int(g["n_meshes"])
    tris = [x["vtx"][x["idx"].astype(np.int64)].reshape(-1, 9) for x in ms]     # resolved triangles: vertex order is unspecified (A.9 ii)
    assert [len(t) for t in tris] == list(g["counts"])
    assert (bits(np.concatenate(tris)) == bits(g["tris"])).all()
    assert (bits(np.array([x["material"].color.tolist() for x in ms])) == bits(g["colors"])).all()
    assert (bits(np.array([x["material"].emission.tolist() for x in ms])) == bits(g["emission"])).all()
    assert (np.stack([np.frombuffer(bytes(x["material"]), np.uint8) for x in ms]) == g["materials"]).all()     # every other Disney field = ctor default
    assert [x["texture_id"] for x in ms] == list(g["texture_ids"])
    for x in ms:
        assert x["idx"].max() < len(x["vtx"])                                # no dangling indices
    H.fovpth_model_free(m)


def test_loadobj_errors_and_variants(H, tmp_path):
    assert not H.fovpth_load_obj(str(tmp_path / "missing.obj").encode())
    assert b"Could not read OBJ model" in H.fovpth_last_error()
    # quads, negative indices, v/vt/vn corners, two materials in one shape, a PPM texture, a face without usemtl
    (tmp_path / "t.mtl").write_text("newmtl a\nKd 0.1 0.2 0.3\nKe 1 2 3\nmap_Kd tex.ppm\nnewmtl b\nKd 0.5 0.5 0.5\n")
    with open(tmp_path / "tex.ppm", "wb") as f:
        f.write(b"P6\n2 2\n255\n" + bytes([255, 0, 0, 0, 255, 0, 0, 0, 255, 9, 8, 7]))
    (tmp_path / "t.obj").write_text(
        "mtllib t.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
        "o bare\nf 1 2 3\no quad\nusemtl a\nf 1/1/1 2/2/1 3/3/1 4/4/1\nusemtl b\nf -4 -3 -2\n")
    m = H.fovpth_load_obj(str(tmp_path / "t.obj").encode())
    ms = model_meshes(H, m)
    assert len(ms) == 3 and [len(x["idx"]) for x in ms] == [1, 2, 1]
    ms = [ms[1], ms[2], ms[0]]                                                # -> quad/a, quad/b, bare
    assert ms[0]["material"].color.tolist() == pytest.approx([0.1, 0.2, 0.3]) and ms[0]["material"].emission.tolist() == [1.0, 2.0, 3.0]
    assert ms[0]["texture_id"] == 0 and ms[1]["texture_id"] == -1 and ms[2]["texture_id"] == -1
    assert ms[0]["texcoord"] is not None and len(ms[0]["vtx"]) == 4           # fan-split quad shares its 4 de-duplicated vertices
    assert (ms[1]["vtx"][ms[1]["idx"][0].astype(int)] == np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0]], np.float32)).all()
    assert ms[2]["material"].color.tolist() == [1.0, 0.0, 0.0]                # no usemtl -> default Material
    for x in ms:
        assert x["idx"].max() < len(x["vtx"])                                 # the reference's per-shape map would dangle here (A.9 i)
    w, h = C.c_int(), C.c_int(); px = np.zeros(4, np.uint32)
    assert H.fovpth_texture_count(m) == 1
    H.fovpth_texture_get(m, 0, C.byref(w), C.byref(h), p(px))
    assert (w.value, h.value) == (2, 2) and px[0] == (0 | (0 << 8) | (255 << 16) | (255 << 24))     # y-mirrored: bottom-left first
    H.fovpth_model_free(m)


def test_texture_decoders_match_reference_loader(H):
    """loadOBJ on an OBJ whose 57 materials name PNG (every colour type, 1-16 bit, tRNS, Adam7), TGA (true colour, colour
    mapped, grey, RLE, both origins), JPEG (baseline / progressive, 4:4:4 4:2:2 4:2:0 4:1:1, restart markers, grey), BMP and PPM
    files: every texture byte-identical to what the REFERENCE's loader (its Model.cpp + vendored decoder, through oracle/_ref)
    made of the same files (tests/golden/textures.npz, written by tests/golden/make_textures.py); undecodable files are
    refused by both (PT_sv5_/Model.cpp:87-136)."""
    g = np.load(os.path.join(GOLDEN, "textures.npz"))
    m = H.fovpth_load_obj(os.path.join(GOLDEN, "textures", "textures.obj").encode())
    assert m, H.fovpth_last_error()
    ms = model_meshes(H, m)
    names = [str(n) for n in g["names"]]
    assert len(ms) == len(names)
    assert [x["texture_id"] for x in ms] == list(g["texture_ids"])             # same files accepted, same ids
    kinds = set()
    for n, x in zip(names, ms):
        if x["texture_id"] < 0:
            assert n in ("broken.png", "not_an_image.tga")
            continue
        w, h = C.c_int(), C.c_int()
        H.fovpth_texture_get(m, x["texture_id"], C.byref(w), C.byref(h), None)
        want = g["px_" + n]
        assert (h.value, w.value) == want.shape, n
        px = np.zeros(want.shape, np.uint32)
        H.fovpth_texture_get(m, x["texture_id"], C.byref(w), C.byref(h), p(px))
        bad = int((px != want).sum())
        assert bad == 0, f"{n}: {bad} of {px.size} texels differ from the reference loader"
        kinds.add(n.rsplit(".", 1)[1])
    assert kinds == {"png", "tga", "jpg", "bmp", "ppm"}
    H.fovpth_model_free(m)
    # the decoder hook reports why a file is refused
    w, h = C.c_int(), C.c_int()
    assert H.fovpth_decode_image(os.path.join(GOLDEN, "textures", "broken.png").encode(), C.byref(w), C.byref(h), None, 0) != 0
    assert b"PNG" in H.fovpth_last_error() or b"truncated" in H.fovpth_last_error()


def test_texture_decoders_survive_damaged_files(H, tmp_path):
    """textures come from disk: every fixture truncated at a dozen places and with random bytes overwritten must be either
    refused with a message or decoded to an image of the size the header states -- never crash, hang or allocate without bound"""
    rng = np.random.default_rng(7)
    tex = os.path.join(GOLDEN, "textures")
    files = sorted(f for f in os.listdir(tex) if f.rsplit(".", 1)[-1] in ("png", "tga", "jpg", "bmp", "ppm"))
    assert len(files) >= 50
    path = str(tmp_path / "damaged.bin").encode()
    w, h = C.c_int(), C.c_int()
    out = np.zeros(1 << 18, np.uint32)
    refused = decoded = 0
    for f in files:
        data = open(os.path.join(tex, f), "rb").read()
        variants = [data[:k] for k in sorted(set(int(x) for x in np.linspace(0, len(data) - 1, 8)))]
        for _ in range(24):
            d = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
            variants.append(bytes(d))
        big = bytearray(data)                                         # blow the size fields up
        if f.endswith(".png"): big[16:24] = b"\x7f\xff\xff\xff\x7f\xff\xff\xff"
        elif f.endswith(".tga"): big[12:16] = b"\xff\xff\xff\xff"
        elif f.endswith(".bmp"): big[18:26] = b"\xff\xff\xff\x7f\xff\xff\xff\x7f"
        variants.append(bytes(big))
        for v in variants:
            open(path, "wb").write(v)
            rc = H.fovpth_decode_image(path, C.byref(w), C.byref(h), p(out), out.size)
            if rc != 0:
                assert H.fovpth_last_error(), f
                refused += 1
            else:
                assert 0 < w.value * h.value <= (1 << 28), f
                decoded += 1
    assert refused > 300 and decoded > 300                            # both outcomes occur (a flipped pixel byte still decodes)


def test_cpp_tile_partition_is_the_python_partition(H):
    """host/MultiGpu.cpp (what a C++ application links) against partition.py (what bench.py uses): byte-identical owner tables
    and equal per-rank loads, for the density-weighted and the interleaved assignment, several frame sizes / gaze points / ranks"""
    from fovpathtracing_optixcodelatest_b200 import native, partition
    cases = [(1920, 1080, (960, 540), 74, 241, 32), (2160, 2160, (1680, 480), 148, 482, 32), (1000, 333, (3, 330), 20, 90, 16), (64, 48, (32, 24), 4, 12, 32)]
    for (w, h, gaze, ri, ro, tile) in cases:
        for n in (1, 2, 3, 4, 8):
            cfg = native.Config(); cfg.foveated = 1; cfg.inner_radius, cfg.outer_radius = ri, ro
            cfg.spp_periphery, cfg.spp_mid, cfg.spp_fovea, cfg.spp_uniform = 8, 16, 32, 4
            want, wloads = partition.partition(w, h, gaze, ri, ro, n, tile=tile)
            own = np.zeros(want.shape, np.uint8); loads = np.zeros(max(1, n), np.float64)
            rc = H.fovpth_partition(w, h, gaze[0], gaze[1], C.byref(cfg), n, tile, 0, p(own), p(loads))
            assert rc == (want.shape[1] | (want.shape[0] << 16)), H.fovpth_last_error()
            assert (own == want).all() and (loads == wloads).all(), (w, h, n)
            want = partition.interleaved(w, h, n, 16)
            own = np.zeros(want.shape, np.uint8)
            assert H.fovpth_partition(w, h, gaze[0], gaze[1], C.byref(cfg), n, 16, 1, p(own), p(loads)) == (want.shape[1] | (want.shape[0] << 16))
            assert (own == want).all()
            assert np.allclose(loads, partition.loads_of(want, w, h, gaze, ri, ro, n, 16))
    cfg = native.Config(); cfg.foveated = 0; cfg.spp_uniform = 4; cfg.spp_periphery = cfg.spp_mid = cfg.spp_fovea = 1
    want, _ = partition.partition(640, 360, (0, 0), 0, 0, 4, foveated=False)
    own = np.zeros(want.shape, np.uint8)
    assert H.fovpth_partition(640, 360, 0, 0, C.byref(cfg), 4, 32, 0, p(own), None) > 0 and (own == want).all()
    assert H.fovpth_partition(640, 360, 0, 0, C.byref(cfg), 300, 32, 0, p(own), None) == -1       # one byte per tile: <= 256 ranks


def test_addbox_matches_reference(H, ref):
    mat = Material.diffuse((.2, .4, .6))
    pos = np.array([1.0, 2.0, -3.0], np.float32); ext = np.array([0.5, 1.5, 2.5], np.float32)
    m = H.fovpth_model_new()
    assert H.fovpth_add_box(m, C.byref(mat), p(pos), p(ext)) == 0
    mine = model_meshes(H, m)[0]
    R = ref.Oracle("reference"); R.add_box(mat, pos, ext)
    theirs = R.meshes()[0]; R.close()
    a = mine["vtx"][mine["idx"].astype(np.int64)]; b = theirs["vtx"][theirs["idx"].astype(np.int64)]
    assert a.shape == b.shape == (12, 3, 3) and (bits(a) == bits(b)).all()
    H.fovpth_model_free(m)


def test_buildcdf_and_camera_match_reference_fixtures(H):
    g = np.load(os.path.join(GOLDEN, "probe.npz"))
    img = np.ascontiguousarray(scenes.sky_probe(64, 32, seed=4), np.float32)
    pdfX = np.zeros((32, 64), np.float32); cdfX = np.zeros((32, 64), np.float32); pdfY = np.zeros(32, np.float32); cdfY = np.zeros(32, np.float32)
    H.fovpth_build_cdf(64, 32, p(img), p(pdfX), p(cdfX), p(pdfY), p(cdfY))
    for a, k in ((pdfX, "pdfX"), (cdfX, "cdfX"), (pdfY, "pdfY"), (cdfY, "cdfY")):
        assert (bits(a) == bits(g[k])).all(), k
    r = np.load(os.path.join(GOLDEN, "rng.npz"))
    out = np.zeros(9, np.float32)
    eye = np.array([-1293.07, 154.681, 0], np.float32); look = np.array([200, 0, 0], np.float32); up = np.array([0, 1, 0], np.float32)
    H.fovpth_camera_uvw(p(eye), p(look), p(up), 45.0, 1920 / 1080, p(out))
    assert (bits(out) == bits(r["sponza_uvw"])).all()


# ------------------------------------------------------------------------------------------- GPU
def cpp_model(H, scene):
    m = H.fovpth_model_new()
    for t in scene.textures:
        a = np.ascontiguousarray(t, np.uint32); H.fovpth_add_texture(m, p(a), a.shape[1], a.shape[0])
    for k in scene.meshes:
        v = np.ascontiguousarray(k.vtx, np.float32); i = np.ascontiguousarray(k.idx, np.uint32)
        tc = None if k.texcoord is None else np.ascontiguousarray(k.texcoord, np.float32)
        H.fovpth_add_mesh(m, p(v), len(v), p(i), len(i), None if tc is None else p(tc), C.byref(k.material), k.texture_id)
    return m


@pytest.mark.gpu
def test_cpp_samplerenderer_matches_oracle_and_python_binding(H, port):
    from fovpathtracing_optixcodelatest_b200 import native
    from fovpathtracing_optixcodelatest_b200.renderer import ProbeData, SampleRenderer
    w, h, ri, ro = 320, 180, 24, 80
    sc, cam = scenes.cornell_box(), scenes.CORNELL_CAMERA
    probe = np.ascontiguousarray(scenes.constant_probe(16, 8, 2.5), np.float32)
    cfg = native.default_config(); cfg.inner_radius, cfg.outer_radius = ri, ro

    m = cpp_model(H, sc)
    r = H.fovpth_renderer_new(m, 0, C.byref(cfg))
    assert r, H.fovpth_last_error()
    assert H.fovpth_render(r) == 0                                            # no-op before resize (SimplePathtracer.cpp:81-82)
    eye = np.array(cam.eye, np.float32); look = np.array(cam.lookat, np.float32); up = np.array(cam.up, np.float32)
    assert H.fovpth_set_camera(r, p(eye), p(look), p(up), cam.fovy) == -1     # before resize: frame.size is 0
    assert H.fovpth_resize(r, w, h) == 0
    assert H.fovpth_set_camera(r, p(eye), p(look), p(up), cam.fovy) == 0
    assert H.fovpth_set_probe(r, 16, 8, p(probe), 0) == -1 and b"Probe Data is not valid" in H.fovpth_last_error()
    assert H.fovpth_set_probe(r, 16, 8, p(probe), 1) == 0
    lp = LaunchParams.from_address(H.fovpth_launch_params(r))                 # the PUBLIC member, poked like main.cpp:366-367 does
    lp.frame.c_x, lp.frame.c_y = w // 2, h // 2
    assert H.fovpth_render(r) == 0, H.fovpth_last_error()
    assert lp.frame.subframe_index == 1
    px = np.zeros((h, w), np.uint32); acc = np.zeros((h, w, 4), np.float32)
    assert H.fovpth_download_pixels(r, p(px)) == 0 and H.fovpth_download_accum(r, p(acc)) == 0

    # (1) the CPU oracle
    o = port.Oracle("port"); o.add_scene(sc); o.commit(); o.set_probe(probe)
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, w / h)
    olp, oacc, ofr = o.make_params(w, h, cam.eye, U, V, W)
    o.render(olp, port.FovCfg.sv5(inner_radius=ri, outer_radius=ro), port.Opts.sv5())
    a = oacc[..., :3].astype(np.float64); g = acc[..., :3].astype(np.float64)
    assert ((a - g) ** 2).mean() / a.mean() ** 2 < 1e-6 and (px == ofr).mean() > 0.99
    assert (bits(np.array(lp.camera.U.tolist())) == bits(U)).all()
    o.close()
    # (2) the python binding over the same C ABI: bit-identical
    pr = SampleRenderer(sc, 0, cfg); pr.resize((w, h)); pr.setCamera(cam); pr.setProbe(ProbeData(probe).BuildCDF())
    pr.launchParams.frame.c_x, pr.launchParams.frame.c_y = w // 2, h // 2
    pr.render()
    assert (pr.downloadPixels() == px).all() and (bits(pr.downloadAccum()) == bits(acc)).all()
    pr.close()
    # (3) render(CUDAOutputBuffer&): same frame into an external display buffer, every flavour of sutil/CUDAOutputBuffer.h:43-49
    written = acc[..., 3] == 1.0          # the few pixels no pass writes (ring-boundary holes) keep the target's old content, as in the reference
    for kind in (0, 2, 3):                # CUDA_DEVICE, ZERO_COPY (kernels store into mapped host memory), CUDA_P2P
        lp.frame.subframe_index = 0
        out = np.zeros((h, w), np.uint32)
        assert H.fovpth_render_to_output_buffer(r, kind, p(out)) == 0, H.fovpth_last_error()
        assert (out[written] == px[written]).all() and written.mean() > 0.999, kind
    assert H.fovpth_render_to_output_buffer(r, 1, p(out)) != 0 and b"GL interop" in H.fovpth_last_error()   # no GL context here
    # after a borrowed target the renderer falls back to its own frame buffer
    lp.frame.subframe_index = 0
    assert H.fovpth_render(r) == 0
    px2 = np.zeros((h, w), np.uint32); H.fovpth_download_pixels(r, p(px2))
    assert (px2 == px).all()
    # (4) saveCheckpoint / loadCheckpoint: a second C++ renderer continues the progressive run of the first, bit for bit
    import tempfile
    cfg2 = native.default_config(); cfg2.inner_radius, cfg2.outer_radius = ri, ro; cfg2.accumulate = 1; cfg2.seed_all_passes_with_subframe = 1
    def cpp_renderer():
        q = H.fovpth_renderer_new(m, 0, C.byref(cfg2))
        assert H.fovpth_resize(q, w, h) == 0 and H.fovpth_set_camera(q, p(eye), p(look), p(up), cam.fovy) == 0 and H.fovpth_set_probe(q, 16, 8, p(probe), 1) == 0
        qlp = LaunchParams.from_address(H.fovpth_launch_params(q)); qlp.frame.c_x, qlp.frame.c_y = w // 2, h // 2
        return q, qlp
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "run.ck").encode()
        r1, lp1 = cpp_renderer()
        assert H.fovpth_render_subframes(r1, 3) == 0 and H.fovpth_save_checkpoint(r1, ck) == 0 and H.fovpth_render_subframes(r1, 2) == 0
        a1 = np.zeros((h, w, 4), np.float32); H.fovpth_download_accum(r1, p(a1))
        r2, lp2 = cpp_renderer()
        assert H.fovpth_load_checkpoint(r2, ck) == 0, H.fovpth_last_error()
        assert lp2.frame.subframe_index == 3
        assert H.fovpth_render_subframes(r2, 2) == 0
        a2 = np.zeros((h, w, 4), np.float32); H.fovpth_download_accum(r2, p(a2))
        assert (bits(a1) == bits(a2)).all() and lp2.frame.subframe_index == 5
        assert H.fovpth_load_checkpoint(r2, os.path.join(td, "missing.ck").encode()) != 0 and b"checkpoint" in H.fovpth_last_error()
        H.fovpth_renderer_free(r1); H.fovpth_renderer_free(r2)
    H.fovpth_renderer_free(r); H.fovpth_model_free(m)


@pytest.mark.gpu
def test_fovpt_demo_renders_writes_and_rereads_a_ppm(H, tmp_path):
    """host/main.cpp (the headless stand-in for the reference's main.cpp:297-481: loadOBJ -> SampleRenderer -> resize ->
    setCamera -> setProbe -> per frame gaze + render() -> downloadPixels -> image file) run as a program: the PPM it writes,
    re-read through the texture decoder, is the frame the python binding renders with the same settings"""
    import subprocess
    from fovpathtracing_optixcodelatest_b200 import native
    from fovpathtracing_optixcodelatest_b200.renderer import ProbeData, SampleRenderer
    demo = os.path.join(ROOT, "fovpathtracing_optixcodelatest_b200", "host", "fovpt_demo")
    assert os.path.exists(demo), "fovpt_demo missing: run __graft_entry__.build()"
    w, h, frames, gx, gy = 320, 180, 3, 200, 70
    out = tmp_path / "demo.ppm"
    run = subprocess.run([demo, os.path.join(ROOT, "assets", "cornell.obj"), str(w), str(h), str(frames), str(gx), str(gy), str(out)],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    assert "Msamples/s" in run.stdout
    ww, hh = C.c_int(), C.c_int(); px = np.zeros((h, w), np.uint32)
    assert H.fovpth_decode_image(str(out).encode(), C.byref(ww), C.byref(hh), p(px), px.size) == 0, H.fovpth_last_error()
    assert (ww.value, hh.value) == (w, h)
    # the same frame through the python binding (main.cpp's camera: bounding box centre, 1.4 extents back along +z)
    m = H.fovpth_load_obj(os.path.join(ROOT, "assets", "cornell.obj").encode())      # the scene as the OBJ file holds it (rounded decimals)
    sc = scenes.Scene()
    for x in model_meshes(H, m):
        sc.add(x["vtx"], x["idx"], x["material"], x["texcoord"], x["texture_id"])
    H.fovpth_model_free(m)
    v = np.concatenate([m.vtx for m in sc.meshes]); lo, hi = v.min(0), v.max(0)
    centre = (0.5 * (lo + hi)).astype(np.float32); extent = np.float32((hi - lo).max())
    cam = scenes.CameraSpec(eye=(centre[0], centre[1], np.float32(centre[2] + np.float32(1.4) * extent)), lookat=tuple(centre), up=(0.0, 1.0, 0.0), fovy=45.0)
    cfg = native.default_config()
    cfg.tmin = np.float32(1e-4) * extent; cfg.accumulate = 1; cfg.seed_all_passes_with_subframe = 1; cfg.exposure = 1.0
    r = SampleRenderer(sc, 0, cfg); r.resize((w, h)); r.setCamera(cam)
    r.setProbe(ProbeData(scenes.constant_probe(64, 32, 2.5)).BuildCDF())
    for _ in range(frames):
        r.launchParams.frame.c_x, r.launchParams.frame.c_y = gx, gy
        r.render()
    want = r.downloadPixels()[::-1] | np.uint32(0xff000000)                # the file stores the top row first, without alpha
    acc = r.downloadAccum()[::-1]
    r.close()
    written = acc[..., 3] == 1.0
    same = (px[written] == want[written]).mean()
    assert same > 0.999, same
    assert len(np.unique(px)) > 50                                          # an image, not a constant


@pytest.mark.gpu
def test_fovpt_multi_cpp_processes_write_the_single_gpu_frame(tmp_path):
    """host/multi_main.cpp: a C++ multi-GPU job -- one process per rank, tile partition from host/MultiGpu.cpp, the frame
    gathered over peer memory into rank 0 (SharedFrame) -- writes byte for byte the PPM the single-process fovpt_demo writes.
    With one GPU in the box the three ranks share it (the peer path is CUDA IPC either way)."""
    import subprocess
    host = os.path.join(ROOT, "fovpathtracing_optixcodelatest_b200", "host")
    obj = os.path.join(ROOT, "assets", "cornell.obj")
    args = ["320", "180", "3", "200", "70"]
    one = tmp_path / "one.ppm"; many = tmp_path / "many.ppm"
    a = subprocess.run([os.path.join(host, "fovpt_demo"), obj] + args + [str(one)], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0, a.stderr
    b = subprocess.run([os.path.join(host, "fovpt_multi"), obj] + args + ["3", str(many)], capture_output=True, text=True, timeout=300)
    assert b.returncode == 0, b.stderr + b.stdout
    assert "rank 0 of 3" in b.stdout
    assert one.read_bytes() == many.read_bytes()
