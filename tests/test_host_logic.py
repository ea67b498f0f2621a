"""CPU: host-side logic of the product (no compute calls): ABI surface, struct layouts, python mirrors
of ProbeData::BuildCDF / Camera::UVWFrame, the three-pass sample count, the tile partition."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from fovpathtracing_optixcodelatest_b200 import native, partition, scenes, structs, workloads
from fovpathtracing_optixcodelatest_b200.renderer import ProbeData


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fovpt.h")).read()
    declared = set(re.findall(r"\b(fovpt_[a-z_]+)\s*\(", hdr))
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    L = C.CDLL(native.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    native.load_library()


def _build_c_abi_smoke(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    pkg = os.path.join(ROOT, "fovpathtracing_optixcodelatest_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                    "-o", exe, "-L", pkg, "-lfovpt", f"-Wl,-rpath,{pkg}"], check=True)
    return exe


def test_header_is_plain_c_and_library_links_from_c(tmp_path):
    """include/fovpt.h compiles as pedantic C99 and a C program drives libfovpt.so (here, without a GPU: clean failure)"""
    import subprocess
    out = subprocess.run([_build_c_abi_smoke(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_bvh_top_level_sah_builder(tmp_path):
    """the BVH builder's host stage (csrc/top_sah.h: binned SAH over the clusters the device-side PLOC phase leaves) as a CPU
    program: every cluster a leaf exactly once, node boxes / counts consistent, result independent of thread timing, separated
    groups split first, coincident boxes balanced (tests/c/top_sah_test.cpp)"""
    import subprocess
    exe = str(tmp_path / "top_sah_test")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-pthread", "-I", os.path.join(ROOT, "fovpathtracing_optixcodelatest_b200", "csrc"),
                    os.path.join(ROOT, "tests", "c", "top_sah_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "top_sah ok" in out.stdout, out.stdout + out.stderr


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        native.Context(0)


def test_default_config_is_sv5():
    k = native.default_config()
    assert (k.max_depth, k.inner_radius, k.outer_radius, k.spp_periphery, k.spp_mid, k.spp_fovea, k.spp_uniform) == (4, 74, 241, 8, 16, 32, 4)
    assert abs(k.tmin - 0.01) < 1e-9 and k.tmax == np.float32(1e16) and k.exposure == 16.0 and k.accumulate == 0 and k.foveated == 1


def test_struct_layouts():
    assert C.sizeof(structs.LaunchParams) == 248 and C.sizeof(structs.Material) == 104 and C.sizeof(structs.Probe) == 64
    m = structs.Material.default()
    assert (m.metallic, m.specular, m.roughness, m.specularTint, m.eta) == (0.5, 1.0, 1.0, 1.0, np.float32(1.4))
    assert abs(m.transmission - 0.4) < 1e-7 and m.color.tolist() == [1.0, 0.0, 0.0] and m.emission.tolist() == [1.0, 1.0, 1.0]


def test_oracle_libs_export_api(port):
    hdr = open(os.path.join(ROOT, "oracle", "orc_api.h")).read()
    declared = set(re.findall(r"\b(orc_[a-z0-9_]+)\s*\(", hdr))
    L = C.CDLL(port.PATHS["port"])
    for name in declared:
        assert hasattr(L, name), name


def test_buildcdf_and_camera_match_oracle(port):
    img = scenes.sky_probe(96, 48, seed=7)
    pd = ProbeData(img).BuildCDF()
    o = port.Oracle("port"); o.set_probe(img)
    for a, b in zip(o.get_probe(), (pd.pdfValuesX, pd.cdfValuesX, pd.pdfValuesY, pd.cdfValuesY)):
        assert (a.view(np.uint32) == b.view(np.uint32)).all()
    assert np.all(np.diff(pd.cdfValuesX, axis=1) >= 0) and np.all(np.diff(pd.cdfValuesY) >= 0)
    assert pd.cdfValuesY[-1] == 1.0 and np.allclose(pd.cdfValuesX[:, -1], 1.0, atol=1e-6)
    for cam, asp in ((scenes.CORNELL_CAMERA, 960 / 540), (scenes.SPONZA_CAMERA, 1920 / 1080), (scenes.CORNELL_CAMERA, 1.0)):
        for a, b in zip(o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, asp), scenes.camera_uvw(cam, asp)):
            assert (a.view(np.uint32) == b.view(np.uint32)).all()
    o.close()
    with pytest.raises(RuntimeError):
        from fovpathtracing_optixcodelatest_b200.renderer import SampleRenderer
        SampleRenderer.setProbe(object.__new__(SampleRenderer), ProbeData(img))     # BuildCDF not called: "Probe Data is not valid"


def ring_samples(w, h, c, ri, ro, spp=(8, 16, 32)):
    return partition.frame_sample_count(w, h, c, ri, ro, spp)


def test_three_pass_sample_counts():
    assert ring_samples(1920, 1080, (960, 540), 74, 241) == 2183384
    assert ring_samples(960, 540, (480, 270), 74, 241) == 1405984
    assert ring_samples(1920, 1080, (960, 540), 148, 482) == 5571224


def test_workloads_are_deterministic():
    a, b = scenes.atrium(detail=0.2), scenes.atrium(detail=0.2)
    assert a.n_tris == b.n_tris and all((x.vtx == y.vtx).all() and (x.idx == y.idx).all() for x, y in zip(a.meshes, b.meshes))
    assert 255000 < scenes.atrium().n_tris < 270000
    assert scenes.cornell_box().n_tris == 34
    assert (scenes.sky_probe(64, 32, 2) == scenes.sky_probe(64, 32, 2)).all()
    wl = workloads.c3()
    assert (wl.width, wl.height, wl.max_depth, wl.frames) == (1920, 1080, 8, 8)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_partition_covers_and_balances(n):
    owner, loads = partition.partition(1920, 1080, (960, 540), 74, 241, n)
    assert owner.shape == (34, 60) and owner.max() == n - 1 and set(np.unique(owner)) == set(range(n))
    assert abs(loads / loads.sum() - 1.0 / n).max() < 0.01
    # moving gaze near a corner still balances
    owner, loads = partition.partition(2160, 2160, (1680, 480), 148, 482, n)
    assert abs(loads / loads.sum() - 1.0 / n).max() < 0.02
    cost = partition.pixel_cost(1920, 1080, (960, 540), 74, 241)
    assert abs(cost.sum() - ring_samples(1920, 1080, (960, 540), 74, 241)) / cost.sum() < 0.05


@pytest.mark.parametrize("n", [2, 4, 8])
def test_interleaved_partition_balances_wherever_the_gaze_is(n):
    """the static partition used for a moving gaze (C4): per-rank load within 2.5 % of the mean along the C4 gaze path"""
    import math
    owner = partition.interleaved(2160, 2160, n, 16)
    assert owner.shape == (135, 135) and set(np.unique(owner)) == set(range(n))
    for k in range(0, 126, 9):
        gaze = (int(1080 + 600 * math.cos(0.05 * k)), int(1080 + 600 * math.sin(0.05 * k)))
        loads = partition.loads_of(owner, 2160, 2160, gaze, 148, 482, n, 16)
        assert loads.max() / loads.mean() < 1.025, (k, loads)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: ONE JSON line on stdout (library banners go to stderr); the reference arm runs the CPU checker on
    rank 0 only, other ranks exit 0 without output"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"]
    env = dict(os.environ); env.pop("RANK", None)
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    env["RANK"] = "1"
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""
