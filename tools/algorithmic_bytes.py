"""Counts the canonical-BVH visits of a workload's rays with the CPU oracle and writes the
ALGORITHMIC bytes per ray (64*n_node + 48*n_tri + 48) to profiles/algorithmic_bytes.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload, run_oracle_frames

out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "algorithmic_bytes.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}


def strided_sample(wl, stride=8, spp=2):
    """C5: a full frame is 265 M samples; count on every `stride`-th pixel of the frame instead (one launch with
    factor = stride, fill 1: launch index (x, y) -> pixel (stride x, stride y), deviceProgram.cu:433)"""
    from oracle import orc
    o = orc.Oracle("port"); o.add_scene(wl.scene()); o.commit(); o.set_probe(wl.probe())
    cam = wl.camera
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, wl.width / wl.height)
    lp, _, _ = o.make_params(wl.width, wl.height, cam.eye, U, V, W)
    lp.frame.factor_x = lp.frame.factor_y = stride; lp.frame.factor_z = 1; lp.frame.fillSize = 1
    lp.frame.r_inner, lp.frame.r_outer = 0.0, 1e9; lp.samples_per_launch = spp
    o.reset_counters()
    t0 = time.perf_counter()
    o.launch(lp, wl.width // stride, wl.height // stride, wl.apply(orc.Opts.sv5()))
    dt = time.perf_counter() - t0
    c = o.counters(); o.close()
    return dt, c, os.cpu_count(), "port"


for name in sys.argv[1:] or ["c3"]:
    wl = make_workload(name)
    strided = wl.name == "C5"
    dt, c, threads, kind = strided_sample(wl) if strided else run_oracle_frames(wl, 1)
    nn, nt = c["closest_nodes"] / c["radiance_rays"], c["closest_tris"] / c["radiance_rays"]
    an, at = c["any_nodes"] / max(1, c["shadow_rays"]), c["any_tris"] / max(1, c["shadow_rays"])
    data[wl.name] = {
        "frame": "every 8th pixel of subframe 0, 2 spp" if strided else "subframe 0, all passes", "samples": c["samples"], "radiance_rays": c["radiance_rays"], "shadow_rays": c["shadow_rays"],
        "closest": {"n_node": nn, "n_tri": nt, "bytes_per_ray": 64 * nn + 48 * nt + 48},
        "any": {"n_node": an, "n_tri": at, "bytes_per_ray": 64 * an + 48 * at + 48},
        "oracle_seconds": dt, "oracle_threads": threads, "oracle_kind": kind,
    }
    print(name, json.dumps(data[wl.name]))
json.dump(data, open(out_path, "w"), indent=1)
