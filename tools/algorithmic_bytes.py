"""Counts the canonical-BVH visits of a workload's rays with the CPU oracle and writes the
ALGORITHMIC bytes per ray (64*n_node + 48*n_tri + 48) to profiles/algorithmic_bytes.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload, run_oracle_frames

out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "algorithmic_bytes.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
for name in sys.argv[1:] or ["c3"]:
    wl = make_workload(name)
    dt, c, threads, kind = run_oracle_frames(wl, 1)
    nn, nt = c["closest_nodes"] / c["radiance_rays"], c["closest_tris"] / c["radiance_rays"]
    an, at = c["any_nodes"] / max(1, c["shadow_rays"]), c["any_tris"] / max(1, c["shadow_rays"])
    data[wl.name] = {
        "frame": "subframe 0, all three passes", "samples": c["samples"], "radiance_rays": c["radiance_rays"], "shadow_rays": c["shadow_rays"],
        "closest": {"n_node": nn, "n_tri": nt, "bytes_per_ray": 64 * nn + 48 * nt + 48},
        "any": {"n_node": an, "n_tri": at, "bytes_per_ray": 64 * an + 48 * at + 48},
        "oracle_seconds": dt, "oracle_threads": threads, "oracle_kind": kind,
    }
    print(name, json.dumps(data[wl.name]))
json.dump(data, open(out_path, "w"), indent=1)
