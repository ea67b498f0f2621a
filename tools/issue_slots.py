"""k_trace constants for bench.py's issue-slot roofline, from one `ncu --set full` capture of the k_trace launches of one timed step
(tools/r2_profile.sh full) and the bench line the same command printed without the profiler (gpurun_out/plain.log).
usage: python tools/issue_slots.py <workload> <ncu summary json> <bench line json> [fraction of the step's rays in the captured launches]
Updates profiles/issue_slots_r02.json in place (one entry per workload)."""
import json, os, sys

wl, ncu_json, bench_json = sys.argv[1], sys.argv[2], sys.argv[3]
d = json.load(open(ncu_json))
line = [l for l in open(bench_json).read().splitlines() if l.startswith("{")][-1]
b = json.loads(line)
frac = float(sys.argv[4]) if len(sys.argv) > 4 else min(1.0, (48 << 20) / b["samples_per_step"]) if wl.lower() == "c5" else 1.0


def val(k, name):
    for key, v in k.items():
        if key.split(" ")[0] == name:
            return float(v)
    raise KeyError(name)


def unit_scale(k, name):
    for key in k:
        if key.split(" ")[0] == name:
            u = key[key.index("[") + 1:-1] if "[" in key else ""
            return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}.get(u, 1.0)
    return 1.0


tr = [k for k in d["kernels"] if "k_trace" in k.get("Kernel Name", "")]
rays = (b["radiance_rays_per_step"] + b["shadow_rays_per_step"]) * frac
inst = sum(val(k, "smsp__inst_executed.sum") for k in tr)
lanes = sum(val(k, "smsp__inst_executed.sum") * val(k, "smsp__thread_inst_executed_per_inst_executed.ratio") for k in tr) / inst
issue = sum(val(k, "smsp__inst_executed.sum") * val(k, "smsp__issue_active.avg.pct_of_peak_sustained_active") for k in tr) / inst
dram = sum((val(k, "dram__bytes_read.sum") * unit_scale(k, "dram__bytes_read.sum") + val(k, "dram__bytes_write.sum") * unit_scale(k, "dram__bytes_write.sum")) for k in tr)
l2 = sum(val(k, "lts__t_sectors.sum") * 32.0 for k in tr)
t = sum(val(k, "gpu__time_duration.sum") * unit_scale(k, "gpu__time_duration.sum") for k in tr)
entry = {"captured": d.get("source", ""), "k_trace_launches": len(tr), "rays_in_captured_launches": rays,
         "warp_inst_per_ray": inst / rays, "lanes_per_warp_inst": lanes, "issue_active_pct_weighted": issue,
         "measured_dram_bytes_per_launch": dram / len(tr), "measured_dram_bytes_per_ray": dram / rays,
         "l2_bytes_per_ray": l2 / rays, "l2_GBps": l2 / t / 1e9, "dram_GBps": dram / t / 1e9, "k_trace_ms_under_ncu": t * 1e3}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "issue_slots_r02.json")
allw = json.load(open(path)) if os.path.exists(path) else {}
allw[wl.upper()] = entry
json.dump(allw, open(path, "w"), indent=1)
print(json.dumps(entry, indent=1))
