#!/usr/bin/env python
"""bench.py -- headline benchmark of the foveated path-tracing hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c3] [--impl reference]

One "step" = one frame of the workload: the three foveation passes (periphery / mid / fovea,
reference SimplePathtracer.cpp:137-209) traced through the whole wavefront and resolved to the
accumulation + RGBA8 frame buffers.  Default workload: BASELINE.json configs[2] (C3: Sponza-class
atrium, 1920x1080 foveated, depth 8, probe light), the configuration the 1/2/4/8-GPU metric is
quoted on; it fits one GPU.  Frames accumulate (subframe_index advances every step), so K steps
are K of the frames that make up the quoted 256 foveal spp.

value   : Msamples/s with scene, BVH, probe and frame buffers resident in HBM (device-timed).
e2e     : the same frames through the reference-facing API -- SampleRenderer.render() then
          downloadPixels() into host memory: launch parameters cross host->device every frame and
          the RGBA8 frame comes back device->host inside the timed region (per-kernel statistics off,
          as in a host that only renders; `value` is measured with them on, ~1 % of a frame).
roofline: traversal kernel k_trace against the HBM roofline, algorithmic bytes per ray from the
          CPU oracle's canonical BVH (profiles/algorithmic_bytes.json).
cpu_baseline / --impl reference: the CPU checker (oracle/) on the same workload on host cores.

Multi-GPU (torchrun, one rank per GPU): tiles weighted by foveal sample density are split across
ranks (partition.py); each rank traces only its tiles and stores their pixels straight into rank 0's frame over NVLink peer
memory (include/fovpt.h fovpt_peer_frame_*; --gather nccl: one NCCL reduce of the RGBA8 frame per step instead).
"""
import argparse
import contextlib
import ctypes as C
import io
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(workload_name):
    p = os.path.join(ROOT, "profiles", "algorithmic_bytes.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if workload_name in d:
            return d[workload_name]
    return None


def issue_profile(workload_name):
    """ncu-measured constants of k_trace per workload (profiles/issue_slots_r02.json: warp instructions per ray, lanes per warp
    instruction, DRAM / L2 bytes per ray) -- what turns the live rays/s of this run into an issue-slot fraction"""
    p = os.path.join(ROOT, "profiles", "issue_slots_r02.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload_name)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["reasons"] = sorted(reasons)
            out["samples"] = len(sm)
        return out


def make_workload(name):
    from fovpathtracing_optixcodelatest_b200 import workloads
    return workloads.BY_NAME[name]()


# --------------------------------------------------------------------------------------- CPU checker legs
def run_oracle_frames(wl, frames, kind="port", threads=0):
    """times `frames` frames of the workload on the host cores with the CPU checker; returns
    (seconds, counters, threads_used, kind)"""
    from oracle import orc
    if kind == "reference" and not (orc.available("reference") and wl.max_depth == 4 and not wl.seed_all_passes_with_subframe):
        kind = "port"      # oracle/_ref carries sv5's literals (depth 4, no accumulation); other configs use the port
    o = orc.Oracle(kind)
    o.add_scene(wl.scene()); o.commit(); o.set_probe(wl.probe())
    cam = wl.camera
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, wl.width / wl.height)
    lp, accum, frame = o.make_params(wl.width, wl.height, cam.eye, U, V, W, gaze=wl.gaze_xy())
    cfg = wl.apply(orc.FovCfg.sv5())
    opts = wl.apply(orc.Opts.sv5(threads=threads, jitter_yx=1 if kind == "reference" else 0))
    o.reset_counters()
    t0 = time.perf_counter()
    for _ in range(frames):
        o.render(lp, cfg, opts)
    dt = time.perf_counter() - t0
    cnt = o.counters()
    o.close()
    return dt, cnt, (threads if threads > 0 else (os.cpu_count() or 1)), kind


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload)
    # each step = one frame of the workload on all host threads (a bounded sample: K+W frames, not the
    # full 8-frame accumulation)
    for _ in range(args.warmup):
        run_oracle_frames(wl, 1)
    dt, cnt, threads, kind = run_oracle_frames(wl, args.steps)
    msamples = cnt["samples"] / dt / 1e6
    line = {
        "impl": "reference", "metric": "Msamples/s (1080p foveated path tracing)", "value": msamples, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, subframes_per_step(wl, args.subframes)),
        "rays_per_s": (cnt["radiance_rays"] + cnt["shadow_rays"]) / dt,
        "cpu_baseline": {"value": msamples, "unit": "Msamples/s", "cores": threads, "kind": kind,
                         "sample": f"bounded sample: each step = ONE full subframe of {wl.name} (the GPU arm's step is "
                                   f"{subframes_per_step(wl, args.subframes)} of them); {args.steps} step(s), {cnt['samples']} samples on {threads} host threads"},
        "e2e": {"value": msamples, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def subframes_per_step(wl, requested=0):
    """how many accumulated subframes one step renders (as ONE wavefront, fovpt_render_frames): all the subframes of the
    quoted spp when they fit the path budget (C2: 2, C3: 8), one otherwise (C5: a subframe is 265 M samples already)"""
    if requested > 0:
        return requested
    if not wl.accumulate or wl.view_at is not None:
        return 1
    passes = [(wl.width // 4) * (wl.height // 4) * wl.spp[0], (wl.outer_radius + 2) ** 2 * wl.spp[1], (2 * (wl.inner_radius + 1)) ** 2 * wl.spp[2]] \
        if wl.foveated else [wl.width * wl.height * wl.spp_uniform]
    return max(1, min(wl.frames, (48 << 20) // max(1, sum(passes))))


def workload_config(wl, subframes=None):
    return {"workload": f"{wl.name}: {wl.description}", "scene_tris": wl.scene().n_tris, "width": wl.width, "height": wl.height,
            "gaze": list(wl.gaze_xy()), "radii": [wl.inner_radius, wl.outer_radius], "spp_per_frame": list(wl.spp),
            "max_depth": wl.max_depth, "frames_for_quoted_spp": wl.frames,
            "subframes_per_step": subframes if subframes is not None else subframes_per_step(wl),
            "l2": "256 MiB memset between steps, outside the per-step CUDA-event pairs"}


# --------------------------------------------------------------------------------------- native arm
def native_arm(args):
    import torch
    import torch.distributed as dist
    from fovpathtracing_optixcodelatest_b200 import native, partition
    from fovpathtracing_optixcodelatest_b200.renderer import ProbeData, SampleRenderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: libfovpt has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = make_workload(args.workload)
    cfg = wl.apply(native.default_config())
    S = subframes_per_step(wl, args.subframes)

    # the product renderer runs on a dedicated torch stream made current, so that torch's CUDA events and
    # NCCL collectives order with libfovpt's kernels (a NULL handle would make libfovpt create its own stream)
    tstream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    r = SampleRenderer.__new__(SampleRenderer)
    from fovpathtracing_optixcodelatest_b200.structs import LaunchParams
    r.ctx = native.Context(local, stream)
    r.launchParams = LaunchParams(); r.model = wl.scene(); r.config = cfg
    r.ctx.set_config(cfg)
    # scene ingest: geometry goes from the caller's arrays straight into device arenas (fovpt_scene_add_mesh),
    # commit rebases indices and builds the BVH on the device; timed on the host, reported as `ingest`
    t_i0 = time.perf_counter()
    r.ctx.scene_begin()
    for tex in r.model.textures:
        r.ctx.add_texture(tex)
    for m in r.model.meshes:
        r.ctx.add_mesh(m.vtx, m.idx, m.material, m.texcoord, m.texture_id)
    t_i1 = time.perf_counter()
    r.launchParams.traversable = r.ctx.commit()
    t_i2 = time.perf_counter()
    ingest = {"triangles": int(r.model.n_tris), "hand_over_s": t_i1 - t_i0, "commit_s": t_i2 - t_i1,
              "device_build_ms": float(r.ctx.stats()["bvh_build_ms"]), "mtris_per_s": r.model.n_tris / max(t_i2 - t_i0, 1e-9) / 1e6,
              "note": "first ingest of the process (cold allocator, kernels loaded lazily): fovpt_scene_add_* from pageable numpy arrays, "
                      "then fovpt_scene_commit = index rebase + BVH build (Morton, radix sort, PLOC, 4-wide collapse) + upload of the small tables"}
    w, h = wl.width, wl.height
    # frame buffers are torch tensors so NCCL can reduce them in place
    accum_t = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    frame_t = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    lp = r.launchParams
    lp.frame.size_x, lp.frame.size_y = w, h
    lp.frame.accum_buffer, lp.frame.frame_buffer = accum_t.data_ptr(), frame_t.data_ptr()
    r.setCamera(wl.camera)
    r.setProbe(ProbeData(wl.probe()).BuildCDF())
    lp.frame.c_x, lp.frame.c_y = wl.gaze_xy()
    gather = args.gather if world > 1 else "none"
    frame_own = None
    tile = args.tile
    if world > 1:
        if wl.view_at is not None:
            # moving gaze (C4): a static interleaved partition balances wherever the gaze is (partition.interleaved); the
            # density-weighted LPT assignment would cost ~150 ms of host time per gaze point
            tile = 16
            owner = partition.interleaved(w, h, world, tile)
        else:
            owner, loads = partition.partition(w, h, wl.gaze_xy(), wl.inner_radius, wl.outer_radius, world, tile, wl.spp, bool(wl.foveated), wl.spp_uniform)
        r.ctx.set_tile_owner(tile, owner, rank)
        if gather == "peer":
            # the frame lives on rank 0; the other ranks map it (CUDA IPC over NVLink) and their resolve kernels store
            # their tiles straight into it -- include/fovpt.h fovpt_peer_frame_*; no collective call per frame
            ok, shared_ptr, why = 1, None, ""
            hb = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                try:
                    handle, shared_ptr = r.ctx.peer_frame_create(w * h, world, 0)
                    hb = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
                except RuntimeError as e:
                    ok, why = 0, str(e)
            dist.broadcast(hb, src=0)
            if rank != 0:
                try:
                    shared_ptr = r.ctx.peer_frame_open(bytes(hb.cpu().tolist()), rank)
                except RuntimeError as e:
                    ok, why = 0, str(e)
            okt = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            if int(okt[0]) == 1:
                lp.frame.frame_buffer = shared_ptr
                r.ctx.peer_frame_accum(True)           # every rank also mirrors the float4 accumulation value of its pixels into rank 0's memory
            else:
                # CUDA IPC / peer access unavailable between these processes: every rank falls back to the NCCL reduce,
                # and the result line says so (config.gather)
                if why:
                    print(f"[bench] rank {rank}: peer frame unavailable ({why}); using --gather nccl", file=sys.stderr)
                r.ctx.peer_frame_close()
                gather = "nccl"
        if gather == "nccl":
            accum_sum = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")       # reduce target: the job's accumulation buffer on rank 0
        if gather == "nccl" and rank == 0:
            # NCCL path: rank 0 renders its tiles into its own buffer; frame_t receives the sum over ranks
            frame_own = torch.zeros((h, w), dtype=torch.int32, device="cuda")
            lp.frame.frame_buffer = frame_own.data_ptr()
    elif args.emulate_world > 1:        # profiling aid: one process renders rank 0's share of an N-rank partition
        owner, loads = partition.partition(w, h, wl.gaze_xy(), wl.inner_radius, wl.outer_radius, args.emulate_world, args.tile, wl.spp, bool(wl.foveated), wl.spp_uniform)
        r.ctx.set_tile_owner(args.tile, owner, 0)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    host_frame = torch.empty((h, w), dtype=torch.int32).pin_memory()

    step_no = [0]

    def step():
        if wl.view_at is not None:                     # stereo / moving gaze: camera and gaze point change every step (host-side only)
            cam_k, gaze_k = wl.view_at(step_no[0]); step_no[0] += 1
            r.setCamera(cam_k); lp.frame.c_x, lp.frame.c_y = gaze_k
        r.ctx.render_frames(lp, S)
        gather_frame()

    def gather_frame(consume=None):
        """the frame's one exchange step: every rank's tiles end up in rank 0's frame; `consume` reads it on rank 0"""
        if gather == "peer":
            if rank == 0:
                r.ctx.peer_frame_wait()                # this stream waits for every contributor's arrival flag
                if consume:
                    consume()
                r.ctx.peer_frame_release()             # contributors may overwrite the frame from here on
            else:
                r.ctx.peer_frame_signal()
        elif gather == "nccl":
            if rank == 0:
                frame_t.copy_(frame_own)
            dist.reduce(frame_t, dst=0, op=dist.ReduceOp.SUM)
            accum_sum.copy_(accum_t)                   # ranks hold disjoint pixels (zero elsewhere): the sum is the whole float image
            dist.reduce(accum_sum, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0 and consume:
                consume()
        elif consume:
            consume()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the frame's own sample count (what one GPU traces for the whole frame); ranks re-trace the few splat blocks
    # that straddle their tile edges, those duplicates must not count as throughput
    if world > 1 or args.emulate_world > 1:
        r.ctx.set_tile_owner(tile, None, 0)
    r.ctx.reset_stats(); lp.frame.subframe_index = 0
    if rank == 0 or world == 1:
        job_fb = lp.frame.frame_buffer
        lp.frame.frame_buffer = frame_t.data_ptr()      # a counting pass must not leave pixels in the job's shared frame
        if wl.view_at is not None and world > 1:       # moving gaze: the count changes with the gaze point -> the K views of the timed region
            for k in range(args.steps):
                cam_k, gaze_k = wl.view_at(k)
                r.setCamera(cam_k); lp.frame.c_x, lp.frame.c_y = gaze_k
                r.ctx.render_frame(lp)
            torch.cuda.synchronize()
        else:
            t_f0 = time.perf_counter()
            r.ctx.render_frame(lp); torch.cuda.synchronize()
            ingest["first_frame_s"] = time.perf_counter() - t_f0            # includes the one-off allocation of the wavefront buffers
            ingest["time_to_first_frame_s"] = ingest["hand_over_s"] + ingest["commit_s"] + ingest["first_frame_s"]
        lp.frame.frame_buffer = job_fb
    frame_samples = float(r.ctx.stats()["samples"])    # (moving gaze at N>1: the total of the K timed frames)
    if world > 1:
        fs = torch.tensor([frame_samples], dtype=torch.float64, device="cuda"); dist.broadcast(fs, src=0); frame_samples = float(fs[0])
        r.ctx.set_tile_owner(tile, owner, rank)
    elif args.emulate_world > 1:
        r.ctx.set_tile_owner(tile, owner, 0)
    # ---- N > 1: before anything is timed, the partitioned + gathered result of one step is compared PER PIXEL with the same
    # step rendered by rank 0 alone (RGBA8 frame and float4 accumulation buffer, bit for bit)
    parity = None
    if world > 1:
        ref_frame = ref_accum = None
        if rank == 0:
            shared_fb = lp.frame.frame_buffer
            lp.frame.frame_buffer = frame_t.data_ptr()          # rank 0 alone, into its private buffers
            r.ctx.set_tile_owner(tile, None, 0)
            accum_t.zero_(); frame_t.zero_(); lp.frame.subframe_index = 0; step_no[0] = 0
            if wl.view_at is not None:
                cam_k, gaze_k = wl.view_at(0); r.setCamera(cam_k); lp.frame.c_x, lp.frame.c_y = gaze_k
            r.ctx.render_frames(lp, S); torch.cuda.synchronize()
            ref_frame = frame_t.cpu().numpy().view(np.uint32).copy(); ref_accum = accum_t.cpu().numpy().copy()
            lp.frame.frame_buffer = shared_fb
            r.ctx.set_tile_owner(tile, owner, rank)
        accum_t.zero_(); frame_t.zero_(); lp.frame.subframe_index = 0; step_no[0] = 0
        if frame_own is not None:
            frame_own.zero_()
        barrier()
        got = {}

        def grab():
            torch.cuda.current_stream().synchronize()
            if gather == "peer":
                got["frame"] = r.ctx.download(w, h); got["accum"] = r.ctx.download_accum(w, h)
            else:
                got["frame"] = frame_t.cpu().numpy().view(np.uint32).copy(); got["accum"] = accum_sum.cpu().numpy().copy()
        if wl.view_at is not None:
            cam_k, gaze_k = wl.view_at(0); r.setCamera(cam_k); lp.frame.c_x, lp.frame.c_y = gaze_k; step_no[0] = 1
        r.ctx.render_frames(lp, S)
        gather_frame(grab)
        barrier()
        if rank == 0:
            dpx = int((got["frame"] != ref_frame).sum())
            dac = int((got["accum"].view(np.uint32) != ref_accum.view(np.uint32)).any(axis=2).sum())
            parity = {"checked": f"one step ({S} subframe(s)) rendered by {world} ranks and gathered vs the same step rendered by rank 0 alone",
                      "frame_pixels_differing": dpx, "accum_pixels_differing_bitwise": dac, "pixels": w * h, "identical": dpx == 0 and dac == 0}
            if not parity["identical"]:
                print(f"[bench] WARNING: multi-GPU result differs from the 1-rank render: {parity}", file=sys.stderr)
        accum_t.zero_(); frame_t.zero_(); lp.frame.subframe_index = 0; step_no[0] = 0
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    r.ctx.reset_stats()
    lp.frame.subframe_index = 0; step_no[0] = 0
    accum_t.zero_(); frame_t.zero_()
    if frame_own is not None:
        frame_own.zero_()

    # ---- device-timed region: K steps, each bracketed by its own event pair, L2 flushed in between
    clocks = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ext_ms = shade_ms = shadow_ms = 0.0
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step()
        ev[k][1].record()
        if args.kernel_times:
            torch.cuda.synchronize()
            st = r.ctx.stats()
            ext_ms += st["last_extend_ms"]; shade_ms += st["last_shade_ms"]; shadow_ms += st["last_shadow_ms"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    st = r.ctx.stats()
    mine = torch.tensor([total_ms, float(st["samples"]), float(st["radiance_rays"]), float(st["shadow_rays"]), float(st["kernel_launches"]),
                         ext_ms, shade_ms, shadow_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = mine.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = mine.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms = float(mx[0]); samples = float(sm[1]); rad = float(sm[2]); shw = float(sm[3]); launches = float(sm[4])
        ext_ms, shade_ms, shadow_ms = float(mx[5]), float(mx[6]), float(mx[7])
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"ms": float(t[0]) / args.steps, "samples": float(t[1]) / args.steps, "rays": float(t[2] + t[3]) / args.steps,
                     "k_trace_ms": float(t[5]) / args.steps, "k_shade_ms": float(t[6]) / args.steps} for t in allr]
    else:
        samples, rad, shw, launches = float(st["samples"]), float(st["radiance_rays"]), float(st["shadow_rays"]), float(st["kernel_launches"])
        per_rank = None
    clk = clocks.stop() if clocks else None

    # ---- end-to-end region: public API with host buffers (params in, pixels out) every step.  A host that renders and
    # downloads does not read per-kernel statistics, so they are off here (include/fovpt.h fovpt_set_kernel_timing; they
    # were on in the device-timed region above because the roofline needs k_trace's time)
    r.ctx.set_kernel_timing(False)
    lp.frame.subframe_index = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_np = host_frame.numpy().view(np.uint32)
    e0.record()
    step_no[0] = 0
    for k in range(args.steps):
        if wl.view_at is not None:
            cam_k, gaze_k = wl.view_at(step_no[0]); step_no[0] += 1
            r.setCamera(cam_k); lp.frame.c_x, lp.frame.c_y = gaze_k
        if world == 1:
            r.renderSubframes(S)                       # SampleRenderer::render() x S as one submission: LaunchParams host -> device, 3 passes per subframe, sync
            r.downloadPixels(host_np)                  # SampleRenderer::downloadPixels(): RGBA8 frame device -> host, blocking
        else:
            r.renderSubframes(S, sync=False)
            if gather == "peer":
                gather_frame(lambda: r.downloadPixels(host_np))          # blocking D2H of the shared frame on rank 0
            else:
                gather_frame(lambda: (host_frame.copy_(frame_t), torch.cuda.current_stream().synchronize()))
            if rank != 0:
                torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms[0])
    checksum = int(host_np.astype(np.uint64).sum()) if rank == 0 else 0
    # the same loop with a PAGEABLE destination, as the reference's main.cpp holds its pixels (a plain uint32_t array): the
    # download stages through the context's pinned buffer and is copied once more on the host
    e2e_pageable_ms = None
    if world == 1:
        pageable = np.empty((h, w), np.uint32)
        lp.frame.subframe_index = 0; step_no[0] = 0
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for k in range(args.steps):
            if wl.view_at is not None:
                cam_k, gaze_k = wl.view_at(step_no[0]); step_no[0] += 1
                r.setCamera(cam_k); lp.frame.c_x, lp.frame.c_y = gaze_k
            r.renderSubframes(S)
            r.downloadPixels(pageable)
        p1.record()
        barrier()
        e2e_pageable_ms = p0.elapsed_time(p1)
        assert int(pageable.astype(np.uint64).sum()) == checksum

    # the reference's own build flag (--use_fast_math, CMakeLists.txt:181) as an extra data point: the same timed loop with
    # fovpt_config.fast_math = 1 (SFU sin / cos / pow / log / reciprocal in shading; parity bar 1e-3 instead of 1e-6)
    fast = None
    if world == 1:
        cfg_fast = wl.apply(native.default_config()); cfg_fast.fast_math = 1
        r.ctx.set_config(cfg_fast)
        lp.frame.subframe_index = 0; step_no[0] = 0
        for _ in range(2):
            step()
        barrier()
        lp.frame.subframe_index = 0; step_no[0] = 0
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for k in range(args.steps):
            step()
        f1.record()
        barrier()
        fast = f0.elapsed_time(f1)
        r.ctx.set_config(cfg)

    if rank == 0:
        hbm, peak_src = load_peaks()
        traced_samples = samples                               # sum over ranks, incl. tile-edge duplicates
        if wl.view_at is None:
            samples = frame_samples * S * args.steps           # unique samples of the subframes rendered
        elif world > 1:                                        # moving gaze: unique samples of the K views, counted by rank 0 before timing
            samples = frame_samples
        per_step_samples = samples / args.steps
        value = samples / (total_ms * 1e-3) / 1e6
        line = {
            "metric": "Msamples/s (1080p foveated path tracing)", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, S),
            "samples_per_step": per_step_samples, "traced_samples_per_step": traced_samples / args.steps,
            "rays_per_s": (rad + shw) / (total_ms * 1e-3), "radiance_rays_per_step": rad / args.steps, "shadow_rays_per_step": shw / args.steps,
            "gpu_launches": int(launches),
            "e2e": {"value": samples / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": C.sizeof(LaunchParams),
                    "d2h_bytes_per_step": w * h * 4, "ms_per_step": e2e_ms / args.steps, "frame_checksum": checksum,
                    "destination": "pinned host buffer",
                    "pageable_destination": None if e2e_pageable_ms is None else
                    {"value": samples / (e2e_pageable_ms * 1e-3) / 1e6, "ms_per_step": e2e_pageable_ms / args.steps}},
            "ingest": ingest,
            "use_fast_math": None if fast is None else
            {"value": samples / (fast * 1e-3) / 1e6, "ms_per_step": fast / args.steps,
             "note": "fovpt_config.fast_math = 1, the reference's --use_fast_math build (CMakeLists.txt:181); not the headline: `value` is IEEE shading"},
            "clocks": clk, "wall_s_timed_region": t_wall, "step_ms": step_ms, "per_rank": per_rank,
        }
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if world > 1:
            line["config"]["parallelism"] = (f"{world} ranks, static interleaved {tile}-px screen tiles (gaze-independent)" if wl.view_at is not None
                                             else f"{world} ranks, screen tiles of {tile} px weighted by foveal sample density")
            line["config"]["gather"] = ("peer stores of the RGBA8 pixel and the float4 accumulation value into rank 0's memory over NVLink (CUDA IPC), fused into k_resolve_px; device-side arrival/release flags"
                                        if gather == "peer" else "NCCL reduce(SUM) of the zero-filled RGBA8 frame and of the float4 accumulation buffer per step")
        ab = algorithmic_bytes(wl.name)
        if args.kernel_times and ext_ms > 0:
            # k_trace serves both ray kinds in one launch (stats field last_extend_ms carries its time)
            line["kernel_ms_per_step"] = {"k_trace": ext_ms / args.steps, "k_shade": shade_ms / args.steps}
            if ab:
                b_closest, b_any = ab["closest"]["bytes_per_ray"], ab["any"]["bytes_per_ray"]
                n_trace_launches = int(st["trace_launches"]) if world == 1 else args.steps * (wl.max_depth + 2)
                alg_bytes = (rad / world) * b_closest + (shw / world) * b_any              # per GPU, whole timed region
                achieved = alg_bytes / (ext_ms * 1e-3) / 1e9
                ip = issue_profile(wl.name)
                rays_in_kernel = ((rad + shw) / world) / (ext_ms * 1e-3)                   # per GPU
                line["roofline"] = {"bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                                    "frac": achieved / hbm,
                                    # measured DRAM bytes per launch (ncu --set full, dram__bytes_read + write) exist for the 1-GPU launch shape only
                                    "traffic": (ip or {}).get("measured_dram_bytes_per_launch") if world == 1 else None,
                                    "peak_source": peak_src, "bytes_per_ray": {"closest": b_closest, "any": b_any},
                                    "launches": n_trace_launches, "algorithmic_bytes_per_launch": alg_bytes / n_trace_launches,
                                    "avg_launch_ms": ext_ms / n_trace_launches,
                                    "rays_per_s_in_kernel": rays_in_kernel,
                                    "note": "algorithmic bytes = oracle canonical-BVH node visits*64 + triangle tests*48 + 48 per ray "
                                            "(DESIGN.md section 5).  Measured DRAM traffic is 58 B/ray on C3 (tree L2-resident) and 109 B/ray on C5 "
                                            "(10 M triangles, tree > L2; ncu, profiles/ncu_r03c_*): L1/L2 absorb > 90 % of the algorithmic bytes on every "
                                            "workload, so HBM is not what binds k_trace -- issue slots are, see roofline_issue"}
                if ip and clk and clk.get("sm_mhz"):
                    # issue-slot roofline: one warp instruction per SM sub-partition and clock.  warp_inst_per_ray is an ncu constant of this
                    # build (smsp__inst_executed.sum of the k_trace launches of one step / rays in them); everything else is live.
                    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
                    peak_issue = 4.0 * n_sm * clk["sm_mhz"] * 1e6
                    ach_issue = rays_in_kernel * ip["warp_inst_per_ray"]
                    line["roofline_issue"] = {"bound": "issue-slots", "kernel": "k_trace", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9,
                                              "unit": "G warp-instructions/s", "frac": ach_issue / peak_issue,
                                              "warp_inst_per_ray": ip["warp_inst_per_ray"], "lanes_per_warp_inst": ip["lanes_per_warp_inst"],
                                              "useful_lane_frac": ach_issue / peak_issue * ip["lanes_per_warp_inst"] / 32.0,
                                              "sm_mhz_under_load": clk["sm_mhz"], "ncu_issue_active_pct": ip["issue_active_pct_weighted"],
                                              "l2_bytes_per_ray": ip["l2_bytes_per_ray"], "l2_GBps": rays_in_kernel * ip["l2_bytes_per_ray"] / 1e9,
                                              "dram_bytes_per_ray": ip["measured_dram_bytes_per_ray"], "dram_GBps": rays_in_kernel * ip["measured_dram_bytes_per_ray"] / 1e9,
                                              "note": "peak = 4 sub-partitions x SMs x SM clock sampled during the timed region; warp_inst_per_ray, "
                                                      "lanes, L2 / DRAM bytes per ray from profiles/issue_slots_r02.json (ncu --set full of one step, N = 1 launch shapes)"}
        if world == 1 and not args.no_cpu_baseline:
            dt1, _, _, _ = run_oracle_frames(wl, 1)                       # sizes the sample: about 12 s of host work
            nfr = int(min(16, max(1, round(12.0 / max(dt1, 1e-3)))))
            dt, cnt, threads, kind = run_oracle_frames(wl, nfr)
            line["cpu_baseline"] = {"value": cnt["samples"] / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
                                    "sample": f"{nfr} full frame(s) of {wl.name} ({cnt['samples']} samples, {cnt['radiance_rays'] + cnt['shadow_rays']} rays) in {dt:.1f} s",
                                    "rays_per_s": (cnt["radiance_rays"] + cnt["shadow_rays"]) / dt}
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--subframes", type=int, default=0,
                    help="accumulated subframes per step, traced as one wavefront (default: all subframes of the workload's quoted spp "
                         "that fit the path budget: C2 2, C3 8, C5 1)")
    ap.add_argument("--tile", type=int, default=32, help="screen tile edge of the multi-GPU partition")
    ap.add_argument("--gather", choices=("peer", "nccl"), default="peer",
                    help="N>1: how the ranks' tiles reach rank 0 -- peer: stores into rank 0's frame over NVLink peer memory, fused into "
                         "the resolve kernel (default); nccl: one dist.reduce(SUM) of zero-filled RGBA8 frames")
    ap.add_argument("--emulate-world", type=int, default=1, help="profiling aid: render only rank 0's tiles of an N-rank partition (single process; value is NOT a benchmark number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-times", dest="kernel_times", action="store_false")
    args = ap.parse_args()
    # contract: ONE JSON line on stdout.  Libraries under us may print banners there (NCCL prints its version when the
    # box exports NCCL_DEBUG), so fd 1 points at stderr while the benchmark runs and is restored for the result line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                reference_arm(args)
            else:
                native_arm(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    sys.stdout.write(buf.getvalue())
    sys.stdout.flush()


if __name__ == "__main__":
    main()
