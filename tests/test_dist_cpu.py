"""CPU, world_size 2 over gloo: the N>1 host path of bench.py -- density-weighted tile partition,
per-rank rendering of owned tiles only, one SUM-reduce of the frame -- reproduces the 1-rank frame.
The per-rank renderer here is the CPU oracle masked to the rank's tiles (the CUDA version of the same
test, with two libfovpt contexts, is tests/test_gpu_parity.py::test_tile_partition_union)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, out):
    sys.path.insert(0, ROOT)
    import torch
    from fovpathtracing_optixcodelatest_b200 import partition, scenes
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, ri, ro, tile = 128, 72, 10, 32, 8
    owner, loads = partition.partition(w, h, (w // 2, h // 2), ri, ro, world, tile)
    o = orc.Oracle("port")
    o.add_scene(scenes.cornell_box()); o.commit(); o.set_probe(scenes.constant_probe(8, 4, 2.5))
    cam = scenes.CORNELL_CAMERA
    U, V, W = o.camera_uvw(cam.eye, cam.lookat, cam.up, cam.fovy, w / h)
    lp, accum, frame = o.make_params(w, h, cam.eye, U, V, W)
    o.render(lp, orc.FovCfg.sv5(inner_radius=ri, outer_radius=ro), orc.Opts.sv5(threads=1))
    mask = np.kron(owner == rank, np.ones((tile, tile), bool))[:h, :w]
    mine = torch.from_numpy(np.where(mask, frame.astype(np.int64), 0))
    cover = torch.from_numpy(mask.astype(np.int64))
    dist.reduce(mine, dst=0, op=dist.ReduceOp.SUM)
    dist.reduce(cover, dst=0, op=dist.ReduceOp.SUM)
    t = torch.tensor([float(loads[rank])], dtype=torch.float64)
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)      # bench.py takes the max over ranks
    if rank == 0:
        out["ok"] = bool((mine.numpy() == frame.astype(np.int64)).all() and (cover.numpy() == 1).all())
        out["max_load"] = float(mx[0]); out["sum_load"] = float(loads.sum())
    dist.destroy_process_group()


def test_two_rank_partition_and_reduce():
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, out), nprocs=2, join=True)
    assert out["ok"]
    assert out["max_load"] / out["sum_load"] < 0.52
