"""world_size-2 gloo test (CPU) of the N > 1 path's host logic: iterations dealt round-robin
(rank g renders k = g mod N), every rank replays the jitter stream and the progressive
schedule for ALL iterations, and one sum all-reduce of the int64 accumulation layers gives
exactly the single-rank result.  The per-iteration renderer here is the CPU oracle (test
infrastructure); on the GPU box bench.py --gpus N does the same with the sm_100a kernels
and NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, PATHS, VPL_PATHS, ITERS = 24, 16, 128, 12, 4


def _render(rank, world):
    import evplp_b200 as E
    from evplp_b200 import _capi as capi
    from evplp_b200 import host_api as HA
    from tests import oracle_api as O

    scene, cam = E.cornell_scene(seed=4, detail=2)
    camera = E.Camera(cam["origin"], cam["lookat"], cam["up"], cam["fovx"], W / H)
    orc = O.OracleScene(scene)
    host = HA.load_host_library()
    jit = np.empty(2 * ITERS, dtype=np.float32)
    host.evplp_host_jitter_stream(0, ITERS, capi.ptr(jit))
    state = np.array([float(scene.bounding_sphere_radius()) * 0.05, 0.02, 0, 0, 0], dtype=np.float32)
    vpl = np.zeros((H, W, 3), dtype=np.int64); photon = np.zeros((H, W, 3), dtype=np.int64); light = np.zeros((H, W), dtype=np.uint32)
    for k in range(ITERS):
        if k % world == rank:
            j = ((2 * jit[2 * k] - 1) / W, (2 * jit[2 * k + 1] - 1) / H)
            P = E.make_params(camera, PATHS, VPL_PATHS, 3, float(state[0]), mis_mode=capi.MIS_GEOMETRY_CLAMP, clamp=float(state[1]),
                              jitter=j, rng_seed=k)
            planes, prims = orc.gbuffer(P, W, H)
            rec = orc.light_trace(P, k, 0, PATHS)
            img, _ = orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL)
            orc.accumulate_fixed(img, vpl)
            orc.photon_splat(P, W, H, planes, prims, rec, 0, len(rec), photon)
            orc.light_pass(P, W, H, light)
        # every rank replays the schedule of every iteration (rtcomphoton.h:1033-1063)
        host.evplp_host_progressive_update(k + 1, 0.7, 0.02, VPL_PATHS, PATHS, 0, capi.ptr(state))
    return vpl, photon, light, state


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vpl, photon, light, state = _render(rank, world)
    tv, tp, tl = torch.from_numpy(vpl), torch.from_numpy(photon), torch.from_numpy(light.astype(np.int64))
    for t in (tv, tp, tl):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), vpl=tv.numpy(), photon=tp.numpy(), light=tl.numpy(), state=state)
    dist.destroy_process_group()


def test_round_robin_iterations_plus_allreduce_equal_single_rank(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "reduced.npz")
    vpl, photon, light, state = _render(0, 1)
    assert np.array_equal(got["vpl"], vpl)
    assert np.array_equal(got["photon"], photon)
    # the light mask is written (un-jittered, the same in every iteration), so N ranks sum to N x mask: resolve tests != 0
    assert np.array_equal(got["light"] != 0, light != 0)
    assert np.array_equal(got["state"], state)
    assert photon.sum() > 0 and vpl.sum() > 0
