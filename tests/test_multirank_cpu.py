"""world_size-2 gloo test (CPU) of the N > 1 path's HOST logic.  What each rank does in each pass of the loop comes from the
C++ RtComPhoton class itself (RtComPhoton::planNext / advanceSchedule, the code iterate() runs, reached device-free through
evplp_host_technique_plan): which iterations it renders, with which jitter / seed / radius / clamp, which light paths it
splats, which 8x4-pixel tiles it gathers, who draws the light and who counts the iteration.  The stages themselves are played
by the CPU oracle here (test infrastructure: no GPU in this suite); the int64 layers and the iteration count are summed with
a gloo all-reduce and must equal the single-rank run exactly, for both partitions (iterations round-robin; image tiles +
light-path ranges).  On the GPU box bench.py --gpus N prints the checksums of the same comparison with the sm_100a kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, ITERS = 40, 24, 4
FAM = {"rngOffset": 2, "numMaxIteration": -1, "timeLimitMs": 600000.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
       "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True, "useStat": False,
       "numLightPaths": 160, "numVplLightPaths": 12, "numMaxBounces": 3, "radiusPercentage": 0.05, "misMode": "geometryClamp",
       "clampingCoeff": 0.02, "DoProgressive": True, "AlphaProgressive": 0.7}


def _tile_mask(stride, offset):
    """pixels of the 8x4-pixel tiles t = offset (mod stride), row-major numbering with the phantom-tile pitch rule
    (include/evplp.h, gather_band_stride)"""
    tiles_x, tiles_y = (W + 7) // 8, (H + 3) // 4
    pitch = tiles_x + 1 if (stride > 1 and tiles_x % stride == 0) else tiles_x
    m = np.zeros((H, W), dtype=bool)
    for t in range(offset, pitch * tiles_y, stride):
        ty, tx = divmod(t, pitch)
        if tx < tiles_x:
            m[ty * 4:ty * 4 + 4, tx * 8:tx * 8 + 8] = True
    return m


def _render(rank, world, image_partition):
    import evplp_b200 as E
    from evplp_b200 import _capi as capi
    from evplp_b200 import host_api as HA
    from tests import oracle_api as O

    hs = HA.HostScene.generate("livingroom", 2, 1, W / H)
    scene = hs.to_scene()
    orc = O.OracleScene(scene)
    plans = HA.plan(hs, FAM, W, H, rank, world, ITERS, image_partition=image_partition)
    vpl = np.zeros((H, W, 3), dtype=np.int64); photon = np.zeros((H, W, 3), dtype=np.int64); light = np.zeros((H, W), dtype=np.uint32)
    count = 0
    paths, vpl_paths = FAM["numLightPaths"], FAM["numVplLightPaths"]
    for p in plans:
        if not p["render"]:
            continue
        P = E.make_params(hs.camera(), paths, vpl_paths, 3, np.float32(p["photonRadius"]), mis_mode=capi.MIS_GEOMETRY_CLAMP,
                          clamp=np.float32(p["clampingValue"]), jitter=(np.float32(p["jitter_x"]), np.float32(p["jitter_y"])),
                          rng_seed=int(p["rngSeed"]))
        P.pdfMc = np.float32(p["pdfMc"])
        planes, prims = orc.gbuffer(P, W, H)
        rec = orc.light_trace(P, int(p["rngSeed"]), 0, paths)
        img, _ = orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL)
        img[~_tile_mask(int(p["tileStride"]), int(p["tileOffset"]))] = 0          # this rank's tiles of the gather
        orc.accumulate_fixed(img, vpl)
        first, num = int(p["splatFirstPath"]) * 4, int(p["splatNumPaths"]) * 4        # this rank's light paths of the splat
        orc.photon_splat(P, W, H, planes, prims, rec, first, num, photon)
        if p["drawLight"]:
            orc.light_pass(P, W, H, light)
        count += int(p["countIteration"])
    return vpl, photon, light, count, plans


def _worker(rank, world, port, out_dir, image_partition):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vpl, photon, light, count, _ = _render(rank, world, image_partition)
    tv, tp, tl = torch.from_numpy(vpl), torch.from_numpy(photon), torch.from_numpy(light.astype(np.int64))
    tc = torch.tensor([count], dtype=torch.int64)
    for t in (tv, tp, tl, tc):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), vpl=tv.numpy(), photon=tp.numpy(), light=tl.numpy(), count=tc.numpy())
    dist.destroy_process_group()


def _check(tmp_path, image_partition):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000) + (7 if image_partition else 0)
    mp.spawn(_worker, args=(2, port, str(tmp_path), image_partition), nprocs=2, join=True)
    got = np.load(tmp_path / "reduced.npz")
    vpl, photon, light, count, plans = _render(0, 1, False)
    assert np.array_equal(got["vpl"], vpl)
    assert np.array_equal(got["photon"], photon)
    # the light mask is written (un-jittered, the same in every iteration), so N ranks sum to N x mask: resolve tests != 0
    assert np.array_equal(got["light"] != 0, light != 0)
    assert int(got["count"][0]) == count == ITERS          # what finish() normalises by
    assert photon.sum() > 0 and vpl.sum() > 0
    return plans


def test_round_robin_iterations_plus_allreduce_equal_single_rank(tmp_path):
    plans = _check(tmp_path, image_partition=False)
    assert all(p["render"] for p in plans)                 # a single rank renders every iteration
    assert plans[1]["photonRadius"] < plans[0]["photonRadius"]  # the progressive schedule shrinks the radius


def test_image_partition_plus_allreduce_equal_single_rank(tmp_path):
    _check(tmp_path, image_partition=True)


def test_plans_partition_the_work():
    from evplp_b200 import host_api as HA

    hs = HA.HostScene.generate("livingroom", 2, 1, W / H)
    for world in (2, 3, 8):
        it = [HA.plan(hs, FAM, W, H, r, world, 16) for r in range(world)]
        for k in range(16):
            assert sum(int(p[k]["render"]) for p in it) == 1                       # exactly one rank renders iteration k
            assert len({(p[k]["jitter_x"], p[k]["photonRadius"], p[k]["rngSeed"]) for p in it}) == 1   # all replay the same schedule
        im = [HA.plan(hs, FAM, W, H, r, world, 2, image_partition=True) for r in range(world)]
        assert sum(p[0]["splatNumPaths"] for p in im) == FAM["numLightPaths"]
        assert sum(int(p[0]["drawLight"]) for p in im) == 1 and sum(int(p[0]["countIteration"]) for p in im) == 1
        cover = sum(_tile_mask(world, r).astype(int) for r in range(world))
        assert (cover == 1).all()                                                   # the tiles are a disjoint cover of the image
    import pytest

    cef = dict(FAM, frameMode="cleareveryframe")
    with pytest.raises(HA.HostError):
        HA.plan(hs, cef, W, H, 0, 2, 2)                                             # one frame cannot be dealt round-robin
    assert HA.plan(hs, cef, W, H, 0, 2, 2, image_partition=True)[0]["render"]
