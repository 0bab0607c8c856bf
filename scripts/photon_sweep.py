"""BASELINE config 5: progressive photon splatting sweep on the conference-like scene at 3840x2160 (PM mode,
numVplLightPaths = 0), light paths streamed through a 32 Mi-path record buffer.  Under torchrun with N ranks the frame is
partitioned by light paths (RtComPhoton::PartitionImage: rank r traces AND splats paths [r P / N, (r + 1) P / N) into a
full-frame photon layer) and one NCCL all-reduce of the int64 layers ends the frame.  Wall clock per frame with a
device synchronise + barrier on both sides, max over ranks; reports paths/s, photon records/s, usable photons/s and
fragments/s of the whole job.

  python scripts/photon_sweep.py [paths,paths,...]            (N = 1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/photon_sweep.py ...
"""
import ctypes as C
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evplp_b200 import _capi as capi  # noqa: E402
from evplp_b200 import host_api as HA  # noqa: E402

W, H = 3840, 2160
sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "262144,4194304,67108864").split(",")]
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
lib = capi.load_library()
hs = HA.HostScene.generate("conference", 1, 8, W / H)


def layers(h):
    out = []
    for layer, ts in ((0, "<i8"), (1, "<i8"), (2, "<i4")):
        p, n = C.c_void_p(), C.c_uint64()
        capi.check(lib, lib.evplp_accum_layer(h, layer, C.byref(p), C.byref(n)), "layer")

        class _W:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": ts, "data": (p.value, False), "version": 2}

        out.append(torch.as_tensor(_W(), device=f"cuda:{local}"))
    return out


for paths in sizes:
    fam = {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
           "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True, "useStat": False,
           "numLightPaths": paths, "numVplLightPaths": 0, "numMaxBounces": 3, "radiusPercentage": 0.003, "DoProgressive": True}
    t = HA.Technique(hs, fam, W, H, device=local, rank=rank, world_size=world, image_partition=world > 1)
    h = t.device_handle()
    if world > 1:   # a rank traces only its own range of paths (runStreamed), in chunks of at most 32 Mi paths
        t.set_max_paths_per_trace(max(1, min(1 << 25, paths // world)))
    ls = layers(h)

    def frame():
        t.iterate()
        capi.check(lib, lib.evplp_synchronize(h), "sync")
        if world > 1:
            for x in ls:
                dist.all_reduce(x, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()

    frame()   # warm-up (allocations, NCCL channels)
    capi.check(lib, lib.evplp_reset_stats(h), "reset")
    if world > 1:
        dist.barrier()
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        frame()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], device=f"cuda:{local}", dtype=torch.float64)
    st = capi.Stats(); capi.check(lib, lib.evplp_stats(h, C.byref(st)), "stats")
    cnt = torch.tensor([float(st.splatPhotons), float(st.splatFragments)], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dts = float(dt.item())
    ms = C.c_float(); stage = {}
    for name, idx in (("light_trace", 2), ("photon_splat", 4)):   # the LAST chunk of the frame on rank 0 (evplp_last_stage_ms)
        capi.check(lib, lib.evplp_last_stage_ms(h, idx, C.byref(ms)), "stage_ms"); stage[name] = round(ms.value, 3)
    if rank == 0:
        print(json.dumps({"res": f"{W}x{H}", "n_gpus": world, "paths": paths, "records_per_frame": paths * 4, "frame_ms": round(dts * 1e3, 2),
                          "paths_per_s": paths / dts, "records_per_s": paths * 4 / dts, "usable_photons_per_s": float(cnt[0].item()) / reps / dts,
                          "fragments_per_s": float(cnt[1].item()) / reps / dts, "radius": t.state()["radius"],
                          "rank0_last_chunk_stage_ms": stage,
                          "timing": "wall clock, device sync + barrier on both sides, max over ranks; includes the all-reduce of the layers"}), flush=True)
    t.close()
if world > 1:
    dist.destroy_process_group()
