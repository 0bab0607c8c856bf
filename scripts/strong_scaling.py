"""Strong scaling of ONE heavy frame under the image partition: every rank traces all
light paths, gathers its interleaved 16-row bands, splats its range of light paths, and one NCCL all-reduce of the int64
layers ends the frame.  Run under torchrun with N ranks (or plainly for N = 1); prints one JSON line with the frame time
(CUDA events, max over ranks) -- divide the N = 1 time by it for the speed-up.

  python scripts/strong_scaling.py [scene] [WxH] [numVplLightPaths] [numLightPaths]
"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evplp_b200 import _capi as capi  # noqa: E402
from evplp_b200 import host_api as HA  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "buddha"
W, H = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "3840x2160").split("x"))
nvpl = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
npaths = int(sys.argv[4]) if len(sys.argv) > 4 else 300000
FAM = {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
       "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True,
       "useStat": False, "numLightPaths": npaths, "numVplLightPaths": nvpl, "numMaxBounces": 3, "radiusPercentage": 0.003,
       "DoProgressive": True, "AlphaProgressive": 0.7}

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
lib = capi.load_library()
for kv in os.environ.get("EVPLP_OPTS", "").split(","):   # e.g. EVPLP_OPTS=gather_chunks=4
    if kv:
        capi.check(lib, lib.evplp_set_option(None, kv.split("=")[0].encode(), int(kv.split("=")[1])), "set_option")
hs = HA.HostScene.generate(scene, 1, 8, W / H)
fake = int(os.environ.get("FAKE_WORLD", 0))   # debugging aid: render rank FAKE_RANK's share of a FAKE_WORLD-way partition on one GPU
if fake:
    t = HA.Technique(hs, FAM, W, H, device=local, rank=int(os.environ.get("FAKE_RANK", 0)), world_size=fake, image_partition=True)
else:
    t = HA.Technique(hs, FAM, W, H, device=local, rank=rank, world_size=world, image_partition=world > 1)
h = t.device_handle()


def layers():
    out = []
    for layer, ts in ((0, "<i8"), (1, "<i8"), (2, "<i4")):
        p, n = C.c_void_p(), C.c_uint64()
        capi.check(lib, lib.evplp_accum_layer(h, layer, C.byref(p), C.byref(n)), "layer")

        class _W:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": ts, "data": (p.value, False), "version": 2}

        out.append(torch.as_tensor(_W(), device=f"cuda:{local}"))
    return out


ls = layers()


split = {"render_ms": 0.0, "reduce_ms": 0.0}


def frame():
    import time as _t
    a = _t.perf_counter()
    t.iterate()
    capi.check(lib, lib.evplp_synchronize(h), "sync")  # the technique runs on its own stream
    b = _t.perf_counter()
    if world > 1:
        for x in ls:
            dist.all_reduce(x, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    c = _t.perf_counter()
    split["render_ms"] += (b - a) * 1e3
    split["reduce_ms"] += (c - b) * 1e3


frame()  # warm-up
torch.cuda.synchronize()
split["render_ms"] = split["reduce_ms"] = 0.0
if world > 1:
    dist.barrier()
reps = 2
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter()
e0.record()
for _ in range(reps):
    frame()
e1.record()
torch.cuda.synchronize()
ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / reps  # the gather is on the handle's stream: wall clock bounds it
tm = torch.tensor([ms], device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
msv = C.c_float(); stage = []
for idx in (1, 2, 3, 4):   # gbuffer, light trace, gather, splat of the LAST frame on this rank
    capi.check(lib, lib.evplp_last_stage_ms(h, idx, C.byref(msv)), "stage_ms"); stage.append(msv.value)
stage_t = torch.tensor(stage, device=f"cuda:{local}")
all_stage = [torch.zeros_like(stage_t) for _ in range(world)]
if world > 1:
    dist.all_gather(all_stage, stage_t)
else:
    all_stage = [stage_t]
st = capi.Stats(); capi.check(lib, lib.evplp_stats(h, C.byref(st)), "stats")
pairs = torch.tensor([float(st.gatherPairs)], device=f"cuda:{local}", dtype=torch.float64)
if world > 1:
    dist.all_reduce(pairs, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"scene": scene, "res": f"{W}x{H}", "numVplLightPaths": nvpl, "numLightPaths": npaths, "n_gpus": world,
                      "partition": "image bands (16 rows, interleaved) + light-path ranges" if world > 1 else "none",
                      "frame_ms": float(tm.item()), "rank0_render_ms": split["render_ms"] / reps, "rank0_reduce_ms": split["reduce_ms"] / reps, "pairs_per_frame_all_ranks": float(pairs.item()) / (reps + 1),
                      "per_rank_stage_ms[gbuffer,trace,gather,splat]": [[round(float(v), 2) for v in x.tolist()] for x in all_stage]}))
t.close()
if world > 1:
    dist.destroy_process_group()
