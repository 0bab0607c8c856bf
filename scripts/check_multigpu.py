"""N-GPU == 1-GPU check (run under torchrun): every rank renders iterations k = rank (mod N) of a
small progressive run with the C++ RtComPhoton technique, the int64 accumulation layers are
all-reduced over NCCL, and rank 0 compares them bit for bit with the same run on one GPU."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evplp_b200 import _capi as capi  # noqa: E402
from evplp_b200 import host_api as HA  # noqa: E402

W, H, ITERS = 320, 180, 6
FAM = {"rngOffset": 3, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
       "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True,
       "useStat": False, "numLightPaths": 20000, "numVplLightPaths": 64, "numMaxBounces": 3, "radiusPercentage": 0.01,
       "misMode": "geometryClamp", "DoProgressive": True, "AlphaProgressive": 0.7}


def render(hs, device, rank, world, image_partition=False):
    lib = capi.load_library()
    t = HA.Technique(hs, FAM, W, H, device=device, rank=rank, world_size=world, image_partition=image_partition)
    h = t.device_handle()
    capi.check(lib, lib.evplp_set_option(h, b"gather_chunks", 1), "opt")
    for _ in range(ITERS):
        t.iterate()
    capi.check(lib, lib.evplp_synchronize(h), "sync")
    return t, h, lib


def layers(lib, h, device):
    out = []
    for layer, ts in ((0, "<i8"), (1, "<i8"), (2, "<i4")):
        p, n = C.c_void_p(), C.c_uint64()
        capi.check(lib, lib.evplp_accum_layer(h, layer, C.byref(p), C.byref(n)), "layer")

        class _W:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": ts, "data": (p.value, False), "version": 2}

        out.append(torch.as_tensor(_W(), device=f"cuda:{device}"))
    return out


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    hs = HA.HostScene.generate("livingroom", 2, 2, W / H)
    ok = True
    ref = None
    for image_partition in (False, True):
        t, h, lib = render(hs, local, rank, world, image_partition)
        ls = layers(lib, h, local)
        for x in ls:
            dist.all_reduce(x, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        if rank == 0:
            if ref is None:
                t1, h1, _ = render(hs, local, 0, 1)
                ref = [x.clone() for x in layers(lib, h1, local)]
                t1.close()
            mode = "image bands + path ranges" if image_partition else "iterations round-robin"
            for name, a, b in zip(("vpl", "photon", "light"), ls, ref):
                # the light layer is a mask that every rendering rank writes (un-jittered, identical): compare what resolve tests
                same = bool(torch.equal(a != 0, b != 0)) if name == "light" else bool(torch.equal(a, b))
                print(f"[{mode}] layer {name}: N={world} reduce == 1-GPU: {same} (sum {int(a.sum())})")
                ok &= same and (name == "light" or int(a.abs().sum()) > 0)  # the light may be outside the view
        t.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
