"""Photon-splat sweep (BASELINE config 5 flavour): conference-like scene at 3840x2160, PM mode
(numVplLightPaths = 0), numLightPaths swept; reports light-trace and splat stage times, photons/s,
fragments/s and the splat's algorithmic HBM GB/s (96 B x records + 64 B x px + 48 B x px)."""
import ctypes as C, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from evplp_b200 import host_api as HA, _capi as capi

W, H = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160").split("x"))
sizes = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "262144,1048576,4194304").split(",")]
groups = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]  # splat_mode values: 0 tiled, 1 scatter
radius_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.003
lib = capi.load_library()
hs = HA.HostScene.generate("conference", 1, 8, W / H)
for paths in sizes:
    fam = {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
           "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True, "useStat": False,
           "numLightPaths": paths, "numVplLightPaths": 0, "numMaxBounces": 3, "radiusPercentage": radius_pct, "DoProgressive": False}
    t = HA.Technique(hs, fam, W, H)
    h = t.device_handle()
    for g in groups:
        capi.check(lib, lib.evplp_set_option(h, b"splat_mode", g), "opt")
        t.iterate()  # warm-up
        capi.check(lib, lib.evplp_reset_stats(h), "reset")
        reps = 3
        tr = sp = 0.0
        for _ in range(reps):
            t.iterate()
            ms = C.c_float()
            lib.evplp_last_stage_ms(h, capi.STAGE_LIGHT_TRACE, C.byref(ms)); tr += ms.value
            lib.evplp_last_stage_ms(h, capi.STAGE_SPLAT, C.byref(ms)); sp += ms.value
        st = capi.Stats(); lib.evplp_stats(h, C.byref(st))
        tr /= reps; sp /= reps
        photons, frags = st.splatPhotons / reps, st.splatFragments / reps
        bytes_ = 96.0 * paths * 4 + 112.0 * W * H
        print(json.dumps({"res": f"{W}x{H}", "paths": paths, "records": paths * 4, "mode": "scatter" if g else "tiled", "trace_ms": round(tr, 3), "splat_ms": round(sp, 3),
                          "paths_per_s": paths / tr * 1e3, "photons_per_s": photons / sp * 1e3, "frags_per_photon": frags / max(photons, 1),
                          "frags_per_s": frags / sp * 1e3, "splat_algo_GBps": bytes_ / sp / 1e6}))
    t.close()
