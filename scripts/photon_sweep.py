"""BASELINE config 5: progressive photon splatting sweep on the conference-like scene at 3840x2160 (PM mode,
numVplLightPaths = 0), light paths streamed through an 8 Mi-path record buffer.  Wall-clock per frame with the
stream synchronised on both sides; reports paths/s, photon records/s, usable photons/s and fragments/s."""
import ctypes as C, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evplp_b200 import host_api as HA, _capi as capi

W, H = 3840, 2160
sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "262144,4194304,67108864").split(",")]
lib = capi.load_library()
hs = HA.HostScene.generate("conference", 1, 8, W / H)
for paths in sizes:
    fam = {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
           "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True, "useStat": False,
           "numLightPaths": paths, "numVplLightPaths": 0, "numMaxBounces": 3, "radiusPercentage": 0.003, "DoProgressive": True}
    t = HA.Technique(hs, fam, W, H)
    h = t.device_handle()
    t.iterate()
    capi.check(lib, lib.evplp_synchronize(h), "sync")
    capi.check(lib, lib.evplp_reset_stats(h), "reset")
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        t.iterate()
    capi.check(lib, lib.evplp_synchronize(h), "sync")
    dt = (time.perf_counter() - t0) / reps
    st = capi.Stats(); capi.check(lib, lib.evplp_stats(h, C.byref(st)), "stats")
    ms = C.c_float(); stage = {}
    for name, idx in (("light_trace", 2), ("photon_splat", 4)):   # the LAST chunk of the frame (evplp_last_stage_ms)
        capi.check(lib, lib.evplp_last_stage_ms(h, idx, C.byref(ms)), "stage_ms"); stage[name] = round(ms.value, 3)
    print(json.dumps({"res": f"{W}x{H}", "paths": paths, "records_per_frame": paths * 4, "frame_ms": round(dt * 1e3, 2),
                      "paths_per_s": paths / dt, "records_per_s": paths * 4 / dt, "usable_photons_per_s": st.splatPhotons / reps / dt,
                      "fragments_per_s": st.splatFragments / reps / dt, "radius": t.state()["radius"], "last_chunk_stage_ms": stage}))
    t.close()
