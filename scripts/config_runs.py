"""One or two iterations of every BASELINE.json configuration family (C1, C3, C4 at reduced VPL counts where a
full iteration would take minutes), printing stage times and counters.  Development / robustness run."""
import ctypes as C, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evplp_b200 import host_api as HA, _capi as capi

lib = capi.load_library()
BASE = {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
        "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True, "useStat": False,
        "numMaxBounces": 3, "DoProgressive": True, "AlphaProgressive": 0.7}


def run(name, scene, detail, W, H, iters, lvc=False, **kw):
    fam = dict(BASE); fam.update(kw)
    t0 = time.time()
    hs = HA.HostScene.generate(scene, 1, detail, W / H)
    t = HA.Technique(hs, fam, W, H, lvc=lvc)
    h = t.device_handle()
    ms = C.c_float(); lib.evplp_last_stage_ms(h, capi.STAGE_BVH, C.byref(ms)); bvh = ms.value
    setup = time.time() - t0
    lib.evplp_reset_stats(h)
    stage = {}
    for it in range(iters):
        t.iterate()
        for nm, st in (("gbuffer", capi.STAGE_GBUFFER), ("trace", capi.STAGE_LIGHT_TRACE), ("gather", capi.STAGE_GATHER), ("splat", capi.STAGE_SPLAT)):
            if lib.evplp_last_stage_ms(h, st, C.byref(ms)) == 0:
                stage[nm] = round(ms.value, 3)
    st = capi.Stats(); capi.check(lib, lib.evplp_stats(h, C.byref(st)), "stats")
    img = t.final(1.0 / iters, 1.0 / iters, 1.0)
    print(json.dumps({"config": name, "tris": hs.num_triangles, "res": f"{W}x{H}", "setup_s": round(setup, 2), "bvh_ms": round(bvh, 2),
                      "last_iter_stage_ms": stage, "pairs": st.gatherPairs, "shadow_rays": st.shadowRays, "photons": st.splatPhotons,
                      "fragments": st.splatFragments, "mean_rgb": [round(float(v), 4) for v in img.mean(axis=(0, 1))],
                      "finite": bool((img == img).all())}))
    t.close()


which = sys.argv[1:] or ["C1", "C3", "C4"]
if "C1" in which:
    run("C1 conference 256x256 4k VPL + 64k photon records, geometryClamp", "conference", 8, 256, 256, 2,
        numLightPaths=16384, numVplLightPaths=1024, radiusPercentage=0.003, misMode="geometryClamp")
if "C3" in which:
    run("C3 livingroom 1920x1080 VSL + splat (reduced: 256 VPL paths)", "livingroom", 8, 1920, 1080, 1,
        numLightPaths=65536, numVplLightPaths=256, radiusPercentage=0.003, forceVsl=True, vslRadiusPercentage=0.05, misMode="one")
if "C4" in which:
    run("C4 buddha ~1M tris 3840x2160 VPL gather (reduced: 256 VPL paths)", "buddha", 8, 3840, 2160, 1,
        numLightPaths=65536, numVplLightPaths=256, radiusPercentage=0.003, misMode="one")
if "C3s" in which:   # profiling size: the VSL gather kernel with few enough VSLs that ncu's ~40 replays fit
    run("C3 (profiling size) livingroom 1920x1080 VSL gather, 32 VPL paths", "livingroom", 8, 1920, 1080, 1,
        numLightPaths=65536, numVplLightPaths=32, radiusPercentage=0.003, forceVsl=True, vslRadiusPercentage=0.05, misMode="one")
if "LVC" in which:
    run("LVC conference 1920x1080, per-pixel window of light paths", "conference", 8, 1920, 1080, 1, lvc=True,
        numLightPaths=65536, numVplLightPaths=64, radiusPercentage=0.003, misMode="balance")
if "PT" in which:
    t0 = time.time()
    hs = HA.HostScene.generate("conference", 1, 8, 1920 / 1080)
    pt = HA.PathTracer(hs, {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "outputFilename": "pt.pfm",
                            "statFilename": "s.json", "useJitter": True, "useStat": False, "numSamplePerPixel": 1, "numMaxBounces": 3}, 1920, 1080)
    for _ in range(2):
        pt.iterate()
    img = pt.final(0.5, 0.0)
    print(json.dumps({"config": "RtPt2 conference 1920x1080, 2 iterations", "s": round(time.time() - t0, 2),
                      "mean_rgb": [round(float(v), 4) for v in img.mean(axis=(0, 1))]}))
    pt.close()
