"""Ad-hoc stage timing on a procedural room (development aid, not the bench contract)."""
import sys, time
import numpy as np
import evplp_b200 as E
from evplp_b200 import _capi as capi

W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 1280, int(sys.argv[2]) if len(sys.argv) > 2 else 720
detail = int(sys.argv[3]) if len(sys.argv) > 3 else 48
nvpl = int(sys.argv[4]) if len(sys.argv) > 4 else 256
npaths = int(sys.argv[5]) if len(sys.argv) > 5 else 65536
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
sc, cam = E.cornell_scene(detail=detail)
camera = E.Camera(cam["origin"], cam["lookat"], cam["up"], cam["fovx"], W / H)
dev = E.Device(W, H)
dev.upload_scene(sc); dev.build_bvh()
print("prims", sc.num_prims, "bvh ms", dev.stage_ms(capi.STAGE_BVH))
radius = float(sc.bounding_sphere_radius()) * 0.003
for it in range(iters):
    P = E.make_params(camera, npaths, nvpl, 3, radius, mis_mode=capi.MIS_GEOMETRY_CLAMP, clamp=0.05, rng_seed=it)
    dev.set_params(P); dev.reset_stats()
    dev.gbuffer(); dev.light_trace(it, 0, npaths); dev.vpl_gather(); dev.photon_splat(0, npaths * 4); dev.light_pass()
    dev.synchronize()
    st = dev.stats()
    ms = [dev.stage_ms(s) for s in (capi.STAGE_GBUFFER, capi.STAGE_LIGHT_TRACE, capi.STAGE_GATHER, capi.STAGE_SPLAT)]
    print(f"it{it} gbuf {ms[0]:.3f} trace {ms[1]:.3f} gather {ms[2]:.3f} splat {ms[3]:.3f} ms | pairs {st.gatherPairs:.3e} "
          f"({st.gatherPairs/ms[2]*1e3:.3e}/s) rays {st.shadowRays:.3e} ({st.shadowRays/ms[2]*1e3:.3e}/s) photons {st.splatPhotons} "
          f"({st.splatPhotons/ms[3]*1e3:.3e}/s) frags {st.splatFragments:.3e} ({st.splatFragments/ms[3]*1e3:.3e}/s) "
          f"closest {st.closestRays:.3e} ({(st.closestRays - W*H)/ms[1]*1e3:.3e}/s)")
img = dev.resolve(1.0 / iters, 1.0 / iters, 1.0)
print("mean", img.mean(axis=(0, 1)))
