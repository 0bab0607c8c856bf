"""Parity at the BASELINE.json configurations AS STATED (numMaxBounces = 3, so "4 k VPL records" means
numVplLightPaths = 1024 etc.), each through the C ABI against the CPU oracle on the same seeded inputs:

  C1  conference 256x256, 4 k VPLs + 64 k photons, one energy-compensated iteration: the WHOLE image, both the balance
      heuristic (the bundled default) and geometryClamp, VPL layer + photon layer + resolved image, bit for bit.
  C3  livingroom (glossy) 1920x1080, VSL gather, 256 k VPL records per iteration (numVplLightPaths = 65536): the stated
      scene, resolution and VPL set; the gather runs on a 64x32-pixel rectangle of the frame and its 16x8-pixel centre is
      compared bit for bit with the oracle (a full VSL frame is 4e11 pairs x up to 303 samples; the kernel is the same).
  C4  buddha (1.06 M triangles) 3840x2160 VPL gather, whole frame on the GPU: a 32x16-pixel crop across the statue's
      silhouette against the oracle, bit for bit in the record-order mode, within the radiance tolerance in the default mode.
  (C2's iteration is what bench.py runs and tests/test_gpu_fullsize.py checks at 1080p; C5's streaming is checked there too.)
"""
import numpy as np
import pytest

import evplp_b200 as E
from evplp_b200 import _capi as capi
from evplp_b200 import host_api as HA
from tests import oracle_api as O

pytestmark = pytest.mark.gpu


def _params(hs, W, H, paths, vpl_paths, mis, seed, **kw):
    radius = float(hs.bounding_sphere_radius) * 0.003          # radiusPercentage of the bundled "ours" configs
    return E.make_params(hs.camera(), paths, vpl_paths, 3, radius, mis_mode=mis, clamp=float(1.0 / hs.total_area),
                         jitter=(0.37 / W, -0.21 / H), rng_seed=seed, **kw)


@pytest.mark.parametrize("mis", [capi.MIS_BALANCE, capi.MIS_GEOMETRY_CLAMP])
def test_config1_conference_256_whole_image_bit_exact(mis):
    W = H = 256
    paths, vpl_paths = 16384, 1024
    hs = HA.HostScene.generate("conference", 1, 8, W / H)
    scene = hs.to_scene()
    P = _params(hs, W, H, paths, vpl_paths, mis, seed=0)
    dev = E.Device(W, H)
    try:
        dev.upload_scene(scene); dev.build_bvh(); dev.set_params(P)
        dev.set_option("gather_chunks", 1)          # every pixel sums its VPLs in record order: the bit-exact mode
        dev.clear_accum()
        dev.gbuffer(); dev.light_trace(0, 0, paths); dev.vpl_gather(capi.GATHER_VPL); dev.photon_splat(0, paths * 4); dev.light_pass()
        vpl, photon, light = dev.download_accum()
        rec = dev.download_records(0, paths * 4)
        img = dev.resolve(1.0, 1.0, 1.0)
        st = dev.stats()
        orc = O.OracleScene(scene)
        orec = orc.light_trace(P, 0, 0, paths)
        assert rec.tobytes() == orec.tobytes()
        assert int((orec["flags"][: vpl_paths * 4] & 1).astype(bool).sum()) > 2000       # ~3 k usable VPLs of 4 k slots
        planes, prims = orc.gbuffer(P, W, H)
        g, cnt = orc.vpl_gather(P, W, H, planes, prims, orec, capi.GATHER_VPL)       # 3.9e8 pairs on the host cores
        ev = np.zeros((H, W, 3), dtype=np.int64); ep = np.zeros((H, W, 3), dtype=np.int64); el = np.zeros((H, W), dtype=np.uint32)
        orc.accumulate_fixed(g, ev)
        orc.photon_splat(P, W, H, planes, prims, orec, 0, len(orec), ep)
        orc.light_pass(P, W, H, el)
        assert np.array_equal(vpl, ev) and ev.any()
        assert np.array_equal(photon, ep) and ep.any()
        assert np.array_equal(light, el)
        assert img.tobytes() == orc.resolve(W, H, ev, ep, el, 1.0, 1.0, 1.0).tobytes()
        assert st.gatherPairs == int(cnt[0]) and st.shadowRays == int(cnt[1])
        # the default gather of this configuration, within the radiance tolerance
        dev.set_option("gather_chunks", 0)
        dev.clear_accum(); dev.vpl_gather(capi.GATHER_VPL)
        fast, _, _ = dev.download_accum()
        a, b = fast.astype(np.float64), ev.astype(np.float64)
        assert (np.abs(a - b) / (np.abs(b) + 1e-3 * np.abs(b).mean())).max() <= 1e-4
    finally:
        dev.close()


def test_config3_livingroom_1080p_vsl_256k_records_crop_bit_exact():
    W, H = 1920, 1080
    vpl_paths = 65536                                # 256 k VPL record slots per iteration
    hs = HA.HostScene.generate("livingroom", 1, 8, W / H)
    scene = hs.to_scene()
    vsl_radius = max(float(hs.bounding_sphere_radius) * 0.05, 0.008)     # vslRadiusPercentage 0.05 (rtcomphoton.h: floor 0.008)
    P = _params(hs, W, H, vpl_paths, vpl_paths, capi.MIS_BALANCE, seed=1, vsl_radius=vsl_radius, accumulate=False)
    dev = E.Device(W, H)
    try:
        dev.upload_scene(scene); dev.build_bvh(); dev.set_params(P)
        dev.gbuffer(); dev.light_trace(1, 0, vpl_paths)
        rec = dev.download_records(0, vpl_paths * 4)
        usable = int((rec["flags"] & 1).astype(bool).sum())
        assert usable > 150000
        x0, y0 = 928, 524
        rect = (x0, y0, x0 + 64, y0 + 32)
        dev.clear_accum(); dev.reset_stats()
        dev.vpl_gather(capi.GATHER_VSL, tile=rect)
        vpl, _, _ = dev.download_accum()
        st = dev.stats()
        assert st.gatherPairs == usable * 64 * 32
        planes, prims = dev.download_gbuffer()
        orc = O.OracleScene(scene)
        cx0, cy0, cw, ch = x0 + 24, y0 + 12, 16, 8
        g, cnt = orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VSL, tile=(cx0, cy0, cx0 + cw, cy0 + ch))
        ev = np.zeros((H, W, 3), dtype=np.int64)
        orc.accumulate_fixed(g, ev)
        assert ev[cy0:cy0 + ch, cx0:cx0 + cw].any()
        assert np.array_equal(vpl[cy0:cy0 + ch, cx0:cx0 + cw], ev[cy0:cy0 + ch, cx0:cx0 + cw])
    finally:
        dev.close()


def test_config4_buddha_4k_gather_crop_across_the_silhouette():
    W, H = 3840, 2160
    paths = vpl_paths = 1024
    hs = HA.HostScene.generate("buddha", 1, 8, W / H)
    scene = hs.to_scene()
    assert scene.num_prims > 1000000
    P = _params(hs, W, H, paths, vpl_paths, capi.MIS_BALANCE, seed=2, accumulate=False)
    dev = E.Device(W, H)
    try:
        dev.upload_scene(scene); dev.build_bvh(); dev.set_params(P)
        dev.gbuffer(); dev.light_trace(2, 0, paths)
        planes, prims = dev.download_gbuffer()
        rec = dev.download_records(0, paths * 4)
        # a 32x16 crop that straddles the statue's silhouette: the row of the image centre, first column where the primitive
        # under the pixel jumps between the statue (the bulk of the 1.06 M triangles) and the room behind it
        row = prims[H // 2]
        big = np.abs(np.diff(row.astype(np.int64))) > 100000
        xs = np.flatnonzero(big)
        assert len(xs) > 0
        x0 = int(np.clip(xs[0] - 16, 0, W - 32)); y0 = H // 2 - 8
        crop = (x0, y0, x0 + 32, y0 + 16)
        assert len(np.unique(prims[y0:y0 + 16, x0:x0 + 32])) > 20
        orc = O.OracleScene(scene)
        g, cnt = orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL, tile=crop)
        ev = np.zeros((H, W, 3), dtype=np.int64)
        orc.accumulate_fixed(g, ev)
        e = ev[y0:y0 + 16, x0:x0 + 32]
        assert e.any()
        dev.set_option("gather_chunks", 1)           # record order: bit for bit (whole 4K frame on the GPU, 2.5e10 pairs)
        dev.clear_accum(); dev.vpl_gather(capi.GATHER_VPL)
        exact, _, _ = dev.download_accum()
        assert np.array_equal(exact[y0:y0 + 16, x0:x0 + 32], e)
        for algo in (1, 2):                          # the default (per-VPL kernel at this VPL count) and the cluster gather
            dev.set_option("gather_chunks", 0); dev.set_option("gather_algo", algo)
            dev.clear_accum(); dev.vpl_gather(capi.GATHER_VPL)
            fast, _, _ = dev.download_accum()
            a, b = fast.astype(np.float64), exact.astype(np.float64)
            assert (np.abs(a - b) / (np.abs(b) + 1e-3 * np.abs(b).mean())).max() <= 1e-4, algo
            assert np.sqrt(np.mean((a - b) ** 2)) / np.abs(b).mean() <= 1e-5, algo
    finally:
        dev.close()
