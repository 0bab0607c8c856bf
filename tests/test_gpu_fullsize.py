"""Full-size GPU tests (BASELINE.json resolutions and scene sizes): size-independent properties
(tile partition == whole, chunked == unchunked within tolerance, tiled splat == scatter splat,
window sums) plus bit-exact oracle parity on a pixel crop of the full-size run."""
import ctypes as C

import numpy as np
import pytest

import evplp_b200 as E
from evplp_b200 import _capi as capi
from evplp_b200 import host_api as HA
from tests import oracle_api as O

pytestmark = pytest.mark.gpu
W, H = 1920, 1080
PATHS, VPL_PATHS = 100000, 48


@pytest.fixture(scope="module")
def conf():
    hs = HA.HostScene.generate("conference", 1, 8, W / H)
    scene = hs.to_scene()
    dev = E.Device(W, H)
    dev.upload_scene(scene)
    dev.build_bvh()
    radius = float(hs.bounding_sphere_radius) * 0.003
    P = E.make_params(hs.camera(), PATHS, VPL_PATHS, 3, radius, mis_mode=capi.MIS_BALANCE, clamp=float(1.0 / hs.total_area),
                      jitter=(0.25 / W, -0.4 / H), rng_seed=2)
    dev.set_params(P)
    dev.gbuffer()
    dev.light_trace(2, 0, PATHS)
    yield dict(hs=hs, scene=scene, dev=dev, P=P)
    dev.set_option("gather_chunks", 0)
    dev.set_option("splat_mode", 0)
    dev.close()


def test_conference_scene_is_baseline_sized(conf):
    info = conf["dev"].bvh_info()
    assert 300000 < info.numPrims < 360000
    st = conf["dev"].stats()
    assert st.emittedPhotons > PATHS * 2 and st.emittedVpls > PATHS * 2


def test_gather_tiles_and_chunks_at_1080p(conf):
    dev = conf["dev"]
    dev.set_option("gather_chunks", 1)
    dev.clear_accum()
    dev.vpl_gather(capi.GATHER_VPL)
    whole, _, _ = dev.download_accum()
    assert whole.any()
    dev.clear_accum()
    for tile in [(0, 0, 1000, 500), (1000, 0, W, 500), (0, 500, 333, H), (333, 500, W, H)]:
        dev.vpl_gather(capi.GATHER_VPL, tile=tile)
    parts, _, _ = dev.download_accum()
    assert np.array_equal(parts, whole)  # ragged image tiles (multi-GPU partition) reproduce the whole frame exactly
    dev.set_option("gather_algo", 0)   # exact-order kernel, VPL list in 3 ranges
    dev.set_option("gather_chunks", 3)
    dev.clear_accum()
    dev.vpl_gather(capi.GATHER_VPL)
    chunked, _, _ = dev.download_accum()
    dev.set_option("gather_chunks", 0)
    dev.set_option("gather_algo", 1)
    a, b = chunked.astype(np.float64), whole.astype(np.float64)
    # 1e-5 relative (north_star bar: 1e-4) + 4 units of the Q31.32 quantum (each chunk rounds its partial sum once)
    assert (np.abs(a - b) <= 4 + 1e-5 * np.abs(b)).all()
    conf["whole_vpl"] = whole
    # the default gather (VPL clusters, Morton summation order, FMA-contracted shading tail): whole frame at 1080p against
    # the exact-order frame -- per-pixel radiance within 1e-4 relative (north_star) and image relative RMSE <= 1e-5
    dev.set_option("gather_algo", 2)
    dev.clear_accum()
    dev.vpl_gather(capi.GATHER_VPL)
    dev.set_option("gather_algo", 1)
    fast, _, _ = dev.download_accum()
    a = fast.astype(np.float64)
    scale = np.abs(b).mean()
    rel = np.abs(a - b) / (np.abs(b) + 1e-3 * scale)
    assert rel.max() <= 1e-4, rel.max()
    assert np.sqrt(np.mean((a - b) ** 2)) / scale <= 1e-5


def test_gather_crop_matches_oracle_bit_for_bit(conf):
    dev, P = conf["dev"], conf["P"]
    planes, prims = dev.download_gbuffer()
    rec = dev.download_records(0, VPL_PATHS * 4)
    orc = O.OracleScene(conf["scene"])
    x0, y0, tw, th = 900, 500, 32, 16
    exp, _ = orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL, tile=(x0, y0, x0 + tw, y0 + th))
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    orc.accumulate_fixed(exp, eacc)
    # the per-ray packet traversal (gather_mode 0) gives the same image bit for bit as the default shaft traversal
    dev.set_option("gather_chunks", 1); dev.set_option("gather_mode", 0)
    dev.clear_accum(); dev.vpl_gather(capi.GATHER_VPL)
    shaft, _, _ = dev.download_accum()
    dev.set_option("gather_chunks", 0); dev.set_option("gather_mode", 1)
    assert np.array_equal(shaft[y0:y0 + th, x0:x0 + tw], eacc[y0:y0 + th, x0:x0 + tw])
    whole = conf.get("whole_vpl")
    if whole is not None:
        assert np.array_equal(shaft, whole)
    if whole is None:
        dev.set_option("gather_chunks", 1)
        dev.clear_accum(); dev.vpl_gather(capi.GATHER_VPL)
        whole, _, _ = dev.download_accum()
        dev.set_option("gather_chunks", 0)
    assert eacc[y0:y0 + th, x0:x0 + tw].any()
    assert np.array_equal(whole[y0:y0 + th, x0:x0 + tw], eacc[y0:y0 + th, x0:x0 + tw])
    # primary hits of the crop: the oracle's own G-buffer of those pixels equals the device's
    rays = np.zeros((tw * th, 8), dtype=np.float32)
    cam = conf["hs"].camera()
    k = 0
    for y in range(y0, y0 + th):
        for x in range(x0, x0 + tw):
            cx = np.float32((np.float32(x) + np.float32(0.5)) / np.float32(W) * np.float32(2) - np.float32(1))
            cy = np.float32((np.float32(y) + np.float32(0.5)) / np.float32(H) * np.float32(2) - np.float32(1))
            nx = np.float32((cx - np.float32(P.jitter[0])) * cam.tan_x)
            ny = np.float32((cy - np.float32(P.jitter[1])) * cam.tan_y)
            rays[k, 0:3] = cam.origin
            rays[k, 3:6] = cam.forward + cam.right * nx + cam.up * ny
            rays[k, 6], rays[k, 7] = 0.1, 100.0
            k += 1
    op, _ = orc.trace_rays(rays, 0)
    assert np.array_equal(op.reshape(th, tw), prims[y0:y0 + th, x0:x0 + tw])


def test_splat_tiled_equals_scatter_equals_oracle_crop(conf):
    dev, P = conf["dev"], conf["P"]
    n = PATHS * 4
    dev.set_option("splat_mode", 0)
    dev.clear_accum(); dev.reset_stats()
    dev.photon_splat(0, n)
    _, tiled, _ = dev.download_accum()
    st0 = dev.stats()
    dev.set_option("splat_mode", 1)
    dev.clear_accum(); dev.reset_stats()
    dev.photon_splat(0, n)
    _, scatter, _ = dev.download_accum()
    st1 = dev.stats()
    assert tiled.any() and np.array_equal(tiled, scatter)
    assert st0.splatFragments == st1.splatFragments and st0.splatPhotons == st1.splatPhotons
    # record windows x image tiles sum to the whole (path partition over GPUs)
    dev.set_option("splat_mode", 0)
    dev.clear_accum()
    cut = (n // 3 // 4) * 4
    for first, count in [(0, cut), (cut, n - cut)]:
        for tile in [(0, 0, 777, H), (777, 0, W, H)]:
            dev.photon_splat(first, count, tile=tile)
    _, parts, _ = dev.download_accum()
    assert np.array_equal(parts, tiled)
    # oracle on a crop
    planes, prims = dev.download_gbuffer()
    rec = dev.download_records(0, n)
    orc = O.OracleScene(conf["scene"], brute_force=True)  # the splat needs no ray tracing
    x0, y0, tw, th = 640, 300, 96, 64
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    orc.photon_splat(P, W, H, planes, prims, rec, 0, n, eacc, tile=(x0, y0, x0 + tw, y0 + th))
    assert eacc.any()
    assert np.array_equal(tiled[y0:y0 + th, x0:x0 + tw], eacc[y0:y0 + th, x0:x0 + tw])


def test_million_triangle_bvh_matches_cpu_twin_and_hits():
    """buddha-like statue, ~1.06 M triangles (BASELINE config 4): Morton codes, order and radix-tree
    topology equal the CPU twin; closest / any hits equal the oracle's on random rays."""
    hs = HA.HostScene.generate("buddha", 1, 8, 16 / 9)
    scene = hs.to_scene()
    assert scene.num_prims > 1000000
    dev = E.Device(64, 64)
    try:
        dev.upload_scene(scene)
        dev.build_bvh()
        codes, order, left, right, parent, bounds = dev.download_bvh()
        orc = O.OracleScene(scene)
        ocodes, oorder, oleft, oright, oparent, obounds, _ = orc.lbvh()
        assert np.array_equal(codes, ocodes) and np.array_equal(order, oorder)
        assert np.array_equal(left, oleft) and np.array_equal(right, oright) and np.array_equal(parent, oparent)
        assert np.array_equal(bounds.view(np.uint32), obounds.view(np.uint32))
        rs = np.random.RandomState(11)
        nr = 20000
        rays = np.zeros((nr, 8), dtype=np.float32)
        rays[:, 0:3] = rs.uniform([-4, -5, -1], [4, 3, 5], (nr, 3))
        tgt = rs.uniform([-1.2, -1.2, -1.2], [1.2, 1.2, 2.2], (nr, 3))  # aim at the statue
        rays[:, 3:6] = tgt - rays[:, 0:3]
        rays[:, 6], rays[:, 7] = 1e-4, 1e27
        gp, gt = dev.trace_rays(rays, 0)
        op, ot = orc.trace_rays(rays, 0)
        assert np.array_equal(gp, op) and np.array_equal(gt.view(np.uint32), ot.view(np.uint32))
        qp, qt = dev.trace_rays(rays, 3)   # quantised nodes (light paths): the same hits on the 1 M-triangle statue
        assert np.array_equal(qp, op) and np.array_equal(qt.view(np.uint32), ot.view(np.uint32))
        rays[:, 7] = 1 - 1e-4
        oa, _ = orc.trace_rays(rays, 1)
        for mode in (1, 2):
            ga, _ = dev.trace_rays(rays, mode)
            assert np.array_equal(ga, oa)
    finally:
        dev.close()


def test_bvh_leaf_size_does_not_change_hits():
    scene, cam = E.cornell_scene(seed=5, detail=6)
    rs = np.random.RandomState(2)
    rays = np.zeros((5000, 8), dtype=np.float32)
    rays[:, 0:3] = rs.uniform([0.5, 0.5, 0.5], [9.5, 9.5, 7.5], (5000, 3))
    rays[:, 3:6] = rs.normal(size=(5000, 3))
    rays[:, 6], rays[:, 7] = 1e-4, 1e27
    lib = capi.load_library()
    ref = None
    try:
        for leaf in (1, 2, 4, 8):
            capi.check(lib, lib.evplp_set_option(None, b"bvh_leaf_max", leaf), "opt")
            dev = E.Device(32, 32)
            dev.upload_scene(scene); dev.build_bvh()
            got = (dev.trace_rays(rays, 0), dev.trace_rays(rays * np.float32([1, 1, 1, 1, 1, 1, 1, 0]) + np.float32([0, 0, 0, 0, 0, 0, 0, 5.0]), 2))
            dev.close()
            if ref is None:
                ref = got
            else:
                assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[0][1], ref[0][1]) and np.array_equal(got[1][0], ref[1][0])
    finally:
        capi.check(lib, lib.evplp_set_option(None, b"bvh_leaf_max", 2), "opt")


def test_streamed_light_paths_equal_one_pass():
    """BASELINE config 5 machinery: light paths traced + splatted in chunks (record buffer smaller than the photon count)
    give exactly the one-pass accumulation layers."""
    W2, H2 = 640, 360
    hs = HA.HostScene.generate("conference", 1, 3, W2 / H2)
    fam = {"rngOffset": 1, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
           "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True, "useStat": False,
           "numLightPaths": 100000, "numVplLightPaths": 40, "numMaxBounces": 3, "radiusPercentage": 0.004, "misMode": "geometryClamp",
           "DoProgressive": True}
    lib = capi.load_library()
    out = []
    for chunk in (None, 30000):
        t = HA.Technique(hs, fam, W2, H2)
        if chunk:
            t.set_max_paths_per_trace(chunk)
        h = t.device_handle()
        capi.check(lib, lib.evplp_set_option(h, b"gather_chunks", 1), "opt")
        for _ in range(2):
            t.iterate()
        vpl = np.empty((H2, W2, 3), dtype=np.int64); ph = np.empty((H2, W2, 3), dtype=np.int64); li = np.empty((H2, W2), dtype=np.uint32)
        capi.check(lib, lib.evplp_download_accum(h, capi.ptr(vpl), capi.ptr(ph), capi.ptr(li)), "download")
        capi.check(lib, lib.evplp_set_option(h, b"gather_chunks", 0), "opt")
        t.close()
        out.append((vpl, ph, li))
    assert out[0][0].any() and out[0][1].any()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
