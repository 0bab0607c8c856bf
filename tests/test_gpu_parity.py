"""GPU parity tests: the sm_100a kernels (through the C ABI) against the CPU oracle on the
same seeded inputs.  Bit-exact for RNG streams, Morton codes, BVH topology, hits, records,
G-buffer and the fixed-point photon layer; the VPL layer is bit-exact when every pixel sums
its VPLs in record order (gather_chunks = 1) and within 1e-6 relative otherwise
(north_star tolerance: 1e-4 relative per pixel).
"""
import numpy as np
import pytest

import evplp_b200 as E
from evplp_b200 import _capi as capi
from tests import oracle_api as O

pytestmark = pytest.mark.gpu

W, H = 96, 64
NUM_PATHS, NUM_VPL_PATHS, BOUNCES = 2048, 96, 3


class Rig:
    def __init__(self, glossy=True, light_exponent=0.0, detail=5, seed=1):
        self.scene, cam = E.cornell_scene(seed=seed, detail=detail, glossy=glossy, light_exponent=light_exponent)
        self.camera = E.Camera(cam["origin"], cam["lookat"], cam["up"], cam["fovx"], W / H)
        self.radius = float(self.scene.bounding_sphere_radius()) * 0.02
        self.dev = E.Device(W, H)
        self.dev.upload_scene(self.scene)
        self.dev.build_bvh()
        self.orc = O.OracleScene(self.scene)

    def params(self, **kw):
        d = dict(num_light_paths=NUM_PATHS, num_vpl_paths=NUM_VPL_PATHS, max_bounces=BOUNCES, radius=self.radius,
                 mis_mode=capi.MIS_BALANCE, clamp=0.05, jitter=(0.3 / W, -0.2 / H), accumulate=True,
                 vsl_radius=float(self.scene.bounding_sphere_radius()) * 0.05, rng_seed=3)
        d.update(kw)
        return E.make_params(self.camera, **d)


@pytest.fixture(scope="module")
def rig():
    r = Rig()
    yield r
    r.dev.close()


def test_xorwow_streams_match_curand_and_oracle(rig):
    for seed, sub in [(0, 0), (1, 0), (0, 1), (3, 2), (12345, 7), (2047, 999), (65535, 123456), (7, 0xFFFFFFFF)]:
        a = rig.dev.debug_uniforms(seed, sub, 64)
        b = rig.dev.debug_curand(seed, sub, 64)
        c = O.uniforms(seed, sub, 64)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (seed, sub)
        assert np.array_equal(a.view(np.uint32), c.view(np.uint32)), (seed, sub)


def test_detmath_device_equals_host(rig):
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.uniform(0, 2 * np.pi, 20000), rs.uniform(-50, 50, 2000)]).astype(np.float32)
    for op in (0, 1):
        assert np.array_equal(rig.dev.debug_math(op, x).view(np.uint32), O.math_op(op, x).view(np.uint32))
    base = np.concatenate([rs.uniform(1e-6, 1, 20000), rs.uniform(1e-30, 1e-6, 1000), [1.0, 0.5, 1e-38]]).astype(np.float32)
    expo = np.concatenate([rs.uniform(0, 200, 10000), rs.uniform(0, 1, 10000), rs.uniform(0, 5000, 1003)]).astype(np.float32)
    assert np.array_equal(rig.dev.debug_math(2, base, expo).view(np.uint32), O.math_op(2, base, expo).view(np.uint32))
    u = rs.uniform(0, 1, 20000).astype(np.float32)
    assert np.array_equal(rig.dev.debug_math(3, u).view(np.uint32), O.math_op(3, u).view(np.uint32))
    assert np.array_equal(rig.dev.debug_math(4, u).view(np.uint32), O.math_op(4, u).view(np.uint32))


def test_lbvh_codes_order_topology_bounds(rig):
    codes, order, left, right, parent, bounds = rig.dev.download_bvh()
    ocodes, oorder, oleft, oright, oparent, obounds, smm = rig.orc.lbvh()
    info = rig.dev.bvh_info()
    assert np.array_equal(np.array(list(info.sceneMin) + list(info.sceneMax), dtype=np.float32), smm)
    assert np.array_equal(codes, ocodes)
    assert np.array_equal(order, oorder)
    assert np.array_equal(left, oleft)
    assert np.array_equal(right, oright)
    assert np.array_equal(parent, oparent)
    assert np.array_equal(bounds.view(np.uint32), obounds.view(np.uint32))


def _random_rays(scene, n, seed):
    rs = np.random.RandomState(seed)
    tris = scene.triangles()
    lo, hi = tris.reshape(-1, 3).min(axis=0), tris.reshape(-1, 3).max(axis=0)
    org = rs.uniform(lo + 0.05, hi - 0.05, size=(n, 3))
    # half the rays aim at random triangle points (grazing edges / vertices included)
    t = tris[rs.randint(0, len(tris), n)]
    bc = rs.dirichlet([1, 1, 1], n)
    bc[: n // 8] = np.eye(3)[rs.randint(0, 3, n // 8)]  # exact vertices
    tgt = (t * bc[:, :, None]).sum(axis=1)
    d = tgt - org
    d[n // 2:] = rs.normal(size=(n - n // 2, 3))
    d[-16:, 0] = 0.0  # axis-parallel components
    d[-8:, 1] = 0.0
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = org
    rays[:, 3:6] = d
    rays[:, 6] = 1e-4
    rays[:, 7] = 1e27
    return rays


def test_closest_and_any_hits_identical(rig):
    rays = _random_rays(rig.scene, 20000, 5)
    gp, gt = rig.dev.trace_rays(rays, 0)
    op, ot = rig.orc.trace_rays(rays, 0)
    assert np.array_equal(gp, op)
    assert np.array_equal(gt.view(np.uint32), ot.view(np.uint32))
    qp, qt = rig.dev.trace_rays(rays, 3)   # the quantised-node traversal of the light paths: looser boxes, the same exact hits
    assert np.array_equal(qp, op) and np.array_equal(qt.view(np.uint32), ot.view(np.uint32))
    srays = rays.copy()
    srays[:, 7] = 1.0 - 1e-4  # parametric shadow segments
    oa, _ = rig.orc.trace_rays(srays, 1)
    for mode in (1, 2):
        ga, _ = rig.dev.trace_rays(srays, mode)
        assert np.array_equal(ga, oa), mode
    assert 0.05 < oa.mean() < 0.95


def test_light_trace_records_bit_exact(rig):
    for seed, first in [(0, 0), (5, 0), (1000, 777)]:
        P = rig.params()
        rig.dev.set_params(P)
        rig.dev.light_trace(seed, first, NUM_PATHS)
        got = rig.dev.download_records(0, NUM_PATHS * (BOUNCES + 1))
        exp = rig.orc.light_trace(P, seed, first, NUM_PATHS)
        assert np.array_equal(got["flags"], exp["flags"])
        assert got.tobytes() == exp.tobytes()
        st = rig.dev.stats()
        assert st.emittedVpls == int(((exp["flags"] & 1) != 0).sum())
        assert st.emittedPhotons == int(((exp["flags"] & 2) != 0).sum())
        assert st.emittedPhotons > NUM_PATHS  # paths do bounce


def test_gbuffer_bit_exact(rig):
    P = rig.params()
    rig.dev.set_params(P)
    rig.dev.gbuffer()
    planes, prims = rig.dev.download_gbuffer()
    oplanes, oprims = rig.orc.gbuffer(P, W, H)
    assert np.array_equal(prims, oprims)
    assert planes.tobytes() == oplanes.tobytes()
    assert (prims >= 0).all()
    ms = rig.scene.mesh_starts()
    assert (prims >= ms[rig.scene.light_mesh]).sum() > 0  # the light is visible


def _setup_iteration(rig, P, seed=3):
    rig.dev.set_params(P)
    rig.dev.gbuffer()
    rig.dev.light_trace(seed, 0, P.numLightPaths)
    planes, prims = rig.orc.gbuffer(P, W, H)
    rec = rig.orc.light_trace(P, seed, 0, P.numLightPaths)
    return planes, prims, rec


@pytest.mark.parametrize("gather_mode", [0, 1])
@pytest.mark.parametrize("mis", [0, 1, 2, 3, 4, 5])
def test_vpl_gather_all_mis_modes(rig, mis, gather_mode):
    """gather_mode 0 = per-ray packet traversal, 1 = shaft traversal of the 32-wide hierarchy: both bit-exact."""
    rig.dev.set_option("gather_mode", gather_mode)
    P = rig.params(mis_mode=mis, accumulate=False)
    planes, prims, rec = _setup_iteration(rig, P)
    exp, cnt = rig.orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    rig.orc.accumulate_fixed(exp, eacc)
    assert eacc.sum() > 0
    # (a) record-order summation: bit-exact
    rig.dev.set_option("gather_chunks", 1)
    rig.dev.reset_stats()
    rig.dev.vpl_gather(capi.GATHER_VPL)
    vpl, _, _ = rig.dev.download_accum()
    assert np.array_equal(vpl, eacc)
    st = rig.dev.stats()
    assert st.gatherPairs == int(cnt[0]) and st.shadowRays == int(cnt[1])
    # (b) VPL list split across blocks (exact-order kernel, gather_algo 0): same image within rounding of the partial sums
    rig.dev.set_option("gather_algo", 0)
    rig.dev.set_option("gather_chunks", 5)
    rig.dev.vpl_gather(capi.GATHER_VPL)
    vpl5, _, _ = rig.dev.download_accum()
    rig.dev.set_option("gather_chunks", 0)
    rig.dev.set_option("gather_algo", 1)
    rig.dev.set_option("gather_mode", 1)  # back to the default
    a, b = vpl5.astype(np.float64), eacc.astype(np.float64)
    assert (np.abs(a - b) <= 6 + 1e-5 * np.abs(b)).all()  # 1e-5 relative + a few Q31.32 quanta (one rounding per chunk)


@pytest.mark.parametrize("chunks", [2, 5])
def test_chunked_gather_equals_the_oracle_chunk_by_chunk(rig, chunks):
    """gather_chunks = N (what balances small image shares in the multi-GPU image partition) splits the usable-VPL list into N
    contiguous ranges of ceil(total / N); each range is summed in record order in float by one thread per pixel, scaled by
    1 / numVplLightPaths, converted to Q31.32 and added with integer atomics.  That is deterministic, and it equals -- bit for
    bit -- the oracle run on each range alone (IsUsableVpl cleared on every other record) with the fixed-point images added."""
    P = rig.params(mis_mode=1, accumulate=False)
    planes, prims, rec = _setup_iteration(rig, P)
    prefix = int(P.numVplLightPaths) * int(P.numPhotonsPerLightPath)
    usable = np.flatnonzero(rec["flags"][:prefix] & 1)
    per = (len(usable) + chunks - 1) // chunks
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    for c in range(chunks):
        sub = rec.copy()
        drop = np.ones(prefix, dtype=bool)
        drop[usable[c * per:(c + 1) * per]] = False
        flags = sub["flags"]
        flags[:prefix][drop] &= np.uint32(0xFFFFFFFE)
        sub["flags"] = flags
        exp, _ = rig.orc.vpl_gather(P, W, H, planes, prims, sub, capi.GATHER_VPL)
        rig.orc.accumulate_fixed(exp, eacc)
    assert eacc.sum() > 0
    rig.dev.set_option("gather_algo", 0)   # the exact-order kernel (the cluster gather sums in Morton order: test_gpu_cluster.py)
    rig.dev.set_option("gather_chunks", chunks)
    try:
        for _ in range(2):  # twice: the result does not depend on the order the atomics land in
            rig.dev.vpl_gather(capi.GATHER_VPL)
            vpl, _, _ = rig.dev.download_accum()
            assert np.array_equal(vpl, eacc)
    finally:
        rig.dev.set_option("gather_chunks", 0)
        rig.dev.set_option("gather_algo", 1)


def test_vpl_gather_accumulates_and_tiles(rig):
    P = rig.params(mis_mode=4, accumulate=True)
    planes, prims, rec = _setup_iteration(rig, P)
    exp, _ = rig.orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    rig.orc.accumulate_fixed(exp, eacc)
    rig.dev.set_option("gather_chunks", 1)
    rig.dev.clear_accum()
    rig.dev.vpl_gather(capi.GATHER_VPL)
    rig.dev.vpl_gather(capi.GATHER_VPL)
    vpl, _, _ = rig.dev.download_accum()
    assert np.array_equal(vpl, 2 * eacc)
    # ragged tiles (multi-GPU partition of the image)
    rig.dev.clear_accum()
    for tile in [(0, 0, 37, H), (37, 0, W, 21), (37, 21, W, H)]:
        rig.dev.vpl_gather(capi.GATHER_VPL, tile=tile)
    vpl, _, _ = rig.dev.download_accum()
    rig.dev.set_option("gather_chunks", 0)
    assert np.array_equal(vpl, eacc)


@pytest.mark.parametrize("gather_mode", [0, 2])
def test_vsl_gather_bit_exact(rig, gather_mode):
    rig.dev.set_option("gather_mode", gather_mode)
    P = rig.params(mis_mode=0, accumulate=False, num_vpl_paths=24)
    planes, prims, rec = _setup_iteration(rig, P)
    exp, cnt = rig.orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VSL)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    rig.orc.accumulate_fixed(exp, eacc)
    assert eacc.sum() > 0
    rig.dev.reset_stats()
    rig.dev.vpl_gather(capi.GATHER_VSL)
    vpl, _, _ = rig.dev.download_accum()
    rig.dev.set_option("gather_mode", 1)
    assert np.array_equal(vpl, eacc)
    st = rig.dev.stats()
    assert st.gatherPairs == int(cnt[0]) and st.shadowRays == int(cnt[1])


def test_lvc_gather_bit_exact(rig):
    P = rig.params(mis_mode=1, accumulate=False, num_vpl_paths=16)
    planes, prims, rec = _setup_iteration(rig, P)
    exp, cnt = rig.orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_LVC)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    rig.orc.accumulate_fixed(exp, eacc)
    assert eacc.sum() > 0
    rig.dev.reset_stats()
    rig.dev.vpl_gather(capi.GATHER_LVC)
    vpl, _, _ = rig.dev.download_accum()
    assert np.array_equal(vpl, eacc)
    st = rig.dev.stats()
    assert st.gatherPairs == int(cnt[0]) and st.shadowRays == int(cnt[1])


@pytest.mark.parametrize("mis", [0, 1, 2, 3, 4, 5])
def test_photon_splat_fixed_point_exact(rig, mis):
    P = rig.params(mis_mode=mis)
    planes, prims, rec = _setup_iteration(rig, P)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    cnt = rig.orc.photon_splat(P, W, H, planes, prims, rec, 0, len(rec), eacc)
    assert cnt[1] > 1000
    rig.dev.clear_accum()
    rig.dev.reset_stats()
    rig.dev.photon_splat(0, len(rec))
    _, photon, _ = rig.dev.download_accum()
    assert np.array_equal(photon, eacc)
    st = rig.dev.stats()
    assert st.splatPhotons == int(cnt[0]) and st.splatFragments == int(cnt[1])


def test_photon_splat_windows_and_tiles_sum_to_whole(rig):
    P = rig.params(mis_mode=4)
    planes, prims, rec = _setup_iteration(rig, P)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    rig.orc.photon_splat(P, W, H, planes, prims, rec, 0, len(rec), eacc, brute_force=True)  # validates the screen-rect cull too
    rig.dev.clear_accum()
    n = len(rec)
    cut = (n // 3 // 4) * 4
    for first, count in [(0, cut), (cut, n - cut)]:
        for tile in [(0, 0, 50, H), (50, 0, W, H)]:
            rig.dev.photon_splat(first, count, tile=tile)
    _, photon, _ = rig.dev.download_accum()
    assert np.array_equal(photon, eacc)


def test_zero_radius_and_zero_vpl_modes(rig):
    # vpl.json: radiusPercentage 0 -> no fragments
    P = rig.params(radius=0.0)
    _setup_iteration(rig, P)
    rig.dev.clear_accum()
    rig.dev.photon_splat(0, NUM_PATHS * (BOUNCES + 1))
    _, photon, _ = rig.dev.download_accum()
    assert not photon.any()
    # pm.json: numVplLightPaths 0 -> the gather is refused like the reference disables it
    P = rig.params(num_vpl_paths=0)
    rig.dev.set_params(P)
    with pytest.raises(capi.EvplpError):
        rig.dev.vpl_gather(capi.GATHER_VPL)


def test_full_iterations_accumulate_and_resolve(rig):
    """Three energy-compensated iterations (gbuffer -> trace -> gather -> splat -> light) with
    per-iteration jitter and seeds, then resolve: final image equals the oracle's bit for bit."""
    rig.dev.set_option("gather_chunks", 1)
    rig.dev.clear_accum()
    ev = np.zeros((H, W, 3), dtype=np.int64); ep = np.zeros((H, W, 3), dtype=np.int64); el = np.zeros((H, W), dtype=np.uint32)
    jit = np.empty(6, dtype=np.float32)
    O.load().orc_jitter_stream(0, 3, capi.ptr(jit))
    for it in range(3):
        j = ((2 * jit[2 * it] - 1) / W, (2 * jit[2 * it + 1] - 1) / H)
        P = rig.params(mis_mode=4, jitter=j, rng_seed=it, num_light_paths=1024)
        rig.dev.set_params(P)
        rig.dev.gbuffer()
        rig.dev.light_trace(it, 0, 1024)
        rig.dev.vpl_gather(capi.GATHER_VPL)
        rig.dev.photon_splat(0, 1024 * (BOUNCES + 1))
        rig.dev.light_pass()
        planes, prims = rig.orc.gbuffer(P, W, H)
        rec = rig.orc.light_trace(P, it, 0, 1024)
        img, _ = rig.orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL)
        rig.orc.accumulate_fixed(img, ev)
        rig.orc.photon_splat(P, W, H, planes, prims, rec, 0, len(rec), ep)
        rig.orc.light_pass(P, W, H, el)
    rig.dev.set_option("gather_chunks", 0)
    vpl, photon, light = rig.dev.download_accum()
    assert np.array_equal(vpl, ev) and np.array_equal(photon, ep) and np.array_equal(light, el)
    s = 1.0 / 3.0
    got = rig.dev.resolve(s, s, 1.0)
    exp = rig.orc.resolve(W, H, ev, ep, el, s, s, 1.0)
    assert got.tobytes() == exp.tobytes()
    assert np.isfinite(got).all() and got.mean() > 0.01
    gg = rig.dev.resolve(s, s, 1.0, gamma=True)
    eg = rig.orc.resolve(W, H, ev, ep, el, s, s, 1.0, gamma=True)
    assert np.allclose(gg, eg, rtol=1e-5, atol=1e-6)


def test_directional_light_and_diffuse_scene():
    """buddha-style emission exponent (intensity.w = 50) and a purely diffuse scene."""
    r = Rig(glossy=False, light_exponent=50.0, detail=3, seed=9)
    try:
        P = r.params(mis_mode=1)
        r.dev.set_params(P)
        r.dev.light_trace(11, 0, 512)
        got = r.dev.download_records(0, 512 * (BOUNCES + 1))
        exp = r.orc.light_trace(P, 11, 0, 512)
        assert got.tobytes() == exp.tobytes()
    finally:
        r.dev.close()


@pytest.mark.parametrize("bounces", [1, 3, 6])
def test_path_tracer_bit_exact(rig, bounces):
    """RtPt2 (pathtracing.cu): one path per pixel with NEE + MIS; per-pixel XORWOW streams; bit-exact incl. ray counts."""
    P = rig.params(accumulate=True, rng_seed=11)
    rig.dev.set_params(P)
    rig.dev.gbuffer()
    planes, prims = rig.orc.gbuffer(P, W, H)
    exp, cnt = rig.orc.path_trace(P, W, H, planes, prims, bounces)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    rig.orc.accumulate_fixed(exp, eacc)
    rig.orc.accumulate_fixed(exp, eacc)
    assert eacc.sum() > 0
    rig.dev.clear_accum()
    rig.dev.reset_stats()
    rig.dev.path_trace(bounces)
    rig.dev.path_trace(bounces, tile=(0, 0, 50, H))   # accumulates (doAccumulate = 1), ragged tiles
    rig.dev.path_trace(bounces, tile=(50, 0, W, H))
    vpl, _, _ = rig.dev.download_accum()
    assert np.array_equal(vpl, eacc)
    assert rig.dev.stats().closestRays == 2 * int(cnt[0])
