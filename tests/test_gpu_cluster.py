"""The default VPL gather (gather_fast.cu: VPL clusters + double shafts, Morton summation order, FMA-contracted shading tail)
against the CPU oracle and against the exact-order kernel.  Visibility is exact (same triangle test); radiance must stay
inside the north_star tolerance: per-pixel linear radiance within 1e-4 relative, image relative RMSE <= 1e-5.
"""
import numpy as np
import pytest

import evplp_b200 as E
from evplp_b200 import _capi as capi
from tests import oracle_api as O
from tests.test_gpu_parity import H, W, Rig, _setup_iteration

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rig():
    r = Rig()
    r.dev.set_option("gather_algo", 2)   # force the cluster gather (by default it starts at 16384 usable VPLs)
    yield r
    r.dev.close()


def _errors(vpl, eacc):
    a, b = vpl.astype(np.float64), eacc.astype(np.float64)
    scale = np.abs(b).mean()
    rel = np.abs(a - b) / (np.abs(b) + 1e-3 * scale)
    return rel.max(), np.sqrt(np.mean((a - b) ** 2)) / scale


def test_fast_pow_accuracy(rig):
    rs = np.random.RandomState(3)
    base = np.concatenate([rs.uniform(1e-6, 1, 200000), 1 - rs.uniform(0, 1e-3, 50000), [1.0, 0.5, 1e-6]]).astype(np.float32)
    expo = np.concatenate([rs.uniform(0, 200, 125000), rs.uniform(0, 4, 125000), [0.0, 1.0, 1000.0]]).astype(np.float32)
    got = rig.dev.debug_math(5, base, expo).astype(np.float64)
    exact = np.power(base.astype(np.float64), expo.astype(np.float64))
    big = exact > 1e-12   # smaller values are invisible next to the diffuse term they are added to
    assert (np.abs(got - exact)[big] / exact[big]).max() < 1e-5
    assert np.abs(got - exact)[~big].max() < 1e-12 * 1.001 + 1e-17


@pytest.mark.parametrize("chunks", [0, 3])
@pytest.mark.parametrize("mis", [0, 1, 2, 3, 4, 5])
def test_cluster_gather_within_tolerance_of_the_oracle(rig, mis, chunks):
    P = rig.params(mis_mode=mis, accumulate=False)
    planes, prims, rec = _setup_iteration(rig, P)
    exp, cnt = rig.orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    rig.orc.accumulate_fixed(exp, eacc)
    rig.dev.set_option("gather_chunks", chunks)
    rig.dev.reset_stats()
    try:
        for rep in range(2):   # the second launch draws the items in longest-first order: same image, bit for bit
            rig.dev.vpl_gather(capi.GATHER_VPL)
            vpl, _, _ = rig.dev.download_accum()
            worst, rmse = _errors(vpl, eacc)
            assert worst <= 1e-4, (mis, worst)
            assert rmse <= 1e-5, (mis, rmse)
            if rep == 0:
                first = vpl
            else:
                assert np.array_equal(vpl, first)
    finally:
        rig.dev.set_option("gather_chunks", 0)
    st = rig.dev.stats()
    assert st.gatherPairs == 2 * int(cnt[0])
    assert st.shadowRays == 2 * int(cnt[1])   # the same pairs pass the cosine test


def test_cluster_gather_accumulates_tiles_and_small_clusters(rig):
    P = rig.params(mis_mode=1, accumulate=True)
    planes, prims, rec = _setup_iteration(rig, P)
    exp, _ = rig.orc.vpl_gather(P, W, H, planes, prims, rec, capi.GATHER_VPL)
    eacc = np.zeros((H, W, 3), dtype=np.int64)
    rig.orc.accumulate_fixed(exp, eacc)
    rig.dev.clear_accum()
    rig.dev.vpl_gather(capi.GATHER_VPL)
    rig.dev.vpl_gather(capi.GATHER_VPL)
    vpl, _, _ = rig.dev.download_accum()
    assert _errors(vpl, 2 * eacc)[0] <= 1e-4
    rig.dev.clear_accum()
    for tile in [(0, 0, 37, H), (37, 0, W, 21), (37, 21, W, H)]:   # ragged rectangles
        rig.dev.vpl_gather(capi.GATHER_VPL, tile=tile)
    vpl, _, _ = rig.dev.download_accum()
    assert _errors(vpl, eacc)[0] <= 1e-4
    for cs in (1, 5, 16):
        rig.dev.set_option("gather_cluster_size", cs)
        rig.dev.clear_accum()
        rig.dev.vpl_gather(capi.GATHER_VPL)
        vpl, _, _ = rig.dev.download_accum()
        assert _errors(vpl, eacc)[0] <= 1e-4, cs
    rig.dev.set_option("gather_cluster_size", 16)
    # cluster layout: runs of 16 cut into halves / quarters where their box is large (1 = every run is cut, 0 = never, 1000 = whole scene)
    for permille in (1, 0, 1000, 70):
        rig.dev.set_option("gather_cluster_extent_permille", permille)
        rig.dev.clear_accum()
        rig.dev.vpl_gather(capi.GATHER_VPL)
        vpl, _, _ = rig.dev.download_accum()
        assert _errors(vpl, eacc)[0] <= 1e-4, permille


def test_two_handles_partition_the_image_between_them(rig):
    """Options are per handle: two handles on one GPU, tile stride 2 / offsets 0 and 1, each render their tiles of the
    frame; the sum of their layers is the frame of a single handle, bit for bit (same item grid per pixel)."""
    P = rig.params(mis_mode=4, accumulate=False)
    devs = []
    try:
        for off in (0, 1):
            d = E.Device(W, H)
            d.upload_scene(rig.scene); d.build_bvh(); d.set_params(P)
            d.set_option("gather_band_stride", 2); d.set_option("gather_band_offset", off); d.set_option("gather_chunks", 2)
            d.set_option("gather_algo", 2)
            d.gbuffer(); d.light_trace(3, 0, P.numLightPaths)
            d.clear_accum(); d.vpl_gather(capi.GATHER_VPL)
            devs.append(d)
        parts = [d.download_accum()[0] for d in devs]
        assert parts[0].any() and parts[1].any()
        assert not (parts[0].astype(bool) & parts[1].astype(bool)).any()   # disjoint pixels
        _setup_iteration(rig, P)
        rig.dev.set_option("gather_chunks", 2)
        rig.dev.clear_accum(); rig.dev.vpl_gather(capi.GATHER_VPL)
        whole, _, _ = rig.dev.download_accum()
        assert np.array_equal(parts[0] + parts[1], whole)
        # and the exact-order kernel under the same per-handle options (16-row bands)
        for d in devs:
            d.set_option("gather_chunks", 1); d.clear_accum(); d.vpl_gather(capi.GATHER_VPL)
        rig.dev.set_option("gather_chunks", 1); rig.dev.clear_accum(); rig.dev.vpl_gather(capi.GATHER_VPL)
        whole1, _, _ = rig.dev.download_accum()
        assert np.array_equal(devs[0].download_accum()[0] + devs[1].download_accum()[0], whole1)
    finally:
        rig.dev.set_option("gather_chunks", 0)
        for d in devs:
            d.close()


def test_options_are_range_checked(rig):
    lib = capi.load_library()
    for name, bad in ((b"gather_chunks", -1), (b"gather_cluster_size", 0), (b"gather_cluster_size", 99), (b"splat_group", 3),
                      (b"bvh_leaf_max", 9), (b"gather_band_stride", -2)):
        assert lib.evplp_set_option(rig.dev.h, name, bad) != 0
    assert lib.evplp_set_option(rig.dev.h, b"no_such_option", 1) != 0
