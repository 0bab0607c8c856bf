"""Physics-level check that does not depend on the oracle: energy compensation.  Clamping the VPL geometry term
(misMode geometryClamp) loses energy; the photon splat weighted by max(G - c, 0) / G restores it, so the converged
"ours" image carries the same total energy as the unclamped VPL image (misMode one), while the clamped gather alone
is darker (lighttracing.cu:340, photonsplatinstanced.frag:222, README of the reference)."""
import numpy as np
import pytest

from evplp_b200 import host_api as HA

pytestmark = pytest.mark.gpu
W, H, ITERS = 160, 90, 96


def _render(hs, **kw):
    fam = {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
           "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": True, "useStat": False,
           "numLightPaths": 40000, "numVplLightPaths": 128, "numMaxBounces": 3, "radiusPercentage": 0.02, "DoProgressive": False}
    fam.update(kw)
    t = HA.Technique(hs, fam, W, H)
    for _ in range(ITERS):
        t.iterate()
    s = 1.0 / ITERS
    vpl = t.final(s, 0.0, 0.0)
    photon = t.final(0.0, s, 0.0)
    t.close()
    return vpl.astype(np.float64), photon.astype(np.float64)


def test_clamp_plus_compensation_conserves_energy():
    hs = HA.HostScene.generate("livingroom", 4, 2, W / H)
    ref_vpl, _ = _render(hs, misMode="one", radiusPercentage=0.0)                       # unclamped VPLs: unbiased
    cl_vpl, cl_photon = _render(hs, misMode="geometryClamp", clampingCoeff=0.02)         # "ours"
    e_ref, e_clamped, e_comp = ref_vpl.sum(), cl_vpl.sum(), cl_photon.sum()
    assert e_ref > 0 and e_comp > 0
    assert e_clamped < 0.97 * e_ref                       # the clamp really removes energy in this scene
    assert abs((e_clamped + e_comp) - e_ref) < 0.05 * e_ref  # and the splat puts it back (density-estimation bias << 5 %)
    # balance-heuristic MIS (the bundled default) combines the two estimators with weights that sum to one as well
    mis_vpl, mis_photon = _render(hs, misMode="balance")
    assert abs((mis_vpl.sum() + mis_photon.sum()) - e_ref) < 0.05 * e_ref


def test_path_tracer_agrees_with_vpl_and_energy_compensated_renders():
    """Independent ground truth: RtPt2 (NEE + MIS path tracer, numMaxBounces 3) and the VPL / EVPLP renders cover the same
    light transport (direct + two indirect bounces) and must converge to the same image energy."""
    hs = HA.HostScene.generate("livingroom", 4, 2, W / H)
    pt = HA.PathTracer(hs, {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "outputFilename": "pt.pfm",
                            "statFilename": "s.json", "useJitter": True, "useStat": False, "numSamplePerPixel": 1, "numMaxBounces": 3}, W, H)
    n = 256
    for _ in range(n):
        pt.iterate()
    img_pt = pt.final(1.0 / n, 0.0).astype(np.float64)
    pt.close()
    ref_vpl, _ = _render(hs, misMode="one", radiusPercentage=0.0)
    ours_vpl, ours_photon = _render(hs, misMode="geometryClamp", clampingCoeff=0.02)
    e_pt, e_vpl, e_ours = img_pt.sum(), ref_vpl.sum(), (ours_vpl + ours_photon).sum()
    assert e_pt > 0
    assert abs(e_vpl - e_pt) < 0.05 * e_pt
    assert abs(e_ours - e_pt) < 0.06 * e_pt
    # and pixel-wise on a 4x4-blocked image (averages out the Monte-Carlo noise): relative RMSE small
    def blocks(a):
        return a[: H // 6 * 6, : W // 8 * 8].reshape(H // 6, 6, W // 8, 8, 3).mean(axis=(1, 3))
    a, b = blocks(img_pt), blocks(ours_vpl + ours_photon)
    rel = np.sqrt(((a - b) ** 2).mean()) / a.mean()
    assert rel < 0.25
