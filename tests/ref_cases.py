"""The fixed inputs on which the oracle is pinned against the reference's own device code, and the comparison
rules.  Shared by tests/test_gpu_reference.py (live, on the B200: oracle vs oracle/_ref/libref_device.so),
tests/test_ref_goldens.py (CPU: oracle vs the outputs that library wrote on a B200, tests/golden/ref_device/)
and scripts/make_ref_goldens.py (the generator of those files)."""
import numpy as np

import evplp_b200 as E
from evplp_b200 import _capi as capi

W, H = 96, 64
NUM_PATHS, NUM_VPL_PATHS, BOUNCES = 512, 64, 3
f32 = np.float32


def rig_scene(glossy=True, light_exponent=0.0):
    scene, cam = E.cornell_scene(seed=1, detail=5, glossy=glossy, light_exponent=light_exponent)
    camera = E.Camera(cam["origin"], cam["lookat"], cam["up"], cam["fovx"], W / H)
    return scene, camera


def rig_params(scene, camera, **kw):
    r = float(scene.bounding_sphere_radius())
    d = dict(num_light_paths=NUM_PATHS, num_vpl_paths=NUM_VPL_PATHS, max_bounces=BOUNCES, radius=r * 0.02,
             mis_mode=capi.MIS_BALANCE, clamp=0.05, jitter=(0.3 / W, -0.2 / H), accumulate=False, vsl_radius=r * 0.05, rng_seed=3)
    d.update(kw)
    return E.make_params(camera, **d)


def _unit(rs, n):
    v = rs.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def brdf_inputs(op, n=2048, seed=11):
    """16 floats per item: a, b, c, refl, exponent, seed, subsequence, ks.x (see oracle/ref_device.cu:k_brdf)."""
    rs = np.random.RandomState(seed + op)
    q = np.zeros((n, 16), dtype=f32)
    nrm = _unit(rs, n)
    q[:, 12] = rs.choice([0.0, 1.0, 5.0, 20.0, 50.0, 200.0], n) + rs.uniform(0, 1, n) * (rs.uniform(size=n) < 0.5)
    q[:, 13] = rs.randint(0, 100000, n)
    q[:, 14] = rs.randint(0, 1000, n)
    q[:, 15] = rs.choice([0.0, 0.3, 1.0], n, p=[0.1, 0.45, 0.45])
    if op in (0, 1):      # Sample(out, pdf, in = a, normal = b, reflectance = refl, e)
        inn = _unit(rs, n)
        inn[(inn * nrm).sum(1) < 0] *= -1   # incoming direction on the normal's side
        q[:, 0:3], q[:, 3:6], q[:, 9:12] = inn, nrm, rs.uniform(0, 1, (n, 3))
    elif op in (2, 3):    # PdfA(n1 = a, n2 = b, v12 = c, in = refl, ks, e)
        q[:, 0:3], q[:, 3:6] = nrm, _unit(rs, n)
        q[:, 6:9] = _unit(rs, n) * rs.uniform(0.05, 8.0, (n, 1))
        q[:, 9:12] = _unit(rs, n)
    elif op == 4:         # PhongEvalF(out = a, in = b, normal = c, e), PhongEval(.., ks = refl, e)
        q[:, 0:3], q[:, 3:6], q[:, 6:9], q[:, 9:12] = _unit(rs, n), _unit(rs, n), nrm, rs.uniform(0, 1, (n, 3))
        q[: n // 8, 9] = 0.0
    else:                 # SquareToBarycentric(x, y), SquareToSolidAngle(x, y, halfAngle), russianProb(b)
        q[:, 0:2] = rs.uniform(0, 1, (n, 2))
        q[:, 2] = rs.uniform(0.001, np.pi / 2, n)
        q[:, 3:6] = rs.uniform(0, 1.5, (n, 3))
    return q


def random_rays(scene, n, seed):
    rs = np.random.RandomState(seed)
    tris = scene.triangles()
    lo, hi = tris.reshape(-1, 3).min(axis=0), tris.reshape(-1, 3).max(axis=0)
    org = rs.uniform(lo + 0.05, hi - 0.05, size=(n, 3))
    t = tris[rs.randint(0, len(tris), n)]
    bc = rs.dirichlet([1, 1, 1], n)
    tgt = (t * bc[:, :, None]).sum(axis=1)
    d = tgt - org
    d[n // 2:] = rs.normal(size=(n - n // 2, 3))
    rays = np.zeros((n, 8), dtype=f32)
    rays[:, 0:3], rays[:, 3:6], rays[:, 6], rays[:, 7] = org, d, 1e-4, 1e27
    return rays


# ---- comparison rules -------------------------------------------------------------------------------------------------
# The reference build contracts multiply-adds into FMAs and calls libdevice powf / sinf / cosf / asinf; the oracle rounds
# every operation (detmath.h).  So: everything integer (flags, counts, RNG consumption, primitive ids) must be EQUAL
# (up to the rounding-decided events listed at each check), floats must agree to 1e-5 relative, widened only where the
# arithmetic itself amplifies a 1-ulp input difference: x^e multiplies the relative error of x by e, and unit-scale results
# computed by cancellation (1 - z*z, cos near 0) carry an absolute error of a few ulp(1).
REL = 1e-5
# per BRDF tap: (relative tolerance, + per unit of Phong exponent, absolute floor)
BRDF_TOL = {0: (1e-5, 0.0, 4e-6), 1: (1e-5, 1e-6, 4e-6), 2: (1e-5, 0.0, 1e-7), 3: (1e-5, 1e-6, 1e-7), 4: (1e-5, 1e-6, 1e-7), 5: (1e-5, 0.0, 4e-6)}


def close(a, b, rel=REL, floor=1e-7):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= rel * np.maximum(np.abs(a), np.abs(b)) + floor


REC_FIELDS = (("position", 1.0), ("normal", 1.0), ("flux", 1e-3), ("fluxDir", 1.0), ("lambertReflectance", 1.0),
              ("phongReflectance", 1.0), ("pSelectLambert", 1.0), ("phongExponent", 1.0))


def compare_records(ref, orc, b1):
    """Per path: are all flag words equal; and, where both wrote a record, does every field agree to 5e-4 of the field's
    magnitude.  Errors grow along a path: every bounce re-samples a direction from values that already differ by ulps, and
    a hit point that moves by 1e-6 across a checker-texture edge changes the bilinear Kd (and with it the flux) by 1e-4."""
    rf, of = ref["flags"].reshape(-1, b1), orc["flags"].reshape(-1, b1)
    same = (rf == of).all(axis=1)
    live = (ref["flags"] != 0) & (orc["flags"] != 0)
    ok = np.ones(len(ref), dtype=bool)
    worst = 0.0
    for name, floor in REC_FIELDS:
        a, b = ref[name].astype(np.float64), orc[name].astype(np.float64)
        if a.ndim == 1:
            a, b = a[:, None], b[:, None]
        mag = np.maximum(np.maximum(np.abs(a).max(axis=1), np.abs(b).max(axis=1)), floor)
        e = np.abs(a - b).max(axis=1) / mag
        worst = max(worst, float(e[live].max()) if live.any() else 0.0)
        ok &= (e <= 5e-4) | ~live
    return same, ok.reshape(-1, b1).all(axis=1), worst


# ---- the cases ---------------------------------------------------------------------------------------------------------
TRACE_CASES = [(0, 0), (5, 0), (1000, 777)]      # (rngSeed, firstPath)
VARIANTS = {"glossy": dict(glossy=True, light_exponent=0.0), "spot": dict(glossy=True, light_exponent=50.0)}
VSL_VPL_PATHS = 12
PT_SPP, PT_BOUNCES = 4, 4


def reference_outputs():
    """Run every case through the reference's device code (needs a GPU).  Returns {name: array}."""
    from tests import oracle_api as O
    from tests import ref_device_api as D

    out = {}
    for op in range(6):
        out["brdf%d" % op] = D.brdf(op, brdf_inputs(op))
    for vname, vkw in VARIANTS.items():
        scene, camera = rig_scene(**vkw)
        ref = D.RefScene(scene)
        orc = O.OracleScene(scene)
        P = rig_params(scene, camera)
        b1 = P.numPhotonsPerLightPath
        for seed, first in TRACE_CASES:
            out["%s_records_%d_%d" % (vname, seed, first)] = ref.trace_photons(seed, first, NUM_PATHS, b1)
        if vname == "glossy":
            rays = random_rays(scene, 4000, 5)
            out["rays_closest_prim"], out["rays_closest_t"] = ref.trace_rays(rays, 0)
            srays = rays.copy(); srays[:, 7] = 1.0 - 1e-4
            out["rays_any"], _ = ref.trace_rays(srays, 1)
        planes, prims = orc.gbuffer(P, W, H)
        recs = out["%s_records_5_0" % vname]
        for mode in range(6):
            Pm = rig_params(scene, camera, mis_mode=mode)
            out["%s_splatColor_mode%d" % (vname, mode)] = ref.gather(Pm, W, H, planes, recs[: NUM_VPL_PATHS * b1])[:, :, :3].copy()
        Pv = rig_params(scene, camera, num_vpl_paths=VSL_VPL_PATHS)
        out["%s_splatSplotch" % vname] = ref.gather(Pv, W, H, planes, recs[: VSL_VPL_PATHS * b1], vsl=True)[:, :, :3].copy()
        # accumulate on top of a previous image (doAccumulate * outputBuffer, lighttracing.cu:378)
        Pa = rig_params(scene, camera, accumulate=True)
        prev = np.full((H, W, 4), 0.25, dtype=f32)
        out["%s_splatColor_accumulate" % vname] = ref.gather(Pa, W, H, planes, recs[: NUM_VPL_PATHS * b1], prev=prev)[:, :, :3].copy()
        ref.close()
        # RtPt2 (pathtracing.cu), PT_SPP independent samples per pixel summed: streams (pixel, rngSeed = 0 .. PT_SPP - 1)
        acc = np.zeros((H, W, 4), dtype=f32)
        for k in range(PT_SPP):
            acc = D.path_trace(scene, rig_params(scene, camera, rng_seed=k, accumulate=True), W, H, planes, PT_BOUNCES, prev=acc)
        out["%s_pathtrace" % vname] = acc[:, :, :3].copy()
    return out


def check_oracle(outputs, report=None):
    """Compare the CPU oracle with `outputs` (from reference_outputs(), live or loaded from tests/golden/ref_device).
    Raises AssertionError on the first broken rule; fills `report` (dict) with the measured agreement."""
    from tests import oracle_api as O

    rep = report if report is not None else {}
    lib = O.load()
    for op in range(6):
        q = brdf_inputs(op)
        mine = np.empty((len(q), 8), dtype=f32)
        lib.orc_brdf(op, capi.ptr(q), len(q), capi.ptr(mine))
        ref = outputs["brdf%d" % op]
        if op in (0, 1):   # o[7] = the uniform drawn AFTER the sampler: equal bits <=> same number of draws, same stream
            assert np.array_equal(mine[:, 7].view(np.uint32), ref[:, 7].view(np.uint32)), "RNG consumption of sampler %d" % op
        rel, per_e, floor = BRDF_TOL[op]
        tol = (rel + per_e * q[:, 12:13].astype(np.float64)) * np.maximum(np.abs(mine), np.abs(ref)) + floor
        err = np.abs(mine.astype(np.float64) - ref)
        rep["brdf%d_max_err_over_tol" % op] = float((err / tol).max())
        assert (err <= tol).all(), ("brdf op %d" % op, int((err > tol).sum()), mine[(err > tol).any(axis=1)][:3], ref[(err > tol).any(axis=1)][:3])
    for vname, vkw in VARIANTS.items():
        scene, camera = rig_scene(**vkw)
        orc = O.OracleScene(scene)
        P = rig_params(scene, camera)
        b1 = P.numPhotonsPerLightPath
        for seed, first in TRACE_CASES:
            ref = outputs["%s_records_%d_%d" % (vname, seed, first)]
            mine = orc.light_trace(P, seed, first, NUM_PATHS)
            same, ok, worst = compare_records(ref, mine, b1)
            rep["%s_records_%d_%d" % (vname, seed, first)] = dict(
                paths=len(same), flags_equal=int(same.sum()), floats_ok=int((same & ok).sum()), worst_field_error=worst,
                vpls=int((ref["flags"] & 1).astype(bool).sum()), photons=int((ref["flags"] & 2).astype(bool).sum()),
                oracle_vpls=int((mine["flags"] & 1).astype(bool).sum()), oracle_photons=int((mine["flags"] & 2).astype(bool).sum()))
            # Every flag word of every path must be equal, except for paths whose fate is decided by a rounding: the known
            # case is a photon leaving the light at a grazing angle -- the sampled point pos1*b + pos2*g + pos3*(1-g-b)
            # (rtlightsource.cuh:74) lands exactly on the light's plane with FMA contraction and 1 ulp beside it without, and
            # from 1 ulp above, the ray re-hits the light's own triangle beyond tmin and ends (lighttracing.cu:124).  At most
            # 1 path in 256 may differ; the float fields of all other paths must agree.
            assert same.sum() >= len(same) - len(same) // 256, (vname, seed, first, int((~same).sum()))
            assert (ok | ~same).all(), (vname, seed, first, int((~ok & same).sum()))
        if vname == "glossy":
            rays = random_rays(scene, 4000, 5)
            p, t = orc.trace_rays(rays, 0)
            rp, rt = outputs["rays_closest_prim"], outputs["rays_closest_t"]
            strict = (p == rp)
            # the rig has coincident faces (box bottoms lying in the floor plane): there, which of the two coplanar triangles
            # reports the smaller t is decided by the last ulp.  Same hit point, different id -- counted separately.
            tied = ~strict & (p >= 0) & (rp >= 0) & close(t, rt, rel=1e-6, floor=0.0)
            rep["rays_closest_prim_equal"] = float(strict.mean())
            rep["rays_closest_agree"] = float((strict | tied).mean())
            assert (strict | tied).mean() >= 0.9999, (strict.mean(), tied.mean())
            assert strict.mean() >= 0.998
            assert close(t[strict], rt[strict], rel=1e-5, floor=0.0).all()
            srays = rays.copy(); srays[:, 7] = 1.0 - 1e-4
            a, _ = orc.trace_rays(srays, 1)
            rep["rays_any_agree"] = float((a == outputs["rays_any"]).mean())
            assert (a == outputs["rays_any"]).mean() >= 0.9999
        planes, prims = orc.gbuffer(P, W, H)
        recs = outputs["%s_records_5_0" % vname]

        def img_check(name, mine, ref, rel, frac):
            scale = float(np.mean(np.abs(ref))) + 1e-12
            err = np.abs(mine.astype(np.float64) - ref) / (np.abs(ref) + 1e-3 * scale)
            good = (err <= rel).all(axis=2)
            rmse = float(np.sqrt(np.mean((mine.astype(np.float64) - ref) ** 2)) / scale)
            rep[name] = dict(pixels_within=float(good.mean()), max_rel=float(err.max()), rel_rmse=rmse)
            assert good.mean() >= frac, (name, rep[name])
            return rmse

        for mode in range(6):
            Pm = rig_params(scene, camera, mis_mode=mode)
            mine, _ = orc.vpl_gather(Pm, W, H, planes, prims, recs[: NUM_VPL_PATHS * b1], mode=0)
            # north_star: per-pixel radiance within 1e-4 relative OR image relative RMSE <= 1e-5.  A shadow ray that grazes a
            # silhouette edge can flip under the reference's FMA roundings (one VPL of one pixel): nearly all pixels, and the RMSE
            rmse = img_check("%s_splatColor_mode%d" % (vname, mode), mine, outputs["%s_splatColor_mode%d" % (vname, mode)], 1e-4, 0.999)
            assert rmse <= 1e-5
        Pv = rig_params(scene, camera, num_vpl_paths=VSL_VPL_PATHS)
        mine, _ = orc.vpl_gather(Pv, W, H, planes, prims, recs[: VSL_VPL_PATHS * b1], mode=1)
        # VSL: the per-pixel stream's consumption depends on float comparisons (numSamples, cone / cosine early-outs,
        # lighttracing.cu:485-491, 560-569, 649); a pixel whose comparison lands differently under the reference's
        # roundings draws a different sample set from there on.  Nearly all pixels must agree tightly, the image closely.
        rmse = img_check("%s_splatSplotch" % vname, mine, outputs["%s_splatSplotch" % vname], 1e-4, 0.99)
        assert rmse <= 1.5e-5
        Pa = rig_params(scene, camera, accumulate=True)
        mine, _ = orc.vpl_gather(Pa, W, H, planes, prims, recs[: NUM_VPL_PATHS * b1], mode=0)
        img_check("%s_splatColor_accumulate" % vname, mine + f32(0.25), outputs["%s_splatColor_accumulate" % vname], 1e-4, 0.999)
        # RtPt2: a path whose Russian roulette / lobe choice / hit lands differently under the reference's roundings is a
        # different path from there on (a different, equally valid sample): nearly all pixels must agree tightly
        acc = np.zeros((H, W, 3), dtype=f32)
        for k in range(PT_SPP):
            mine, _ = orc.path_trace(rig_params(scene, camera, rng_seed=k), W, H, planes, prims, PT_BOUNCES)
            acc = acc + mine
        img_check("%s_pathtrace" % vname, acc, outputs["%s_pathtrace" % vname], 1e-4, 0.99)
    return rep
