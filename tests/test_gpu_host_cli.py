"""The C++ drop-in path end to end on the GPU: evplp_render (the headless reflectcuts.exe) renders an exported
scene JSON through RtComPhoton::render() and writes the three PFM files + stat JSON of rtcomphoton.h:1107-1132;
the result equals stepping the same technique through the C API."""
import json
import os
import subprocess

import numpy as np
import pytest

from evplp_b200 import host_api as HA

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "evplp_b200", "bin", "evplp_render")


def _read_pfm_rows(path):
    raw = open(path, "rb").read()
    parts = raw.split(b"\n", 3)
    assert parts[0] == b"PF" and parts[2] == b"-1"
    w, h = (int(v) for v in parts[1].split())
    return np.frombuffer(parts[3], dtype="<f4").reshape(h, w, 3)  # file rows = bottom-up = glReadPixels order


@pytest.mark.parametrize("variant", ["ours_clamp", "vsl", "pm_progressive"])
def test_cli_render_writes_the_reference_outputs(tmp_path, variant):
    d = str(tmp_path)
    HA.export_scene("livingroom", d, seed=3, detail=2, res_x=320, res_y=180)
    jpath = os.path.join(d, f"livingroom_{variant}.json")
    j = json.load(open(jpath))
    fam = j["photonfam"]
    fam["numMaxIteration"] = 4
    fam["numLightPaths"] = min(fam["numLightPaths"], 40000)
    fam["numVplLightPaths"] = min(fam["numVplLightPaths"], fam["numLightPaths"])
    fam["timeLimitMs"] = 600000.0
    json.dump(j, open(jpath, "w"))
    r = subprocess.run([EXE, jpath], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "4 iterations" in r.stdout
    comb = _read_pfm_rows(os.path.join(d, fam["combinedFilename"]))
    wvpl = _read_pfm_rows(os.path.join(d, fam["weightedVplFilename"]))
    wpm = _read_pfm_rows(os.path.join(d, fam["weightedPhotonFilename"]))
    stat = json.load(open(os.path.join(d, fam["statFilename"])))
    assert stat["numIterations"] == 4 and stat["time"] > 0
    assert comb.shape == (180, 320, 3) and np.isfinite(comb).all() and comb.mean() > 1e-3
    # the same run stepped through the C API of the host library
    hs = HA.HostScene.load(jpath)
    t = HA.Technique(hs, fam, 320, 180)
    for _ in range(4):
        assert t.iterate() or True
    param = np.float32(1.0 / 4)
    light = t.final(0.0, 0.0, 1.0)
    photon = t.final(0.0, 1.0, 0.0) * param
    vpl = t.final(1.0, 0.0, 0.0) * param
    t.close()
    assert np.array_equal(wpm, photon)
    assert np.array_equal(wvpl, light + vpl)
    assert np.array_equal(comb, (light + vpl) + photon)
    if variant == "pm_progressive":
        assert not vpl.any() and photon.any()      # numVplLightPaths 0 disables the gather (rtcomphoton.h:200-203)
    if variant == "vsl":
        assert vpl.any() and not photon.any()      # radiusPercentage 0: no splat fragments


def _fam(**kw):
    fam = {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "combinedFilename": "a.pfm",
           "weightedPhotonFilename": "b.pfm", "weightedVplFilename": "c.pfm", "statFilename": "s.json", "useJitter": False, "useStat": False,
           "numLightPaths": 20000, "numVplLightPaths": 32, "numMaxBounces": 3, "radiusPercentage": 0.01, "misMode": "geometryClamp",
           "clampingCoeff": 0.05, "DoProgressive": False}
    fam.update(kw)
    return fam


def test_cleareveryframe_keeps_only_the_last_iteration():
    """frameMode cleareveryframe (doAccumulate = 0, lighttracing.cu:378; glClear per frame, rtcomphoton.h:978-981)."""
    hs = HA.HostScene.generate("livingroom", 3, 2, 320 / 180)
    a = HA.Technique(hs, _fam(frameMode="cleareveryframe"), 320, 180)
    for _ in range(3):
        a.iterate()                                   # iterations use rngSeed 0, 1, 2; the buffers are cleared each time
    img_a = a.final(1.0, 1.0, 1.0)
    a.close()
    b = HA.Technique(hs, _fam(rngOffset=2), 320, 180)  # one accumulated iteration with rngSeed 2
    b.iterate()
    img_b = b.final(1.0, 1.0, 1.0)
    b.close()
    assert img_a.any() and np.array_equal(img_a, img_b)


def test_lvc_technique_runs_through_the_host_class(tmp_path):
    """RtLvcComPhoton ("lvcphotonfam", main.cpp:116-120): per-pixel window of light paths (lvclighttracing.cu:348-387)."""
    d = str(tmp_path)
    HA.export_scene("livingroom", d, seed=3, detail=2, res_x=160, res_y=90)
    jpath = os.path.join(d, "livingroom_ours.json")
    j = json.load(open(jpath))
    fam = j.pop("photonfam")
    fam.update(numMaxIteration=2, numLightPaths=4096, numVplLightPaths=16, timeLimitMs=600000.0)
    j["lvcphotonfam"] = fam
    json.dump(j, open(jpath, "w"))
    r = subprocess.run([EXE, jpath], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "lvcphotonfam: 2 iterations" in r.stdout
    comb = _read_pfm_rows(os.path.join(d, fam["combinedFilename"]))
    hs = HA.HostScene.load(jpath)
    t = HA.Technique(hs, fam, 160, 90, lvc=True)
    for _ in range(2):
        t.iterate()
    half = np.float32(0.5)
    exp = (t.final(0.0, 0.0, 1.0) + t.final(1.0, 0.0, 0.0) * half) + t.final(0.0, 1.0, 0.0) * half
    t.close()
    assert comb.mean() > 1e-4 and np.array_equal(comb, exp)


def test_cli_path_tracer_json(tmp_path):
    """scene/*/*_pt.json -> RtPt2::render (main.cpp:105-109): output = light + pt / N, flipped, PFM."""
    d = str(tmp_path)
    HA.export_scene("livingroom", d, seed=3, detail=2, res_x=160, res_y=90)
    jpath = os.path.join(d, "livingroom_pt.json")
    j = json.load(open(jpath))
    j["pt"].update(numMaxIteration=5, timeLimitMs=600000.0)
    json.dump(j, open(jpath, "w"))
    r = subprocess.run([EXE, jpath], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pt: 5 iterations" in r.stdout
    img = _read_pfm_rows(os.path.join(d, j["pt"]["outputFilename"]))
    hs = HA.HostScene.load(jpath)
    t = HA.PathTracer(hs, j["pt"], 160, 90)
    for _ in range(5):
        t.iterate()
    exp = t.final(0.0, 1.0) + t.final(1.0, 0.0) * np.float32(1.0 / 5)
    t.close()
    assert img.mean() > 1e-4 and np.array_equal(img, exp)
    assert json.load(open(os.path.join(d, j["pt"]["statFilename"])))["numIterations"] == 5


def test_cli_path_tracer_write_every_frame(tmp_path):
    """writeEveryFrame in RtPt2 (rtpt2.h:669-689, the after-swap callback of the loop): <output>_<k>.pfm after iteration k holds
    light + pt / k; the last one equals the final output."""
    d = str(tmp_path)
    HA.export_scene("livingroom", d, seed=3, detail=2, res_x=96, res_y=54)
    jpath = os.path.join(d, "livingroom_pt.json")
    j = json.load(open(jpath))
    j["pt"].update(numMaxIteration=3, timeLimitMs=600000.0, writeEveryFrame=True)
    json.dump(j, open(jpath, "w"))
    r = subprocess.run([EXE, jpath], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = j["pt"]["outputFilename"]
    stem, ext = os.path.splitext(out)
    frames = [_read_pfm_rows(os.path.join(d, f"{stem}_{k}{ext}")) for k in (1, 2, 3)]
    final = _read_pfm_rows(os.path.join(d, out))
    assert np.array_equal(frames[2], final)
    hs = HA.HostScene.load(jpath)
    t = HA.PathTracer(hs, j["pt"], 96, 54)
    t.iterate()
    first = t.final(0.0, 1.0) + t.final(1.0, 0.0) * np.float32(1.0)
    t.close()
    assert np.array_equal(frames[0], first) and not np.array_equal(frames[0], frames[1])
    assert not os.path.exists(os.path.join(d, f"{stem}_4{ext}"))


def test_png_textured_scene_renders_like_its_decoded_twin(tmp_path):
    """MTL map_Kd -> PNG: the host decodes it (host/pngdecode.h, byte-exact with stb_image) and the render equals, bit for bit, the
    same scene with the texture replaced by a PPM of the pixels the reference's decoder returns (tests/golden/png/expected.npz)."""
    name = "rgb8_large_dynamic"
    exp = np.load(os.path.join(ROOT, "tests", "golden", "png", "expected.npz"))[name + "__want3"]
    imgs = []
    for variant in ("png", "ppm"):
        d = str(tmp_path / variant)
        os.makedirs(d)
        HA.export_scene("conference", d, seed=2, detail=1, res_x=96, res_y=54)
        mtl = [f for f in os.listdir(d) if f.endswith(".mtl") and "light" not in f][0]
        txt = open(os.path.join(d, mtl)).read()
        assert "map_Kd" in txt
        if variant == "png":
            with open(os.path.join(ROOT, "tests", "golden", "png", name + ".png"), "rb") as f:
                open(os.path.join(d, "tex.png"), "wb").write(f.read())
            tex = "tex.png"
        else:
            with open(os.path.join(d, "tex.ppm"), "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (exp.shape[1], exp.shape[0]) + exp.tobytes())
            tex = "tex.ppm"
        import re as _re
        txt = _re.sub(r"map_Kd .*", "map_Kd " + tex, txt)
        open(os.path.join(d, mtl), "w").write(txt)
        jpath = os.path.join(d, "conference_ours.json")
        j = json.load(open(jpath))
        j["photonfam"].update(numMaxIteration=1, timeLimitMs=600000.0, numLightPaths=4096, numVplLightPaths=64)
        json.dump(j, open(jpath, "w"))
        r = subprocess.run([EXE, jpath], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        imgs.append(_read_pfm_rows(os.path.join(d, j["photonfam"]["combinedFilename"])))
    assert imgs[0].mean() > 1e-4 and np.array_equal(imgs[0], imgs[1])


def test_jpeg_textured_scene_renders_like_its_decoded_twin(tmp_path):
    """A real-asset style scene (MTL map_Kd -> JPEG, as the reference's livingroom ships) loaded by the C++ host and
    rendered on the GPU equals, bit for bit, the same scene with the texture replaced by a PPM holding the pixels the
    reference's decoder (stb_image, tests/golden/jpeg/expected.npz) produces for that file."""
    import shutil
    gold = os.path.join(ROOT, "tests", "golden", "jpeg")
    exp = np.load(os.path.join(gold, "expected.npz"))
    images = []
    for kind in ("jpeg", "ppm"):
        d = os.path.join(str(tmp_path), kind)
        os.makedirs(d)
        HA.export_scene("livingroom", d, seed=3, detail=2, res_x=320, res_y=180)
        mtl = [f for f in os.listdir(d) if f.endswith(".mtl") and "light" not in f][0]
        text = open(os.path.join(d, mtl)).read()
        assert "map_Kd livingroom_wood.ppm" in text
        if kind == "jpeg":
            shutil.copy(os.path.join(gold, "prog_420_q95.jpg"), os.path.join(d, "wood.jpg"))
            text = text.replace("map_Kd livingroom_wood.ppm", "map_Kd wood.jpg")
        else:
            px = exp["prog_420_q95"]
            with open(os.path.join(d, "wood.ppm"), "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (px.shape[1], px.shape[0]) + px.tobytes())
            text = text.replace("map_Kd livingroom_wood.ppm", "map_Kd wood.ppm")
        open(os.path.join(d, mtl), "w").write(text)
        jpath = os.path.join(d, "livingroom_ours.json")
        fam = json.load(open(jpath))["photonfam"]
        fam.update(numMaxIteration=2, numLightPaths=20000, numVplLightPaths=64, timeLimitMs=600000.0)
        hs = HA.HostScene.load(jpath)
        t = HA.Technique(hs, fam, 320, 180)
        for _ in range(2):
            t.iterate()
        images.append(t.final(1.0, 1.0, 1.0))
        t.close()
    assert images[0].mean() > 1e-3
    assert np.array_equal(images[0], images[1])
