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
