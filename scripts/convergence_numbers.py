"""Prints the image energies of the independent renders compared in tests/test_gpu_energy.py (documentation aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from evplp_b200 import host_api as HA
from tests.test_gpu_energy import _render, W, H
hs = HA.HostScene.generate("livingroom", 4, 2, W / H)
pt = HA.PathTracer(hs, {"rngOffset": 0, "numMaxIteration": -1, "timeLimitMs": -1.0, "frameMode": "accumulate", "outputFilename": "pt.pfm",
                        "statFilename": "s.json", "useJitter": True, "useStat": False, "numSamplePerPixel": 1, "numMaxBounces": 3}, W, H)
n = 256
for _ in range(n):
    pt.iterate()
e_pt = pt.final(1.0 / n, 0.0).astype(np.float64).sum(); pt.close()
ref, _ = _render(hs, misMode="one", radiusPercentage=0.0)
cv, cp = _render(hs, misMode="geometryClamp", clampingCoeff=0.02)
bv, bp = _render(hs, misMode="balance")
print(f"path tracer {e_pt:.1f} | VPL unclamped {ref.sum():.1f} ({ref.sum()/e_pt:.4f}) | clamped gather {cv.sum():.1f} ({cv.sum()/e_pt:.4f}) "
      f"+ compensation {cp.sum():.1f} = {(cv.sum()+cp.sum())/e_pt:.4f} | balance MIS {bv.sum():.1f} + {bp.sum():.1f} = {(bv.sum()+bp.sum())/e_pt:.4f}")
