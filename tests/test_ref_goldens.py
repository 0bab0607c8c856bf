"""CPU: the oracle against what the reference's own device code returned on a B200 (tests/golden/ref_device/,
written by scripts/make_ref_goldens.py from oracle/_ref/libref_device.so).  This is the pin of the oracle."""
import json
import os

import numpy as np

from tests import ref_cases as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_device")


def test_oracle_matches_reference_device_goldens():
    outs = dict(np.load(os.path.join(GOLD, "reference_outputs.npz")))
    rep = R.check_oracle(outs)
    assert rep["rays_closest_agree"] >= 0.9999


def test_golden_report_says_the_oracle_agreed_on_the_gpu_box():
    with open(os.path.join(GOLD, "report.json")) as f:
        rep = json.load(f)
    assert rep["oracle_agrees"] is True
    assert "lighttracing.cu" in rep["sources"]
