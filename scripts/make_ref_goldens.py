"""Generator of tests/golden/ref_device/*.npz: runs the reference's own device programs (oracle/_ref/libref_device.so,
built by `make -C oracle refdevice` from the unmodified sources under /root/reference) on a B200 and stores what they
returned for the fixed cases of tests/ref_cases.py, then checks the CPU oracle against them and writes the agreement report.

    gpurun -- python scripts/make_ref_goldens.py        # writes gpurun_out/ref_device/{reference_outputs.npz,report.json}
    cp gpurun_out/ref_device/* tests/golden/ref_device/
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import ref_cases as R  # noqa: E402
from tests import ref_device_api as D  # noqa: E402


def main():
    outdir = os.path.join(ROOT, "gpurun_out", "ref_device")
    os.makedirs(outdir, exist_ok=True)
    outs = R.reference_outputs()
    np.savez_compressed(os.path.join(outdir, "reference_outputs.npz"), **outs)
    report = {"sources": D.load().ref_sources().decode()}
    try:
        R.check_oracle(outs, report)
        report["oracle_agrees"] = True
    except AssertionError as e:
        report["oracle_agrees"] = False
        report["first_failure"] = repr(e)[:2000]
    with open(os.path.join(outdir, "report.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
