"""The oracle against the REFERENCE'S OWN device code, live on the B200.

oracle/_ref/libref_device.so holds the unmodified reference sources lighttracing.cu, triangleintersect.cu,
rtmaterial.cuh, rtmath.cuh and rtlightsource.cuh compiled with nvcc for sm_100a against stand-in OptiX headers
(oracle/ref_shim/; built by `make -C oracle refdevice` where /root/reference exists, shipped to the GPU box as a
git-ignored binary).  These tests run the reference's tracePhotons / rtMaterialClosestHit / meshFineIntersect /
splatColor + vplSplat (6 MIS modes) / splatSplotch + vslSplat and its BRDF library with the real curand_kernel.h and
compare the CPU oracle with them: flags, emitted counts, RNG consumption and primitive ids exactly, floats to 1e-5
(the reference build contracts FMAs and uses libdevice).  The product is then compared with the oracle bit for bit
(tests/test_gpu_parity.py), which closes the chain reference -> oracle -> product.
"""
import numpy as np
import pytest

from tests import ref_cases as R
from tests import ref_device_api as D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def outputs():
    if not D.available():
        pytest.skip("oracle/_ref/libref_device.so not built (needs /root/reference at build time)")
    return R.reference_outputs()


def test_oracle_equals_reference_device_code(outputs):
    rep = R.check_oracle(outputs)
    for k in ("glossy_records_5_0", "spot_records_5_0"):
        assert rep[k]["flags_equal"] >= rep[k]["paths"] - rep[k]["paths"] // 256   # (see tests/ref_cases.py: rounding-decided paths)
        assert rep[k]["vpls"] > R.NUM_PATHS and rep[k]["photons"] > R.NUM_PATHS // 2


def test_reference_outputs_equal_the_committed_goldens(outputs):
    """the fixtures the CPU suite uses are what the reference code returns (same GPU family, same build)"""
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "ref_device", "reference_outputs.npz")
    if not os.path.exists(path):
        pytest.skip("goldens not generated yet")
    gold = np.load(path)
    assert set(gold.files) == set(outputs.keys())
    for k in gold.files:
        a, b = gold[k], outputs[k]
        if a.dtype.names:
            assert a.tobytes() == b.tobytes(), k
        else:
            assert np.array_equal(a, b, equal_nan=True), k
