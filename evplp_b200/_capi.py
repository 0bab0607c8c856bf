"""ctypes view of include/evplp.h (the C ABI of libevplp_b200.so).

PyTorch / numpy are plumbing only: every compute call below runs the sm_100a kernels of
libevplp_b200.so.  There is no CPU fallback -- if the library is missing or no B200 is
present the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EVPLP_LIB") or os.path.join(_HERE, "lib", "libevplp_b200.so")   # EVPLP_LIB: a tuning build of the same library

EVPLP_OK = 0
FLAG_USABLE_VPL, FLAG_USABLE_PHOTON, FLAG_LAMBERT_ONLY, FLAG_PHONG_ONLY = 1, 2, 4, 8
MIS_ONE, MIS_BALANCE, MIS_MAX, MIS_POWER2, MIS_GEOMETRY_CLAMP, MIS_GEOMETRY_BRDF_CLAMP = range(6)
MIS_BY_NAME = {"one": 0, "balance": 1, "max": 2, "power2": 3, "geometryClamp": 4, "geometryBrdfClamp": 5}
GATHER_VPL, GATHER_VSL, GATHER_LVC = 0, 1, 2
STAGE_BVH, STAGE_GBUFFER, STAGE_LIGHT_TRACE, STAGE_GATHER, STAGE_SPLAT, STAGE_RESOLVE = range(6)

RECORD_DTYPE = np.dtype(
    [
        ("position", "<f4", 3), ("flags", "<u4"),
        ("normal", "<f4", 3), ("pSelectLambert", "<f4"),
        ("flux", "<f4", 3), ("padding1", "<f4"),
        ("fluxDir", "<f4", 3), ("padding2", "<f4"),
        ("lambertReflectance", "<f4", 3), ("padding3", "<f4"),
        ("phongReflectance", "<f4", 3), ("phongExponent", "<f4"),
    ]
)
assert RECORD_DTYPE.itemsize == 96


class MeshDesc(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p), ("texcoords", C.c_void_p), ("indices", C.c_void_p),
        ("numVertices", C.c_int32), ("numTriangles", C.c_int32), ("matIndex", C.c_int32),
    ]


class MaterialDesc(C.Structure):
    _fields_ = [
        ("lambertReflectance", C.c_void_p), ("lambertW", C.c_int32), ("lambertH", C.c_int32),
        ("phongReflectance", C.c_void_p), ("phongW", C.c_int32), ("phongH", C.c_int32),
        ("phongExponent", C.c_void_p), ("exponentW", C.c_int32), ("exponentH", C.c_int32),
        ("lightIntensity", C.c_float * 4),
    ]


class Params(C.Structure):
    _fields_ = [
        ("cameraPosition", C.c_float * 3), ("camForward", C.c_float * 3), ("camRight", C.c_float * 3),
        ("camUp", C.c_float * 3), ("tanHalfFovX", C.c_float), ("tanHalfFovY", C.c_float),
        ("jitter", C.c_float * 2), ("nearDist", C.c_float), ("farDist", C.c_float),
        ("numLightPaths", C.c_uint32), ("numVplLightPaths", C.c_uint32), ("numPhotonsPerLightPath", C.c_uint32),
        ("radius", C.c_float), ("pdfMc", C.c_float), ("misMode", C.c_uint32), ("clampingValue", C.c_float),
        ("doAccumulate", C.c_uint32), ("vslRadius", C.c_float), ("vslInvPiRadius2", C.c_float),
        ("rngSeed", C.c_uint32),
    ]

    def copy(self):
        p = Params()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Params))
        return p


class Tile(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [
        ("emittedVpls", C.c_uint64), ("emittedPhotons", C.c_uint64), ("gatherPairs", C.c_uint64),
        ("shadowRays", C.c_uint64), ("splatPhotons", C.c_uint64), ("splatFragments", C.c_uint64),
        ("closestRays", C.c_uint64),
    ]


class BvhInfo(C.Structure):
    _fields_ = [
        ("numPrims", C.c_uint32), ("numInternal", C.c_uint32), ("wideNodeBytes", C.c_uint32),
        ("sceneMin", C.c_float * 3), ("sceneMax", C.c_float * 3),
    ]


# every symbol include/evplp.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "evplp_last_error": (C.c_char_p, []),
    "evplp_version": (C.c_char_p, []),
    "evplp_device_count": (C.c_int, []),
    "evplp_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "evplp_destroy": (C.c_int, [_P]),
    "evplp_upload_scene": (C.c_int, [_P, C.POINTER(MeshDesc), C.c_int32, C.POINTER(MaterialDesc), C.c_int32, C.c_int32,
                                     C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "evplp_build_bvh": (C.c_int, [_P]),
    "evplp_set_params": (C.c_int, [_P, C.POINTER(Params)]),
    "evplp_clear_accum": (C.c_int, [_P]),
    "evplp_gbuffer": (C.c_int, [_P]),
    "evplp_light_trace": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32]),
    "evplp_vpl_gather": (C.c_int, [_P, C.POINTER(Tile), C.c_int]),
    "evplp_path_trace": (C.c_int, [_P, C.POINTER(Tile), C.c_uint32]),
    "evplp_photon_splat": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.POINTER(Tile)]),
    "evplp_light_pass": (C.c_int, [_P]),
    "evplp_add_iterations": (C.c_int, [_P, C.c_int64]),
    "evplp_iterations": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "evplp_reduce": (C.c_int, [_P, _P]),
    "evplp_accum_layer": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "evplp_resolve": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_int, _P]),
    "evplp_download_records": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P]),
    "evplp_upload_records": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64]),
    "evplp_download_gbuffer": (C.c_int, [_P, _P, _P]),
    "evplp_upload_gbuffer": (C.c_int, [_P, _P, _P]),
    "evplp_bvh_info": (C.c_int, [_P, C.POINTER(BvhInfo)]),
    "evplp_download_bvh": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "evplp_trace_rays": (C.c_int, [_P, _P, C.c_uint64, C.c_int, _P, _P]),
    "evplp_download_accum": (C.c_int, [_P, _P, _P, _P]),
    "evplp_debug_uniforms": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, _P]),
    "evplp_debug_curand": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, _P]),
    "evplp_debug_xorwow_tables": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32]),
    "evplp_debug_math": (C.c_int, [_P, C.c_int, _P, _P, C.c_uint32, _P]),
    "evplp_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "evplp_reset_stats": (C.c_int, [_P]),
    "evplp_last_stage_ms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "evplp_synchronize": (C.c_int, [_P]),
    "evplp_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "evplp_debug_counters": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "evplp_debug_cluster_hist": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "evplp_event_record": (C.c_int, [_P, C.c_int]),
    "evplp_event_elapsed_ms": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "evplp_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
}

_lib = None


def load_library(path=None):
    """Load libevplp_b200.so and bind every symbol; raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(evplp_b200 has no CPU fallback)"
        )
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class EvplpError(RuntimeError):
    pass


def check(lib, rc, what):
    if rc != EVPLP_OK:
        msg = lib.evplp_last_error()
        raise EvplpError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def ptr(a):
    """Pointer to a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
