"""ctypes view of oracle/_ref/libref_device.so -- the reference's OWN OptiX device programs
(lighttracing.cu, triangleintersect.cu, rtmaterial.cuh, rtmath.cuh, rtlightsource.cuh), compiled
unmodified from /root/reference with nvcc for sm_100a against the stand-in headers of
oracle/ref_shim/ (`make -C oracle refdevice`).  TEST INFRASTRUCTURE: it pins the oracle; it needs a GPU
to run, so CPU tests use the outputs it wrote on a B200 (tests/golden/ref_device/, generator
scripts/make_ref_goldens.py).
"""
import ctypes as C
import os

import numpy as np

from evplp_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_device.so")
_P = C.c_void_p
_lib = None
f32 = np.float32


class GatherArgs(C.Structure):
    _fields_ = [("cameraPosition", C.c_float * 3), ("misMode", C.c_uint32), ("pdfMc", C.c_float), ("clampingValue", C.c_float),
                ("numVplLightPaths", C.c_uint32), ("numPhotonsPerLightPath", C.c_uint32), ("doAccumulate", C.c_uint32),
                ("rngSeed", C.c_uint32), ("vslRadius", C.c_float), ("vslInvPiRadius2", C.c_float),
                ("W", C.c_int32), ("H", C.c_int32), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32)]


def available():
    return os.path.exists(REF_SO)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(REF_SO)
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_sources.restype = C.c_char_p
        lib.ref_scene_create.restype = _P
        lib.ref_scene_create.argtypes = [C.POINTER(capi.MeshDesc), C.c_int, C.POINTER(capi.MaterialDesc), C.c_int, C.c_int,
                                         C.POINTER(C.c_float), _P, C.c_float]
        lib.ref_scene_destroy.argtypes = [_P]
        lib.ref_trace_photons.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P]
        lib.ref_gather.argtypes = [_P, C.c_int, C.POINTER(GatherArgs), _P, _P, _P]
        lib.ref_trace_rays.argtypes = [_P, _P, C.c_uint32, C.c_int, _P, _P]
        lib.ref_brdf.argtypes = [C.c_int, _P, C.c_uint32, _P]
        _lib = lib
    return _lib


def light_cdf(scene):
    """RtAreaLight::createOptixCdf (rtcommon.h:501-531): sequential f32 sums of Triangle::ComputeArea
    (trianglemesh.cpp:13-19), then normalised -- restated in numpy, independently of the oracle."""
    m = scene.meshes[scene.light_mesh]
    t = m.vertices[m.indices]
    ab, ac = t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]
    cx = ab[:, 1] * ac[:, 2] - ab[:, 2] * ac[:, 1]
    cy = ab[:, 2] * ac[:, 0] - ab[:, 0] * ac[:, 2]
    cz = ab[:, 0] * ac[:, 1] - ab[:, 1] * ac[:, 0]
    area = np.sqrt(((cx * cx).astype(f32) + (cy * cy).astype(f32)).astype(f32) + (cz * cz).astype(f32)).astype(f32) / f32(2.0)
    cdf = np.empty(len(area), dtype=f32)
    s = f32(0.0)
    for i, a in enumerate(area):
        s = f32(s + a)
        cdf[i] = s
    return (cdf / s).astype(f32), float(s)


def _ck(rc, what):
    if rc != 0:
        raise RuntimeError("%s: %s" % (what, load().ref_last_error().decode()))


class RefScene:
    def __init__(self, scene):
        self.lib = load()
        md, mt, pre, _ = scene.descriptors()
        self._keep = (md, mt, scene)
        cdf, area = light_cdf(scene)
        self.cdf, self.area = cdf, area
        self.h = self.lib.ref_scene_create(md, len(scene.meshes), mt, len(scene.materials), scene.light_mesh, pre, capi.ptr(cdf), area)

    def close(self):
        if self.h:
            self.lib.ref_scene_destroy(self.h)
            self.h = None

    def trace_photons(self, rng_seed, first_path, num_paths, b1):
        out = np.zeros(num_paths * b1, dtype=capi.RECORD_DTYPE)
        _ck(self.lib.ref_trace_photons(self.h, rng_seed, first_path, num_paths, b1, capi.ptr(out)), "ref_trace_photons")
        return out

    def gather(self, params, W, H, planes, records, vsl=False, tile=None, prev=None):
        a = GatherArgs()
        for k in range(3):
            a.cameraPosition[k] = params.cameraPosition[k]
        a.misMode, a.pdfMc, a.clampingValue = params.misMode, params.pdfMc, params.clampingValue
        a.numVplLightPaths, a.numPhotonsPerLightPath = params.numVplLightPaths, params.numPhotonsPerLightPath
        a.doAccumulate, a.rngSeed = params.doAccumulate, params.rngSeed
        a.vslRadius, a.vslInvPiRadius2 = params.vslRadius, params.vslInvPiRadius2
        a.W, a.H = W, H
        a.x0, a.y0, a.x1, a.y1 = tile if tile is not None else (0, 0, W, H)
        out = np.zeros((H, W, 4), dtype=f32) if prev is None else np.ascontiguousarray(prev, dtype=f32)
        planes = np.ascontiguousarray(planes, dtype=f32)
        records = np.ascontiguousarray(records)
        _ck(self.lib.ref_gather(self.h, 1 if vsl else 0, C.byref(a), capi.ptr(planes), capi.ptr(records), capi.ptr(out)), "ref_gather")
        return out

    def trace_rays(self, rays, any_hit=0):
        rays = np.ascontiguousarray(rays, dtype=f32).reshape(-1, 8)
        prim = np.empty(len(rays), dtype=np.int32)
        t = np.empty(len(rays), dtype=f32)
        _ck(self.lib.ref_trace_rays(self.h, capi.ptr(rays), len(rays), any_hit, capi.ptr(prim), capi.ptr(t)), "ref_trace_rays")
        return prim, t


REF_PT_SO = os.path.join(ROOT, "oracle", "_ref", "libref_device_pt.so")
_ptlib = None


class PtArgs(C.Structure):
    _fields_ = [("cameraPosition", C.c_float * 3), ("doAccumulate", C.c_uint32), ("maxBounces", C.c_uint32), ("rngSeed", C.c_uint32),
                ("W", C.c_int32), ("H", C.c_int32), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32)]


def load_pt():
    global _ptlib
    if _ptlib is None:
        lib = C.CDLL(REF_PT_SO)
        lib.refpt_last_error.restype = C.c_char_p
        lib.refpt_sources.restype = C.c_char_p
        lib.refpt_scene_create.restype = _P
        lib.refpt_scene_create.argtypes = [C.POINTER(capi.MeshDesc), C.c_int, C.POINTER(capi.MaterialDesc), C.c_int, C.c_int,
                                           C.POINTER(C.c_float), _P, C.c_float]
        lib.refpt_scene_destroy.argtypes = [_P]
        lib.refpt_path_trace.argtypes = [_P, C.POINTER(PtArgs), _P, _P]
        _ptlib = lib
    return _ptlib


def path_trace(scene, params, W, H, planes, max_bounces, prev=None):
    """pathtracing.cu splatColor over the whole image (one sample per pixel of stream (pixel, params.rngSeed))."""
    lib = load_pt()
    md, mt, pre, _ = scene.descriptors()
    cdf, area = light_cdf(scene)
    h = lib.refpt_scene_create(md, len(scene.meshes), mt, len(scene.materials), scene.light_mesh, pre, capi.ptr(cdf), area)
    a = PtArgs()
    for k in range(3):
        a.cameraPosition[k] = params.cameraPosition[k]
    a.doAccumulate, a.maxBounces, a.rngSeed = params.doAccumulate, max_bounces, params.rngSeed
    a.W, a.H, a.x0, a.y0, a.x1, a.y1 = W, H, 0, 0, W, H
    out = np.zeros((H, W, 4), dtype=f32) if prev is None else np.ascontiguousarray(prev, dtype=f32)
    planes = np.ascontiguousarray(planes, dtype=f32)
    rc = lib.refpt_path_trace(h, C.byref(a), capi.ptr(planes), capi.ptr(out))
    lib.refpt_scene_destroy(h)
    if rc != 0:
        raise RuntimeError("refpt_path_trace: %s" % lib.refpt_last_error().decode())
    return out


def brdf(op, in16):
    in16 = np.ascontiguousarray(in16, dtype=f32).reshape(-1, 16)
    out = np.empty((len(in16), 8), dtype=f32)
    _ck(load().ref_brdf(op, capi.ptr(in16), len(in16), capi.ptr(out)), "ref_brdf")
    return out
