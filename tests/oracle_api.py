"""ctypes view of oracle/liboracle.so -- the CPU oracle (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from evplp_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")

_P = C.c_void_p
_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, capture_output=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_SO):
        build()
    lib = C.CDLL(ORACLE_SO)
    lib.orc_scene_create.restype = _P
    lib.orc_scene_create.argtypes = [C.POINTER(capi.MeshDesc), C.c_int32, C.POINTER(capi.MaterialDesc), C.c_int32, C.c_int32,
                                     C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32]
    lib.orc_scene_destroy.argtypes = [_P]
    lib.orc_num_prims.restype = C.c_int32
    lib.orc_num_prims.argtypes = [_P]
    lib.orc_light_area.restype = C.c_float
    lib.orc_light_area.argtypes = [_P]
    lib.orc_light_cdf.argtypes = [_P, _P]
    lib.orc_num_threads.restype = C.c_int32
    lib.orc_set_threads.argtypes = [C.c_int32]
    lib.orc_brdf.argtypes = [C.c_int32, _P, C.c_uint32, _P]
    lib.orc_total_area.restype = C.c_float
    lib.orc_total_area.argtypes = [_P, _P, C.c_int32]
    lib.orc_uniforms.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _P]
    lib.orc_raw_u32.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _P]
    lib.orc_math.argtypes = [C.c_int, _P, _P, C.c_uint32, _P]
    lib.orc_mt19937.argtypes = [C.c_uint32, C.c_uint32, _P]
    lib.orc_jitter_stream.argtypes = [C.c_uint32, C.c_uint32, _P]
    lib.orc_progressive_update.argtypes = [C.c_int32, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_int32, _P]
    lib.orc_light_trace.argtypes = [_P, C.POINTER(capi.Params), C.c_uint32, C.c_uint32, C.c_uint32, _P]
    lib.orc_gbuffer.argtypes = [_P, C.POINTER(capi.Params), C.c_int32, C.c_int32, _P, _P]
    lib.orc_vpl_gather.argtypes = [_P, C.POINTER(capi.Params), C.c_int32, C.c_int32, _P, _P, _P, C.c_int32,
                                   C.POINTER(capi.Tile), _P, _P]
    lib.orc_path_trace.argtypes = [_P, C.POINTER(capi.Params), C.c_int32, C.c_int32, _P, _P, C.c_uint32, C.POINTER(capi.Tile), _P, _P]
    lib.orc_accumulate_fixed.argtypes = [_P, C.c_int64, _P]
    lib.orc_photon_splat.argtypes = [C.POINTER(capi.Params), C.c_int32, C.c_int32, _P, _P, _P, C.c_uint64, C.c_uint64,
                                     C.POINTER(capi.Tile), _P, _P, C.c_int32]
    lib.orc_light_pass.argtypes = [_P, C.POINTER(capi.Params), C.c_int32, C.c_int32, _P]
    lib.orc_resolve.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_int32, _P]
    lib.orc_trace_rays.argtypes = [_P, _P, C.c_uint64, C.c_int32, _P, _P]
    lib.orc_lbvh.argtypes = [_P, _P, _P, _P, _P, _P, _P, _P]
    _lib = lib
    return lib


class OracleScene:
    def __init__(self, scene, brute_force=False):
        self.lib = load()
        self.scene = scene
        md, mt, pre, disp = scene.descriptors()
        self._keep = (md, mt)
        self.h = self.lib.orc_scene_create(md, len(scene.meshes), mt, len(scene.materials), scene.light_mesh, pre, disp,
                                           1 if brute_force else 0)

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def light_trace(self, params, rng_seed, first_path, num_paths):
        out = np.zeros(num_paths * params.numPhotonsPerLightPath, dtype=capi.RECORD_DTYPE)
        self.lib.orc_light_trace(self.h, C.byref(params), rng_seed, first_path, num_paths, capi.ptr(out))
        return out

    def gbuffer(self, params, W, H):
        planes = np.empty((4, H, W, 4), dtype=np.float32)
        prims = np.empty((H, W), dtype=np.int32)
        self.lib.orc_gbuffer(self.h, C.byref(params), W, H, capi.ptr(planes), capi.ptr(prims))
        return planes, prims

    def vpl_gather(self, params, W, H, planes, prims, records, mode=0, tile=None):
        out = np.empty((H, W, 3), dtype=np.float32)
        counters = np.zeros(2, dtype=np.uint64)
        t = C.byref(capi.Tile(*tile)) if tile is not None else None
        records = np.ascontiguousarray(records)
        self.lib.orc_vpl_gather(self.h, C.byref(params), W, H, capi.ptr(planes), capi.ptr(prims), capi.ptr(records), mode, t,
                                capi.ptr(out), capi.ptr(counters))
        return out, counters

    def path_trace(self, params, W, H, planes, prims, max_bounces, tile=None):
        out = np.empty((H, W, 3), dtype=np.float32)
        counters = np.zeros(1, dtype=np.uint64)
        t = C.byref(capi.Tile(*tile)) if tile is not None else None
        self.lib.orc_path_trace(self.h, C.byref(params), W, H, capi.ptr(planes), capi.ptr(prims), max_bounces, t, capi.ptr(out),
                                capi.ptr(counters))
        return out, counters

    def accumulate_fixed(self, rgb, accum):
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        self.lib.orc_accumulate_fixed(capi.ptr(rgb), rgb.size, capi.ptr(accum))

    def photon_splat(self, params, W, H, planes, prims, records, first, count, accum, tile=None, brute_force=False):
        counters = np.zeros(2, dtype=np.uint64)
        t = C.byref(capi.Tile(*tile)) if tile is not None else None
        records = np.ascontiguousarray(records)
        self.lib.orc_photon_splat(C.byref(params), W, H, capi.ptr(planes), capi.ptr(prims), capi.ptr(records), first, count, t,
                                  capi.ptr(accum), capi.ptr(counters), 1 if brute_force else 0)
        return counters

    def light_pass(self, params, W, H, light):
        self.lib.orc_light_pass(self.h, C.byref(params), W, H, capi.ptr(light))

    def resolve(self, W, H, vpl, photon, light, vs, ps, ls, gamma=False):
        out = np.empty((H, W, 3), dtype=np.float32)
        self.lib.orc_resolve(self.h, W, H, capi.ptr(vpl), capi.ptr(photon), capi.ptr(light), vs, ps, ls, 1 if gamma else 0,
                             capi.ptr(out))
        return out

    def trace_rays(self, rays, any_hit=0):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        prim = np.empty(len(rays), dtype=np.int32)
        t = np.empty(len(rays), dtype=np.float32)
        self.lib.orc_trace_rays(self.h, capi.ptr(rays), len(rays), any_hit, capi.ptr(prim), capi.ptr(t))
        return prim, t

    def lbvh(self):
        n = self.lib.orc_num_prims(self.h)
        ni = max(n - 1, 0)
        codes = np.empty(n, dtype=np.uint64); order = np.empty(n, dtype=np.uint32)
        left = np.empty(ni, dtype=np.int32); right = np.empty(ni, dtype=np.int32); parent = np.empty(ni, dtype=np.int32)
        bounds = np.empty((ni, 6), dtype=np.float32); smm = np.empty(6, dtype=np.float32)
        self.lib.orc_lbvh(self.h, capi.ptr(codes), capi.ptr(order), capi.ptr(left), capi.ptr(right), capi.ptr(parent),
                          capi.ptr(bounds), capi.ptr(smm))
        return codes, order, left, right, parent, bounds, smm


def uniforms(seed, subsequence, n):
    out = np.empty(n, dtype=np.float32)
    load().orc_uniforms(seed, subsequence, n, capi.ptr(out))
    return out


def math_op(op, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros_like(x) if y is None else np.ascontiguousarray(y, dtype=np.float32)
    out = np.empty_like(x)
    load().orc_math(op, capi.ptr(x), capi.ptr(y), len(x), capi.ptr(out))
    return out
