"""Thin Python handle over the C ABI (one Device per GPU).  Plumbing for tests and bench.py;
the C++ host classes in evplp_b200/host/ drive the same entry points.
"""
import ctypes as C

import numpy as np

from . import _capi as capi


class Device:
    def __init__(self, width, height, device=0):
        self.lib = capi.load_library()
        self.W, self.H = int(width), int(height)
        h = C.c_void_p()
        capi.check(self.lib, self.lib.evplp_create(device, self.W, self.H, C.byref(h)), "evplp_create")
        self.h = h
        self._keep = None

    def close(self):
        if self.h:
            self.lib.evplp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        capi.check(self.lib, rc, what)

    # ---- setup
    def upload_scene(self, scene):
        md, mt, pre, disp = scene.descriptors()
        self._keep = (scene, md, mt)
        self._ck(self.lib.evplp_upload_scene(self.h, md, len(scene.meshes), mt, len(scene.materials), scene.light_mesh,
                                             pre, disp), "evplp_upload_scene")

    def build_bvh(self):
        self._ck(self.lib.evplp_build_bvh(self.h), "evplp_build_bvh")

    def set_params(self, params):
        self._ck(self.lib.evplp_set_params(self.h, C.byref(params)), "evplp_set_params")

    def set_option(self, name, value):
        self._ck(self.lib.evplp_set_option(self.h, name.encode(), int(value)), "evplp_set_option")

    # ---- stages
    def clear_accum(self):
        self._ck(self.lib.evplp_clear_accum(self.h), "evplp_clear_accum")

    def gbuffer(self):
        self._ck(self.lib.evplp_gbuffer(self.h), "evplp_gbuffer")

    def light_trace(self, rng_seed, first_path, num_paths):
        self._ck(self.lib.evplp_light_trace(self.h, rng_seed, first_path, num_paths), "evplp_light_trace")

    @staticmethod
    def _tile(tile):
        if tile is None:
            return None
        return C.byref(capi.Tile(*tile))

    def vpl_gather(self, mode=capi.GATHER_VPL, tile=None):
        self._ck(self.lib.evplp_vpl_gather(self.h, self._tile(tile), mode), "evplp_vpl_gather")

    def path_trace(self, max_bounces, tile=None):
        self._ck(self.lib.evplp_path_trace(self.h, self._tile(tile), max_bounces), "evplp_path_trace")

    def photon_splat(self, first_record, num_records, tile=None):
        self._ck(self.lib.evplp_photon_splat(self.h, first_record, num_records, self._tile(tile)), "evplp_photon_splat")

    def light_pass(self):
        self._ck(self.lib.evplp_light_pass(self.h), "evplp_light_pass")

    def resolve(self, vpl_scale, photon_scale, light_scale, gamma=False, out=None):
        if out is None:
            out = np.empty((self.H, self.W, 3), dtype=np.float32)
        self._ck(self.lib.evplp_resolve(self.h, vpl_scale, photon_scale, light_scale, 1 if gamma else 0, capi.ptr(out)),
                 "evplp_resolve")
        return out

    def accum_layer(self, layer):
        p = C.c_void_p()
        n = C.c_uint64()
        self._ck(self.lib.evplp_accum_layer(self.h, layer, C.byref(p), C.byref(n)), "evplp_accum_layer")
        return p.value, n.value

    def synchronize(self):
        self._ck(self.lib.evplp_synchronize(self.h), "evplp_synchronize")

    # ---- taps
    def download_records(self, first, count):
        out = np.zeros(count, dtype=capi.RECORD_DTYPE)
        self._ck(self.lib.evplp_download_records(self.h, first, count, capi.ptr(out)), "evplp_download_records")
        return out

    def upload_records(self, records, first_path=0):
        records = np.ascontiguousarray(records)
        self._ck(self.lib.evplp_upload_records(self.h, first_path, capi.ptr(records), len(records)), "evplp_upload_records")

    def download_gbuffer(self):
        planes = np.empty((4, self.H, self.W, 4), dtype=np.float32)
        prims = np.empty((self.H, self.W), dtype=np.int32)
        self._ck(self.lib.evplp_download_gbuffer(self.h, capi.ptr(planes), capi.ptr(prims)), "evplp_download_gbuffer")
        return planes, prims

    def upload_gbuffer(self, planes, prims):
        planes = np.ascontiguousarray(planes, dtype=np.float32)
        prims = np.ascontiguousarray(prims, dtype=np.int32)
        self._ck(self.lib.evplp_upload_gbuffer(self.h, capi.ptr(planes), capi.ptr(prims)), "evplp_upload_gbuffer")

    def bvh_info(self):
        info = capi.BvhInfo()
        self._ck(self.lib.evplp_bvh_info(self.h, C.byref(info)), "evplp_bvh_info")
        return info

    def download_bvh(self, topology=True):
        info = self.bvh_info()
        n, ni = info.numPrims, info.numInternal
        codes = np.empty(n, dtype=np.uint64)
        order = np.empty(n, dtype=np.uint32)
        if topology and ni:
            left = np.empty(ni, dtype=np.int32); right = np.empty(ni, dtype=np.int32); parent = np.empty(ni, dtype=np.int32)
            bounds = np.empty((ni, 6), dtype=np.float32)
            self._ck(self.lib.evplp_download_bvh(self.h, capi.ptr(codes), capi.ptr(order), capi.ptr(left), capi.ptr(right),
                                                 capi.ptr(parent), capi.ptr(bounds)), "evplp_download_bvh")
            return codes, order, left, right, parent, bounds
        self._ck(self.lib.evplp_download_bvh(self.h, capi.ptr(codes), capi.ptr(order), None, None, None, None), "evplp_download_bvh")
        return codes, order, None, None, None, None

    def trace_rays(self, rays, any_hit=0):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        prim = np.empty(len(rays), dtype=np.int32)
        t = np.empty(len(rays), dtype=np.float32)
        self._ck(self.lib.evplp_trace_rays(self.h, capi.ptr(rays), len(rays), any_hit, capi.ptr(prim), capi.ptr(t)), "evplp_trace_rays")
        return prim, t

    def download_accum(self):
        vpl = np.empty((self.H, self.W, 3), dtype=np.int64)
        photon = np.empty((self.H, self.W, 3), dtype=np.int64)
        light = np.empty((self.H, self.W), dtype=np.uint32)
        self._ck(self.lib.evplp_download_accum(self.h, capi.ptr(vpl), capi.ptr(photon), capi.ptr(light)), "evplp_download_accum")
        return vpl, photon, light

    def debug_uniforms(self, seed, subsequence, n):
        out = np.empty(n, dtype=np.float32)
        self._ck(self.lib.evplp_debug_uniforms(self.h, seed, subsequence, n, capi.ptr(out)), "evplp_debug_uniforms")
        return out

    def debug_curand(self, seed, subsequence, n):
        out = np.empty(n, dtype=np.float32)
        self._ck(self.lib.evplp_debug_curand(self.h, seed, subsequence, n, capi.ptr(out)), "evplp_debug_curand")
        return out

    def debug_math(self, op, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float32)
        out = np.empty_like(x)
        self._ck(self.lib.evplp_debug_math(self.h, op, capi.ptr(x), capi.ptr(y), len(x), capi.ptr(out)), "evplp_debug_math")
        return out

    def stats(self):
        s = capi.Stats()
        self._ck(self.lib.evplp_stats(self.h, C.byref(s)), "evplp_stats")
        return s

    def reset_stats(self):
        self._ck(self.lib.evplp_reset_stats(self.h), "evplp_reset_stats")

    def stage_ms(self, stage):
        ms = C.c_float()
        self._ck(self.lib.evplp_last_stage_ms(self.h, stage, C.byref(ms)), "evplp_last_stage_ms")
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        self._ck(self.lib.evplp_launch_count(self.h, C.byref(n)), "evplp_launch_count")
        return n.value
