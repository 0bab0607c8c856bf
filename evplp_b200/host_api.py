"""ctypes view of libevplp_host.so: the C++ host classes (RtScene, RtComPhoton, scene
generator) of evplp_b200/host/.  Scene data preparation only; rendering goes through
libevplp_b200.so."""
import ctypes as C
import json
import os

import numpy as np

from . import _capi as capi
from .scene import Material, Mesh, Scene

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libevplp_host.so")
SCENE_LIB_PATH = os.path.join(_HERE, "lib", "libevplp_scene.so")
_P = C.c_void_p
_lib = None
_scene_lib = None


def _bind_scene_symbols(lib):
    lib.evplp_host_last_error.restype = C.c_char_p
    lib.evplp_host_export_scene.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    lib.evplp_host_generate_scene.restype = _P
    lib.evplp_host_generate_scene.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_float]
    lib.evplp_host_load_scene.restype = _P
    lib.evplp_host_load_scene.argtypes = [C.c_char_p]
    lib.evplp_host_scene_destroy.argtypes = [_P]
    lib.evplp_host_scene_descriptors.argtypes = [_P, C.POINTER(C.POINTER(capi.MeshDesc)), C.POINTER(C.c_int32),
                                                 C.POINTER(C.POINTER(capi.MaterialDesc)), C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.evplp_host_scene_info.argtypes = [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]


def load_scene_library():
    """libevplp_scene.so: scene generation / loading only, linked against nothing of the product (the CPU reference arm
    of bench.py builds its inputs with it, so that it never maps libevplp_b200.so)."""
    global _scene_lib
    if _scene_lib is None:
        if not os.path.exists(SCENE_LIB_PATH):
            raise RuntimeError(f"{SCENE_LIB_PATH} not found: run __graft_entry__.build()")
        _scene_lib = C.CDLL(SCENE_LIB_PATH)
        _bind_scene_symbols(_scene_lib)
    return _scene_lib


def load_host_library():
    global _lib
    if _lib is not None:
        return _lib
    capi.load_library()  # libevplp_b200.so first (RTLD_GLOBAL)
    if not os.path.exists(HOST_LIB_PATH):
        raise RuntimeError(f"{HOST_LIB_PATH} not found: run __graft_entry__.build()")
    lib = C.CDLL(HOST_LIB_PATH)
    lib.evplp_host_last_error.restype = C.c_char_p
    lib.evplp_host_export_scene.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    lib.evplp_host_generate_scene.restype = _P
    lib.evplp_host_generate_scene.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_float]
    lib.evplp_host_load_scene.restype = _P
    lib.evplp_host_load_scene.argtypes = [C.c_char_p]
    lib.evplp_host_scene_destroy.argtypes = [_P]
    lib.evplp_host_scene_descriptors.argtypes = [_P, C.POINTER(C.POINTER(capi.MeshDesc)), C.POINTER(C.c_int32),
                                                 C.POINTER(C.POINTER(capi.MaterialDesc)), C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.evplp_host_scene_info.argtypes = [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.evplp_host_technique_create.restype = _P
    lib.evplp_host_technique_create.argtypes = [_P, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.evplp_host_technique_set_max_paths_per_trace.argtypes = [_P, C.c_uint64]
    lib.evplp_host_technique_plan.argtypes = [_P, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]
    lib.evplp_host_technique_handle.restype = _P
    lib.evplp_host_technique_handle.argtypes = [_P]
    lib.evplp_host_technique_iterate.argtypes = [_P]
    lib.evplp_host_technique_state.argtypes = [_P, C.POINTER(C.c_float)]
    lib.evplp_host_technique_final.argtypes = [_P, C.c_float, C.c_float, C.c_float, C.c_int, _P]
    lib.evplp_host_technique_destroy.argtypes = [_P]
    lib.evplp_host_pt_create.restype = _P
    lib.evplp_host_pt_create.argtypes = [_P, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.evplp_host_pt_iterate.argtypes = [_P]
    lib.evplp_host_pt_final.argtypes = [_P, C.c_float, C.c_float, C.c_int, _P]
    lib.evplp_host_pt_handle.restype = _P
    lib.evplp_host_pt_handle.argtypes = [_P]
    lib.evplp_host_pt_destroy.argtypes = [_P]
    lib.evplp_host_render_json.argtypes = [C.c_char_p, C.c_int]
    lib.evplp_host_progressive_update.argtypes = [C.c_int32, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_int32, _P]
    lib.evplp_host_jitter_stream.argtypes = [C.c_uint32, C.c_uint32, _P]
    lib.evplp_host_save_pfm.argtypes = [C.c_char_p, _P, C.c_int, C.c_int]
    lib.evplp_host_pfm_relmse.restype = C.c_float
    lib.evplp_host_pfm_relmse.argtypes = [C.c_char_p, C.c_char_p]
    lib.evplp_host_config_check.argtypes = [_P, C.c_char_p, C.POINTER(C.c_double)]
    lib.evplp_host_jpeg_info.argtypes = [_P, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.evplp_host_jpeg_decode.argtypes = [_P, C.c_uint64, _P, C.c_uint64]
    lib.evplp_host_png_info.argtypes = [_P, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.evplp_host_png_decode.argtypes = [_P, C.c_uint64, C.c_int, _P, C.c_uint64]
    lib.evplp_host_realtime_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.evplp_host_realtime_probe.restype = None
    lib.evplp_host_texture_load.argtypes = [C.c_char_p, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, C.c_uint64]
    _lib = lib
    return lib


class HostError(RuntimeError):
    pass


def _err(lib, what):
    raise HostError(f"{what}: {lib.evplp_host_last_error().decode()}")


class HostScene:
    """An RtScene living in the C++ host library."""

    def __init__(self, handle, lib=None):
        self.lib = lib or load_host_library()
        self.h = handle
        s = (C.c_float * 3)()
        cam = (C.c_float * 14)()
        if self.lib.evplp_host_scene_info(self.h, s, cam) != 0:
            _err(self.lib, "evplp_host_scene_info")
        self.bounding_sphere_radius = np.float32(s[0])
        self.total_area = np.float32(s[1])
        self.num_triangles = int(s[2])
        c = np.array(list(cam), dtype=np.float32)
        self.cam_origin, self.cam_forward, self.cam_right, self.cam_up = c[0:3], c[3:6], c[6:9], c[9:12]
        self.tan_x, self.tan_y = c[12], c[13]

    @classmethod
    def generate(cls, name, seed=1, detail=8, aspect=16 / 9, scene_only=False):
        lib = load_scene_library() if scene_only else load_host_library()
        h = lib.evplp_host_generate_scene(name.encode(), seed, detail, aspect)
        if not h:
            _err(lib, "evplp_host_generate_scene")
        return cls(h, lib)

    @classmethod
    def load(cls, json_path):
        lib = load_host_library()
        h = lib.evplp_host_load_scene(json_path.encode())
        if not h:
            _err(lib, "evplp_host_load_scene")
        return cls(h)

    def close(self):
        if self.h:
            self.lib.evplp_host_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def to_scene(self):
        """Copy into the numpy Scene container (what the oracle and Device.upload_scene take)."""
        md = C.POINTER(capi.MeshDesc)(); mt = C.POINTER(capi.MaterialDesc)()
        nm = C.c_int32(); nt = C.c_int32(); li = C.c_int32()
        pre = (C.c_float * 4)(); disp = (C.c_float * 4)()
        self.lib.evplp_host_scene_descriptors(self.h, C.byref(md), C.byref(nm), C.byref(mt), C.byref(nt), C.byref(li), pre, disp)
        sc = Scene()

        def arr(p, n, dt):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(dt)), shape=(n,)).copy()

        for k in range(nt.value):
            m = mt[k]
            mat = Material.__new__(Material)
            mat.lambert = arr(m.lambertReflectance, m.lambertW * m.lambertH * 4, C.c_float).reshape(m.lambertH, m.lambertW, 4)
            mat.phong = arr(m.phongReflectance, m.phongW * m.phongH * 4, C.c_float).reshape(m.phongH, m.phongW, 4)
            mat.exponent = arr(m.phongExponent, m.exponentW * m.exponentH * 4, C.c_float).reshape(m.exponentH, m.exponentW, 4)
            mat.lightIntensity = np.array(list(m.lightIntensity), dtype=np.float32)
            sc.materials.append(mat)
        for k in range(nm.value):
            m = md[k]
            v = arr(m.vertices, m.numVertices * 3, C.c_float)
            t = arr(m.texcoords, m.numVertices * 2, C.c_float) if m.texcoords else None
            i = arr(m.indices, m.numTriangles * 3, C.c_int32)
            sc.meshes.append(Mesh(v, i, m.matIndex, t))
        sc.light_mesh = li.value
        sc.light_intensity = np.array(list(disp), dtype=np.float32)
        return sc

    def camera(self):
        class _Cam:
            pass

        c = _Cam()
        c.origin, c.forward, c.right, c.up, c.tan_x, c.tan_y = self.cam_origin, self.cam_forward, self.cam_right, self.cam_up, self.tan_x, self.tan_y
        return c


class Technique:
    """RtComPhoton (or RtLvcComPhoton) of the C++ host library, stepped one iteration at a time."""

    def __init__(self, host_scene, photonfam, res_x, res_y, device=0, lvc=False, rank=0, world_size=1, image_partition=False):
        self.lib = load_host_library()
        self.W, self.H = res_x, res_y
        text = json.dumps(photonfam).encode()
        self.h = self.lib.evplp_host_technique_create(host_scene.h, text, res_x, res_y, device, 1 if lvc else 0, rank, world_size,
                                                      1 if image_partition else 0)
        if not self.h:
            _err(self.lib, "evplp_host_technique_create")
        self._scene = host_scene

    def set_max_paths_per_trace(self, n):
        self.lib.evplp_host_technique_set_max_paths_per_trace(self.h, n)

    def device_handle(self):
        return C.c_void_p(self.lib.evplp_host_technique_handle(self.h))

    def iterate(self):
        rc = self.lib.evplp_host_technique_iterate(self.h)
        if rc < 0:
            _err(self.lib, "evplp_host_technique_iterate")
        return rc == 1

    def state(self):
        s = (C.c_float * 6)()
        self.lib.evplp_host_technique_state(self.h, s)
        return dict(radius=s[0], clamping=s[1], pdfMc=s[2], vslRadius=s[3], vslInvPiRadius2=s[4], numIterations=int(s[5]))

    def final(self, vpl_scale, photon_scale, light_scale, gamma=False, out=None):
        if out is None:
            out = np.empty((self.H, self.W, 3), dtype=np.float32)
        if self.lib.evplp_host_technique_final(self.h, vpl_scale, photon_scale, light_scale, 1 if gamma else 0, capi.ptr(out)) != 0:
            _err(self.lib, "evplp_host_technique_final")
        return out

    def close(self):
        if self.h:
            self.lib.evplp_host_technique_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PathTracer:
    """RtPt2 of the C++ host library (the reference's ground-truth technique), stepped one iteration at a time."""

    def __init__(self, host_scene, pt, res_x, res_y, device=0, rank=0, world_size=1):
        self.lib = load_host_library()
        self.W, self.H = res_x, res_y
        self.h = self.lib.evplp_host_pt_create(host_scene.h, json.dumps(pt).encode(), res_x, res_y, device, rank, world_size)
        if not self.h:
            _err(self.lib, "evplp_host_pt_create")
        self._scene = host_scene

    def iterate(self):
        rc = self.lib.evplp_host_pt_iterate(self.h)
        if rc < 0:
            _err(self.lib, "evplp_host_pt_iterate")
        return rc == 1

    def final(self, pt_scale, light_scale, gamma=False):
        out = np.empty((self.H, self.W, 3), dtype=np.float32)
        if self.lib.evplp_host_pt_final(self.h, pt_scale, light_scale, 1 if gamma else 0, capi.ptr(out)) != 0:
            _err(self.lib, "evplp_host_pt_final")
        return out

    def close(self):
        if self.h:
            self.lib.evplp_host_pt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PLAN_FIELDS = ("iteration", "render", "jitter_x", "jitter_y", "rngSeed", "photonRadius", "clampingValue", "pdfMc", "vslRadius",
               "vslInvPiRadius2", "splatFirstPath", "splatNumPaths", "tileStride", "tileOffset", "drawLight", "countIteration")


def plan(host_scene, photonfam, res_x, res_y, rank, world_size, num_iterations, image_partition=False, lvc=False):
    """What rank `rank` of `world_size` does in each of the first `num_iterations` passes of RtComPhoton's loop (the host logic
    of iterate(), no device): a list of dicts with PLAN_FIELDS."""
    lib = load_host_library()
    out = np.zeros((num_iterations, 16), dtype=np.float64)
    if lib.evplp_host_technique_plan(host_scene.h, json.dumps(photonfam).encode(), res_x, res_y, 1 if lvc else 0, rank, world_size,
                                     1 if image_partition else 0, num_iterations, capi.ptr(out)) != 0:
        _err(lib, "evplp_host_technique_plan")
    return [dict(zip(PLAN_FIELDS, row)) for row in out]


def export_scene(name, out_dir, seed=1, detail=8, res_x=1280, res_y=720):
    lib = load_host_library()
    if lib.evplp_host_export_scene(name.encode(), out_dir.encode(), seed, detail, res_x, res_y) != 0:
        _err(lib, "evplp_host_export_scene")


MIS_NAMES = ("one", "balance", "max", "power2", "geometryClamp", "geometryBrdfClamp")
_FAM_KEYS = ("numLightPaths", "numVplLightPaths", "numMaxBounces", "radiusPercentage", "misMode", "DoProgressive", "AlphaProgressive",
             "forceVsl", "vplSplat", "photonSplat", "frameMode", "rngOffset", "numMaxIteration", "timeLimitMs", "clampingValue",
             "vslRadiusPercentage", "useJitter")
_PT_KEYS = ("rngOffset", "numMaxIteration", "timeLimitMs", "numMaxBounces", "numSamplePerPixel", "frameMode", "useJitter")


def config_check(host_scene, json_path):
    """What the C++ host (main.cpp dispatch, RtStableCamera, the techniques' parse()) understands of a scene JSON file,
    evaluated against `host_scene` instead of the OBJ files the JSON names.  No device is touched."""
    lib = load_host_library()
    out = (C.c_double * 80)()
    if lib.evplp_host_config_check(host_scene.h, json_path.encode(), out) != 0:
        _err(lib, "evplp_host_config_check")
    r = {"resX": out[3], "resY": out[4], "camera": {"origin": list(out[5:8]), "lookAt": list(out[8:11]), "up": list(out[11:14]), "fovy": out[14]}}
    for f, name in enumerate(("photonfam", "lvcphotonfam")):
        if out[1 + f]:
            r[name] = dict(zip(_FAM_KEYS, out[16 + 24 * f:16 + 24 * f + len(_FAM_KEYS)]))
    if out[0]:
        r["pt"] = dict(zip(_PT_KEYS, out[64:64 + len(_PT_KEYS)]))
    return r


def decode_jpeg(data):
    """JPEG bytes -> (uint8 [H, W, 3] top-down, file channel count): what stbi_load(path, .., 3) gives the reference
    (rtcommon.h:144) before its vertical flip."""
    lib = load_host_library()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    w, h, ch = C.c_int32(), C.c_int32(), C.c_int32()
    if lib.evplp_host_jpeg_info(buf, len(data), C.byref(w), C.byref(h), C.byref(ch)) != 0:
        _err(lib, "evplp_host_jpeg_info")
    out = np.empty((h.value, w.value, 3), dtype=np.uint8)
    if lib.evplp_host_jpeg_decode(buf, len(data), out.ctypes.data_as(_P), out.size) != 0:
        _err(lib, "evplp_host_jpeg_decode")
    return out, ch.value


def decode_png(data, want_channels=0):
    """PNG bytes -> (uint8 [H, W, C] top-down, file channel count): what stbi_load_from_memory(.., want_channels) returns."""
    lib = load_host_library()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    w, h, ch = C.c_int32(), C.c_int32(), C.c_int32()
    if lib.evplp_host_png_info(buf, len(data), C.byref(w), C.byref(h), C.byref(ch)) != 0:
        _err(lib, "evplp_host_png_info")
    c = want_channels or ch.value
    out = np.empty((h.value, w.value, c), dtype=np.uint8)
    if lib.evplp_host_png_decode(buf, len(data), want_channels, out.ctypes.data_as(_P), out.size) != 0:
        _err(lib, "evplp_host_png_decode")
    return out, ch.value


def realtime_probe(before_true, after_true, close_after=-1):
    """(loop passes, frames presented, afterSwap calls) of the headless RealTime::loop for scripted callbacks."""
    lib = load_host_library()
    out = (C.c_uint64 * 3)()
    lib.evplp_host_realtime_probe(before_true, after_true, close_after, out)
    return tuple(out)


def pfm_masked_error(a, b, mask_png, relative=True):
    lib = load_host_library()
    lib.evplp_host_pfm_masked_error.restype = C.c_float
    lib.evplp_host_pfm_masked_error.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    v = lib.evplp_host_pfm_masked_error(a.encode(), b.encode(), mask_png.encode(), 1 if relative else 0)
    if v < 0:
        _err(lib, "evplp_host_pfm_masked_error")
    return v


def load_texture(path, gamma=1.0):
    """RtTexture(filepath, gamma) (rtcommon.h:139-194): float32 [H, W, 4], row 0 = bottom, alpha 0."""
    lib = load_host_library()
    w, h = C.c_int32(), C.c_int32()
    if lib.evplp_host_texture_load(path.encode(), gamma, C.byref(w), C.byref(h), None, 0) != 0:
        _err(lib, "evplp_host_texture_load")
    out = np.empty((h.value, w.value, 4), dtype=np.float32)
    if lib.evplp_host_texture_load(path.encode(), gamma, C.byref(w), C.byref(h), out.ctypes.data_as(_P), out.size) != 0:
        _err(lib, "evplp_host_texture_load")
    return out


def render_json(json_path, device=0):
    lib = load_host_library()
    if lib.evplp_host_render_json(json_path.encode(), device) != 0:
        _err(lib, "evplp_host_render_json")
