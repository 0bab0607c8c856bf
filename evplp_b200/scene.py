"""Host-side scene arrays in the shape the C ABI takes (RtScene / RtMesh / RtMaterial /
RtAreaLight / RtStableCamera of the reference, rtcommon.h:464-467, 772-798, 548-598), plus
small procedural test scenes.  Pure numpy; no compute happens here.
"""
import ctypes as C
import math

import numpy as np

from . import _capi as capi

f32 = np.float32


class Material:
    """RtMaterial: three RGBA32F textures (rows bottom-up) + lightIntensity."""

    def __init__(self, kd=(0.5, 0.5, 0.5), ks=(0.0, 0.0, 0.0), exponent=1.0, light=(0, 0, 0, 0)):
        self.lambert = self._tex(kd)
        self.phong = self._tex(ks)
        self.exponent = self._tex(exponent)
        self.lightIntensity = np.asarray(light, dtype=f32)

    @staticmethod
    def _tex(v):
        a = np.asarray(v, dtype=f32)
        if a.ndim == 0:  # constant exponent -> 1x1 texture, .x used
            return np.array([[[a, a, a, 0.0]]], dtype=f32)
        if a.ndim == 1:  # constant colour -> 1x1 texture (rtcommon.h:80-90)
            return np.array([[[a[0], a[1], a[2], 0.0]]], dtype=f32)
        assert a.ndim == 3 and a.shape[2] == 4
        return np.ascontiguousarray(a)


class Mesh:
    def __init__(self, vertices, indices, mat, texcoords=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=f32).reshape(-1, 3)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        self.texcoords = None if texcoords is None else np.ascontiguousarray(texcoords, dtype=f32).reshape(-1, 2)
        self.mat = int(mat)


class Scene:
    """Meshes + materials + one area light mesh (RtScene)."""

    def __init__(self):
        self.meshes = []
        self.materials = []
        self.light_mesh = -1
        self.light_intensity = np.zeros(4, dtype=f32)  # un-scaled JSON intensity

    # RtScene::addAreaLight (rtcommon.h:772-798): intensity.rgb * pi, zero reflectance material
    def add_area_light(self, vertices, indices, intensity):
        self.light_intensity = np.asarray(intensity, dtype=f32)
        pre = self.light_precomputed()
        self.materials.append(Material(kd=(0, 0, 0), ks=(0, 0, 0), exponent=1.0, light=pre))
        self.meshes.append(Mesh(vertices, indices, len(self.materials) - 1))
        self.light_mesh = len(self.meshes) - 1

    def light_precomputed(self):
        i = self.light_intensity
        pi = f32(3.14159265358979323846)
        return np.array([i[0] * pi, i[1] * pi, i[2] * pi, i[3]], dtype=f32)

    @property
    def num_prims(self):
        return sum(len(m.indices) for m in self.meshes)

    def mesh_starts(self):
        s = [0]
        for m in self.meshes:
            s.append(s[-1] + len(m.indices))
        return np.asarray(s, dtype=np.int32)

    def triangles(self):
        """(numPrims, 3, 3) float32 in global primitive order."""
        return np.concatenate([m.vertices[m.indices] for m in self.meshes], axis=0)

    # RtScene::findBoundingSphereRadius (rtcommon.h:805-814): half the AABB diagonal, light included
    def bounding_sphere_radius(self):
        v = np.concatenate([m.vertices for m in self.meshes], axis=0)
        d = (v.max(axis=0) - v.min(axis=0)).astype(f32)
        return f32(math.sqrt(float(f32(d[0] * d[0]) + f32(d[1] * d[1]) + f32(d[2] * d[2])))) * f32(0.5)

    def descriptors(self):
        """ctypes arrays for evplp_upload_scene / the oracle; keeps the numpy buffers alive."""
        md = (capi.MeshDesc * len(self.meshes))()
        for k, m in enumerate(self.meshes):
            md[k].vertices = capi.ptr(m.vertices)
            md[k].texcoords = capi.ptr(m.texcoords) if m.texcoords is not None else None
            md[k].indices = capi.ptr(m.indices)
            md[k].numVertices = len(m.vertices)
            md[k].numTriangles = len(m.indices)
            md[k].matIndex = m.mat
        mt = (capi.MaterialDesc * len(self.materials))()
        for k, m in enumerate(self.materials):
            mt[k].lambertReflectance = capi.ptr(m.lambert); mt[k].lambertH, mt[k].lambertW = m.lambert.shape[:2]
            mt[k].phongReflectance = capi.ptr(m.phong); mt[k].phongH, mt[k].phongW = m.phong.shape[:2]
            mt[k].phongExponent = capi.ptr(m.exponent); mt[k].exponentH, mt[k].exponentW = m.exponent.shape[:2]
            for j in range(4):
                mt[k].lightIntensity[j] = float(m.lightIntensity[j])
        pre = (C.c_float * 4)(*[float(x) for x in self.light_precomputed()])
        disp = (C.c_float * 4)(*[float(x) for x in self.light_intensity])
        return md, mt, pre, disp


class Camera:
    """RtStableCamera (rtcommon.h:548-598): JSON "direction" is a look-at POINT; fovx -> fovy."""

    def __init__(self, origin, lookat, up, fovx_deg, aspect):
        o = np.asarray(origin, dtype=np.float64)
        c = np.asarray(lookat, dtype=np.float64)
        u = np.asarray(up, dtype=np.float64)
        f = c - o
        f /= np.linalg.norm(f)
        s = np.cross(f, u)
        s /= np.linalg.norm(s)
        uu = np.cross(s, f)
        self.origin = o.astype(f32)
        self.forward = f.astype(f32)
        self.right = s.astype(f32)
        self.up = uu.astype(f32)
        fovy = 2.0 * math.atan2(math.tan(math.radians(fovx_deg) * 0.5), aspect)
        self.tan_y = f32(math.tan(fovy * 0.5))
        self.tan_x = f32(float(self.tan_y) * aspect)


def make_params(cam, num_light_paths, num_vpl_paths, max_bounces, radius, mis_mode=capi.MIS_BALANCE, clamp=0.0,
                jitter=(0.0, 0.0), accumulate=True, vsl_radius=0.0, rng_seed=0):
    p = capi.Params()
    for k in range(3):
        p.cameraPosition[k] = float(cam.origin[k])
        p.camForward[k] = float(cam.forward[k])
        p.camRight[k] = float(cam.right[k])
        p.camUp[k] = float(cam.up[k])
    p.tanHalfFovX = float(cam.tan_x)
    p.tanHalfFovY = float(cam.tan_y)
    p.jitter[0], p.jitter[1] = float(jitter[0]), float(jitter[1])
    p.nearDist, p.farDist = 0.1, 100.0
    p.numLightPaths = num_light_paths
    p.numVplLightPaths = num_vpl_paths
    p.numPhotonsPerLightPath = max_bounces + 1
    p.radius = float(radius)
    inv_pi = f32(0.318309886183790671537767526745028724068919291480912897495)
    r = f32(radius)
    with np.errstate(divide="ignore", invalid="ignore"):
        # mPrecomptedPdfMc (rtcomphoton.h:120)
        p.pdfMc = float(f32(num_vpl_paths) / f32(max(num_light_paths, 1)) * inv_pi / (r * r)) if num_light_paths else 0.0
    p.misMode = mis_mode
    p.clampingValue = float(clamp)
    p.doAccumulate = 1 if accumulate else 0
    p.vslRadius = float(vsl_radius)
    vr = f32(vsl_radius)
    p.vslInvPiRadius2 = float(inv_pi / (vr * vr)) if vsl_radius > 0 else 0.0
    p.rngSeed = rng_seed
    return p


# ----------------------------------------------------------------------------------------
# procedural test scenes
# ----------------------------------------------------------------------------------------
def _quad(p0, p1, p2, p3):
    """two CCW triangles p0 p1 p2, p0 p2 p3"""
    return [p0, p1, p2, p3], [[0, 1, 2], [0, 2, 3]]


def _box(lo, hi, inward=False):
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    v = [(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), (x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)]
    # outward-facing CCW
    faces = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]
    idx = []
    for a, b, c, d in faces:
        if inward:
            a, b, c, d = d, c, b, a
        idx += [[a, b, c], [a, c, d]]
    return np.asarray(v, dtype=f32), np.asarray(idx, dtype=np.int32)


def _grid(origin, du, dv, nu, nv):
    """(nu x nv) quad grid; normal = cross(du, dv)."""
    o = np.asarray(origin, dtype=np.float64)
    du = np.asarray(du, dtype=np.float64)
    dv = np.asarray(dv, dtype=np.float64)
    verts = np.array([o + du * (i / nu) + dv * (j / nv) for j in range(nv + 1) for i in range(nu + 1)])
    uv = np.array([(i / nu, j / nv) for j in range(nv + 1) for i in range(nu + 1)])
    idx = []
    for j in range(nv):
        for i in range(nu):
            a = j * (nu + 1) + i
            b, c, d = a + 1, a + nu + 2, a + nu + 1
            idx += [[a, b, c], [a, c, d]]
    return verts.astype(f32), np.asarray(idx, dtype=np.int32), uv.astype(f32)


def checker_texture(w, h, c0, c1, cells=4):
    t = np.zeros((h, w, 4), dtype=f32)
    for j in range(h):
        for i in range(w):
            c = c0 if ((i * cells // w) + (j * cells // h)) % 2 == 0 else c1
            t[j, i, :3] = c
    return t


def cornell_scene(seed=1, detail=4, glossy=True, light_exponent=0.0, light_grid=2):
    """A closed room (Z up) with two boxes, a textured floor and a ceiling area light.
    `detail` subdivides the walls so that the BVH has a few thousand triangles."""
    rng = np.random.RandomState(seed)
    sc = Scene()
    white = Material(kd=(0.7, 0.7, 0.7), ks=(0.05, 0.05, 0.05) if glossy else (0, 0, 0), exponent=10.0)
    red = Material(kd=(0.6, 0.1, 0.1))
    green = Material(kd=(0.1, 0.6, 0.1))
    shiny = Material(kd=(0.2, 0.2, 0.25), ks=(0.5, 0.5, 0.5) if glossy else (0, 0, 0), exponent=40.0)
    floor = Material(kd=checker_texture(16, 16, (0.7, 0.7, 0.7), (0.3, 0.3, 0.5)),
                     ks=(0.1, 0.1, 0.1) if glossy else (0, 0, 0), exponent=20.0)
    sc.materials += [white, red, green, shiny, floor]
    X, Y, Zh = 10.0, 10.0, 8.0
    n = detail
    # floor (normal +z), ceiling (-z), back wall (y = Y, normal -y), left (x=0, +x), right (x=X, -x), front (y=0, +y)
    v, i, uv = _grid((0, 0, 0), (X, 0, 0), (0, Y, 0), n, n); sc.meshes.append(Mesh(v, i, 4, uv * 2.0))
    v, i, uv = _grid((0, Y, Zh), (X, 0, 0), (0, -Y, 0), n, n); sc.meshes.append(Mesh(v, i, 0, uv))
    v, i, uv = _grid((0, Y, 0), (X, 0, 0), (0, 0, Zh), n, n); sc.meshes.append(Mesh(v, i, 0, uv))
    v, i, uv = _grid((0, 0, 0), (0, Y, 0), (0, 0, Zh), n, n); sc.meshes.append(Mesh(v, i, 1, uv))
    v, i, uv = _grid((X, Y, 0), (0, -Y, 0), (0, 0, Zh), n, n); sc.meshes.append(Mesh(v, i, 2, uv))
    v, i, uv = _grid((X, 0, 0), (-X, 0, 0), (0, 0, Zh), n, n); sc.meshes.append(Mesh(v, i, 0, uv))
    # two boxes, one rotated
    v, i = _box((1.5, 5.0, 0.0), (4.0, 7.5, 5.0)); sc.meshes.append(Mesh(v, i, 3))
    v, i = _box((-1.2, -1.2, 0.0), (1.2, 1.2, 2.4))
    ang = 0.5
    rot = np.array([[math.cos(ang), -math.sin(ang), 0], [math.sin(ang), math.cos(ang), 0], [0, 0, 1]])
    v = (v.astype(np.float64) @ rot.T + np.array([6.5, 3.5, 0.0])).astype(f32)
    sc.meshes.append(Mesh(v, i, 0))
    # some random small triangles floating around (irregular geometry for the BVH)
    nt = 40 * detail
    c = rng.uniform([1, 1, 1], [9, 9, 6], size=(nt, 1, 3))
    tri = (c + rng.uniform(-0.25, 0.25, size=(nt, 3, 3))).astype(f32)
    sc.meshes.append(Mesh(tri.reshape(-1, 3), np.arange(nt * 3, dtype=np.int32).reshape(-1, 3), 3))
    # ceiling light: light_grid x light_grid quads, facing down
    v, i, _ = _grid((3.5, 6.5, Zh - 0.02), (3.0, 0, 0), (0, -3.0, 0), light_grid, light_grid)
    sc.add_area_light(v, i, (17.0, 12.0, 4.0, light_exponent))
    cam = dict(origin=(5.0, 0.3, 3.0), lookat=(5.0, 9.0, 6.0), up=(0, 0, 1), fovx=70.0)
    return sc, cam
