"""Test-side ctypes bindings of the two CHECKERS (test infrastructure, never used by the product):

  * Ref    -- oracle/_ref/libvdbref.so : the unmodified OpenVDB/NanoVDB reference behind oracle/ref_driver.cc
  * Oracle -- oracle/libvdbrt_oracle.so: the CPU restatement (oracle/vdbrt_oracle.cc)

Both speak the PODs of include/vdbrt.h (mirrored in openvdb_b200/_abi.py).
"""
import ctypes as C
import os

import numpy as np

from openvdb_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvdbref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "libvdbrt_oracle.so")

FLT_MAX = float(np.finfo(np.float32).max)


class CameraDesc(C.Structure):
    """constructor arguments of tools::PerspectiveCamera / OrthographicCamera (+ optional lookAt)"""
    _fields_ = [("kind", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("use_lookat", C.c_uint32),
                ("rotation", C.c_double * 3), ("translation", C.c_double * 3), ("focal_or_frame", C.c_double),
                ("aperture", C.c_double), ("near_plane", C.c_double), ("far_plane", C.c_double),
                ("target", C.c_double * 3), ("up", C.c_double * 3)]


def camera_desc(width, height, translation=(0, 0, 0), rotation=(0, 0, 0), lookat=None, up=(0, 1, 0), kind=0,
                focal=None, aperture=None, frame=1.0, near=None, far=None):
    """vdb_render's values by default: floats widened to double"""
    d = CameraDesc()
    d.kind, d.width, d.height = kind, width, height
    d.rotation = abi.vec3(rotation)
    d.translation = abi.vec3(translation)
    if kind == abi.CAMERA_PERSPECTIVE:
        d.focal_or_frame = float(np.float32(50.0)) if focal is None else focal
    else:
        d.focal_or_frame = frame
    d.aperture = float(np.float32(41.2136)) if aperture is None else aperture
    d.near_plane = float(np.float32(1e-3)) if near is None else near
    d.far_plane = FLT_MAX if far is None else far
    d.use_lookat = 0 if lookat is None else 1
    d.target = abi.vec3(lookat if lookat is not None else (0, 0, 0))
    d.up = abi.vec3(up)
    return d


def shader(kind=abi.SHADER_DIFFUSE, rgba=(1, 1, 1, 1), bbox_min=(0, 0, 0), inv_dim=(1, 1, 1)):
    s = abi.Shader()
    s.kind = kind
    s.rgba = (C.c_float * 4)(*rgba)
    s.bbox_min = abi.vec3(bbox_min)
    s.inv_dim = abi.vec3(inv_dim)
    return s


def new_film(width, height, rgba=(0, 0, 0, 1)):
    f = np.empty((height, width, 4), np.float32)
    f[...] = np.asarray(rgba, np.float32)
    return f


class AuxArrays:
    def __init__(self, width, height):
        n = width * height
        self.hit = np.zeros(n, np.uint8)
        self.ijk = np.zeros((n, 3), np.int32)
        self.t_index = np.zeros(n, np.float64)
        self.t_world = np.zeros(n, np.float64)
        self.xyz = np.zeros((n, 3), np.float64)
        self.nml = np.zeros((n, 3), np.float64)

    def pod(self):
        a = abi.Aux()
        for name in ("hit", "ijk", "t_index", "t_world", "xyz", "nml"):
            setattr(a, name, getattr(self, name).ctypes.data)
        return a


def rays_array(n):
    return (abi.Ray * n)()


def make_rays(eyes, dirs, t0=1e-9, t1=np.finfo(np.float64).max):
    eyes = np.atleast_2d(np.asarray(eyes, np.float64))
    dirs = np.atleast_2d(np.asarray(dirs, np.float64))
    n = len(eyes)
    flat = np.empty((n, 8), np.float64)                      # vdbrt_ray: eye[3], dir[3], t0, t1
    flat[:, 0:3] = eyes; flat[:, 3:6] = dirs
    flat[:, 6] = np.broadcast_to(np.asarray(t0, np.float64), (n,)); flat[:, 7] = np.broadcast_to(np.asarray(t1, np.float64), (n,))
    r = rays_array(n)
    C.memmove(r, flat.ctypes.data, flat.nbytes)
    return r


def hits_to_dict(h, n):
    a = np.frombuffer(h, dtype=np.dtype([("hit", "<i4"), ("ijk", "<i4", 3), ("t_index", "<f8"), ("t_world", "<f8"),
                                         ("xyz_index", "<f8", 3), ("xyz_world", "<f8", 3), ("nml", "<f8", 3)]), count=n)
    return a


class Ref:
    """the unmodified reference"""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (build with `make -C oracle ref`; needs /root/reference)")
        L = self.L = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.vdbref_last_error.restype = C.c_char_p
        for name in ("vdbref_grid_sphere", "vdbref_grid_nano_sphere"):
            getattr(L, name).restype = vp
            getattr(L, name).argtypes = [C.c_double] * 6
        L.vdbref_grid_torus.restype = vp
        L.vdbref_grid_torus.argtypes = [C.c_double] * 7
        L.vdbref_grid_fog_from_levelset.restype = vp
        L.vdbref_grid_fog_from_levelset.argtypes = [vp]
        L.vdbref_grid_spheres_union.restype = vp
        L.vdbref_grid_spheres_union.argtypes = [vp, C.c_uint32, C.c_double, C.c_double]
        L.vdbref_grid_spheres_union_mt.restype = vp
        L.vdbref_grid_spheres_union_mt.argtypes = [vp, C.c_uint32, C.c_double, C.c_double, C.c_int]
        L.vdbref_grid_custom.restype = vp
        L.vdbref_grid_custom.argtypes = [C.c_float, C.c_uint32, C.c_double, vp, vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64]
        L.vdbref_grid_from_nanovdb.restype = vp
        L.vdbref_grid_from_nanovdb.argtypes = [vp, C.c_uint64]
        L.vdbref_grid_free.argtypes = [vp]
        L.vdbref_grid_nanovdb.restype = C.c_uint64
        L.vdbref_grid_nanovdb.argtypes = [vp, C.POINTER(vp)]
        L.vdbref_grid_nanovdb_quantized.restype = C.c_uint64
        L.vdbref_grid_nanovdb_quantized.argtypes = [vp, C.c_uint32, C.c_int, C.c_float, C.POINTER(vp)]
        L.vdbref_grid_stats.argtypes = [vp, vp, vp, vp, vp]
        L.vdbref_grid_probe.argtypes = [vp, vp, C.c_uint64, vp, vp]
        L.vdbref_camera_pod.argtypes = [C.POINTER(CameraDesc), C.POINTER(abi.Camera)]
        L.vdbref_camera_rays.argtypes = [C.POINTER(CameraDesc), vp, vp, C.c_uint64, vp]
        L.vdbref_jitter_table.argtypes = [C.c_uint, vp]
        L.vdbref_render_levelset.restype = C.c_double
        L.vdbref_render_levelset.argtypes = [vp, C.POINTER(CameraDesc), C.POINTER(abi.Shader), C.c_float, C.c_uint32,
                                             C.c_uint, C.c_int, vp]
        L.vdbref_levelset_records.restype = C.c_int64
        L.vdbref_levelset_records.argtypes = [vp, C.POINTER(CameraDesc), C.c_float, C.POINTER(abi.Aux),
                                              C.POINTER(abi.Counters)]
        L.vdbref_intersect_levelset.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_float, vp]
        L.vdbref_volume_spans.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp]
        L.vdbref_render_volume.restype = C.c_double
        L.vdbref_render_volume.argtypes = [vp, C.POINTER(CameraDesc), C.POINTER(abi.VolOpts), C.c_int, vp]
        L.vdbref_vol_defaults.argtypes = [C.POINTER(abi.VolOpts)]
        L.vdbref_set_threads.argtypes = [C.c_int]
        L.vdbref_hardware_threads.restype = C.c_int

    def err(self):
        return (self.L.vdbref_last_error() or b"").decode()

    # ---- grids
    def _chk(self, h):
        if not h:
            raise RuntimeError("reference: " + self.err())
        return h

    def sphere(self, radius, center=(0, 0, 0), voxel=1.0, half_width=3.0):
        return self._chk(self.L.vdbref_grid_sphere(radius, *[float(c) for c in center], voxel, half_width))

    def nano_sphere(self, radius, center=(0, 0, 0), voxel=1.0, half_width=3.0):
        return self._chk(self.L.vdbref_grid_nano_sphere(radius, *[float(c) for c in center], voxel, half_width))

    def torus(self, R, r, center=(0, 0, 0), voxel=1.0, half_width=3.0):
        return self._chk(self.L.vdbref_grid_torus(R, r, *[float(c) for c in center], voxel, half_width))

    def fog_from_levelset(self, ls):
        return self._chk(self.L.vdbref_grid_fog_from_levelset(ls))

    def spheres_union(self, spheres, voxel=1.0, half_width=3.0):
        s = np.ascontiguousarray(spheres, np.float64).reshape(-1, 4)
        return self._chk(self.L.vdbref_grid_spheres_union(s.ctypes.data, len(s), voxel, half_width))

    def spheres_union_mt(self, spheres, threads, voxel=1.0, half_width=3.0):
        """the same union folded by `threads` workers (order-independent; builds BASELINE config 4 in seconds)"""
        s = np.ascontiguousarray(spheres, np.float64).reshape(-1, 4)
        return self._chk(self.L.vdbref_grid_spheres_union_mt(s.ctypes.data, len(s), voxel, half_width, int(threads)))

    def custom(self, background=0.0, grid_class=abi.GRID_CLASS_FOG_VOLUME, voxel=1.0, translation=(0, 0, 0), voxels=(), boxes=()):
        """voxels: [((i,j,k), value)], boxes: [((min xyz),(max xyz), value, active)]"""
        ijk = np.ascontiguousarray([v[0] for v in voxels], np.int32).reshape(-1, 3)
        val = np.ascontiguousarray([v[1] for v in voxels], np.float32)
        bx = np.ascontiguousarray([list(b[0]) + list(b[1]) for b in boxes], np.int32).reshape(-1, 6)
        bv = np.ascontiguousarray([b[2] for b in boxes], np.float32)
        ba = np.ascontiguousarray([1 if b[3] else 0 for b in boxes], np.uint8)
        t = np.ascontiguousarray(translation, np.float64)
        return self._chk(self.L.vdbref_grid_custom(background, grid_class, voxel, t.ctypes.data, ijk.ctypes.data, val.ctypes.data, len(ijk),
                                                   bx.ctypes.data, bv.ctypes.data, ba.ctypes.data, len(bx)))

    def from_nanovdb(self, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        return self._chk(self.L.vdbref_grid_from_nanovdb(buf.ctypes.data, buf.size))

    def free(self, g):
        self.L.vdbref_grid_free(g)

    def nanovdb(self, g):
        """serialised NanoGrid<float> as a 32-byte aligned uint8 array (a copy)"""
        p = C.c_void_p()
        n = self.L.vdbref_grid_nanovdb(g, C.byref(p))
        if n == 0:
            raise RuntimeError("reference: " + self.err())
        return aligned_copy(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)))

    def nanovdb_quantized(self, g, grid_type, dither=False, tolerance=-1.0):
        """createNanoGrid<FloatGrid, Fp4|Fp8|Fp16|FpN> (grid_type 13..16) as a 32-byte aligned uint8 array (a copy)"""
        p = C.c_void_p()
        n = self.L.vdbref_grid_nanovdb_quantized(g, grid_type, 1 if dither else 0, tolerance, C.byref(p))
        if n == 0:
            raise RuntimeError("reference: " + self.err())
        return aligned_copy(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)))

    def stats(self, g):
        st = np.zeros(3, np.uint64); nb = np.zeros(6, np.int32); vb = np.zeros(6, np.int32); bg = np.zeros(1, np.float32)
        self.L.vdbref_grid_stats(g, st.ctypes.data, nb.ctypes.data, vb.ctypes.data, bg.ctypes.data)
        return {"active_voxels": int(st[0]), "leaf_count": int(st[1]), "active_tiles": int(st[2]),
                "node_bbox": nb, "voxel_bbox": vb, "background": float(bg[0])}

    def probe(self, g, ijk):
        ijk = np.ascontiguousarray(ijk, np.int32).reshape(-1, 3)
        v = np.zeros(len(ijk), np.float32); a = np.zeros(len(ijk), np.uint8)
        self.L.vdbref_grid_probe(g, ijk.ctypes.data, len(ijk), v.ctypes.data, a.ctypes.data)
        return v, a

    # ---- cameras
    def camera_pod(self, desc):
        cam = abi.Camera()
        if self.L.vdbref_camera_pod(C.byref(desc), C.byref(cam)):
            raise RuntimeError("reference: " + self.err())
        return cam

    def camera_rays(self, desc, ij, offsets=None):
        ij = np.ascontiguousarray(ij, np.uint32).reshape(-1, 2)
        r = rays_array(len(ij))
        off = None if offsets is None else np.ascontiguousarray(offsets, np.float64)
        self.L.vdbref_camera_rays(C.byref(desc), ij.ctypes.data, None if off is None else off.ctypes.data, len(ij), r)
        return r

    def jitter_table(self, seed):
        out = np.zeros(16, np.float64)
        self.L.vdbref_jitter_table(seed, out.ctypes.data)
        return out

    # ---- colour grids (GridT = Vec3SGrid forms of the shaders)
    def color_grid(self, ls, voxel=1.0, translation=(0, 0, 0)):
        self.L.vdbref_color_grid.restype = C.c_void_p
        self.L.vdbref_color_grid.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        return self._chk(self.L.vdbref_color_grid(ls, voxel, abi.vec3(translation)))

    def color_nanovdb(self, c):
        self.L.vdbref_color_nanovdb.restype = C.c_uint64
        self.L.vdbref_color_nanovdb.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        p = C.c_void_p()
        n = self.L.vdbref_color_nanovdb(c, C.byref(p))
        if n == 0:
            raise RuntimeError("reference: " + self.err())
        return aligned_copy(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)))

    def render_levelset_color(self, g, c, desc, sh, film, iso=0.0, spp=1, seed=0, threaded=False):
        self.L.vdbref_render_levelset_color.restype = C.c_double
        self.L.vdbref_render_levelset_color.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CameraDesc), C.POINTER(abi.Shader), C.c_float,
                                                        C.c_uint32, C.c_uint, C.c_int, C.c_void_p]
        t = self.L.vdbref_render_levelset_color(g, c, C.byref(desc), C.byref(sh), iso, spp, seed, int(threaded), film.ctypes.data)
        if t < 0:
            raise RuntimeError(self.err())
        return t

    # ---- the path
    def render_levelset(self, g, desc, sh, film, iso=0.0, spp=1, seed=0, threaded=False):
        assert film.dtype == np.float32 and film.flags.c_contiguous
        t = self.L.vdbref_render_levelset(g, C.byref(desc), C.byref(sh), iso, spp, seed, int(threaded), film.ctypes.data)
        if t < 0:
            raise RuntimeError(self.err())
        return t

    def levelset_records(self, g, desc, iso=0.0):
        aux = AuxArrays(desc.width, desc.height)
        pod = aux.pod()
        ctr = abi.Counters()
        mism = self.L.vdbref_levelset_records(g, C.byref(desc), iso, C.byref(pod), C.byref(ctr))
        if mism < 0:
            raise RuntimeError(self.err())
        return aux, ctr, mism

    def intersect(self, g, rays, space=abi.SPACE_WORLD, iso=0.0):
        n = len(rays)
        h = (abi.Hit * n)()
        if self.L.vdbref_intersect_levelset(g, rays, n, space, iso, h):
            raise RuntimeError(self.err())
        return hits_to_dict(h, n)

    def intersect_iter(self, g, rays, iterations, space=abi.SPACE_WORLD, iso=0.0):
        """the stock intersector with LinearSearchImpl<FloatGrid, iterations> (0..3); no first-hit voxel; index-space rays: index outputs only"""
        n = len(rays)
        h = (abi.Hit * n)()
        self.L.vdbref_intersect_levelset_iter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_float, C.c_uint32, C.c_void_p]
        if self.L.vdbref_intersect_levelset_iter(g, rays, n, space, iso, iterations, h):
            raise RuntimeError(self.err())
        return hits_to_dict(h, n)

    def render_levelset_iter(self, g, desc, sh, film, iterations, iso=0.0, spp=1, seed=0, threaded=False):
        self.L.vdbref_render_levelset_iter.restype = C.c_double
        self.L.vdbref_render_levelset_iter.argtypes = [C.c_void_p, C.POINTER(CameraDesc), C.POINTER(abi.Shader), C.c_float, C.c_uint32, C.c_uint, C.c_int,
                                                       C.c_uint32, C.c_void_p]
        t = self.L.vdbref_render_levelset_iter(g, C.byref(desc), C.byref(sh), iso, spp, seed, int(threaded), iterations, film.ctypes.data)
        if t < 0:
            raise RuntimeError(self.err())
        return t

    def volume_spans(self, g, rays, space=abi.SPACE_WORLD, max_spans=16):
        n = len(rays)
        spans = np.zeros((n, max_spans, 2), np.float64); counts = np.zeros(n, np.int32)
        if self.L.vdbref_volume_spans(g, rays, n, space, max_spans, spans.ctypes.data, counts.ctypes.data):
            raise RuntimeError(self.err())
        return spans, counts

    def vol_defaults(self):
        o = abi.VolOpts()
        self.L.vdbref_vol_defaults(C.byref(o))
        return o

    def render_volume(self, g, desc, opts, film, threaded=False):
        t = self.L.vdbref_render_volume(g, C.byref(desc), C.byref(opts), int(threaded), film.ctypes.data)
        if t < 0:
            raise RuntimeError(self.err())
        return t

    def set_threads(self, n):
        self.L.vdbref_set_threads(int(n))

    def hardware_threads(self):
        return self.L.vdbref_hardware_threads()


def aligned_copy(a, align=32):
    """NanoVDB buffers must be 32-byte aligned"""
    a = np.asarray(a, np.uint8)
    raw = np.empty(a.size + align, np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + a.size]
    out[:] = a
    return out


class Oracle:
    """the CPU restatement"""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            raise FileNotFoundError(ORACLE_SO + " (build with `make -C oracle port`)")
        L = self.L = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_grid_open.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
        L.oracle_grid_close.argtypes = [vp]
        L.oracle_grid_get_info.argtypes = [vp, C.POINTER(abi.GridInfo)]
        L.oracle_grid_probe.argtypes = [vp, vp, C.c_uint64, vp, vp]
        L.oracle_render_levelset.argtypes = [vp, C.POINTER(abi.Camera), C.POINTER(abi.Shader), C.POINTER(abi.LsOpts),
                                             C.POINTER(abi.Film), C.POINTER(abi.Aux), C.POINTER(abi.Counters), C.c_int]
        L.oracle_render_volume.argtypes = [vp, C.POINTER(abi.Camera), C.POINTER(abi.VolOpts), C.POINTER(abi.Film),
                                           C.POINTER(abi.Counters), C.c_int]
        L.oracle_intersect_levelset.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_float, vp]
        L.oracle_volume_spans.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp]
        L.oracle_camera_rays.argtypes = [C.POINTER(abi.Camera), vp, vp, C.c_uint64, vp]
        L.oracle_film_over.argtypes = [vp, vp, C.c_uint64]
        L.oracle_dda_trace.argtypes = [C.POINTER(abi.Ray), C.c_int, C.c_int, vp]
        L.oracle_ray_clip.argtypes = [C.POINTER(abi.Ray), vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        self._keep = {}

    def err(self):
        return (self.L.oracle_last_error() or b"").decode()

    def check(self, code):
        if code != 0:
            raise OracleError(code, self.err())

    def open(self, buf):
        assert buf.ctypes.data % 32 == 0
        h = C.c_void_p()
        self.check(self.L.oracle_grid_open(buf.ctypes.data, buf.size, C.byref(h)))
        self._keep[h.value] = buf
        return h.value

    def close(self, g):
        self.L.oracle_grid_close(g)
        self._keep.pop(g, None)

    def info(self, g):
        i = abi.GridInfo()
        self.L.oracle_grid_get_info(g, C.byref(i))
        return i

    def probe(self, g, ijk):
        ijk = np.ascontiguousarray(ijk, np.int32).reshape(-1, 3)
        v = np.zeros(len(ijk), np.float32); a = np.zeros(len(ijk), np.uint8)
        self.L.oracle_grid_probe(g, ijk.ctypes.data, len(ijk), v.ctypes.data, a.ctypes.data)
        return v, a

    def open_color(self, buf):
        """a serialised NanoGrid<Vec3f> for the colour-grid shaders (the buffer must stay alive)"""
        self.L.oracle_color_open.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        h = C.c_void_p()
        self.check(self.L.oracle_color_open(buf.ctypes.data, buf.size, C.byref(h)))
        self._keep[h.value] = buf
        return h.value

    def render_levelset(self, g, cam, sh, film, iso=0.0, spp=1, jitter=None, part=None, aux=False, counters=False,
                        threads=1, color=None, iterations=0):
        H, W = film.shape[:2]
        o = abi.LsOpts()
        o.iso, o.spp = iso, spp
        o.iterations = iterations
        if jitter is not None:
            o.jitter = (C.c_double * 16)(*jitter)
        if part is not None:
            o.part = part
        f = abi.Film(film.ctypes.data, W, H, abi.MEM_HOST)
        ax = AuxArrays(W, H) if aux else None
        pod = ax.pod() if aux else None
        ctr = abi.Counters() if counters else None
        self.L.oracle_render_levelset_color.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.Shader),
                                                        C.POINTER(abi.LsOpts), C.POINTER(abi.Film), C.c_void_p, C.c_void_p, C.c_int]
        self.check(self.L.oracle_render_levelset_color(g, color, C.byref(cam), C.byref(sh), C.byref(o), C.byref(f),
                                                       C.byref(pod) if aux else None, C.byref(ctr) if counters else None,
                                                       threads))
        return ax, ctr

    def render_volume(self, g, cam, opts, film, counters=False, threads=1):
        H, W = film.shape[:2]
        f = abi.Film(film.ctypes.data, W, H, abi.MEM_HOST)
        ctr = abi.Counters() if counters else None
        self.check(self.L.oracle_render_volume(g, C.byref(cam), C.byref(opts), C.byref(f),
                                               C.byref(ctr) if counters else None, threads))
        return ctr

    def set_lazy_init(self, on):
        """test knob: tester.init's value is evaluated on demand (the CUDA kernels' evaluation order); same results, fewer refills"""
        self.L.oracle_set_lazy_init(1 if on else 0)

    def film_over(self, top, bottom):
        """top = top.over(bottom), Film::RGBA::over per pixel (tools/RayTracer.h:252-259)"""
        assert top.shape == bottom.shape and top.dtype == np.float32 and bottom.dtype == np.float32
        self.check(self.L.oracle_film_over(top.ctypes.data, np.ascontiguousarray(bottom).ctypes.data, top.shape[0] * top.shape[1]))

    def intersect(self, g, rays, space=abi.SPACE_WORLD, iso=0.0, iterations=0):
        n = len(rays)
        h = (abi.Hit * n)()
        self.L.oracle_intersect_levelset_iter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_float, C.c_uint32, C.c_void_p]
        self.check(self.L.oracle_intersect_levelset_iter(g, rays, n, space, iso, iterations, h))
        return hits_to_dict(h, n)

    def volume_spans(self, g, rays, space=abi.SPACE_WORLD, max_spans=16):
        n = len(rays)
        spans = np.zeros((n, max_spans, 2), np.float64); counts = np.zeros(n, np.int32)
        self.check(self.L.oracle_volume_spans(g, rays, n, space, max_spans, spans.ctypes.data, counts.ctypes.data))
        return spans, counts

    def camera_rays(self, cam, ij, offsets=None):
        ij = np.ascontiguousarray(ij, np.uint32).reshape(-1, 2)
        r = rays_array(len(ij))
        off = None if offsets is None else np.ascontiguousarray(offsets, np.float64)
        self.L.oracle_camera_rays(C.byref(cam), ij.ctypes.data, None if off is None else off.ctypes.data, len(ij), r)
        return r

    def dda_trace(self, ray, log2dim, max_steps=64):
        out = np.zeros((max_steps, 5), np.float64)
        n = self.L.oracle_dda_trace(C.byref(ray), log2dim, max_steps, out.ctypes.data)
        return out[:n]

    def ray_clip(self, ray, bbox):
        b = np.ascontiguousarray(bbox, np.int32)
        t0, t1 = C.c_double(), C.c_double()
        hit = self.L.oracle_ray_clip(C.byref(ray), b.ctypes.data, C.byref(t0), C.byref(t1))
        return bool(hit), t0.value, t1.value


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (abi.ERR_NAMES.get(code, code), msg))
        self.code = code


def rays_to_numpy(rays):
    return np.frombuffer(rays, dtype=np.dtype([("eye", "<f8", 3), ("dir", "<f8", 3), ("t0", "<f8"), ("t1", "<f8")]),
                         count=len(rays))
