"""Python binding of the C ABI in include/vdbrt.h (libvdbrt.so, hand-written sm_100a CUDA).

This is plumbing for tests and bench.py: the product is the shared library and its C / C++ interface
(include/vdbrt.h, include/vdbrt/RayTracer.h).  There is no CPU fallback here -- if the library is missing or no
CUDA device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VDBRT_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvdbrt.so")

# every symbol include/vdbrt.h declares
SYMBOLS = [
    "vdbrt_create", "vdbrt_destroy", "vdbrt_last_error", "vdbrt_device_count", "vdbrt_set_stream", "vdbrt_synchronize",
    "vdbrt_host_alloc", "vdbrt_host_free", "vdbrt_host_register", "vdbrt_host_unregister", "vdbrt_upload_grid", "vdbrt_free_grid", "vdbrt_grid_get_info",
    "vdbrt_grid_download", "vdbrt_camera_perspective", "vdbrt_camera_orthographic", "vdbrt_camera_look_at", "vdbrt_camera_get_rays",
    "vdbrt_jitter_table", "vdbrt_vol_opts_default", "vdbrt_render_levelset", "vdbrt_render_volume",
    "vdbrt_intersect_levelset", "vdbrt_volume_spans", "vdbrt_count_levelset", "vdbrt_count_volume",
    "vdbrt_last_kernel_ms", "vdbrt_build_levelset_sphere", "vdbrt_build_levelset_torus",
    "vdbrt_build_levelset_spheres", "vdbrt_build_fog_from_levelset", "vdbrt_random_spheres",
    "vdbrt_device_alloc", "vdbrt_device_free", "vdbrt_ipc_export", "vdbrt_ipc_import", "vdbrt_ipc_close", "vdbrt_memcpy",
    "vdbrt_upload_color_grid", "vdbrt_nvdb_list", "vdbrt_nvdb_read", "vdbrt_nvdb_read_typed", "vdbrt_nvdb_write", "vdbrt_buffer_free", "vdbrt_film_save_ppm", "vdbrt_film_over", "vdbrt_set_tuning", "vdbrt_intersect_levelset_ex", "vdbrt_volume_clip",
]


class VdbrtError(RuntimeError):
    """Non-zero status from the C ABI; `.code` is the VDBRT_ERR_* value.  The reference raises
    openvdb::RuntimeError / ValueError for the same conditions (tools/RayIntersector.h:100-112,305-311,533-539)."""

    def __init__(self, code, msg):
        super().__init__("%s: %s" % (abi.ERR_NAMES.get(code, code), msg))
        self.code = code


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            LIB_PATH + " is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C openvdb_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_double
    P = C.POINTER
    L.vdbrt_last_error.restype = C.c_char_p
    L.vdbrt_create.argtypes = [C.c_int, P(vp)]
    L.vdbrt_destroy.argtypes = [vp]
    L.vdbrt_destroy.restype = None
    L.vdbrt_set_stream.argtypes = [vp, vp]
    L.vdbrt_synchronize.argtypes = [vp]
    L.vdbrt_host_alloc.argtypes = [C.c_size_t, P(vp)]
    L.vdbrt_host_free.argtypes = [vp]
    L.vdbrt_host_register.argtypes = [vp, C.c_size_t]
    L.vdbrt_host_unregister.argtypes = [vp]
    L.vdbrt_upload_grid.argtypes = [vp, vp, u64, u32, P(vp)]
    L.vdbrt_free_grid.argtypes = [vp, vp]
    L.vdbrt_upload_color_grid.argtypes = [vp, vp, u64, u32, P(vp)]
    L.vdbrt_grid_get_info.argtypes = [vp, P(abi.GridInfo)]
    L.vdbrt_grid_download.argtypes = [vp, vp, vp, u64]
    L.vdbrt_camera_perspective.argtypes = [P(abi.Camera), u32, u32, P(dbl), P(dbl), dbl, dbl, dbl, dbl]
    L.vdbrt_camera_orthographic.argtypes = [P(abi.Camera), u32, u32, P(dbl), P(dbl), dbl, dbl, dbl]
    L.vdbrt_camera_look_at.argtypes = [P(abi.Camera), P(dbl), P(dbl)]
    L.vdbrt_camera_get_rays.argtypes = [P(abi.Camera), vp, vp, u64, vp]
    L.vdbrt_jitter_table.argtypes = [C.c_uint, P(dbl)]
    L.vdbrt_vol_opts_default.argtypes = [P(abi.VolOpts)]
    L.vdbrt_render_levelset.argtypes = [vp, vp, P(abi.Camera), P(abi.Shader), P(abi.LsOpts), P(abi.Film), P(abi.Aux)]
    L.vdbrt_render_volume.argtypes = [vp, vp, P(abi.Camera), P(abi.VolOpts), P(abi.Film)]
    L.vdbrt_intersect_levelset.argtypes = [vp, vp, vp, u64, u32, C.c_float, vp, u32]
    L.vdbrt_intersect_levelset_ex.argtypes = [vp, vp, vp, u64, u32, C.c_float, u32, vp, u32]
    L.vdbrt_volume_spans.argtypes = [vp, vp, vp, u64, u32, u32, vp, vp, u32]
    L.vdbrt_count_levelset.argtypes = [vp, vp, P(abi.Camera), P(abi.LsOpts), P(abi.Counters)]
    L.vdbrt_count_volume.argtypes = [vp, vp, P(abi.Camera), P(abi.VolOpts), P(abi.Counters)]
    L.vdbrt_last_kernel_ms.argtypes = [vp, P(C.c_float), P(u32)]
    L.vdbrt_set_tuning.argtypes = [vp, C.c_char_p, u32]
    L.vdbrt_build_levelset_sphere.argtypes = [vp, dbl, P(dbl), dbl, dbl, P(vp)]
    L.vdbrt_build_levelset_torus.argtypes = [vp, dbl, dbl, P(dbl), dbl, dbl, P(vp)]
    L.vdbrt_build_levelset_spheres.argtypes = [vp, vp, u32, dbl, dbl, P(vp)]
    L.vdbrt_build_fog_from_levelset.argtypes = [vp, vp, P(vp)]
    L.vdbrt_random_spheres.argtypes = [u64, u32, dbl, dbl, dbl, vp]
    L.vdbrt_device_alloc.argtypes = [vp, C.c_size_t, P(vp)]
    L.vdbrt_device_free.argtypes = [vp, vp]
    L.vdbrt_ipc_export.argtypes = [vp, vp, vp]
    L.vdbrt_ipc_import.argtypes = [vp, vp, P(vp)]
    L.vdbrt_ipc_close.argtypes = [vp, vp]
    L.vdbrt_memcpy.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
    L.vdbrt_nvdb_list.argtypes = [C.c_char_p, P(abi.NvdbMeta), u32, P(u32)]
    L.vdbrt_nvdb_read.argtypes = [C.c_char_p, C.c_char_p, P(vp), P(u64)]
    L.vdbrt_nvdb_read_typed.argtypes = [C.c_char_p, C.c_char_p, u32, P(vp), P(u64)]
    L.vdbrt_nvdb_write.argtypes = [C.c_char_p, vp, u64, u32]
    L.vdbrt_buffer_free.argtypes = [vp]
    L.vdbrt_film_save_ppm.argtypes = [C.c_char_p, vp, u32, u32]
    L.vdbrt_film_over.argtypes = [vp, P(abi.Film), P(abi.Film)]
    _lib = L
    return L


def _check(code):
    if code != 0:
        raise VdbrtError(code, (load_library().vdbrt_last_error() or b"").decode())


FLT_MAX = float(np.finfo(np.float32).max)


# ---- host-side helpers (no GPU needed) --------------------------------------------------------------------
def perspective_camera(width, height, rotation=(0, 0, 0), translation=(0, 0, 0), focal_length=50.0, aperture=41.2136,
                       near=1e-3, far=np.finfo(np.float64).max, lookat=None, up=(0, 1, 0)):
    """tools::PerspectiveCamera (tools/RayTracer.h:418-476) (+ optional BaseCamera::lookAt)"""
    cam = abi.Camera()
    _check(load_library().vdbrt_camera_perspective(C.byref(cam), width, height, abi.vec3(rotation), abi.vec3(translation),
                                                   focal_length, aperture, near, far))
    if lookat is not None:
        _check(load_library().vdbrt_camera_look_at(C.byref(cam), abi.vec3(lookat), abi.vec3(up)))
    return cam


def orthographic_camera(width, height, rotation=(0, 0, 0), translation=(0, 0, 0), frame_width=1.0, near=1e-3,
                        far=np.finfo(np.float64).max, lookat=None, up=(0, 1, 0)):
    """tools::OrthographicCamera (tools/RayTracer.h:479-513)"""
    cam = abi.Camera()
    _check(load_library().vdbrt_camera_orthographic(C.byref(cam), width, height, abi.vec3(rotation), abi.vec3(translation),
                                                    frame_width, near, far))
    if lookat is not None:
        _check(load_library().vdbrt_camera_look_at(C.byref(cam), abi.vec3(lookat), abi.vec3(up)))
    return cam


def vdb_render_camera(width, height, translation, lookat, rotation=(0, 0, 0), focal=None):
    """the perspective camera exactly as vdb_render builds it: float options widened to double (SURVEY 0.8,
    openvdb_cmd/vdb_render/main.cc:62,83-87,425-436)"""
    return perspective_camera(width, height, rotation, translation, float(np.float32(50.0 if focal is None else focal)),
                              float(np.float32(41.2136)), float(np.float32(1e-3)), FLT_MAX, lookat=lookat)


def jitter_table(seed=0):
    out = (C.c_double * 16)()
    _check(load_library().vdbrt_jitter_table(seed, out))
    return np.array(out[:], np.float64)


def vol_opts_default(spp=1, seed=0):
    """VolumeRender's defaults (tools/RayTracer.h:929-936); spp > 1 is this library's extension (include/vdbrt.h)"""
    o = abi.VolOpts()
    _check(load_library().vdbrt_vol_opts_default(C.byref(o)))
    if spp > 1:
        o.spp = spp
        o.jitter = (C.c_double * 16)(*jitter_table(seed))
    return o


def make_shader(kind=abi.SHADER_DIFFUSE, rgba=(1, 1, 1, 1), bbox_min=(0, 0, 0), inv_dim=(1, 1, 1), color_grid=None):
    """color_grid: a Grid from Context.upload_color -> the GridT = Vec3SGrid form of the shader (keep the Grid alive)"""
    s = abi.Shader()
    s.kind = kind
    s.rgba = (C.c_float * 4)(*rgba)
    s.bbox_min = abi.vec3(bbox_min)
    s.inv_dim = abi.vec3(inv_dim)
    s.color_grid = color_grid.handle if color_grid is not None else None
    return s


def camera_rays(cam, ij, offsets=None):
    """BaseCamera::getRay for an (n, 2) array of pixel indices (and optional (n, 2) offsets in [0, 1]); returns (abi.Ray * n)"""
    ij = np.ascontiguousarray(ij, dtype=np.uint32).reshape(-1, 2)
    off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64).reshape(-1, 2)
    rays = (abi.Ray * len(ij))()
    _check(load_library().vdbrt_camera_get_rays(C.byref(cam), ij.ctypes.data, None if off is None else off.ctypes.data, len(ij), rays))
    return rays


def partition(rank=0, count=1, tile_w=0, tile_h=0):
    return abi.Partition(tile_w, tile_h, rank, count)


class PinnedArray:
    """numpy view over cudaHostAlloc memory (vdbrt_host_alloc)"""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _check(load_library().vdbrt_host_alloc(max(n, 1), C.byref(p)))
        self.ptr = p.value
        buf = (C.c_uint8 * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            load_library().vdbrt_host_free(self.ptr)
            self.ptr = None


class Grid:
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle
        self.info = abi.GridInfo()
        _check(ctx.L.vdbrt_grid_get_info(handle, C.byref(self.info)))

    def download(self):
        out = np.empty(self.info.bytes + 32, np.uint8)
        off = (-out.ctypes.data) % 32
        out = out[off:off + self.info.bytes]
        _check(self.ctx.L.vdbrt_grid_download(self.ctx.handle, self.handle, out.ctypes.data, out.size))
        return out

    def free(self):
        if self.handle:
            self.ctx.L.vdbrt_free_grid(self.ctx.handle, self.handle)
            self.handle = None


class Context:
    """one GPU (vdbrt_ctx)"""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        _check(self.L.vdbrt_create(device, C.byref(h)))
        self.handle = h.value
        self.device = device

    def close(self):
        if self.handle:
            self.L.vdbrt_destroy(self.handle)
            self.handle = None

    def set_stream(self, cuda_stream_ptr):
        _check(self.L.vdbrt_set_stream(self.handle, cuda_stream_ptr))

    def synchronize(self):
        _check(self.L.vdbrt_synchronize(self.handle))

    # ---- grids
    def upload(self, buf):
        """buf: uint8 numpy array holding a serialised NanoGrid<float>"""
        buf = np.ascontiguousarray(buf, np.uint8)
        g = C.c_void_p()
        _check(self.L.vdbrt_upload_grid(self.handle, buf.ctypes.data, buf.size, abi.MEM_HOST, C.byref(g)))
        return Grid(self, g.value)

    def upload_color(self, buf):
        """buf: uint8 numpy array holding a serialised NanoGrid<Vec3f> (the colour grid of the colour-grid shaders)"""
        buf = np.ascontiguousarray(buf, np.uint8)
        g = C.c_void_p()
        _check(self.L.vdbrt_upload_color_grid(self.handle, buf.ctypes.data, buf.size, abi.MEM_HOST, C.byref(g)))
        return Grid(self, g.value)

    def upload_device(self, dev_ptr, nbytes):
        g = C.c_void_p()
        _check(self.L.vdbrt_upload_grid(self.handle, dev_ptr, nbytes, abi.MEM_DEVICE, C.byref(g)))
        return Grid(self, g.value)

    def build_sphere(self, radius, center=(0, 0, 0), voxel=1.0, half_width=3.0):
        g = C.c_void_p()
        _check(self.L.vdbrt_build_levelset_sphere(self.handle, radius, abi.vec3(center), voxel, half_width, C.byref(g)))
        return Grid(self, g.value)

    def build_torus(self, major, minor, center=(0, 0, 0), voxel=1.0, half_width=3.0):
        g = C.c_void_p()
        _check(self.L.vdbrt_build_levelset_torus(self.handle, major, minor, abi.vec3(center), voxel, half_width, C.byref(g)))
        return Grid(self, g.value)

    def build_spheres(self, spheres, voxel=1.0, half_width=3.0):
        s = np.ascontiguousarray(spheres, np.float64).reshape(-1, 4)
        g = C.c_void_p()
        _check(self.L.vdbrt_build_levelset_spheres(self.handle, s.ctypes.data, len(s), voxel, half_width, C.byref(g)))
        return Grid(self, g.value)

    def build_fog(self, levelset):
        g = C.c_void_p()
        _check(self.L.vdbrt_build_fog_from_levelset(self.handle, levelset.handle, C.byref(g)))
        return Grid(self, g.value)

    # ---- the hot path
    @staticmethod
    def _film_pod(film, width=None, height=None, memspace=abi.MEM_HOST, bg=(0, 0, 0, 1)):
        if isinstance(film, np.ndarray):
            assert film.dtype == np.float32 and film.flags.c_contiguous and film.shape[2] == 4
            height, width = film.shape[:2]
            ptr = film.ctypes.data
        else:
            ptr = int(film)
        f = abi.Film(ptr, width, height, memspace)
        f.bg_rgba = (C.c_float * 4)(*bg)
        return f

    def ls_opts(self, iso=0.0, spp=1, seed=0, part=None, uniform_bg=False, jitter=None, rounds=None, order=None, iterations=0):
        o = abi.LsOpts()
        o.iso, o.spp = iso, spp
        o.iterations = iterations       # LinearSearchImpl<GridT, Iterations>
        if spp > 1:
            j = jitter_table(seed) if jitter is None else jitter
            o.jitter = (C.c_double * 16)(*j)
        if part is not None:
            o.part = part
        o.flags = abi.LS_UNIFORM_BG if uniform_bg else 0
        if rounds is not None:      # long-ray rounds: None = library default (on for partitioned frames)
            o.flags |= abi.LS_ROUNDS_ON if rounds else abi.LS_ROUNDS_OFF
        if order is not None:       # heavy tiles first: None = library default (on when a warp gets two tiles or more)
            o.flags |= abi.LS_ORDER_ON if order else abi.LS_ORDER_OFF
        return o

    def render_levelset(self, grid, cam, shader, film, iso=0.0, spp=1, seed=0, part=None, aux=None, uniform_bg=False,
                        width=None, height=None, memspace=abi.MEM_HOST, bg=(0, 0, 0, 1), opts=None):
        """LevelSetRayTracer::render.  film: float32 (H,W,4) numpy array (host) or a device pointer (memspace=DEVICE)"""
        o = opts if opts is not None else self.ls_opts(iso, spp, seed, part, uniform_bg)
        f = self._film_pod(film, width, height, memspace, bg)
        _check(self.L.vdbrt_render_levelset(self.handle, grid.handle, C.byref(cam), C.byref(shader), C.byref(o), C.byref(f),
                                            C.byref(aux) if aux is not None else None))

    def film_over(self, top, bottom, width=None, height=None, memspace=abi.MEM_HOST):
        """top = top.over(bottom), Film::RGBA::over per pixel"""
        t, b = self._film_pod(top, width, height, memspace), self._film_pod(bottom, width, height, memspace)
        _check(self.L.vdbrt_film_over(self.handle, C.byref(t), C.byref(b)))

    def render_volume(self, grid, cam, opts, film, width=None, height=None, memspace=abi.MEM_HOST):
        f = self._film_pod(film, width, height, memspace)
        _check(self.L.vdbrt_render_volume(self.handle, grid.handle, C.byref(cam), C.byref(opts), C.byref(f)))

    def intersect(self, grid, rays, space=abi.SPACE_WORLD, iso=0.0, iterations=0):
        n = len(rays)
        hits = (abi.Hit * n)()
        _check(self.L.vdbrt_intersect_levelset_ex(self.handle, grid.handle, rays, n, space, iso, iterations, hits, abi.MEM_HOST))
        return hits

    def volume_spans(self, grid, rays, space=abi.SPACE_WORLD, max_spans=16):
        n = len(rays)
        spans = np.zeros((n, max_spans, 2), np.float64)
        counts = np.zeros(n, np.int32)
        _check(self.L.vdbrt_volume_spans(self.handle, grid.handle, rays, n, space, max_spans, spans.ctypes.data,
                                         counts.ctypes.data, abi.MEM_HOST))
        return spans, counts

    def count_levelset(self, grid, cam, iso=0.0, spp=1, seed=0):
        o = self.ls_opts(iso, spp, seed)
        c = abi.Counters()
        _check(self.L.vdbrt_count_levelset(self.handle, grid.handle, C.byref(cam), C.byref(o), C.byref(c)))
        return c

    def count_volume(self, grid, cam, opts):
        c = abi.Counters()
        _check(self.L.vdbrt_count_volume(self.handle, grid.handle, C.byref(cam), C.byref(opts), C.byref(c)))
        return c

    def set_tuning(self, **kw):
        """scheduling knobs (vdbrt_set_tuning): ls_strip, ls_refill, ls_eager, ls_order, ls_probe_cap, ls_probe_b, ..."""
        for k, v in kw.items():
            _check(self.L.vdbrt_set_tuning(self.handle, k.encode(), int(v)))

    def last_kernel_ms(self):
        ms, n = C.c_float(), C.c_uint32()
        _check(self.L.vdbrt_last_kernel_ms(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def nvdb_list(path):
    """nanovdb::io::readGridMetaData: one abi.NvdbMeta per grid of the file"""
    L = load_library()
    n = C.c_uint32(0)
    _check(L.vdbrt_nvdb_list(os.fsencode(path), None, 0, C.byref(n)))
    out = (abi.NvdbMeta * max(n.value, 1))()
    _check(L.vdbrt_nvdb_list(os.fsencode(path), out, n.value, C.byref(n)))
    return list(out[:n.value])


def nvdb_read(path, name=None, grid_type=1):
    """nanovdb::io::readGrid: the serialised grid as a 32-byte aligned uint8 array (first grid of that type when name is None;
    grid_type 1 = Float, 6 = Vec3f, 0 = any)"""
    L = load_library()
    p, n = C.c_void_p(), C.c_uint64(0)
    _check(L.vdbrt_nvdb_read_typed(os.fsencode(path), name.encode() if name else None, grid_type, C.byref(p), C.byref(n)))
    try:
        raw = np.empty(n.value + 32, np.uint8)
        off = (-raw.ctypes.data) % 32
        buf = raw[off:off + n.value]
        C.memmove(buf.ctypes.data, p, n.value)
    finally:
        L.vdbrt_buffer_free(p)
    return buf


def nvdb_write(path, buf, codec=abi.CODEC_NONE):
    """nanovdb::io::writeGrid of one serialised grid"""
    buf = np.ascontiguousarray(buf, np.uint8)
    _check(load_library().vdbrt_nvdb_write(os.fsencode(path), buf.ctypes.data, buf.size, codec))


def film_save_ppm(path, film):
    """tools::Film::savePPM"""
    f = np.ascontiguousarray(film, np.float32)
    _check(load_library().vdbrt_film_save_ppm(os.fsencode(path), f.ctypes.data, f.shape[1], f.shape[0]))


def random_spheres(n=10000, seed=20240607, extent=1988.0, rmin=10.0, rmax=60.0):
    """the sphere set of BASELINE configs 4/5: n x (cx, cy, cz, r)"""
    out = np.zeros((n, 4), np.float64)
    _check(load_library().vdbrt_random_spheres(seed, n, extent, rmin, rmax, out.ctypes.data))
    return out


class SharedFilm:
    """A device film that every rank of one node can write: rank 0 owns it (vdbrt_device_alloc), the other ranks map it through
    a CUDA IPC handle and render their tiles straight into it over NVLink.  `exchange` broadcasts the 64-byte handle."""

    def __init__(self, ctx, height, width, rank, exchange):
        self.ctx, self.rank, self.nbytes = ctx, rank, height * width * 16
        handle = np.zeros(64, np.uint8)
        p = C.c_void_p()
        if rank == 0:
            _check(ctx.L.vdbrt_device_alloc(ctx.handle, self.nbytes, C.byref(p)))
            _check(ctx.L.vdbrt_ipc_export(ctx.handle, p, handle.ctypes.data))
        handle = exchange(handle)
        if rank != 0:
            _check(ctx.L.vdbrt_ipc_import(ctx.handle, handle.ctypes.data, C.byref(p)))
        self.ptr = p.value

    def close(self):
        if self.ptr:
            if self.rank == 0:
                self.ctx.L.vdbrt_device_free(self.ctx.handle, self.ptr)
            else:
                self.ctx.L.vdbrt_ipc_close(self.ctx.handle, self.ptr)
            self.ptr = None


class SharedHostFilm:
    """One HOST film for all ranks of a node: a POSIX shared-memory segment (rank 0 creates it, `exchange` broadcasts its name)
    that every rank maps and page-locks (vdbrt_host_register).  Each rank renders its tiles with the film as an ordinary host
    film: its kernels store the pixels it owns over its own PCIe link, nobody copies or gathers anything."""

    def __init__(self, height, width, rank, exchange):
        from multiprocessing import shared_memory
        self.rank, self.nbytes = rank, height * width * 16
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
            exchange(self.shm.name)
        else:
            self.shm = shared_memory.SharedMemory(name=exchange(None))
            try:    # Python < 3.13 also registers ATTACHED segments with the resource tracker, which would unlink them at exit
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.array = np.ndarray((height, width, 4), np.float32, buffer=self.shm.buf)
        self.ptr = self.array.ctypes.data
        _check(load_library().vdbrt_host_register(self.ptr, self.nbytes))

    def close(self):
        if self.shm is not None:
            load_library().vdbrt_host_unregister(self.ptr)
            self.array = None
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
            self.shm = None


def memcpy(ctx, dst, src, nbytes, kind):
    """async copy on the context's stream; kind 0 H2D, 1 D2H, 2 D2D"""
    _check(ctx.L.vdbrt_memcpy(ctx.handle, dst, src, nbytes, kind))
