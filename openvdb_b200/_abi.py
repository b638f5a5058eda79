"""ctypes mirrors of the POD structs in include/vdbrt.h (the C ABI of libvdbrt.so).

Used by the product binding (openvdb_b200/api.py); the test-side bindings of the checkers reuse the same PODs.
"""
import ctypes as C

MEM_HOST, MEM_DEVICE = 0, 1
CAMERA_PERSPECTIVE, CAMERA_ORTHOGRAPHIC = 0, 1
SHADER_MATTE, SHADER_NORMAL, SHADER_POSITION, SHADER_DIFFUSE = 0, 1, 2, 3
SPACE_WORLD, SPACE_INDEX = 0, 1
GRID_CLASS_UNKNOWN, GRID_CLASS_LEVEL_SET, GRID_CLASS_FOG_VOLUME = 0, 1, 2
LS_UNIFORM_BG = 1
ASYNC = 2
LS_ROUNDS_ON = 4
LS_ROUNDS_OFF = 8
LS_ORDER_ON = 16
LS_ORDER_OFF = 32

ERR_NAMES = {
    0: "OK", 1: "INVALID_ARG", 2: "BAD_GRID", 3: "NOT_FLOAT", 4: "NOT_LEVELSET", 5: "NONUNIFORM",
    6: "EMPTY_GRID", 7: "ISO_RANGE", 8: "SPP_ZERO", 9: "CUDA", 10: "UNSUPPORTED", 11: "NOMEM", 12: "IO",
}
ERR_INVALID_ARG, ERR_BAD_GRID, ERR_NOT_FLOAT, ERR_UNSUPPORTED, ERR_IO = 1, 2, 3, 10, 12


class Ray(C.Structure):
    _fields_ = [("eye", C.c_double * 3), ("dir", C.c_double * 3), ("t0", C.c_double), ("t1", C.c_double)]


class Camera(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("reserved", C.c_uint32),
                ("m", C.c_double * 16), ("eye", C.c_double * 3), ("dir", C.c_double * 3),
                ("scale_w", C.c_double), ("scale_h", C.c_double), ("t0", C.c_double), ("t1", C.c_double)]


class Shader(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("rgba", C.c_float * 4), ("reserved", C.c_uint32),
                ("bbox_min", C.c_double * 3), ("inv_dim", C.c_double * 3), ("color_grid", C.c_void_p)]


class Partition(C.Structure):
    _fields_ = [("tile_w", C.c_uint32), ("tile_h", C.c_uint32), ("rank", C.c_uint32), ("count", C.c_uint32)]


class LsOpts(C.Structure):
    _fields_ = [("iso", C.c_float), ("spp", C.c_uint32), ("jitter", C.c_double * 16), ("part", Partition),
                ("flags", C.c_uint32), ("iterations", C.c_uint32)]


class VolOpts(C.Structure):
    _fields_ = [("primary_step", C.c_double), ("shadow_step", C.c_double), ("cutoff", C.c_double),
                ("light_gain", C.c_double), ("light_dir", C.c_double * 3), ("light_color", C.c_double * 3),
                ("absorption", C.c_double * 3), ("scattering", C.c_double * 3), ("part", Partition),
                ("flags", C.c_uint32), ("spp", C.c_uint32), ("jitter", C.c_double * 16)]


class Film(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("memspace", C.c_uint32),
                ("bg_rgba", C.c_float * 4)]


class Aux(C.Structure):
    _fields_ = [("hit", C.c_void_p), ("ijk", C.c_void_p), ("t_index", C.c_void_p), ("t_world", C.c_void_p),
                ("xyz", C.c_void_p), ("nml", C.c_void_p)]


class Hit(C.Structure):
    _fields_ = [("hit", C.c_int32), ("ijk", C.c_int32 * 3), ("t_index", C.c_double), ("t_world", C.c_double),
                ("xyz_index", C.c_double * 3), ("xyz_world", C.c_double * 3), ("nml", C.c_double * 3)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "root_probes", "upper_probes", "lower_probes", "voxel_probes",
                                           "stencil_refills", "primary_samples", "shadow_samples", "shadow_rays",
                                           "hits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class GridInfo(C.Structure):
    _fields_ = [("bytes", C.c_uint64), ("active_voxels", C.c_uint64), ("leaf_count", C.c_uint32),
                ("lower_count", C.c_uint32), ("upper_count", C.c_uint32), ("root_tiles", C.c_uint32),
                ("index_bbox", C.c_int32 * 6), ("node_bbox", C.c_int32 * 6), ("voxel_size", C.c_double * 3),
                ("translation", C.c_double * 3), ("background", C.c_float), ("grid_class", C.c_uint32),
                ("source_type", C.c_uint32), ("leaf_kind", C.c_uint32), ("resident_bytes", C.c_uint64)]


class NvdbMeta(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("grid_bytes", C.c_uint64), ("file_bytes", C.c_uint64), ("active_voxels", C.c_uint64),
                ("grid_type", C.c_uint32), ("grid_class", C.c_uint32), ("codec", C.c_uint32), ("pad", C.c_uint32),
                ("index_bbox", C.c_int32 * 6), ("world_bbox", C.c_double * 6), ("voxel_size", C.c_double * 3)]


CODEC_NONE, CODEC_ZIP = 0, 1


def vec3(v):
    return (C.c_double * 3)(*[float(x) for x in v])
