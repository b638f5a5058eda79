/* vdbrt.h -- C ABI of the B200-native VDB ray tracer (libvdbrt.so).
 *
 * This is the drop-in boundary for ONE hot path of OpenVDB 13.0.1: the per-pixel loops behind
 * openvdb::tools::LevelSetRayTracer / LevelSetRayIntersector / VolumeRender as driven by vdb_render,
 * executed by hand-written sm_100a CUDA kernels over a NanoVDB-serialised NanoGrid<float>.
 * Every entry point names the reference interface it replaces (paths relative to the OpenVDB tree).
 * Plain C: POD structs, raw pointers and sizes only.  There is NO CPU fallback behind this API: without a
 * CUDA device vdbrt_create() fails with VDBRT_ERR_CUDA.
 *
 * Conventions
 *   - every function returns an int status (VDBRT_OK == 0); vdbrt_last_error() gives a thread-local message.
 *     The reference reports the same conditions as C++ exceptions thrown at construction time
 *     (tools/RayIntersector.h:100-112,305-311,533-539; tools/RayTracer.h:877-879); the C++ facade
 *     (include/vdbrt/RayTracer.h) turns the codes back into exceptions of the same kind.
 *   - "memspace" says where a caller buffer lives: host memory (copied through the library's stream, fastest
 *     when allocated with vdbrt_host_alloc) or device memory of the context's GPU (used in place).
 *   - a context drives exactly one GPU; multi-GPU rendering is one context (one process) per GPU, each rendering
 *     the film tiles it owns (vdbrt_partition), with the frame gathered by the caller (NCCL / peer copy).
 */
#ifndef VDBRT_H_INCLUDED
#define VDBRT_H_INCLUDED

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDBRT_VERSION 100

/* ---- status codes ---------------------------------------------------------------------------------------- */
enum {
    VDBRT_OK = 0,
    VDBRT_ERR_INVALID_ARG   = 1,  /* null pointer, zero-sized film, bad enum ...                               */
    VDBRT_ERR_BAD_GRID      = 2,  /* not a NanoVDB buffer: magic / version / size (NanoVDB.h:2002-2013)          */
    VDBRT_ERR_NOT_FLOAT     = 3,  /* grid value type is not GridType::Float                                     */
    VDBRT_ERR_NOT_LEVELSET  = 4,  /* RuntimeError "only supports level sets"   (RayIntersector.h:105-109)       */
    VDBRT_ERR_NONUNIFORM    = 5,  /* RuntimeError "only supports uniform voxels" (RayIntersector.h:101-104,305) */
    VDBRT_ERR_EMPTY_GRID    = 6,  /* RuntimeError "does not supports empty grids" (RayIntersector.h:309,533)    */
    VDBRT_ERR_ISO_RANGE     = 7,  /* ValueError "iso-value must be inside the narrow-band" (RayIntersector.h:536)*/
    VDBRT_ERR_SPP_ZERO      = 8,  /* ValueError "pixelSamples must be larger than zero" (RayTracer.h:877-879)   */
    VDBRT_ERR_CUDA          = 9,  /* CUDA runtime error or no usable device                                     */
    VDBRT_ERR_UNSUPPORTED   = 10, /* e.g. a map that is not scale(+translate)                                   */
    VDBRT_ERR_NOMEM         = 11,
    VDBRT_ERR_IO            = 12  /* file cannot be opened / read / written, or has no such grid                     */
};

enum { VDBRT_MEM_HOST = 0, VDBRT_MEM_DEVICE = 1 };
enum { VDBRT_CAMERA_PERSPECTIVE = 0, VDBRT_CAMERA_ORTHOGRAPHIC = 1 };
enum { VDBRT_SHADER_MATTE = 0, VDBRT_SHADER_NORMAL = 1, VDBRT_SHADER_POSITION = 2, VDBRT_SHADER_DIFFUSE = 3 };
enum { VDBRT_SPACE_WORLD = 0, VDBRT_SPACE_INDEX = 1 };
enum { VDBRT_GRID_CLASS_UNKNOWN = 0, VDBRT_GRID_CLASS_LEVEL_SET = 1, VDBRT_GRID_CLASS_FOG_VOLUME = 2 };

typedef struct vdbrt_ctx  vdbrt_ctx;
typedef struct vdbrt_grid vdbrt_grid;

/* ---- POD descriptions ------------------------------------------------------------------------------------ */

/* math::Ray<double> (math/Ray.h:26-295): eye, direction, [t0,t1].  invDir is recomputed as 1/dir (Ray.h:67-71). */
typedef struct vdbrt_ray {
    double eye[3];
    double dir[3];
    double t0, t1;
} vdbrt_ray;

/* Flattened tools::BaseCamera (tools/RayTracer.h:351-415): the screen->world AffineMap as the 16 doubles of its
 * Mat4d (row-vector convention, element [r*4+c]), the base ray (RayTracer.h:404-409), and the two screen scales
 * (RayTracer.h:357-358).  Fill it with vdbrt_camera_perspective / _orthographic / _look_at.                    */
typedef struct vdbrt_camera {
    uint32_t kind;            /* VDBRT_CAMERA_*                                                                 */
    uint32_t width, height;   /* film size the camera was built for                                             */
    uint32_t reserved;
    double   m[16];           /* mScreenToWorld                                                                 */
    double   eye[3];          /* base ray eye  = applyMap(0,0,0)                                                */
    double   dir[3];          /* base ray dir  = applyJacobian(0,0,-1)                                          */
    double   scale_w, scale_h;/* mScaleWidth, mScaleHeight                                                      */
    double   t0, t1;          /* near / far                                                                     */
} vdbrt_camera;

/* The four constant-colour shaders (tools/RayTracer.h:565-581,614-630,671-690,728-753). `rgba` is the colour
 * the user passed to the shader's constructor (NormalShader halves it itself, RayTracer.h:618).                */
typedef struct vdbrt_shader {
    uint32_t kind;            /* VDBRT_SHADER_*                                                                 */
    float    rgba[4];
    uint32_t reserved;
    double   bbox_min[3];     /* PositionShader: bbox.min()            (RayTracer.h:675)                        */
    double   inv_dim[3];      /* PositionShader: 1.0 / bbox.extents()                                           */
    /* NULL: the constant-colour shaders above.  Else a grid from vdbrt_upload_color_grid: the GridT = Vec3SGrid forms of the
     * four shaders (RayTracer.h:542-562, 591-611, 640-668, 702-725) with the default PointSampler -- the colour is the
     * grid's value at the voxel nearest to the hit position (tools/Interpolation.h:600-617), rgba[] is not used.   */
    const struct vdbrt_grid* color_grid;
} vdbrt_shader;

/* Which film tiles this context renders (multi-GPU): the film is cut into tile_w x tile_h tiles numbered
 * row-major; the context renders tiles with  tile_id % count == rank  and leaves all other pixels untouched.
 * {0,0,0,1} or a zeroed struct means "everything".                                                             */
typedef struct vdbrt_partition {
    uint32_t tile_w, tile_h;  /* 0 -> library default (8x4 warp tiles grouped 64x64)                            */
    uint32_t rank, count;
} vdbrt_partition;

/* LevelSetRayTracer parameters (tools/RayTracer.h:100-126,872-889).                                             */
typedef struct vdbrt_ls_opts {
    float    iso;             /* LevelSetRayIntersector isoValue                                                */
    uint32_t spp;             /* pixelSamples (>= 1)                                                            */
    double   jitter[16];      /* mRand[16] from math::Rand01<double>(seed); see vdbrt_jitter_table              */
    vdbrt_partition part;
    uint32_t flags;           /* VDBRT_LS_*                                                                     */
    uint32_t iterations;      /* LinearSearchImpl<GridT, Iterations>: secant refinements of the hit time after the
                               * zero crossing (tools/RayIntersector.h:630-636).  0 = what tools::rayTrace and
                               * vdb_render use (LevelSetRayIntersector's default search)                         */
} vdbrt_ls_opts;
#define VDBRT_LS_UNIFORM_BG 1u /* every film pixel currently equals bg_rgba: skip the host->device film copy    */
#define VDBRT_ASYNC         2u /* device-memory film only: enqueue on the context's stream and return at once   */
/* Long-ray rounds (csrc/vdbrt_kernels.cuh): rays still running when their 8x4 tile has used up its iteration budget
 * are finished by scout / march kernels that spread the leaf visits of one ray over many threads.  Same pixels either
 * way.  Default: on for a small share of a partitioned frame (part.count > 1 and few tiles per resident warp, where
 * the slowest tile bounds the frame time), off otherwise; one sample per pixel only.                            */
#define VDBRT_LS_ROUNDS_ON  4u
#define VDBRT_LS_ROUNDS_OFF 8u
/* Heavy tiles first (csrc/vdbrt_kernels.cuh, k_probe_levelset): one budgeted probe ray per 8x4 tile estimates its cost and
 * the work queue hands out the expensive strips of tiles before the rest, so that no silhouette tile (every ray grazing
 * the surface) starts when the queue is nearly empty.  Same pixels either way.  Default: on when a resident warp gets
 * two tiles or more.                                                                                            */
#define VDBRT_LS_ORDER_ON   16u
#define VDBRT_LS_ORDER_OFF  32u

/* VolumeRender parameters (tools/RayTracer.h:162-207; defaults :929-936).                                      */
typedef struct vdbrt_vol_opts {
    double primary_step, shadow_step, cutoff, light_gain;
    double light_dir[3];      /* already normalised (setLightDir normalises, RayTracer.h:173)                   */
    double light_color[3];
    double absorption[3];
    double scattering[3];
    vdbrt_partition part;
    uint32_t flags;           /* VDBRT_ASYNC                                                                    */
    /* EXTENSION (BASELINE config 5; the reference's VolumeRender takes one sample per pixel): samples per pixel, 0 or 1 =
     * the reference's behaviour.  Sample 0 goes through the pixel centre, the others through the jittered offsets of
     * LevelSetRayTracer::operator() (RayTracer.h:903-913, index n(i,j) = 2*(spp-1)*(j*W+i)); the pixel is the sum of the
     * samples' RGBA (a sample that misses the volume's bbox is (0,0,0,0)) times float(1/spp), accumulated in float in
     * sample order.                                                                                             */
    uint32_t spp;
    double   jitter[16];
} vdbrt_vol_opts;

/* tools::Film (tools/RayTracer.h:226-345): row-major RGBA float4, pixel (w,h) at [w + h*width].               */
typedef struct vdbrt_film {
    float*   pixels;          /* 4*width*height floats, in/out for level sets (misses keep the old pixel)       */
    uint32_t width, height;
    uint32_t memspace;        /* VDBRT_MEM_*                                                                    */
    float    bg_rgba[4];      /* only read when VDBRT_LS_UNIFORM_BG is set                                      */
} vdbrt_film;

/* Optional per-pixel records of the PRIMARY ray, for parity checks; any pointer may be NULL.  Same memspace as
 * the film.  `ijk` is the voxel handed to the LinearSearchImpl call that returned true.             */
typedef struct vdbrt_aux {
    uint8_t* hit;             /* [W*H]    1 = intersectsWS returned true                                        */
    int32_t* ijk;             /* [W*H*3]                                                                        */
    double*  t_index;         /* [W*H]    LinearSearchImpl::getIndexTime                                        */
    double*  t_world;         /* [W*H]    getWorldTime                                                          */
    double*  xyz;             /* [W*H*3]  world position                                                        */
    double*  nml;             /* [W*H*3]  world normal                                                          */
} vdbrt_aux;

/* One result of LevelSetRayIntersector::intersectsWS/IS (tools/RayIntersector.h:119-240).                      */
typedef struct vdbrt_hit {
    int32_t hit;
    int32_t ijk[3];
    double  t_index, t_world;
    double  xyz_index[3];
    double  xyz_world[3];
    double  nml[3];
} vdbrt_hit;

/* Per-launch work counters (averages feed the roofline's algorithmic bytes).                        */
typedef struct vdbrt_counters {
    uint64_t rays;
    uint64_t root_probes;     /* n_R: hasNode<Upper> / root-level DDA probes                                    */
    uint64_t upper_probes;    /* n_U: probes inside upper nodes                                                 */
    uint64_t lower_probes;    /* n_L: probes inside lower nodes                                                 */
    uint64_t voxel_probes;    /* n_V: LinearSearchImpl::operator() probeValue calls                             */
    uint64_t stencil_refills; /* n_S: BoxStencil refills (8 fetches each)                                       */
    uint64_t primary_samples; /* n_P: fog primary trilinear samples                                             */
    uint64_t shadow_samples;  /* n_Sh: fog shadow trilinear samples                                             */
    uint64_t shadow_rays;
    uint64_t hits;
} vdbrt_counters;

typedef struct vdbrt_grid_info {
    uint64_t bytes;
    uint64_t active_voxels;
    uint32_t leaf_count, lower_count, upper_count, root_tiles;
    int32_t  index_bbox[6];   /* NanoVDB voxel-tight bbox (min xyz, max xyz)                                    */
    int32_t  node_bbox[6];    /* leaf/tile-granular bbox == RootNode::evalActiveBoundingBox(bbox,false)          */
    double   voxel_size[3];
    double   translation[3];
    float    background;
    uint32_t grid_class;      /* VDBRT_GRID_CLASS_*                                                             */
    uint32_t source_type;     /* nanovdb::GridType of the buffer that was uploaded: 1 Float, 13 Fp4, 14 Fp8, 15 Fp16,
                               * 16 FpN, 6 Vec3f (colour grids)                                                  */
    uint32_t leaf_kind;       /* how the kernels read the leaves: 0 float (Float sources, and Fp4 / FpN -- or Fp8 / Fp16
                               * with the quant_native knob off -- expanded at upload), 1 Fp8 codes, 2 Fp16 codes     */
    uint64_t resident_bytes;  /* device memory the grid occupies: the buffer, its halo blocks, its lower-node masks  */
} vdbrt_grid_info;

/* ---- context / memory ------------------------------------------------------------------------------------ */
int  vdbrt_create(int device, vdbrt_ctx** out);          /* replaces: nothing (reference is in-process TBB)     */
void vdbrt_destroy(vdbrt_ctx* ctx);
const char* vdbrt_last_error(void);
int  vdbrt_device_count(void);
/* use an existing CUDA stream (e.g. torch's current stream) for every launch/copy of this context; 0 = own stream */
int  vdbrt_set_stream(vdbrt_ctx* ctx, void* cuda_stream);
int  vdbrt_synchronize(vdbrt_ctx* ctx);
int  vdbrt_host_alloc(size_t bytes, void** out);         /* pinned host memory (cuda::DeviceBuffer semantics,   */
int  vdbrt_host_free(void* p);
/* page-lock memory the caller already owns (cudaHostRegister, portable + mapped), e.g. a POSIX shared-memory film that
 * the processes of all GPUs of a node have mapped: each rank's kernels then store the pixels they own straight into
 * that one HOST film over their own PCIe link -- the frame is assembled in host memory without any gather.      */
int  vdbrt_host_register(void* p, size_t bytes);
int  vdbrt_host_unregister(void* p);                           /*  nanovdb/cuda/DeviceBuffer.h:316-344)               */

/* Device buffers that can be shared between the processes of one node (one process per GPU): rank 0 allocates its film with
 * vdbrt_device_alloc, exports it, every other rank imports the handle and renders the tiles it owns STRAIGHT INTO rank 0's
 * film over NVLink (peer stores from the render kernel) -- the frame gather of SURVEY 8e without a separate collective.
 * Thin wrappers of cudaMalloc / cudaIpcGetMemHandle / cudaIpcOpenMemHandle(cudaIpcMemLazyEnablePeerAccess).                  */
#define VDBRT_IPC_HANDLE_BYTES 64
int  vdbrt_device_alloc(vdbrt_ctx* ctx, size_t bytes, void** out);
int  vdbrt_device_free(vdbrt_ctx* ctx, void* p);
int  vdbrt_ipc_export(vdbrt_ctx* ctx, const void* device_ptr, unsigned char handle[VDBRT_IPC_HANDLE_BYTES]);
int  vdbrt_ipc_import(vdbrt_ctx* ctx, const unsigned char handle[VDBRT_IPC_HANDLE_BYTES], void** device_ptr);
int  vdbrt_ipc_close(vdbrt_ctx* ctx, void* device_ptr);
/* cudaMemcpyAsync on the context's stream; kind: 0 host->device, 1 device->host, 2 device->device                            */
int  vdbrt_memcpy(vdbrt_ctx* ctx, void* dst, const void* src, size_t bytes, int kind);

/* ---- grids ------------------------------------------------------------------------------------------------
 * replaces nanovdb::GridHandle<cuda::DeviceBuffer>::deviceUpload (nanovdb/cuda/DeviceBuffer.h:411-456) plus the
 * constructor-time work of LinearSearchImpl / VolumeRayIntersector (validation, node-granular bbox:
 * tools/RayIntersector.h:299-319,527-541).  The buffer is a complete NanoGrid<float> (GridData first), or a quantised
 * NanoGrid<Fp4|Fp8|Fp16|FpN> (nanovdb/NanoVDB.h:3752-3980): its leaves are expanded on the device, once, with
 * LeafData<FpX>::getValue's arithmetic (float(code) * mQuantum + mMinimum) -- what nanovdb::tools::nanoToOpenVDB does before
 * the reference can ray-trace such a grid (nanovdb/tools/NanoToOpenVDB.h:511-518) -- and the grid then is a NanoGrid<float>
 * in every respect (vdbrt_grid_info::bytes and vdbrt_grid_download give the expanded buffer).                      */
int  vdbrt_upload_grid(vdbrt_ctx* ctx, const void* nanovdb_buffer, uint64_t bytes, uint32_t memspace, vdbrt_grid** out);
/* A NanoGrid<Vec3f> (GridType::Vec3f; createNanoGrid of an openvdb::Vec3SGrid) used as vdbrt_shader::color_grid; replaces
 * the `const GridT& grid` argument of the colour-grid shader constructors.  Scale(+translate) maps only.  Released with
 * vdbrt_free_grid; cannot be rendered itself.                                                                   */
int  vdbrt_upload_color_grid(vdbrt_ctx* ctx, const void* nanovdb_buffer, uint64_t bytes, uint32_t memspace, vdbrt_grid** out);
int  vdbrt_free_grid(vdbrt_ctx* ctx, vdbrt_grid* grid);
int  vdbrt_grid_get_info(const vdbrt_grid* grid, vdbrt_grid_info* info);
/* copy the serialised grid back (device -> host), e.g. after vdbrt_build_* */
int  vdbrt_grid_download(vdbrt_ctx* ctx, const vdbrt_grid* grid, void* dst, uint64_t bytes);

/* ---- host-side helpers that flatten the reference's classes ------------------------------------------------ */
/* tools::PerspectiveCamera ctor (RayTracer.h:436-445) / OrthographicCamera ctor (:494-502): rotation in degrees
 * (x,y,z order), then translation.                                                                             */
int  vdbrt_camera_perspective(vdbrt_camera* cam, uint32_t width, uint32_t height, const double rotation[3],
                              const double translation[3], double focal_length, double aperture,
                              double near_plane, double far_plane);
int  vdbrt_camera_orthographic(vdbrt_camera* cam, uint32_t width, uint32_t height, const double rotation[3],
                               const double translation[3], double frame_width, double near_plane, double far_plane);
/* BaseCamera::lookAt (RayTracer.h:379-389); like the reference it silently keeps the camera on failure.         */
int  vdbrt_camera_look_at(vdbrt_camera* cam, const double xyz[3], const double up[3]);
/* PerspectiveCamera::getRay (RayTracer.h:452-462) / OrthographicCamera::getRay (:505-512) for n pixels, on the host: `pixels` holds n
 * (i, j) pairs, `offsets` n (iOffset, jOffset) pairs or NULL for the pixel centres.  World-space rays, as the render kernels build them. */
int  vdbrt_camera_get_rays(const vdbrt_camera* cam, const uint32_t* pixels, const double* offsets, uint64_t n, vdbrt_ray* rays);
/* the 16 doubles LevelSetRayTracer::setPixelSamples draws (RayTracer.h:883-885)                                 */
int  vdbrt_jitter_table(unsigned int seed, double out[16]);
/* VolumeRender defaults (RayTracer.h:929-936)                                                                   */
int  vdbrt_vol_opts_default(vdbrt_vol_opts* opts);

/* ---- the hot path ------------------------------------------------------------------------------------------ */
/* tools::rayTrace(grid, LevelSetRayIntersector(grid, iso), shader, camera, spp, seed, threaded)
 * == LevelSetRayTracer::render (RayTracer.h:48-64,891-918).  Synchronous: the film is complete on return.     */
int  vdbrt_render_levelset(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam,
                           const vdbrt_shader* shader, const vdbrt_ls_opts* opts, vdbrt_film* film, vdbrt_aux* aux);
/* VolumeRender<VolumeRayIntersector<FloatGrid>, BoxSampler>::render (RayTracer.h:985-1070).                     */
/* Film::RGBA::over (tools/RayTracer.h:252-259) per pixel: top = top.over(bottom), i.e. s = bottom.a*(1-top.a);
 * rgb = top.a*top.rgb + s*bottom.rgb; a = top.a + s.  Both films in the same memory space and of the same size.
 * (BASELINE config 5: the fog film over the level-set film.)                                                    */
int  vdbrt_film_over(vdbrt_ctx* ctx, vdbrt_film* top, const vdbrt_film* bottom);
int  vdbrt_render_volume(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam,
                         const vdbrt_vol_opts* opts, vdbrt_film* film);
/* LevelSetRayIntersector::intersectsWS / intersectsIS on a batch of arbitrary rays (RayIntersector.h:119-240).  */
int  vdbrt_intersect_levelset(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_ray* rays, uint64_t n,
                              uint32_t space, float iso, vdbrt_hit* hits, uint32_t memspace);
/* the same with LinearSearchImpl<GridT, Iterations> as the search (LevelSetRayIntersector<GridT, LinearSearchImpl<GridT, N>>,
 * tools/RayIntersector.h:79-82,630-636): `iterations` secant refinements of every hit time                        */
int  vdbrt_intersect_levelset_ex(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_ray* rays, uint64_t n,
                                 uint32_t space, float iso, uint32_t iterations, vdbrt_hit* hits, uint32_t memspace);
/* VolumeRayIntersector::setIndexRay/setWorldRay + hits() (RayIntersector.h:368-432): spans[i*max_spans*2 ..],
 * counts[i] = number of spans (may exceed max_spans: only the first max_spans are stored); counts[i] = -1 when the
 * ray misses the bbox.                                                                                          */
int  vdbrt_volume_spans(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_ray* rays, uint64_t n, uint32_t space,
                        uint32_t max_spans, double* spans, int32_t* counts, uint32_t memspace);
/* VolumeRayIntersector::setIndexRay / setWorldRay (tools/RayIntersector.h:368-390) on the host: the ray mapped to index space
 * (space == VDBRT_SPACE_WORLD; math/Ray.h:150-159) and clipped against the node-granular bbox with its max padded by one
 * (:318, math/Ray.h:233-267).  *hit = 0 when it misses; `scale` receives the grid's index->world scale.  The same arithmetic,
 * in the same order, as the device kernels use: the clipped ray can be handed to vdbrt_volume_spans with VDBRT_SPACE_INDEX.  */
int  vdbrt_volume_clip(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_ray* ray, uint32_t space, vdbrt_ray* clipped,
                       int* hit, double scale[3]);

/* ---- measurement ------------------------------------------------------------------------------------------- */
/* counters of the most recent render with counting enabled (a separate instrumented launch, never timed)       */
int  vdbrt_count_levelset(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam,
                          const vdbrt_ls_opts* opts, vdbrt_counters* out);
int  vdbrt_count_volume(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam,
                        const vdbrt_vol_opts* opts, vdbrt_counters* out);
/* device time (ms, CUDA events on the context's stream) of the kernel(s) of the last render call, and how many
 * kernels of this library that call launched                                                                   */
int  vdbrt_last_kernel_ms(vdbrt_ctx* ctx, float* ms, uint32_t* launches);
/* Scheduling knobs of one context (the environment variables VDBRT_LS_* / VDBRT_FOG_* set the same values when the
 * context is created; this call is for tests and for measuring one setting against another in one process).  Keys:
 * "ls_strip", "ls_strip_ratio", "ls_refill", "ls_eager", "ls_order", "ls_history", "ls_hist_a", "ls_hist_b", "ls_probe_cap", "ls_probe_b", "ls_tail", "ls_budget", "ls_factor", "ls_rounds", "ls_dense", "ls_dense_factor", "ls_leaves0" .. "ls_leaves7", "fog_wave", "fog_refill", "fog_rec_per_ray", "fog_cap_mb".
 * None of them changes a pixel.  Unknown key -> VDBRT_ERR_INVALID_ARG.                                           */
int  vdbrt_set_tuning(vdbrt_ctx* ctx, const char* key, uint32_t value);

/* ---- grid construction on the GPU (inputs for benches) ------------------------------------
 * Same voxel values/topology as nanovdb::tools::createLevelSetSphere/Torus (nanovdb/tools/CreatePrimitives.h:
 * 598-664) and openvdb::tools::sdfToFogVolume (tools/LevelSetUtil.h:2190), written straight into NanoVDB layout. */
int  vdbrt_build_levelset_sphere(vdbrt_ctx* ctx, double radius, const double center[3], double voxel_size,
                                 double half_width, vdbrt_grid** out);
int  vdbrt_build_levelset_torus(vdbrt_ctx* ctx, double major_radius, double minor_radius, const double center[3],
                                double voxel_size, double half_width, vdbrt_grid** out);
/* union (voxel-wise min, tools/Composite.h:886) of n spheres: spheres[i] = {cx,cy,cz,r} in world units          */
int  vdbrt_build_levelset_spheres(vdbrt_ctx* ctx, const double* spheres, uint32_t n, double voxel_size,
                                  double half_width, vdbrt_grid** out);
/* the sphere set of BASELINE configs 4/5: std::mt19937_64 rng(seed); per sphere, in this order, cx,cy,cz ~
 * U(-extent,extent) and r ~ U(rmin,rmax) drawn with std::uniform_real_distribution<double>; out = n x {cx,cy,cz,r}           */
int  vdbrt_random_spheres(uint64_t seed, uint32_t n, double extent, double rmin, double rmax, double* out);
/* sdfToFogVolume of an existing level-set grid (cutoff = background)                                            */
int  vdbrt_build_fog_from_levelset(vdbrt_ctx* ctx, const vdbrt_grid* levelset, vdbrt_grid** out);

/* ---- ingestion: NanoVDB files (replaces nanovdb::io::readGrid / readGridMetaData / writeGrid, nanovdb/io/IO.h,
 *      for the grid this path renders; segment layout NanoVDB.h:5860-5929).  Codecs NONE and ZIP; host-only code. ---- */
enum { VDBRT_CODEC_NONE = 0, VDBRT_CODEC_ZIP = 1 };
typedef struct vdbrt_nvdb_meta {   /* io::FileGridMetaData (NanoVDB.h:5913-5929)                                   */
    char     name[256];
    uint64_t grid_bytes, file_bytes, active_voxels;
    uint32_t grid_type;       /* nanovdb::GridType, 1 = Float                                                    */
    uint32_t grid_class;      /* VDBRT_GRID_CLASS_*                                                              */
    uint32_t codec, pad;
    int32_t  index_bbox[6];
    double   world_bbox[6];
    double   voxel_size[3];
} vdbrt_nvdb_meta;
/* all grids of all segments of the file (or the one grid of a raw grid buffer); *count = number found           */
int  vdbrt_nvdb_list(const char* path, vdbrt_nvdb_meta* out, uint32_t capacity, uint32_t* count);
/* grid_name NULL or "": the first float grid (vdb_render's rule, openvdb_cmd/vdb_render/main.cc:771-786; the quantised
 * float types Fp4/Fp8/Fp16/FpN count as float: vdbrt_upload_grid takes them).  The buffer
 * is a 32-byte aligned host allocation owned by the caller: pass it to vdbrt_upload_grid, release with vdbrt_buffer_free */
int  vdbrt_nvdb_read(const char* path, const char* grid_name, void** buffer, uint64_t* bytes);
/* the same for another value type: grid_type = nanovdb::GridType (1 Float, 6 Vec3f: the colour grid of vdb_render's
 * -color option, main.cc:788-795), 0 = any                                                                      */
int  vdbrt_nvdb_read_typed(const char* path, const char* grid_name, uint32_t grid_type, void** buffer, uint64_t* bytes);
int  vdbrt_nvdb_write(const char* path, const void* buffer, uint64_t bytes, uint32_t codec);
int  vdbrt_buffer_free(void* buffer);
/* tools::Film::savePPM (tools/RayTracer.h:300-335): P6, channel = (unsigned char)(255.0f * value), ".ppm" appended
 * when the name has no extension                                                                                */
int  vdbrt_film_save_ppm(const char* file_name, const float* rgba, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif
#endif /* VDBRT_H_INCLUDED */
