// vdbrt.cu -- host side of libvdbrt.so: the C ABI declared in include/vdbrt.h.
//
// Everything that computes pixels runs in the CUDA kernels of vdbrt_kernels.cuh; this file validates inputs the way
// the reference's constructors do, flattens cameras/shaders into PODs, moves buffers and launches.  There is no CPU
// implementation of the hot path in this library: without a CUDA device vdbrt_create() fails.
#include "../../include/vdbrt.h"
#include "vdbrt_kernels.cuh"
#include "vdbrt_fog.cuh"
#include "vdbrt_host.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <vector>

using namespace vdbrt;

namespace vdbrt {
thread_local std::string g_error;
int setError(int code, const std::string& msg) { g_error = msg; return code; }
int cudaFail(cudaError_t e, const char* what) { return setError(VDBRT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
} // namespace vdbrt

#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cudaFail(e_, #expr); } while (0)

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ensureBuffer(void** p, size_t* cap, size_t bytes)
{
    if (*cap >= bytes) return VDBRT_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CUDA_TRY(cudaMalloc(p, bytes));
    *cap = bytes;
    return VDBRT_OK;
}

constexpr uint64_t MAGIC_NUMB = 0x304244566f6e614eULL, MAGIC_GRID = 0x314244566f6e614eULL; // nanovdb/NanoVDB.h:139-140
constexpr size_t GRID_SIZE = 672, TREE_SIZE = 64;
constexpr size_t OFF_VERSION = 16, OFF_GRIDSIZE = 32, OFF_MATD = 384, OFF_VECD = 528, OFF_CLASS = 632, OFF_TYPE = 636;

template<typename T> T rd(const uint8_t* p) { T v; std::memcpy(&v, p, sizeof(T)); return v; }

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// grid registration: header parsing (host) + node-granular bbox (device)
// ---------------------------------------------------------------------------------------------------------------
void vdbrt::destroyGrid(vdbrt_grid* grid)
{
    if (!grid) return;
    cudaFree(grid->dev);
    cudaFree(grid->halo);
    cudaFree(grid->lowmask);
    delete grid;
}

int vdbrt::finishGrid(vdbrt_ctx* ctx, vdbrt_grid* grid)
{
    // header: GridData + TreeData, then RootData
    uint8_t head[GRID_SIZE + TREE_SIZE];
    if (grid->bytes < sizeof(head)) return setError(VDBRT_ERR_BAD_GRID, "buffer smaller than GridData+TreeData");
    CUDA_TRY(cudaMemcpyAsync(head, grid->dev, sizeof(head), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const uint64_t magic = rd<uint64_t>(head);
    if (magic != MAGIC_NUMB && magic != MAGIC_GRID) return setError(VDBRT_ERR_BAD_GRID, "not a NanoVDB grid (bad magic number)");
    if ((rd<uint32_t>(head + OFF_VERSION) >> 21) != 32) return setError(VDBRT_ERR_BAD_GRID, "incompatible NanoVDB major version (need 32)");
    if (rd<uint64_t>(head + OFF_GRIDSIZE) > grid->bytes) return setError(VDBRT_ERR_BAD_GRID, "grid size exceeds the buffer");
    const uint32_t gridType = rd<uint32_t>(head + OFF_TYPE);
    const uint32_t wantType = grid->leaf_kind == kLeafFp8 ? 14u : (grid->leaf_kind == kLeafFp16 ? 15u : 1u);      // nanovdb::GridType Float / Fp8 / Fp16
    if (gridType != wantType) return setError(VDBRT_ERR_NOT_FLOAT, "grid value type is not float");
    const uint64_t leafBytes = grid->leaf_kind == kLeafFp8 ? LeafKind<kLeafFp8>::bytes : (grid->leaf_kind == kLeafFp16 ? LeafKind<kLeafFp16>::bytes : LeafKind<kLeafFloat>::bytes);
    const uint8_t* tree = head + GRID_SIZE;
    const uint64_t rootOff = GRID_SIZE + uint64_t(rd<int64_t>(tree + 24));
    if (rootOff + 64 > grid->bytes) return setError(VDBRT_ERR_BAD_GRID, "root offset outside the buffer");
    uint8_t rootHead[64];
    CUDA_TRY(cudaMemcpyAsync(rootHead, grid->dev + rootOff, 64, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));

    vdbrt_grid_info& info = grid->info;
    std::memset(&info, 0, sizeof(info));
    info.bytes = grid->bytes;
    info.leaf_count = rd<uint32_t>(tree + 32); info.lower_count = rd<uint32_t>(tree + 36); info.upper_count = rd<uint32_t>(tree + 40);
    info.active_voxels = rd<uint64_t>(tree + 56);
    info.root_tiles = rd<uint32_t>(rootHead + kRootTableSize);
    info.background = rd<float>(rootHead + kRootBackground);
    for (int i = 0; i < 6; ++i) info.index_bbox[i] = rd<int32_t>(rootHead + 4 * i);
    info.grid_class = rd<uint32_t>(head + OFF_CLASS);
    info.source_type = gridType;                            // (vdbrt_upload_grid overrides it for sources it expanded)
    info.leaf_kind = uint32_t(grid->leaf_kind);
    if (rootOff + 64 + uint64_t(info.root_tiles) * kTileSize > grid->bytes) return setError(VDBRT_ERR_BAD_GRID, "root table outside the buffer");

    double m[9];
    for (int i = 0; i < 9; ++i) m[i] = rd<double>(head + OFF_MATD + 8 * i);
    DevGrid& d = grid->dgrid;
    std::memset(&d, 0, sizeof(d));
    // off-diagonal terms: an AffineMap (rotation / shear).  The kernels then evaluate NanoVDB's stored matrix and inverse -- the tolerance
    // path of SURVEY 0.7 (the reference multiplies through its own 4x4s); scale(+translate) maps keep the bit-exact single multiplies.
    d.general = (m[1] != 0 || m[2] != 0 || m[3] != 0 || m[5] != 0 || m[6] != 0 || m[7] != 0) ? 1u : 0u;
    for (int i = 0; i < 9; ++i) { d.mat[i] = m[i]; d.imat[i] = rd<double>(head + OFF_MATD + 72 + 8 * i); }
    d.base = grid->dev; d.root_off = rootOff; d.tiles = grid->dev + rootOff + kRootTiles;
    d.table_size = info.root_tiles; d.background = info.background; d.grid_class = info.grid_class;
    for (int a = 0; a < 3; ++a) {
        d.scale[a] = m[4 * a];
        d.inv[a] = 1.0 / d.scale[a];                    // ScaleMap: mScaleValuesInverse = 1.0 / mScaleValues (math/Maps.h:674)
        d.trans[a] = rd<double>(head + OFF_VECD + 8 * a);
        // mVoxelSize = |scale| (math/Maps.h:667); AffineMap: length of the image of the unit vector (:630-633)
        info.voxel_size[a] = d.general ? std::sqrt(m[a] * m[a] + m[3 + a] * m[3 + a] + m[6 + a] * m[6 + a]) : std::fabs(d.scale[a]);
        info.translation[a] = d.trans[a];
    }
    d.has_translation = (d.trans[0] != 0 || d.trans[1] != 0 || d.trans[2] != 0) ? 1u : 0u;
    d.voxel_size0 = info.voxel_size[0];

    // node-granular bbox on the device; the same pass checks that every child offset lands on a node of the right level inside the
    // buffer (counted in init[6]) -- the derived structures below and the render kernels follow those links without looking again
    const uint64_t leafOff = GRID_SIZE + uint64_t(rd<int64_t>(tree + 0)), lowerOff = GRID_SIZE + uint64_t(rd<int64_t>(tree + 8));
    const uint64_t upperOff = GRID_SIZE + uint64_t(rd<int64_t>(tree + 16));
    if ((info.leaf_count && leafOff + uint64_t(info.leaf_count) * leafBytes > grid->bytes) || (info.lower_count && lowerOff + uint64_t(info.lower_count) * 33856ull > grid->bytes) ||
        (info.upper_count && upperOff + uint64_t(info.upper_count) * 270400ull > grid->bytes) || ((leafOff | lowerOff | upperOff | rootOff) & 31))
        return setError(VDBRT_ERR_BAD_GRID, "node arrays outside the buffer or misaligned");
    int init[7] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0};
    CUDA_TRY(cudaMemcpyAsync(ctx->scratch, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    if (info.root_tiles) {
        const NodeAreas ar = {upperOff, info.upper_count, lowerOff, info.lower_count, leafOff, info.leaf_count, leafBytes};
        const unsigned long long threads = (unsigned long long)info.root_tiles << 15;
        k_node_bbox<<<unsigned((threads + 255) / 256), 256, 0, ctx->stream>>>(grid->dev, rootOff, info.root_tiles, reinterpret_cast<int*>(ctx->scratch), ar);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(init, ctx->scratch, sizeof(init), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (init[6] != 0) return setError(VDBRT_ERR_BAD_GRID, "corrupt NanoVDB tree: " + std::to_string(init[6]) + " child offsets do not point at a node of their level inside the buffer");
    }
    // "leaf or active tile" masks of the lower nodes (DevGrid::lowmask), for the volume walk; VDBRT_LOWMASK=0 keeps them off
    static const bool useLowmask = [] { const char* e = std::getenv("VDBRT_LOWMASK"); return !(e && *e == '0'); }();
    cudaFree(grid->lowmask); grid->lowmask = nullptr;
    if (useLowmask && info.lower_count && info.root_tiles && !(lowerOff & 31) && lowerOff + uint64_t(info.lower_count) * 33856ull <= grid->bytes) {
        if (cudaMalloc(&grid->lowmask, 512 * size_t(info.lower_count)) != cudaSuccess) { cudaGetLastError(); grid->lowmask = nullptr; }
        else {
            const unsigned long long words = (unsigned long long)info.lower_count << 6;
            k_build_lowmask<<<unsigned((words + 255) / 256), 256, 0, ctx->stream>>>(grid->dev, lowerOff, info.lower_count, grid->lowmask);
            CUDA_TRY(cudaGetLastError());
            d.lowmask = grid->lowmask; d.lower0 = uint32_t(lowerOff >> 5); d.lower_count = info.lower_count;
        }
    }
    // halo blocks of the leaves (DevGrid::halo): an acceleration structure like the node bbox, built once here.
    // Needs 2944 B per leaf next to the grid; without the memory (or with VDBRT_HALO=0) the stencil walks the leaves instead.
    static const bool useHalo = [] { const char* e = std::getenv("VDBRT_HALO"); return !(e && *e == '0'); }();
    cudaFree(grid->halo); grid->halo = nullptr;
    const size_t blockBytes = grid->leaf_kind == kLeafFp8 ? LeafKind<kLeafFp8>::block : (grid->leaf_kind == kLeafFp16 ? LeafKind<kLeafFp16>::block : LeafKind<kLeafFloat>::block);
    if ((useHalo || grid->leaf_kind != kLeafFloat) && info.leaf_count && info.root_tiles && !(leafOff & 31) &&
        leafOff + uint64_t(info.leaf_count) * leafBytes <= grid->bytes) {
        if (cudaMalloc(&grid->halo, blockBytes * size_t(info.leaf_count)) != cudaSuccess) { cudaGetLastError(); grid->halo = nullptr; }
        else {
            d.leaf0 = uint32_t(leafOff >> 5); d.leaf_count = info.leaf_count;
            const unsigned cap = unsigned(ctx->sm_count > 0 ? ctx->sm_count : 148) * 32u;          // a multiple of the SM count, grid-stride loop
            const unsigned blocks = info.leaf_count < cap ? info.leaf_count : cap;
            if (grid->leaf_kind == kLeafFp8) k_build_halo_q<kLeafFp8><<<blocks, 256, 0, ctx->stream>>>(d, leafOff, reinterpret_cast<uint8_t*>(grid->halo));
            else if (grid->leaf_kind == kLeafFp16) k_build_halo_q<kLeafFp16><<<blocks, 256, 0, ctx->stream>>>(d, leafOff, reinterpret_cast<uint8_t*>(grid->halo));
            else k_build_halo<<<blocks, 256, 0, ctx->stream>>>(d, leafOff, grid->halo);
            CUDA_TRY(cudaGetLastError());
            d.halo = grid->halo;
        }
    }
    info.resident_bytes = grid->bytes + (grid->halo ? blockBytes * uint64_t(info.leaf_count) : 0) + (grid->lowmask ? 512ull * info.lower_count : 0);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 6; ++i) info.node_bbox[i] = init[i];
    for (int a = 0; a < 3; ++a) { d.bbox_min[a] = init[a]; d.bbox_max[a] = init[3 + a]; }
    return VDBRT_OK;
}

namespace {

// validation the reference performs when the intersectors are constructed
int checkLevelSet(const vdbrt_grid* g, float iso)
{
    const vdbrt_grid_info& i = g->info;
    if (g->is_color) return setError(VDBRT_ERR_NOT_FLOAT, "grid value type is not float");
    // the member LinearSearchImpl is constructed before the intersector's own checks run (tools/RayIntersector.h:98,533-539)
    if (i.root_tiles == 0) return setError(VDBRT_ERR_EMPTY_GRID, "LinearSearchImpl does not supports empty grids"); // :533-535
    if (iso <= -i.background || iso >= i.background)
        return setError(VDBRT_ERR_ISO_RANGE, "The iso-value must be inside the narrow-band!");                 // :536-539
    if (std::fabs(i.voxel_size[0] - i.voxel_size[1]) > 5e-7 || std::fabs(i.voxel_size[0] - i.voxel_size[2]) > 5e-7)
        return setError(VDBRT_ERR_NONUNIFORM, "LevelSetRayIntersector only supports uniform voxels!");          // :101-104
    if (i.grid_class != VDBRT_GRID_CLASS_LEVEL_SET)
        return setError(VDBRT_ERR_NOT_LEVELSET, "LevelSetRayIntersector only supports level sets!");           // :105-109
    return VDBRT_OK;
}
int checkVolume(const vdbrt_grid* g)
{
    const vdbrt_grid_info& i = g->info;
    if (g->is_color) return setError(VDBRT_ERR_NOT_FLOAT, "grid value type is not float");
    if (std::fabs(i.voxel_size[0] - i.voxel_size[1]) > 5e-7 || std::fabs(i.voxel_size[0] - i.voxel_size[2]) > 5e-7)
        return setError(VDBRT_ERR_NONUNIFORM, "VolumeRayIntersector only supports uniform voxels!");           // :305-308
    if (i.root_tiles == 0) return setError(VDBRT_ERR_EMPTY_GRID, "LinearSearchImpl does not supports empty grids"); // :309-311
    return VDBRT_OK;
}

DevCamera toDev(const vdbrt_camera& c)
{
    DevCamera d;
    d.kind = c.kind; d.width = c.width; d.height = c.height; d.pad = 0;
    std::memcpy(d.m, c.m, sizeof(d.m));
    for (int a = 0; a < 3; ++a) { d.eye[a] = c.eye[a]; d.dir[a] = c.dir[a]; }
    d.scale_w = c.scale_w; d.scale_h = c.scale_h; d.t0 = c.t0; d.t1 = c.t1;
    return d;
}

int persistentGrid(vdbrt_ctx* ctx, const void* kernel, unsigned items)
{
    int perSm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kBlockThreads, 0) != cudaSuccess || perSm < 1) perSm = 1;
    const unsigned full = unsigned(ctx->sm_count) * unsigned(perSm);           // a multiple of the SM count
    const unsigned need = (items + (kBlockThreads / 32) - 1) / (kBlockThreads / 32);
    return int(need < full ? (need ? need : 1u) : full);
}

void ls_params(const vdbrt_grid* grid, const vdbrt_ls_opts* o, const vdbrt_film* film, LsParams& p)
{
    p.iso = o->iso;
    p.vmin = o->iso - float(2 * grid->info.voxel_size[0]);      // LinearSearchImpl ctor (tools/RayIntersector.h:530-531)
    p.vmax = o->iso + float(2 * grid->info.voxel_size[0]);
    p.sub = o->spp - 1;
    p.iters = o->iterations; p.pad = 0;
    p.frac = 1.0f / (1.0f + float(p.sub));                       // tools/RayTracer.h:905
    p.uniform_bg = (o->flags & VDBRT_LS_UNIFORM_BG) ? 1u : 0u;
    for (int i = 0; i < 4; ++i) p.bg[i] = film ? film->bg_rgba[i] : 0.f;
    for (int i = 0; i < 16; ++i) p.jitter[i] = o->jitter[i];
    p.bg_film = nullptr;
}

// A pinned host film is written by the kernels directly (unified addressing: the stores travel over PCIe while the rest of
// the frame is still being traced), so no device->host copy of the film follows the render.  Returns the device alias of
// the host pointer, or null for pageable memory.
// 0: pageable (stage it), 1: the whole range is one page-locked, mapped allocation / registration (*alias = its device address),
// 2: only a part of it is page-locked -- CUDA can neither map nor copy such a range in one piece, the caller gets an error
int classifyHostFilm(const void* host, size_t bytes, float4** alias)
{
    // the kernels read and write the WHOLE film in place: the first and the last byte must belong to the same registration
    *alias = nullptr;
    cudaPointerAttributes a, z;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return 0; }
    const bool firstPinned = a.type == cudaMemoryTypeHost && a.devicePointer;
    if (bytes > 1) {
        const char* last = static_cast<const char*>(host) + (bytes - 1);
        if (cudaPointerGetAttributes(&z, last) != cudaSuccess) { cudaGetLastError(); return firstPinned ? 2 : 0; }
        const bool lastPinned = z.type == cudaMemoryTypeHost && z.devicePointer;
        if (firstPinned != lastPinned) return 2;
        if (firstPinned && static_cast<const char*>(z.devicePointer) - static_cast<const char*>(a.devicePointer) != ptrdiff_t(bytes - 1)) return 2;
    }
    if (!firstPinned) return 0;
    *alias = static_cast<float4*>(a.devicePointer);
    return 1;
}
const char* kPartlyPinned = "the host film is only partly page-locked: cudaHostRegister / vdbrt_host_register must cover the whole film";

void vol_params(const vdbrt_vol_opts* o, VolParams& p)
{
    p.pstep = o->primary_step; p.sstep = o->shadow_step; p.cutoff = o->cutoff; p.gain = o->light_gain;
    for (int a = 0; a < 3; ++a) {
        p.light[a] = o->light_dir[a];
        p.ext[a] = -o->scattering[a] - o->absorption[a];                                        // tools/RayTracer.h:996
        p.albedo[a] = o->light_color[a] * o->scattering[a] / (o->scattering[a] + o->absorption[a]); // :997
    }
    p.sub = o->spp > 1 ? o->spp - 1 : 0u;
    p.frac = 1.0f / (1.0f + float(p.sub));
    for (int i = 0; i < 16; ++i) p.jitter[i] = o->jitter[i];
}

} // namespace

extern "C" {

const char* vdbrt_last_error(void) { return g_error.c_str(); }

int vdbrt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int vdbrt_create(int device, vdbrt_ctx** out)
{
    if (!out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return setError(VDBRT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)"); }
    if (device < 0 || device >= n) return setError(VDBRT_ERR_INVALID_ARG, "device index out of range");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return setError(VDBRT_ERR_CUDA, "kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
    auto* ctx = new vdbrt_ctx;
    ctx->device = device; ctx->sm_count = prop.multiProcessorCount;
    // a context that cannot be completed is taken apart again (vdbrt_destroy copes with the members that were never created)
    auto fail = [&](cudaError_t e, const char* what) { const int rc = cudaFail(e, what); vdbrt_destroy(ctx); return rc; };
    cudaError_t ce;
    if ((ce = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(ce, "cudaStreamCreateWithFlags");
    ctx->stream = ctx->own_stream;
    if ((ce = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return fail(ce, "cudaEventCreate");
    if ((ce = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return fail(ce, "cudaEventCreate");
    if (std::getenv("VDBRT_TIME_PROBE") && (ce = cudaEventCreate(&ctx->evp)) != cudaSuccess) return fail(ce, "cudaEventCreate");
    if ((ce = cudaMalloc(&ctx->scratch, 4096)) != cudaSuccess) return fail(ce, "cudaMalloc(scratch)");
    if ((ce = cudaMemset(ctx->scratch, 0, 4096)) != cudaSuccess) return fail(ce, "cudaMemset(scratch)");
    // tuning knobs of the long-ray rounds (see vdbrt_kernels.cuh); the defaults were measured on the B200
    const char* ev = std::getenv("VDBRT_LS_BUDGET");
    ctx->ls_budget = ev ? uint32_t(std::strtoul(ev, nullptr, 10)) : kDefaultBudget;
    ev = std::getenv("VDBRT_LS_TAIL");
    ctx->ls_tail = ev ? uint32_t(std::strtoul(ev, nullptr, 10)) : kDefaultTail;
    ev = std::getenv("VDBRT_LS_VOXEL_ONLY");
    ctx->ls_voxel_only = ev ? uint32_t(std::strtoul(ev, nullptr, 10)) : 0u;
    ev = std::getenv("VDBRT_LS_FACTOR");
    ctx->ls_factor = ev ? uint32_t(std::strtoul(ev, nullptr, 10)) : kDefaultFactor;
    ev = std::getenv("VDBRT_LS_ROUNDS");
    ctx->ls_rounds = ev ? uint32_t(std::strtoul(ev, nullptr, 10)) : uint32_t(kDefaultRounds);
    if (ctx->ls_rounds > uint32_t(kMaxRounds)) ctx->ls_rounds = kMaxRounds;
    // leaf visits per ray and round.  Every round costs the latency of one scout walk and one leaf march whatever the number of rays.
    // With the tail rule (only the rays of the last tiles are suspended, and late) ONE round of 64 leaf visits measured best -- C2 at 1/8 of
    // the frame 0.56 -> 0.51 ms, at 1/4 0.80 -> 0.73 ms against the two rounds (8, then 128) that were best under round 1's per-tile budget;
    // what is still alive after it is walked in line by k_long_finish (profiles/r02_summary.md 8c).
    static const uint32_t kLeaves[kMaxRounds] = {64, 256, 256, 256, 256, 256, 256, 256};
    for (int r = 0; r < kMaxRounds; ++r) ctx->ls_leaves[r] = kLeaves[r];
    // feeding of the render warps (Sched, vdbrt_kernels.cuh); defaults measured on the B200 (profiles/r02_summary.md)
    auto envU = [](const char* name, uint32_t dflt) { const char* e = std::getenv(name); return e ? uint32_t(std::strtoul(e, nullptr, 10)) : dflt; };
    ctx->ls_strip = envU("VDBRT_LS_STRIP", 1);            // 8x4 tiles per strip (measured: anything above 1 costs L2 locality)
    ctx->ls_strip_ratio = envU("VDBRT_LS_STRIP_RATIO", 4);  // ... but at least this many strips per resident warp (0: no such rule)
    ctx->ls_refill = envU("VDBRT_LS_REFILL", 32);         // idle lanes that trigger a refill from the strip (32: whole tiles)
    ctx->ls_affine = envU("VDBRT_LS_AFFINE", 0);          // SM-affine queues: strips per chunk (0: one global queue)
    ctx->ls_eager = envU("VDBRT_LS_EAGER", 0);            // take the next strip while lanes of the old one are still running
    ctx->ls_order = envU("VDBRT_LS_ORDER", 0);            // heavy strips first: 0 off, 1 on, 2 when a warp gets >= 2 tiles
    ctx->ls_history = envU("VDBRT_LS_HISTORY", 1);        // heavy tiles first from the previous frame's tile costs
    ctx->ls_hist_a = envU("VDBRT_LS_HIST_A", 250);        // list A: tiles that took at least this many percent of the mean tile
    ctx->ls_hist_b = envU("VDBRT_LS_HIST_B", 105);        // list B
    ctx->ls_probe_cap = envU("VDBRT_LS_PROBE_CAP", 128);  // steps a probe ray may take; unfinished = list A
    ctx->ls_probe_b = envU("VDBRT_LS_PROBE_B", 64);       // steps from which a strip goes to list B
    ctx->quant_native = envU("VDBRT_QUANT_NATIVE", 1);     // NanoGrid<Fp8|Fp16> rendered as they are; 0: expanded to float leaves at upload
    ctx->fog_wave = envU("VDBRT_FOG_WAVE", 1);            // VolumeRender as a wavefront of three kernels (vdbrt_fog.cuh); 0: the one-loop kernel
    ctx->fog_refill = envU("VDBRT_FOG_REFILL", 8);        // shadow kernel: idle lanes that trigger a refill from the record queue
    ctx->fog_rec_per_ray = envU("VDBRT_FOG_REC_PER_RAY", 12);   // record budget per primary ray (average over a batch of tiles)
    ctx->fog_cap_mb = envU("VDBRT_FOG_CAP_MB", 4096);     // device memory for the records of one batch
    if ((ev = std::getenv("VDBRT_LS_LEAVES"))) {
        int r = 0;
        for (const char* q = ev; *q && r < kMaxRounds; ++r) { ctx->ls_leaves[r] = uint32_t(std::strtoul(q, const_cast<char**>(&q), 10)); if (*q == ',') ++q; }
        if (r > 0) { for (int k = r; k < kMaxRounds; ++k) ctx->ls_leaves[k] = ctx->ls_leaves[r - 1]; if (!std::getenv("VDBRT_LS_ROUNDS")) ctx->ls_rounds = uint32_t(r); }
    }
    *out = ctx;
    return VDBRT_OK;
}

void vdbrt_destroy(vdbrt_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->film) cudaFree(ctx->film);
    if (ctx->aux) cudaFree(ctx->aux);
    if (ctx->io) cudaFree(ctx->io);
    if (ctx->lng) cudaFree(ctx->lng);
    if (ctx->ord) cudaFree(ctx->ord);
    if (ctx->hist) cudaFree(ctx->hist);
    if (ctx->fog) cudaFree(ctx->fog);
    if (ctx->hist_host) cudaFreeHost(ctx->hist_host);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->evp) cudaEventDestroy(ctx->evp);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    cudaGetLastError();
    delete ctx;
}

int vdbrt_set_stream(vdbrt_ctx* ctx, void* stream)
{
    if (!ctx) return setError(VDBRT_ERR_INVALID_ARG, "null context");
    ctx->stream = stream ? static_cast<cudaStream_t>(stream) : ctx->own_stream;
    return VDBRT_OK;
}

int vdbrt_synchronize(vdbrt_ctx* ctx)
{
    if (!ctx) return setError(VDBRT_ERR_INVALID_ARG, "null context");
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VDBRT_OK;
}

int vdbrt_host_alloc(size_t bytes, void** out)
{
    if (!out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return VDBRT_OK;
}
int vdbrt_host_free(void* p) { CUDA_TRY(cudaFreeHost(p)); return VDBRT_OK; }
int vdbrt_host_register(void* p, size_t bytes)
{
    if (!p || !bytes) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return VDBRT_OK;
}
int vdbrt_host_unregister(void* p) { CUDA_TRY(cudaHostUnregister(p)); return VDBRT_OK; }

int vdbrt_device_alloc(vdbrt_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaMalloc(out, bytes));
    return VDBRT_OK;
}
int vdbrt_device_free(vdbrt_ctx* ctx, void* p)
{
    if (!ctx) return setError(VDBRT_ERR_INVALID_ARG, "null context");
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaFree(p));
    return VDBRT_OK;
}
int vdbrt_ipc_export(vdbrt_ctx* ctx, const void* devicePtr, unsigned char handle[VDBRT_IPC_HANDLE_BYTES])
{
    if (!ctx || !devicePtr || !handle) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == VDBRT_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(ctx->device);
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void*>(devicePtr)));
    std::memcpy(handle, &h, sizeof(h));
    return VDBRT_OK;
}
int vdbrt_ipc_import(vdbrt_ctx* ctx, const unsigned char handle[VDBRT_IPC_HANDLE_BYTES], void** devicePtr)
{
    if (!ctx || !devicePtr || !handle) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(ctx->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(devicePtr, h, cudaIpcMemLazyEnablePeerAccess));
    return VDBRT_OK;
}
int vdbrt_ipc_close(vdbrt_ctx* ctx, void* devicePtr)
{
    if (!ctx) return setError(VDBRT_ERR_INVALID_ARG, "null context");
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaIpcCloseMemHandle(devicePtr));
    return VDBRT_OK;
}

int vdbrt_memcpy(vdbrt_ctx* ctx, void* dst, const void* src, size_t bytes, int kind)
{
    if (!ctx || !dst || !src) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(ctx->device);
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
    return VDBRT_OK;
}

int vdbrt_upload_grid(vdbrt_ctx* ctx, const void* buffer, uint64_t bytes, uint32_t memspace, vdbrt_grid** out)
{
    if (!ctx || !buffer || !out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (bytes < GRID_SIZE + TREE_SIZE) return setError(VDBRT_ERR_BAD_GRID, "buffer smaller than GridData+TreeData");
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    const bool onDevice = memspace == VDBRT_MEM_DEVICE;
    uint8_t head[GRID_SIZE + TREE_SIZE];
    if (onDevice) {
        CUDA_TRY(cudaMemcpyAsync(head, buffer, sizeof(head), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    } else std::memcpy(head, buffer, sizeof(head));
    const uint64_t magic = rd<uint64_t>(head);
    const bool quantised = (magic == MAGIC_NUMB || magic == MAGIC_GRID) && (rd<uint32_t>(head + OFF_VERSION) >> 21) == 32 &&
                           isQuantisedType(rd<uint32_t>(head + OFF_TYPE));
    auto* g = new vdbrt_grid;
    g->device = ctx->device;
    const uint32_t srcType = rd<uint32_t>(head + OFF_TYPE);
    const bool native = quantised && ctx->quant_native && (srcType == 14u || srcType == 15u);
    if (native) g->leaf_kind = srcType == 14u ? kLeafFp8 : kLeafFp16;        // rendered as it is: the Fp8 / Fp16 instantiations of the kernels
    if (quantised && !native) {
        // NanoGrid<Fp4|FpN> (and Fp8 / Fp16 with quant_native = 0): the leaves are expanded to floats once, on the device (vdbrt_quant.cu)
        uint8_t* staged = nullptr;
        if (!onDevice) {
            cudaError_t e = cudaMalloc(&staged, bytes);
            if (e == cudaSuccess) e = cudaMemcpyAsync(staged, buffer, bytes, cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) { cudaFree(staged); delete g; return cudaFail(e, "staging the quantised grid"); }
        }
        const int rc = expandQuantised(ctx, onDevice ? static_cast<const uint8_t*>(buffer) : staged, bytes, head, &g->dev, &g->bytes);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(staged);
        if (rc != VDBRT_OK) { delete g; return rc; }
    } else {
        g->bytes = bytes;
        cudaError_t e = cudaMalloc(&g->dev, bytes);
        if (e != cudaSuccess) { delete g; return cudaFail(e, "cudaMalloc(grid)"); }
        e = cudaMemcpyAsync(g->dev, buffer, bytes, onDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { cudaFree(g->dev); delete g; return cudaFail(e, "cudaMemcpyAsync(grid)"); }
    }
    const int rc = finishGrid(ctx, g);
    if (rc != VDBRT_OK) { destroyGrid(g); return rc; }
    g->info.source_type = rd<uint32_t>(head + OFF_TYPE);
    *out = g;
    return VDBRT_OK;
}

int vdbrt_upload_color_grid(vdbrt_ctx* ctx, const void* buffer, uint64_t bytes, uint32_t memspace, vdbrt_grid** out)
{
    if (!ctx || !buffer || !out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (bytes < GRID_SIZE + TREE_SIZE) return setError(VDBRT_ERR_BAD_GRID, "buffer smaller than GridData+TreeData");
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    auto* g = new vdbrt_grid;
    g->bytes = bytes; g->device = ctx->device; g->is_color = true;
    cudaError_t e = cudaMalloc(&g->dev, bytes);
    if (e != cudaSuccess) { delete g; return cudaFail(e, "cudaMalloc(grid)"); }
    auto fail = [&](int rc) { cudaFree(g->dev); delete g; return rc; };
    e = cudaMemcpyAsync(g->dev, buffer, bytes, memspace == VDBRT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return fail(cudaFail(e, "cudaMemcpyAsync(grid)"));
    uint8_t head[GRID_SIZE + TREE_SIZE];
    if (cudaMemcpyAsync(head, g->dev, sizeof(head), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return fail(setError(VDBRT_ERR_CUDA, "reading the grid header back failed"));
    const uint64_t magic = rd<uint64_t>(head);
    if (magic != MAGIC_NUMB && magic != MAGIC_GRID) return fail(setError(VDBRT_ERR_BAD_GRID, "not a NanoVDB grid (bad magic number)"));
    if ((rd<uint32_t>(head + OFF_VERSION) >> 21) != 32) return fail(setError(VDBRT_ERR_BAD_GRID, "incompatible NanoVDB major version (need 32)"));
    if (rd<uint32_t>(head + OFF_TYPE) != 6) return fail(setError(VDBRT_ERR_NOT_FLOAT, "colour grid value type is not Vec3f"));
    const uint8_t* tree = head + GRID_SIZE;
    const uint64_t rootOff = GRID_SIZE + uint64_t(rd<int64_t>(tree + 24));
    if (rootOff + kColRootTiles > bytes) return fail(setError(VDBRT_ERR_BAD_GRID, "root offset outside the buffer"));
    uint8_t rootHead[kColRootTiles];
    if (cudaMemcpyAsync(rootHead, g->dev + rootOff, sizeof(rootHead), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return fail(setError(VDBRT_ERR_CUDA, "reading the root header back failed"));
    double m[9];
    for (int i = 0; i < 9; ++i) m[i] = rd<double>(head + OFF_MATD + 8 * i);
    if (m[1] != 0 || m[2] != 0 || m[3] != 0 || m[5] != 0 || m[6] != 0 || m[7] != 0)
        return fail(setError(VDBRT_ERR_UNSUPPORTED, "only scale(+translate) index->world maps are supported"));
    vdbrt_grid_info& info = g->info;
    std::memset(&info, 0, sizeof(info));
    info.bytes = bytes;
    info.leaf_count = rd<uint32_t>(tree + 32); info.lower_count = rd<uint32_t>(tree + 36); info.upper_count = rd<uint32_t>(tree + 40);
    info.active_voxels = rd<uint64_t>(tree + 56);
    info.root_tiles = rd<uint32_t>(rootHead + kRootTableSize);
    info.background = rd<float>(rootHead + kRootBackground);
    info.grid_class = rd<uint32_t>(head + OFF_CLASS);
    info.source_type = 6;                                   // GridType::Vec3f
    for (int i = 0; i < 6; ++i) info.index_bbox[i] = info.node_bbox[i] = rd<int32_t>(rootHead + 4 * i);
    DevColor& c = g->dcolor;
    std::memset(&c, 0, sizeof(c));
    std::memset(&g->dgrid, 0, sizeof(g->dgrid));
    c.base = g->dev; c.root_off = rootOff; c.tiles = g->dev + rootOff + kColRootTiles; c.table_size = info.root_tiles;
    for (int a = 0; a < 3; ++a) {
        c.background[a] = rd<float>(rootHead + kRootBackground + 4 * a);
        const double scale = m[4 * a];
        c.inv[a] = 1.0 / scale;                                  // ScaleMap::mScaleValuesInverse (math/Maps.h:674)
        c.trans[a] = rd<double>(head + OFF_VECD + 8 * a);
        info.voxel_size[a] = std::fabs(scale); info.translation[a] = c.trans[a];
    }
    c.has_translation = (c.trans[0] != 0 || c.trans[1] != 0 || c.trans[2] != 0) ? 1u : 0u;
    *out = g;
    return VDBRT_OK;
}

int vdbrt_free_grid(vdbrt_ctx* ctx, vdbrt_grid* grid)
{
    if (!grid) return VDBRT_OK;
    DeviceGuard guard(grid->device);
    if (ctx) cudaStreamSynchronize(ctx->stream);
    destroyGrid(grid);
    return VDBRT_OK;
}

int vdbrt_grid_get_info(const vdbrt_grid* grid, vdbrt_grid_info* info)
{
    if (!grid || !info) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    *info = grid->info;
    return VDBRT_OK;
}

int vdbrt_grid_download(vdbrt_ctx* ctx, const vdbrt_grid* grid, void* dst, uint64_t bytes)
{
    if (!ctx || !grid || !dst) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (bytes < grid->bytes) return setError(VDBRT_ERR_INVALID_ARG, "destination too small");
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaMemcpyAsync(dst, grid->dev, grid->bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VDBRT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// LevelSetRayTracer::render
// ---------------------------------------------------------------------------------------------------------------
// carve the long-ray buffers out of one allocation sized for `slots` pixel slots
static int longBuffers(vdbrt_ctx* ctx, size_t slots, LongBufs& lb)
{
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    static const size_t capDiv = [] { const char* e = std::getenv("VDBRT_LS_CAPDIV"); const long v = e ? std::atol(e) : 4; return size_t(v < 1 ? 1 : v); }();
    // tail rule: only rays that are in flight when the queue runs dry can be suspended (32 per resident warp)
    const size_t inFlight = size_t(ctx->sm_count) * VDBRT_MINBLOCKS * kBlockThreads;
    const size_t capLong = ctx->ls_tail ? std::max<size_t>(std::min(slots, 2 * inFlight), 65536) : std::max<size_t>(slots / capDiv, 65536);
    const size_t capSeg = capLong * (ctx->ls_tail ? 32 : 16);
    const size_t oCtl = 0, oA = up(sizeof(LongCtl)), oB = oA + up(capLong * 4), oR = oB + up(capLong * 4), oI = oR + up(capLong * sizeof(LongRay));
    const size_t oO = oI + up(capSeg * sizeof(SegIn)), total = oO + up(capSeg * sizeof(SegOut));
    if (int rc = ensureBuffer(&ctx->lng, &ctx->lng_cap, total)) return rc;
    uint8_t* b = static_cast<uint8_t*>(ctx->lng);
    lb.ctl = reinterpret_cast<LongCtl*>(b + oCtl); lb.liveA = reinterpret_cast<uint32_t*>(b + oA); lb.liveB = reinterpret_cast<uint32_t*>(b + oB);
    lb.rays = reinterpret_cast<LongRay*>(b + oR); lb.segIn = reinterpret_cast<SegIn*>(b + oI); lb.segOut = reinterpret_cast<SegOut*>(b + oO);
    lb.capLong = uint32_t(capLong); lb.capSeg = uint32_t(capSeg); lb.budget = ctx->ls_budget; lb.factor = ctx->ls_factor; lb.tail = ctx->ls_tail; lb.voxel_only = ctx->ls_voxel_only;
    return VDBRT_OK;
}

static const char* kQuantNativeOnly = "not available on a grid that is rendered from its Fp8 / Fp16 leaves: upload it with quant_native = 0 (VDBRT_QUANT_NATIVE=0) to have the leaves expanded";

static int launchLevelSet(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam, const vdbrt_shader* shader,
                          const vdbrt_ls_opts* opts, const vdbrt_film* film, float4* dFilm, const float4* dBg, const AuxOut& aux, bool wantAux,
                          unsigned long long* dCounters)
{
    if (grid->leaf_kind != kLeafFloat && (dCounters || opts->iterations > 0)) return setError(VDBRT_ERR_UNSUPPORTED, std::string("work counters / search iterations are ") + kQuantNativeOnly);
    LsParams p; ls_params(grid, opts, film, p);
    p.bg_film = dBg;
    const TileMap tm = makeTileMap(film->width, film->height, opts->part.tile_w, opts->part.tile_h, opts->part.rank, opts->part.count);
    DevShader sh;
    sh.kind = shader->kind; sh.r = shader->rgba[0]; sh.g = shader->rgba[1]; sh.b = shader->rgba[2]; sh.a = shader->rgba[3];
    for (int a = 0; a < 3; ++a) { sh.bmin[a] = shader->bbox_min[a]; sh.inv[a] = shader->inv_dim[a]; }
    std::memset(&sh.col, 0, sizeof(sh.col));
    if (shader->color_grid) sh.col = shader->color_grid->dcolor;
    const DevCamera dc = toDev(*cam);
    unsigned int* queue = reinterpret_cast<unsigned int*>(ctx->scratch + 64);
    // long-ray rounds: one sample per pixel only (with more, the samples of a pixel are accumulated in order by one thread).
    // They pay when the launch is bounded by its slowest tile, i.e. when it has few tiles per resident warp -- a small frame, or a
    // small share of a frame that is split over several GPUs (C2 at 1/8 of the frame: 1.22 -> 0.68 ms per rank, 1/4: 1.41 -> 0.95).
    // With many tiles per warp the tail is a small part of the launch and the rounds cost about what they save (C2 whole frame
    // 2.72 -> 2.63 ms), and where the long rays cross empty space rather than graze a surface they lose outright (C4: 34.4 -> 36.7 ms,
    // 1/8 share 5.2 -> 5.5 ms: the scout walks empty cells no faster than the render kernel): profiles/r02_summary.md.  Default: fewer
    // than kRoundsMaxTilesPerSm tiles per SM.  VDBRT_LS_ROUNDS_ON / _OFF override.  WHICH rays are suspended is the tail rule
    // (ctx->ls_tail, vdbrt_kernels.cuh) unless VDBRT_LS_TAIL=0 selects round 1's per-tile budget.
    const double tilesPerWarp = double(tm.items) / (double(ctx->sm_count) * VDBRT_MINBLOCKS * (kBlockThreads / 32));
    const bool automatic = double(tm.items) / double(ctx->sm_count) < kRoundsMaxTilesPerSm && !(opts->flags & VDBRT_LS_ROUNDS_OFF);
    const bool rounds = ((opts->flags & VDBRT_LS_ROUNDS_ON) || automatic) && !dCounters && opts->spp == 1 && opts->iterations == 0 && grid->leaf_kind == kLeafFloat && (ctx->ls_tail != 0 || ctx->ls_budget != 0) && ctx->ls_rounds != 0;
    LongBufs lb = {};
    lb.budget = 0xffffffffu;
    if (rounds) {
        if (int rc = longBuffers(ctx, size_t(tm.items) * 32, lb)) return rc;
        CUDA_TRY(cudaMemsetAsync(lb.ctl, 0, sizeof(LongCtl), ctx->stream));
    }
    // how the warps are fed (Sched, vdbrt_kernels.cuh): strips of 8x4 tiles, lanes re-fed from the warp's own strip, heavy strips first
    Sched sc = {};
    // strips as long as the launch still has at least four of them per resident warp (a small share of a partitioned frame gets single tiles)
    const uint32_t residentWarps = uint32_t(ctx->sm_count) * VDBRT_MINBLOCKS * (kBlockThreads / 32);
    sc.strip_tiles = ctx->ls_strip ? ctx->ls_strip : 1u;
    if (ctx->ls_strip_ratio) sc.strip_tiles = std::max(1u, std::min(sc.strip_tiles, tm.items / (ctx->ls_strip_ratio * residentWarps)));
    sc.refill = ctx->ls_refill; sc.eager = ctx->ls_eager;
    if (sc.refill < 1u || sc.refill > 32u) sc.refill = 32u;
    const uint32_t nStrips = (tm.items + sc.strip_tiles - 1u) / sc.strip_tiles;
    // heavy strips first: worth a probe launch when a warp gets more than a couple of tiles (otherwise everything starts at once anyway)
    const bool orderAuto = ctx->ls_order == 1u || (ctx->ls_order == 2u && tilesPerWarp >= 2.0);
    const bool order = !dCounters && nStrips > 1u && grid->leaf_kind == kLeafFloat && !(opts->flags & VDBRT_LS_ORDER_OFF) && ((opts->flags & VDBRT_LS_ORDER_ON) || orderAuto);
    CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(unsigned int), ctx->stream));
    if (ctx->ls_affine && !order) {
        sc.affine = ctx->ls_affine; sc.nq = uint32_t(ctx->sm_count > 0 ? std::min(ctx->sm_count, 256) : 1);
        sc.smq = reinterpret_cast<unsigned int*>(ctx->scratch + 1024);               // 256 x 4 bytes of the 4 KB scratch block
        CUDA_TRY(cudaMemsetAsync(sc.smq, 0, sizeof(unsigned int) * sc.nq, ctx->stream));
    }
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    uint32_t launches = 0;
    // Heavy tiles first from the previous frame's measured tile costs (k_order_from_history): on by default (ctx->ls_history), for
    // any sequence of frames with the same film / tiles / partition through one context.  The costs are hints for the ORDER in
    // which the work queue hands out tiles -- every ray of every frame is traced; a stale history only means a worse order.
    const bool history = ctx->ls_history && !dCounters && !order && sc.strip_tiles == 1u && tm.items > 1u && !(opts->flags & VDBRT_LS_ORDER_OFF);
    if (history) {
        auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
        const size_t n = tm.items, oCost = 256, total = oCost + up(4 * n);
        const bool same = ctx->hist && ctx->hist_valid && std::memcmp(ctx->hist_tm, &tm, sizeof(tm)) == 0 && ctx->hist_grid == grid && ctx->hist_spp == opts->spp;
        if (!same) {
            if (int rc = ensureBuffer(&ctx->hist, &ctx->hist_cap, total)) return rc;
            ctx->hist_valid = 0;
        }
        uint8_t* hb = static_cast<uint8_t*>(ctx->hist);
        sc.cost_sum = reinterpret_cast<unsigned long long*>(hb); sc.cost_max = reinterpret_cast<uint32_t*>(hb + 8); sc.cost_out = reinterpret_cast<uint32_t*>(hb + oCost);
        if (!ctx->hist_host) CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&ctx->hist_host), 16, cudaHostAllocDefault));
        if (!same) ctx->hist_host[0] = ctx->hist_host[1] = 0ull;
        if (same) {
            const size_t oCls = 256, oA = oCls + up(n), oB = oA + up(4 * n), totalB = oB + up(4 * n);
            if (int rc = ensureBuffer(&ctx->ord, &ctx->ord_cap, totalB)) return rc;
            uint8_t* b = static_cast<uint8_t*>(ctx->ord);
            OrderBufs ob = {};
            ob.ctl = reinterpret_cast<uint32_t*>(b); ob.cls = b + oCls; ob.listA = reinterpret_cast<uint32_t*>(b + oA); ob.listB = reinterpret_cast<uint32_t*>(b + oB);
            CUDA_TRY(cudaMemsetAsync(b, 0, oA, ctx->stream));
            k_order_from_history<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(sc.cost_out, sc.cost_sum, tm.items, ob, 0.01f * float(ctx->ls_hist_a), 0.01f * float(ctx->ls_hist_b));
            CUDA_TRY(cudaGetLastError());
            sc.ctl = ob.ctl; sc.listA = ob.listA; sc.listB = ob.listB; sc.cls = ob.cls;
            ++launches;
        }
        CUDA_TRY(cudaMemsetAsync(sc.cost_sum, 0, 16, ctx->stream));                    // sum and max
        static_assert(sizeof(TileMap) <= sizeof(ctx->hist_tm), "history key");
        std::memcpy(ctx->hist_tm, &tm, sizeof(tm)); ctx->hist_grid = grid; ctx->hist_spp = opts->spp; ctx->hist_valid = 1;
    }
    if (order) {
        auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
        const size_t n = nStrips, oCost = 256, oDone = oCost + up(4 * n), oCls = oDone + up(4 * n), oA = oCls + up(n), oB = oA + up(4 * n), totalB = oB + up(4 * n);
        if (int rc = ensureBuffer(&ctx->ord, &ctx->ord_cap, totalB)) return rc;
        uint8_t* b = static_cast<uint8_t*>(ctx->ord);
        OrderBufs ob;
        ob.ctl = reinterpret_cast<uint32_t*>(b); ob.cost = reinterpret_cast<uint32_t*>(b + oCost); ob.done = reinterpret_cast<uint32_t*>(b + oDone);
        ob.cls = b + oCls; ob.listA = reinterpret_cast<uint32_t*>(b + oA); ob.listB = reinterpret_cast<uint32_t*>(b + oB);
        CUDA_TRY(cudaMemsetAsync(b, 0, oA, ctx->stream));
        const unsigned cap = unsigned(ctx->sm_count) * 16u, need = (tm.items + kBlockThreads - 1) / kBlockThreads;
        k_probe_levelset<<<need < cap ? need : cap, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dc, p, tm, ob, sc.strip_tiles, ctx->ls_probe_cap,
                                                                                     ctx->ls_probe_cap, ctx->ls_probe_b);
        CUDA_TRY(cudaGetLastError());
        sc.ctl = ob.ctl; sc.listA = ob.listA; sc.listB = ob.listB; sc.cls = ob.cls;
        ++launches;
        if (ctx->evp) CUDA_TRY(cudaEventRecord(ctx->evp, ctx->stream));
    }
    // LONG = true compiles the suspension of over-budget rays into the kernel; MULTI = more than one sample per pixel
    const bool multi = opts->spp > 1;
    const bool refine = opts->iterations > 0;      // LinearSearchImpl<.., Iterations > 0>: its own instantiations (never with the rounds)
    using KernT = void (*)(DevGrid, DevCamera, DevShader, LsParams, TileMap, float4*, AuxOut, unsigned int*, unsigned long long*, LongBufs, Sched);
    KernT kern;
    if (grid->leaf_kind != kLeafFloat) {
        // NanoGrid<Fp8|Fp16> as they are: the quantised instantiations (frames with or without records, one or several samples)
        const bool q8 = grid->leaf_kind == kLeafFp8;
        kern = wantAux ? (multi ? (q8 ? (KernT)k_render_levelset<true, false, false, true, false, kLeafFp8> : (KernT)k_render_levelset<true, false, false, true, false, kLeafFp16>)
                                : (q8 ? (KernT)k_render_levelset<true, false, false, false, false, kLeafFp8> : (KernT)k_render_levelset<true, false, false, false, false, kLeafFp16>))
                       : (multi ? (q8 ? (KernT)k_render_levelset<false, false, false, true, false, kLeafFp8> : (KernT)k_render_levelset<false, false, false, true, false, kLeafFp16>)
                                : (q8 ? (KernT)k_render_levelset<false, false, false, false, false, kLeafFp8> : (KernT)k_render_levelset<false, false, false, false, false, kLeafFp16>));
    } else
    kern =
        dCounters ? k_render_levelset<false, true, false, true>
        : refine ? (wantAux ? (multi ? k_render_levelset<true, false, false, true, true> : k_render_levelset<true, false, false, false, true>)
                            : (multi ? k_render_levelset<false, false, false, true, true> : k_render_levelset<false, false, false, false, true>))
        : wantAux ? (rounds ? k_render_levelset<true, false, true, false> : multi ? k_render_levelset<true, false, false, true> : k_render_levelset<true, false, false, false>)
                  : (rounds ? k_render_levelset<false, false, true, false> : multi ? k_render_levelset<false, false, false, true> : k_render_levelset<false, false, false, false>);
    // The throughput-bound instantiation (6 CTAs per SM, DENSE).  It pays unless the launch is as long as the critical path of its heaviest
    // tile, which more warps per SM only stretch (C4 whole and at 1/2, 1/4, 1/8 of the frame: -6 %; C1: -4 %; C2 whole / at 1/2: +5 / +3 %).
    // With the tile costs of the previous frame of this sequence (sum and max, copied back asynchronously after every frame and read here
    // without a sync: a hint, possibly one frame old) the rule is just that: a warp's share of the summed tile times against the longest tile
    // (measured ratios: C2 at 1/2 0.50, C2 0.84 | C1 1.16, C4 at 1/8 1.42, 1/4 2.75, 1/2 5.6, whole 11.0 -- the bar is the threshold, 1.0).
    // Without them: very many tiles per SM only.
    bool dense = false;
    if (ctx->ls_dense && grid->leaf_kind == kLeafFloat && !dCounters && !refine && !rounds && !multi) {
        const unsigned long long hsum = history && ctx->hist_host ? *reinterpret_cast<volatile unsigned long long*>(ctx->hist_host) : 0ull;
        const unsigned long long hmax = history && ctx->hist_host ? (*reinterpret_cast<volatile unsigned long long*>(ctx->hist_host + 1) & 0xffffffffull) : 0ull;
        if (hsum && hmax && hmax < 0x7fffffffull)
            dense = double(hsum) / (double(ctx->sm_count) * VDBRT_MINBLOCKS_DENSE * (kBlockThreads / 32)) >= 0.01 * double(ctx->ls_dense_factor) * double(hmax);
        else dense = double(tm.items) / double(ctx->sm_count) >= double(ctx->ls_dense);
        static const bool debugDense = std::getenv("VDBRT_DEBUG_DENSE") != nullptr;
        if (debugDense)
            std::fprintf(stderr, "[vdbrt] dense? tiles %u, previous frame: sum %llu max %llu -> a warp's share / heaviest tile = %.3f -> %s\n", tm.items, hsum, hmax,
                         hmax ? double(hsum) / (double(ctx->sm_count) * VDBRT_MINBLOCKS_DENSE * (kBlockThreads / 32)) / double(hmax) : 0.0, dense ? "6 CTAs" : "5 CTAs");
    }
    if (dense)
        kern = wantAux ? (KernT)k_render_levelset<true, false, false, false, false, kLeafFloat, true> : (KernT)k_render_levelset<false, false, false, false, false, kLeafFloat, true>;
    const int blocks = persistentGrid(ctx, (const void*)kern, nStrips);
    static const bool debugExit = std::getenv("VDBRT_DEBUG_EXIT") != nullptr;
    const size_t nWarps = size_t(blocks) * (kBlockThreads / 32);
    if (debugExit) {
        if (int rc = ensureBuffer(&ctx->io, &ctx->io_cap, 8 * (nWarps + 1))) return rc;
        sc.warp_exit = static_cast<unsigned long long*>(ctx->io);
    }
    kern<<<blocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dc, sh, p, tm, dFilm, aux, queue, dCounters, lb, sc);
    if (history && ctx->hist_host) CUDA_TRY(cudaMemcpyAsync(ctx->hist_host, sc.cost_sum, 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (debugExit) {
        // when did the warps leave?  (the tail of the frame: time between the mean and the last exit)
        std::vector<unsigned long long> t(nWarps + 1);
        CUDA_TRY(cudaMemcpyAsync(t.data(), ctx->io, 8 * (nWarps + 1), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        std::vector<double> us(nWarps);
        for (size_t i = 0; i < nWarps; ++i) us[i] = double(t[i + 1] - t[0]) * 1e-3;
        std::sort(us.begin(), us.end());
        double mean = 0; for (double v : us) mean += v; mean /= double(nWarps);
        std::fprintf(stderr, "[vdbrt] warp exits (us after launch): first %.0f  p10 %.0f  median %.0f  mean %.0f  p90 %.0f  p99 %.0f  last %.0f  (%zu warps, %u tiles)\n",
                     us.front(), us[nWarps / 10], us[nWarps / 2], mean, us[nWarps * 9 / 10], us[nWarps * 99 / 100], us.back(), nWarps, tm.items);
    }
    CUDA_TRY(cudaGetLastError());
    ++launches;
    ctx->last_launches = launches;
    if (rounds) {
        // K leaf visits per ray and round: small first (most suspended rays hit soon), then growing
        // K is clamped so that the 32-bit segment counter cannot wrap (every live ray adds K per round, at most capLong rays are live)
        uint32_t kLeaves[kMaxRounds];
        for (int r = 0; r < kMaxRounds; ++r) kLeaves[r] = std::max(1u, std::min(ctx->ls_leaves[r], uint32_t(0xffffffffu / std::max(1u, lb.capLong)) - 1u));
        const int wide = ctx->sm_count * 8;
        const int nr = int(ctx->ls_rounds);
        for (int r = 0; r < nr; ++r) {
            if (wantAux) k_long_scout<true><<<wide * 2, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, sh, p, dFilm, aux, lb, r, kLeaves[r]);
            else k_long_scout<false><<<wide * 2, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, sh, p, dFilm, aux, lb, r, kLeaves[r]);
            k_long_march<<<wide * 2, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, lb, r, kLeaves[r], p.iso, p.vmin, p.vmax);
        }
        if (wantAux) k_long_finish<true><<<wide, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, sh, p, dFilm, aux, lb, nr);
        else k_long_finish<false><<<wide, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, sh, p, dFilm, aux, lb, nr);
        CUDA_TRY(cudaGetLastError());
        ctx->last_launches = launches + 1 + 2 * uint32_t(nr);
    }
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    if (rounds && std::getenv("VDBRT_DEBUG_LONG")) {
        LongCtl h;
        CUDA_TRY(cudaMemcpyAsync(&h, lb.ctl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        std::fprintf(stderr, "[vdbrt] %u tiles, %u finished with mean %.1f iterations, %d blocks; long rays %u (cap %u), segment cap %u; live per round:",
                     tm.items, h.tiles, h.tiles ? double(h.spent) / h.tiles : 0.0, blocks, h.nLong, lb.capLong, lb.capSeg);
        for (int r = 0; r < int(ctx->ls_rounds); ++r) std::fprintf(stderr, " %u", h.live[r]);
        std::fprintf(stderr, "; segment slots:");
        for (int r = 0; r < int(ctx->ls_rounds); ++r) std::fprintf(stderr, " %u", h.segCount[r]);
        std::fprintf(stderr, "\n");
    }
    return VDBRT_OK;
}

int vdbrt_render_levelset(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam, const vdbrt_shader* shader,
                          const vdbrt_ls_opts* opts, vdbrt_film* film, vdbrt_aux* aux)
{
    if (!ctx || !grid || !cam || !shader || !opts || !film || !film->pixels) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (film->width == 0 || film->height == 0) return setError(VDBRT_ERR_INVALID_ARG, "empty film");
    if (cam->width != film->width || cam->height != film->height) return setError(VDBRT_ERR_INVALID_ARG, "camera was built for a different film size");
    if (shader->kind > VDBRT_SHADER_DIFFUSE) return setError(VDBRT_ERR_UNSUPPORTED, "only the matte, normal, position and diffuse shaders run on the device");
    if (shader->color_grid && (!shader->color_grid->is_color || shader->color_grid->device != ctx->device))
        return setError(VDBRT_ERR_INVALID_ARG, "color_grid is not a Vec3f grid uploaded to this device with vdbrt_upload_color_grid");
    if (grid->is_color) return setError(VDBRT_ERR_NOT_FLOAT, "a colour grid cannot be rendered itself");
    if (opts->spp == 0) return setError(VDBRT_ERR_SPP_ZERO, "pixelSamples must be larger than zero!");
    if (opts->part.count && opts->part.rank >= opts->part.count) return setError(VDBRT_ERR_INVALID_ARG, "partition rank must be smaller than the partition count");
    if (int rc = checkLevelSet(grid, opts->iso)) return rc;
    std::lock_guard<std::mutex> lock(ctx->mx);         // one render at a time per context: the work queue, counters and staging buffers are the context's
    DeviceGuard guard(ctx->device);
    const size_t npx = size_t(film->width) * film->height;
    const bool host = film->memspace == VDBRT_MEM_HOST;
    float4* dFilm = reinterpret_cast<float4*>(film->pixels);
    const float4* dBg = dFilm;
    bool copyBack = false;
    if (host) {
        // level-set misses keep the previous pixel, so the film is an input too (tools/RayTracer.h:908).  A PINNED host film is
        // used in place: a miss reads its old pixel over PCIe right before the same thread overwrites it (what the reference's
        // `bg = mCamera->pixel(i,j)` does), every owned pixel is stored straight to host memory while the rest of the frame is
        // still being traced, and the pixels of other ranks are never touched -- no film copy in either direction.  A pageable
        // film is staged on the device: one copy in (unless the old film is known to be uniform), one copy out.
        const bool uniform = (opts->flags & VDBRT_LS_UNIFORM_BG) != 0;
        float4* alias = nullptr;
        const int kind = classifyHostFilm(film->pixels, npx * 16, &alias);
        if (kind == 2) return setError(VDBRT_ERR_INVALID_ARG, kPartlyPinned);
        if (kind == 1) { dFilm = alias; dBg = uniform ? nullptr : alias; }
        else {
            if (int rc = ensureBuffer(&ctx->film, &ctx->film_cap, npx * 16)) return rc;
            if (!uniform || opts->part.count > 1) CUDA_TRY(cudaMemcpyAsync(ctx->film, film->pixels, npx * 16, cudaMemcpyHostToDevice, ctx->stream));
            dFilm = static_cast<float4*>(ctx->film); dBg = uniform ? nullptr : dFilm; copyBack = true;
        }
    }
    AuxOut a = {};
    const bool wantAux = aux && (aux->hit || aux->ijk || aux->t_index || aux->t_world || aux->xyz || aux->nml);
    // aux layout in the staging buffer: hit[npx] | pad | ijk[3npx] | t_index | t_world | xyz | nml
    size_t offIjk = 0, offTi = 0, offTw = 0, offXyz = 0, offNml = 0, auxBytes = 0;
    if (wantAux) {
        if (host) {
            offIjk = (npx + 255) & ~size_t(255); offTi = offIjk + npx * 12; offTi = (offTi + 255) & ~size_t(255);
            offTw = offTi + npx * 8; offXyz = offTw + npx * 8; offNml = offXyz + npx * 24; auxBytes = offNml + npx * 24;
            if (int rc = ensureBuffer(&ctx->aux, &ctx->aux_cap, auxBytes)) return rc;
            CUDA_TRY(cudaMemsetAsync(ctx->aux, 0, auxBytes, ctx->stream));
            uint8_t* b = static_cast<uint8_t*>(ctx->aux);
            a.hit = aux->hit ? b : nullptr; a.ijk = aux->ijk ? reinterpret_cast<int32_t*>(b + offIjk) : nullptr;
            a.t_index = aux->t_index ? reinterpret_cast<double*>(b + offTi) : nullptr;
            a.t_world = aux->t_world ? reinterpret_cast<double*>(b + offTw) : nullptr;
            a.xyz = aux->xyz ? reinterpret_cast<double*>(b + offXyz) : nullptr;
            a.nml = aux->nml ? reinterpret_cast<double*>(b + offNml) : nullptr;
        } else {
            a.hit = aux->hit; a.ijk = aux->ijk; a.t_index = aux->t_index; a.t_world = aux->t_world; a.xyz = aux->xyz; a.nml = aux->nml;
        }
    }
    if (int rc = launchLevelSet(ctx, grid, cam, shader, opts, film, dFilm, dBg, a, wantAux, nullptr)) return rc;
    if (host) {
        if (copyBack) CUDA_TRY(cudaMemcpyAsync(film->pixels, dFilm, npx * 16, cudaMemcpyDeviceToHost, ctx->stream));
        if (wantAux) {
            const uint8_t* b = static_cast<const uint8_t*>(ctx->aux);
            if (aux->hit) CUDA_TRY(cudaMemcpyAsync(aux->hit, b, npx, cudaMemcpyDeviceToHost, ctx->stream));
            if (aux->ijk) CUDA_TRY(cudaMemcpyAsync(aux->ijk, b + offIjk, npx * 12, cudaMemcpyDeviceToHost, ctx->stream));
            if (aux->t_index) CUDA_TRY(cudaMemcpyAsync(aux->t_index, b + offTi, npx * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (aux->t_world) CUDA_TRY(cudaMemcpyAsync(aux->t_world, b + offTw, npx * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (aux->xyz) CUDA_TRY(cudaMemcpyAsync(aux->xyz, b + offXyz, npx * 24, cudaMemcpyDeviceToHost, ctx->stream));
            if (aux->nml) CUDA_TRY(cudaMemcpyAsync(aux->nml, b + offNml, npx * 24, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    if (host || !(opts->flags & VDBRT_ASYNC)) CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VDBRT_OK;
}

int vdbrt_count_levelset(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam, const vdbrt_ls_opts* opts, vdbrt_counters* out)
{
    if (!ctx || !grid || !cam || !opts || !out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (opts->spp == 0) return setError(VDBRT_ERR_SPP_ZERO, "pixelSamples must be larger than zero!");
    if (int rc = checkLevelSet(grid, opts->iso)) return rc;
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    const size_t npx = size_t(cam->width) * cam->height;
    if (int rc = ensureBuffer(&ctx->io, &ctx->io_cap, npx * 16)) return rc;      // scratch film, discarded
    CUDA_TRY(cudaMemsetAsync(ctx->io, 0, npx * 16, ctx->stream));
    unsigned long long* dC = reinterpret_cast<unsigned long long*>(ctx->scratch + 128);
    CUDA_TRY(cudaMemsetAsync(dC, 0, 40 * sizeof(unsigned long long), ctx->stream));
    vdbrt_film film = {}; film.width = cam->width; film.height = cam->height;
    vdbrt_shader sh = {}; sh.kind = VDBRT_SHADER_DIFFUSE; sh.rgba[0] = sh.rgba[1] = sh.rgba[2] = sh.rgba[3] = 1.f;
    AuxOut a = {};
    if (int rc = launchLevelSet(ctx, grid, cam, &sh, opts, &film, static_cast<float4*>(ctx->io), static_cast<const float4*>(ctx->io), a, false, dC)) return rc;
    unsigned long long h[40];
    CUDA_TRY(cudaMemcpyAsync(h, dC, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    out->rays = h[0]; out->root_probes = h[1]; out->upper_probes = h[2]; out->lower_probes = h[3]; out->voxel_probes = h[4];
    out->stencil_refills = h[5]; out->primary_samples = h[6]; out->shadow_samples = h[7]; out->shadow_rays = h[8]; out->hits = h[9];
    if (std::getenv("VDBRT_DEBUG_TILES"))
        std::fprintf(stderr, "[vdbrt] tiles %llu: max cycles %llu, max iterations %llu, mean cycles %.0f; %llu tiles > 1.5M cycles with mean active lanes %.2f\n", h[13], h[10], h[11],
                     h[13] ? double(h[12]) / double(h[13]) : 0.0, h[14], h[14] ? double(h[15]) / 100.0 / double(h[14]) : 0.0);
    if (std::getenv("VDBRT_DEBUG_TILES")) {
        unsigned long long its = 0;
        for (int k = 0; k < 9; ++k) its += h[16 + k];
        std::fprintf(stderr, "[vdbrt] warp iterations %llu; running lanes 0 / 1-4 / ... / 29-32:", its);
        for (int k = 0; k < 9; ++k) std::fprintf(stderr, " %.1f%%", its ? 100.0 * double(h[16 + k]) / double(its) : 0.0);
        static const char* names[4] = {"node probe", "voxel probe", "stencil", "step"};
        std::fprintf(stderr, "\n[vdbrt] phase: share of the iterations that ran it, lanes in it when it ran:");
        for (int k = 0; k < 4; ++k)
            std::fprintf(stderr, " %s %.1f%% x %.1f;", names[k], its ? 100.0 * double(h[36 + k]) / double(its) : 0.0, h[36 + k] ? double(h[32 + k]) / double(h[36 + k]) : 0.0);
        std::fprintf(stderr, "\n");
    }
    return VDBRT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// VolumeRender::render
// ---------------------------------------------------------------------------------------------------------------
static int launchVolume(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam, const vdbrt_vol_opts* opts,
                        uint32_t width, uint32_t height, float4* dFilm, unsigned long long* dCounters)
{
    VolParams p; vol_params(opts, p);
    {
        const DevGrid& g = grid->dgrid;
        double jx = p.light[0], jy = p.light[1], jz = p.light[2];
        if (g.general) {
            const double a = jx * g.imat[0] + jy * g.imat[1] + jz * g.imat[2], b = jx * g.imat[3] + jy * g.imat[4] + jz * g.imat[5], c = jx * g.imat[6] + jy * g.imat[7] + jz * g.imat[8];
            jx = a; jy = b; jz = c;
        } else { jx *= g.inv[0]; jy *= g.inv[1]; jz *= g.inv[2]; }
        const double len = std::sqrt(jx * jx + jy * jy + jz * jz);
        const double dx = jx / len, dy = jy / len, dz = jz / len;
        p.sb[0] = dx; p.sb[1] = dy; p.sb[2] = dz; p.sb[3] = 1 / dx; p.sb[4] = 1 / dy; p.sb[5] = 1 / dz; p.sb[6] = len * 1e-9; p.sb[7] = len * DBL_MAX;
    }
    const TileMap tm = makeTileMap(width, height, opts->part.tile_w, opts->part.tile_h, opts->part.rank, opts->part.count);
    const DevCamera dc = toDev(*cam);
    unsigned int* queue = reinterpret_cast<unsigned int*>(ctx->scratch + 64);
    CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(unsigned int), ctx->stream));
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    const VolTiles all = {0u, tm.items, nullptr, nullptr};
    const int kind = grid->leaf_kind;
    using OneLoopT = void (*)(DevGrid, DevCamera, VolParams, TileMap, float4*, unsigned int*, unsigned long long*, VolTiles);
    const OneLoopT oneLoop = kind == kLeafFp8 ? (OneLoopT)k_render_volume<false, kLeafFp8> : kind == kLeafFp16 ? (OneLoopT)k_render_volume<false, kLeafFp16> : (OneLoopT)k_render_volume<false>;
    if (dCounters) {
        if (kind != kLeafFloat) return setError(VDBRT_ERR_UNSUPPORTED, std::string("work counters are ") + kQuantNativeOnly);
        const int blocks = persistentGrid(ctx, (const void*)k_render_volume<true>, tm.items);
        k_render_volume<true><<<blocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dc, p, tm, dFilm, queue, dCounters, all);
        ctx->last_launches = 1;
    } else if (!ctx->fog_wave || tm.items == 0) {
        const int blocks = persistentGrid(ctx, (const void*)oneLoop, tm.items);
        oneLoop<<<blocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dc, p, tm, dFilm, queue, nullptr, all);
        ctx->last_launches = 1;
    } else {
        // Wavefront (vdbrt_fog.cuh): primary rays -> records of the dense samples -> shadow rays -> pixels, in batches of tiles whose
        // records fit the buffer (fog_rec_per_ray records per primary ray on average; a tile that runs out is re-rendered by the
        // one-loop kernel, so the budget is a performance knob, not a limit).
        auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
        const uint32_t spp = p.sub + 1u;
        const size_t raysPerTile = 32u * size_t(spp);
        const size_t perRay = ctx->fog_rec_per_ray ? ctx->fog_rec_per_ray : 12u;
        const size_t budget = size_t(ctx->fog_cap_mb ? ctx->fog_cap_mb : 4096u) << 20;
        const size_t bytesPerTile = raysPerTile * (perRay * sizeof(FogRec) + sizeof(FogRayRec)) + 1;
        const uint32_t batchTiles = uint32_t(std::max<size_t>(1, std::min<size_t>(tm.items, budget / bytesPerTile)));
        const size_t cap = std::min<size_t>(size_t(batchTiles) * raysPerTile * perRay, 0xfffffff0u);
        const size_t oCtl = 0, oFlag = 256, oRays = oFlag + up(batchTiles), oRecs = oRays + up(size_t(batchTiles) * raysPerTile * sizeof(FogRayRec));
        const size_t total = oRecs + up(cap * sizeof(FogRec));
        if (int rc = ensureBuffer(&ctx->fog, &ctx->fog_cap, total)) return rc;
        uint8_t* b = static_cast<uint8_t*>(ctx->fog);
        FogWave fw;
        fw.ctl = reinterpret_cast<unsigned int*>(b + oCtl); fw.tileFlag = b + oFlag; fw.rays = reinterpret_cast<FogRayRec*>(b + oRays);
        fw.recs = reinterpret_cast<FogRec*>(b + oRecs); fw.cap = uint32_t(cap);
        fw.refill = ctx->fog_refill >= 1u && ctx->fog_refill <= 32u ? ctx->fog_refill : 32u;
        using PrimT = void (*)(DevGrid, DevCamera, VolParams, TileMap, FogWave);
        using ShadT = void (*)(DevGrid, VolParams, FogWave);
        const PrimT primary = kind == kLeafFp8 ? (PrimT)k_fog_primary<kLeafFp8> : kind == kLeafFp16 ? (PrimT)k_fog_primary<kLeafFp16> : (PrimT)k_fog_primary<kLeafFloat>;
        const ShadT shadow = kind == kLeafFp8 ? (ShadT)k_fog_shadow<kLeafFp8> : kind == kLeafFp16 ? (ShadT)k_fog_shadow<kLeafFp16> : (ShadT)k_fog_shadow<kLeafFloat>;
        int perSm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, (const void*)shadow, kBlockThreads, 0) != cudaSuccess || perSm < 1) perSm = 1;
        const int shadowBlocks = ctx->sm_count * perSm;
        uint32_t launches = 0;
        for (uint32_t t0 = 0; t0 < tm.items; t0 += batchTiles) {
            fw.tile0 = t0; fw.tile1 = std::min(tm.items, t0 + batchTiles);
            CUDA_TRY(cudaMemsetAsync(b, 0, oRays, ctx->stream));                       // queue words, counters, tile flags
            const int blocks = persistentGrid(ctx, (const void*)primary, fw.tile1 - fw.tile0);
            primary<<<blocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dc, p, tm, fw);
            shadow<<<shadowBlocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, p, fw);
            const unsigned long long slots = (unsigned long long)(fw.tile1 - fw.tile0) * 32ull;
            k_fog_resolve<<<unsigned(std::min<unsigned long long>((slots + 255) / 256, (unsigned long long)ctx->sm_count * 32ull)), 256, 0, ctx->stream>>>(p, tm, fw, dFilm);
            // tiles that ran out of record space: the one-loop kernel (returns at once when there is none)
            CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(unsigned int), ctx->stream));
            const VolTiles flagged = {fw.tile0, fw.tile1, fw.tileFlag, fw.ctl + 2};
            oneLoop<<<persistentGrid(ctx, (const void*)oneLoop, fw.tile1 - fw.tile0), kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dc, p, tm, dFilm, queue, nullptr, flagged);
            launches += 4;
        }
        ctx->last_launches = launches;
        if (std::getenv("VDBRT_DEBUG_FOG")) {
            unsigned int h[4];
            CUDA_TRY(cudaMemcpyAsync(h, fw.ctl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            std::fprintf(stderr, "[vdbrt] fog wavefront: %u tiles in batches of %u, last batch: %u records of %zu (%.2f per ray), %u tiles flagged\n", tm.items, batchTiles,
                         h[1], cap, double(h[1]) / double((fw.tile1 - fw.tile0) * raysPerTile), h[2]);
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    return VDBRT_OK;
}

int vdbrt_render_volume(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam, const vdbrt_vol_opts* opts, vdbrt_film* film)
{
    if (!ctx || !grid || !cam || !opts || !film || !film->pixels) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (film->width == 0 || film->height == 0) return setError(VDBRT_ERR_INVALID_ARG, "empty film");
    if (cam->width != film->width || cam->height != film->height) return setError(VDBRT_ERR_INVALID_ARG, "camera was built for a different film size");
    if (opts->part.count && opts->part.rank >= opts->part.count) return setError(VDBRT_ERR_INVALID_ARG, "partition rank must be smaller than the partition count");
    if (int rc = checkVolume(grid)) return rc;
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    const size_t npx = size_t(film->width) * film->height;
    const bool host = film->memspace == VDBRT_MEM_HOST;
    float4* dFilm = reinterpret_cast<float4*>(film->pixels);
    bool copyBack = false;
    if (host) {
        // every owned pixel is overwritten (tools/RayTracer.h:1020) and nothing is read: a pinned film is written in place by
        // the kernel; a pageable one is staged (a partitioned render then needs the old film for the pixels of other ranks)
        float4* alias = nullptr;
        const int kind = classifyHostFilm(film->pixels, npx * 16, &alias);
        if (kind == 2) return setError(VDBRT_ERR_INVALID_ARG, kPartlyPinned);
        if (kind == 1) dFilm = alias;
        else {
            if (int rc = ensureBuffer(&ctx->film, &ctx->film_cap, npx * 16)) return rc;
            dFilm = static_cast<float4*>(ctx->film); copyBack = true;
            if (opts->part.count > 1) CUDA_TRY(cudaMemcpyAsync(dFilm, film->pixels, npx * 16, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    if (int rc = launchVolume(ctx, grid, cam, opts, film->width, film->height, dFilm, nullptr)) return rc;
    if (copyBack) CUDA_TRY(cudaMemcpyAsync(film->pixels, dFilm, npx * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (host || !(opts->flags & VDBRT_ASYNC)) CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VDBRT_OK;
}

int vdbrt_film_over(vdbrt_ctx* ctx, vdbrt_film* top, const vdbrt_film* bottom)
{
    if (!ctx || !top || !bottom || !top->pixels || !bottom->pixels) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (top->width != bottom->width || top->height != bottom->height || top->memspace != bottom->memspace)
        return setError(VDBRT_ERR_INVALID_ARG, "films differ in size or memory space");
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    const size_t npx = size_t(top->width) * top->height;
    float4* dTop = reinterpret_cast<float4*>(top->pixels);
    const float4* dBot = reinterpret_cast<const float4*>(bottom->pixels);
    const bool host = top->memspace == VDBRT_MEM_HOST;
    if (host) {
        if (int rc = ensureBuffer(&ctx->film, &ctx->film_cap, npx * 16)) return rc;
        if (int rc = ensureBuffer(&ctx->io, &ctx->io_cap, npx * 16)) return rc;
        CUDA_TRY(cudaMemcpyAsync(ctx->film, top->pixels, npx * 16, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->io, bottom->pixels, npx * 16, cudaMemcpyHostToDevice, ctx->stream));
        dTop = static_cast<float4*>(ctx->film); dBot = static_cast<const float4*>(ctx->io);
    }
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    k_film_over<<<unsigned((npx + 255) / 256), 256, 0, ctx->stream>>>(dTop, dBot, npx);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->last_launches = 1;
    if (host) {
        CUDA_TRY(cudaMemcpyAsync(top->pixels, dTop, npx * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return VDBRT_OK;
}

int vdbrt_count_volume(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_camera* cam, const vdbrt_vol_opts* opts, vdbrt_counters* out)
{
    if (!ctx || !grid || !cam || !opts || !out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (int rc = checkVolume(grid)) return rc;
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    const size_t npx = size_t(cam->width) * cam->height;
    if (int rc = ensureBuffer(&ctx->io, &ctx->io_cap, npx * 16)) return rc;
    unsigned long long* dC = reinterpret_cast<unsigned long long*>(ctx->scratch + 128);
    CUDA_TRY(cudaMemsetAsync(dC, 0, 10 * sizeof(unsigned long long), ctx->stream));
    if (int rc = launchVolume(ctx, grid, cam, opts, cam->width, cam->height, static_cast<float4*>(ctx->io), dC)) return rc;
    unsigned long long h[10];
    CUDA_TRY(cudaMemcpyAsync(h, dC, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    out->rays = h[0]; out->root_probes = h[1]; out->upper_probes = h[2]; out->lower_probes = h[3]; out->voxel_probes = h[4];
    out->stencil_refills = h[5]; out->primary_samples = h[6]; out->shadow_samples = h[7]; out->shadow_rays = h[8]; out->hits = h[9];
    return VDBRT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// arbitrary-ray batches
// ---------------------------------------------------------------------------------------------------------------
int vdbrt_intersect_levelset(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso,
                             vdbrt_hit* hits, uint32_t memspace)
{
    return vdbrt_intersect_levelset_ex(ctx, grid, rays, n, space, iso, 0u, hits, memspace);
}

int vdbrt_intersect_levelset_ex(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso,
                                uint32_t iterations, vdbrt_hit* hits, uint32_t memspace)
{
    if (!ctx || !grid || (n && (!rays || !hits))) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (int rc = checkLevelSet(grid, iso)) return rc;
    if (n == 0) return VDBRT_OK;
    static_assert(sizeof(vdbrt_ray) == sizeof(RayIn) && sizeof(vdbrt_hit) == sizeof(HitOut), "POD mismatch");
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    const RayIn* dR = reinterpret_cast<const RayIn*>(rays);
    HitOut* dH = reinterpret_cast<HitOut*>(hits);
    const size_t rb = n * sizeof(RayIn), hb = n * sizeof(HitOut);
    if (memspace == VDBRT_MEM_HOST) {
        if (int rc = ensureBuffer(&ctx->io, &ctx->io_cap, rb + hb + 256)) return rc;
        uint8_t* b = static_cast<uint8_t*>(ctx->io);
        CUDA_TRY(cudaMemcpyAsync(b, rays, rb, cudaMemcpyHostToDevice, ctx->stream));
        dR = reinterpret_cast<const RayIn*>(b);
        dH = reinterpret_cast<HitOut*>(b + ((rb + 255) & ~size_t(255)));
    }
    const float vmin = iso - float(2 * grid->info.voxel_size[0]), vmax = iso + float(2 * grid->info.voxel_size[0]);
    const unsigned blocks = unsigned(std::min<uint64_t>((n + kBlockThreads - 1) / kBlockThreads, uint64_t(ctx->sm_count) * 16));
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    if (grid->leaf_kind == kLeafFp8) k_intersect_levelset<kLeafFp8><<<blocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dR, n, space, iso, vmin, vmax, dH, int(iterations));
    else if (grid->leaf_kind == kLeafFp16) k_intersect_levelset<kLeafFp16><<<blocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dR, n, space, iso, vmin, vmax, dH, int(iterations));
    else k_intersect_levelset<kLeafFloat><<<blocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dR, n, space, iso, vmin, vmax, dH, int(iterations));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->last_launches = 1;
    if (memspace == VDBRT_MEM_HOST) CUDA_TRY(cudaMemcpyAsync(hits, dH, hb, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VDBRT_OK;
}

int vdbrt_volume_spans(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_ray* rays, uint64_t n, uint32_t space, uint32_t maxSpans,
                       double* spans, int32_t* counts, uint32_t memspace)
{
    if (!ctx || !grid || (n && (!rays || !spans || !counts))) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (int rc = checkVolume(grid)) return rc;
    if (n == 0) return VDBRT_OK;
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    const RayIn* dR = reinterpret_cast<const RayIn*>(rays);
    double* dS = spans; int32_t* dC = counts;
    const size_t rb = (n * sizeof(RayIn) + 255) & ~size_t(255), sb = (n * maxSpans * 16 + 255) & ~size_t(255), cb = n * 4;
    if (memspace == VDBRT_MEM_HOST) {
        if (int rc = ensureBuffer(&ctx->io, &ctx->io_cap, rb + sb + cb)) return rc;
        uint8_t* b = static_cast<uint8_t*>(ctx->io);
        CUDA_TRY(cudaMemcpyAsync(b, rays, n * sizeof(RayIn), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(b + rb, 0, sb + cb, ctx->stream));
        dR = reinterpret_cast<const RayIn*>(b); dS = reinterpret_cast<double*>(b + rb); dC = reinterpret_cast<int32_t*>(b + rb + sb);
    }
    const unsigned blocks = unsigned(std::min<uint64_t>((n + kBlockThreads - 1) / kBlockThreads, uint64_t(ctx->sm_count) * 16));
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    k_volume_spans<<<blocks, kBlockThreads, 0, ctx->stream>>>(grid->dgrid, dR, n, space, maxSpans, dS, dC);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->last_launches = 1;
    if (memspace == VDBRT_MEM_HOST) {
        CUDA_TRY(cudaMemcpyAsync(spans, dS, n * maxSpans * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(counts, dC, cb, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VDBRT_OK;
}

int vdbrt_set_tuning(vdbrt_ctx* ctx, const char* key, uint32_t value)
{
    if (!ctx || !key) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    const std::string k(key);
    struct { const char* name; uint32_t* field; } table[] = {
        {"ls_strip", &ctx->ls_strip}, {"ls_strip_ratio", &ctx->ls_strip_ratio}, {"ls_refill", &ctx->ls_refill}, {"ls_eager", &ctx->ls_eager}, {"ls_affine", &ctx->ls_affine}, {"ls_order", &ctx->ls_order}, {"ls_history", &ctx->ls_history}, {"ls_hist_a", &ctx->ls_hist_a}, {"ls_hist_b", &ctx->ls_hist_b},
        {"ls_probe_cap", &ctx->ls_probe_cap}, {"ls_probe_b", &ctx->ls_probe_b}, {"ls_budget", &ctx->ls_budget}, {"ls_tail", &ctx->ls_tail}, {"ls_voxel_only", &ctx->ls_voxel_only}, {"ls_factor", &ctx->ls_factor},
        {"ls_rounds", &ctx->ls_rounds}, {"ls_dense", &ctx->ls_dense}, {"ls_dense_factor", &ctx->ls_dense_factor}, {"fog_wave", &ctx->fog_wave}, {"quant_native", &ctx->quant_native}, {"fog_refill", &ctx->fog_refill}, {"fog_rec_per_ray", &ctx->fog_rec_per_ray}, {"fog_cap_mb", &ctx->fog_cap_mb},
    };
    for (auto& t : table) if (k == t.name) {
        if (t.field == &ctx->ls_rounds && value > uint32_t(kMaxRounds)) value = kMaxRounds;
        *t.field = value;
        return VDBRT_OK;
    }
    if (k.size() == 10 && k.compare(0, 9, "ls_leaves") == 0 && k[9] >= '0' && k[9] < '0' + kMaxRounds) {      // leaf visits per ray in round r
        ctx->ls_leaves[k[9] - '0'] = value ? value : 1u;
        return VDBRT_OK;
    }
    return setError(VDBRT_ERR_INVALID_ARG, "unknown tuning key: " + k);
}

int vdbrt_volume_clip(vdbrt_ctx* ctx, const vdbrt_grid* grid, const vdbrt_ray* ray, uint32_t space, vdbrt_ray* clipped, int* hit, double scale[3])
{
    if (!ctx || !grid || !ray || !clipped || !hit) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (int rc = checkVolume(grid)) return rc;
    const DevGrid& g = grid->dgrid;
    // worldToIndex and clipRay of vdbrt_device.cuh restated for the host (this file is compiled without contraction; / and sqrt are IEEE)
    double ex = ray->eye[0], ey = ray->eye[1], ez = ray->eye[2], dx = ray->dir[0], dy = ray->dir[1], dz = ray->dir[2], t0 = ray->t0, t1 = ray->t1;
    if (space == VDBRT_SPACE_WORLD) {
        double jx, jy, jz;
        if (g.general) {
            const double qx = ex - g.trans[0], qy = ey - g.trans[1], qz = ez - g.trans[2];
            ex = qx * g.imat[0] + qy * g.imat[1] + qz * g.imat[2]; ey = qx * g.imat[3] + qy * g.imat[4] + qz * g.imat[5]; ez = qx * g.imat[6] + qy * g.imat[7] + qz * g.imat[8];
            jx = dx * g.imat[0] + dy * g.imat[1] + dz * g.imat[2]; jy = dx * g.imat[3] + dy * g.imat[4] + dz * g.imat[5]; jz = dx * g.imat[6] + dy * g.imat[7] + dz * g.imat[8];
        } else {
            if (g.has_translation) { ex = (ex - g.trans[0]) * g.inv[0]; ey = (ey - g.trans[1]) * g.inv[1]; ez = (ez - g.trans[2]) * g.inv[2]; }
            else { ex = ex * g.inv[0]; ey = ey * g.inv[1]; ez = ez * g.inv[2]; }
            jx = dx * g.inv[0]; jy = dy * g.inv[1]; jz = dz * g.inv[2];
        }
        const double len = std::sqrt(jx * jx + jy * jy + jz * jz);
        dx = jx / len; dy = jy / len; dz = jz / len;
        t0 = len * t0; t1 = len * t1;
    }
    const double e[3] = {ex, ey, ez}, inv[3] = {1 / dx, 1 / dy, 1 / dz};
    double a0 = t0, a1 = t1;
    *hit = 1;
    for (int k = 0; k < 3 && *hit; ++k) {
        double a = (g.bbox_min[k] - e[k]) * inv[k], b = ((g.bbox_max[k] + 1) - e[k]) * inv[k];
        if (a > b) { const double t = a; a = b; b = t; }
        if (a > a0) a0 = a;
        if (b < a1) a1 = b;
        if (a0 > a1) *hit = 0;
    }
    clipped->eye[0] = ex; clipped->eye[1] = ey; clipped->eye[2] = ez; clipped->dir[0] = dx; clipped->dir[1] = dy; clipped->dir[2] = dz;
    clipped->t0 = *hit ? a0 : t0; clipped->t1 = *hit ? a1 : t1;
    if (scale) for (int k = 0; k < 3; ++k) scale[k] = g.scale[k];
    return VDBRT_OK;
}

int vdbrt_last_kernel_ms(vdbrt_ctx* ctx, float* ms, uint32_t* launches)
{
    if (!ctx) return setError(VDBRT_ERR_INVALID_ARG, "null context");
    DeviceGuard guard(ctx->device);
    if (ms) { CUDA_TRY(cudaEventSynchronize(ctx->ev1)); CUDA_TRY(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1)); }
    if (ms && ctx->evp && std::getenv("VDBRT_TIME_PROBE")) {      // diagnostics: the probe launch alone (only valid right after an ordered level-set render)
        float pm = 0.f;
        if (cudaEventElapsedTime(&pm, ctx->ev0, ctx->evp) == cudaSuccess) std::fprintf(stderr, "[vdbrt] probe %.3f ms of %.3f ms\n", pm, *ms); else cudaGetLastError();
    }
    if (launches) *launches = ctx->last_launches;
    return VDBRT_OK;
}

} // extern "C"
