// vdbrt_host.h -- host-side state shared by the translation units of libvdbrt.so
#pragma once
#include "../../include/vdbrt.h"
#include "vdbrt_device.cuh"
#include <cuda_runtime.h>
#include <mutex>
#include <string>

struct vdbrt_ctx {
    int device = 0, sm_count = 0;
    std::mutex mx;                              // serialises the entry points that use the context's queue word, counters and staging buffers
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp = nullptr;   // bracket the kernel(s) of the last call on `stream`
    uint32_t last_launches = 0;
    uint8_t* scratch = nullptr;                 // 4 KB: [0,64) bbox reduction, [64,128) work queue, [128,..) counters
    void* film = nullptr;  size_t film_cap = 0; // device staging for host films
    void* aux = nullptr;   size_t aux_cap = 0;  // device staging for per-pixel records
    void* io = nullptr;    size_t io_cap = 0;   // device staging for ray batches / scratch films
    void* lng = nullptr;   size_t lng_cap = 0;  // long-ray records, segment lists and their control block (vdbrt_kernels.cuh)
    void* hist = nullptr;  size_t hist_cap = 0; // per-tile costs of the last level-set frame (Sched::cost_out) and what they belong to
    uint32_t hist_tm[16] = {}; const void* hist_grid = nullptr; uint32_t hist_spp = 0, hist_valid = 0, ls_history = 1, ls_hist_a = 250, ls_hist_b = 105;
    void* ord = nullptr;   size_t ord_cap = 0;  // tile-ordering buffers of the level-set render (OrderBufs, vdbrt_kernels.cuh)
    uint32_t ls_strip = 1, ls_strip_ratio = 4, ls_refill = 32, ls_eager = 0, ls_affine = 0, ls_order = 0, ls_probe_cap = 128, ls_probe_b = 64;   // Sched (vdbrt_kernels.cuh)
    void* fog = nullptr;   size_t fog_cap = 0;  // records / ray table of the fog wavefront (vdbrt_fog.cuh)
    uint32_t quant_native = 1;                  // NanoGrid<Fp8|Fp16> rendered as they are (their own kernel instantiations) instead of expanded to float leaves
    uint32_t fog_wave = 1, fog_refill = 8, fog_rec_per_ray = 12, fog_cap_mb = 4096;
    uint32_t ls_voxel_only = 0;                 // tail rule: only rays that are marching voxels are suspended
    uint32_t ls_dense_factor = 100;             // percent, see launchLevelSet
    unsigned long long* hist_host = nullptr;    // pinned: {sum, max} of the previous frame's tile costs, copied back asynchronously (a hint, read without a sync)
    uint32_t ls_dense = 800;                    // launches with at least this many 8x4 tiles per SM use the 6-CTA instantiation (0: never; kDenseMinTilesPerSm)
    uint32_t ls_tail = 0;                       // tail rule: iterations a tile may still spend once the work queue has run dry (0: per-tile rule below)
    uint32_t ls_budget = 0;                     // warp iterations a tile may spend before its running rays are suspended (0 = never)
    uint32_t ls_factor = 0;                     // ... or this many percent of a warp's share of the launch, if that is more
    uint32_t ls_rounds = 0;                     // long-ray rounds per frame
    uint32_t ls_leaves[8] = {};                 // leaf visits per ray in round r
};

struct vdbrt_grid {
    uint8_t* dev = nullptr;                     // serialised NanoGrid<float> in device memory
    uint64_t bytes = 0;
    int device = 0;
    vdbrt_grid_info info;
    vdbrt::DevGrid dgrid;
    float* halo = nullptr;                      // DevGrid::halo (9^3 values per leaf)
    unsigned long long* lowmask = nullptr;      // DevGrid::lowmask (512 B per lower node)
    int leaf_kind = 0;                          // vdbrt::kLeafFloat / kLeafFp8 / kLeafFp16: how the kernels read this grid's leaves
    bool is_color = false;                      // a NanoGrid<Vec3f> for the colour-grid shaders (dcolor instead of dgrid)
    vdbrt::DevColor dcolor;
};

namespace vdbrt {
int setError(int code, const std::string& msg);
int cudaFail(cudaError_t e, const char* what);
// parse the header of grid->dev, validate it, fill info/dgrid and compute the node-granular bbox on the device
int finishGrid(vdbrt_ctx* ctx, vdbrt_grid* grid);
// cudaFree of everything a grid owns on the device + delete
void destroyGrid(vdbrt_grid* grid);
// quantised grids (vdbrt_quant.cu): NanoGrid<Fp4|Fp8|Fp16|FpN> in device memory -> freshly allocated NanoGrid<float>
bool isQuantisedType(uint32_t gridType);
int expandQuantised(vdbrt_ctx* ctx, const uint8_t* src, uint64_t srcBytes, const uint8_t* head, uint8_t** outDev, uint64_t* outBytes);
}
