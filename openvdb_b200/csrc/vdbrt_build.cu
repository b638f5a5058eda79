// vdbrt_build.cu -- GPU construction of narrow-band level sets (and their fog volumes) straight into the NanoVDB
// layout, so that benches can create the BASELINE.json grids on the GPU box without the reference.
//
// Same voxel values, active states and tree topology as the reference generators (checked voxel by voxel in
// tests/test_gpu_builders.py through nanovdb::tools::nanoToOpenVDB):
//   sphere : openvdb::tools::createLevelSetSphere / nanovdb initSphere  (tools/LevelSetSphere.h:126-176,
//            nanovdb/tools/CreatePrimitives.h:598-664): v = sqrt(x2y2 + (k-cz)^2) - r0 in FLOAT, active iff |v| < w,
//            value dx*v; everything else +-background by sign (signed flood fill == analytic sign for closed shapes)
//   torus  : nanovdb initTorus (nanovdb/tools/CreatePrimitives.h:666-738)
//   union  : voxel-wise min of spheres with the state of the minimum (openvdb::tools::csgUnion, tools/Composite.h:886)
//   fog    : openvdb::tools::sdfToFogVolume (tools/LevelSetUtil.h:2190, SDFVoxelsToFogVolume :477-508)
// A node exists iff it contains a band voxel, exactly what setValue-on-demand + pruneLevelSet leave behind.
//
// Pipeline: (1) conservative cull of 128^3 blocks by the distance at the block centre, (2) one CTA per surviving
// block classifies its 4096 leaf slots (centre cull, then exact band test) into a child mask, (3) tiny host scan
// over blocks -> node indices, (4) fill kernels write upper / lower / leaf nodes in place.  No sort, no atomics on
// voxel data; HBM traffic is essentially the bytes of the grid written once.
#include "vdbrt_host.h"
#include <mutex>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

using namespace vdbrt;

#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cudaFail(e_, #expr); } while (0)

namespace {

constexpr uint64_t kGridBytes = 672, kTreeBytes = 64, kRootBytes = 64, kUpperBytes = 270400, kLowerBytes = 33856, kLeafBytes = 2144;

// ---- primitives: eval() is the reference's float formula (voxel units); dist() a double distance for culling ----
struct PrimSphere {
    float cx, cy, cz, r0;
    __device__ float eval(int i, int j, int k, uint32_t) const {
        const float a = float(i) - cx, x2 = a * a;
        const float b = float(j) - cy, x2y2 = b * b + x2;
        const float c = float(k) - cz;
        return sqrtf(x2y2 + c * c) - r0;
    }
    __device__ double dist(double x, double y, double z, uint32_t) const {
        const double a = x - cx, b = y - cy, c = z - cz;
        return sqrt(a * a + b * b + c * c) - double(r0);
    }
};
struct PrimTorus {
    float cx, cy, cz, r1, r2;
    __device__ float eval(int i, int j, int k, uint32_t) const {
        const float a = float(i) - cx, x2 = a * a;
        const float c = float(k) - cz;
        const float q = sqrtf(c * c + x2) - r1, x2z2 = q * q;
        const float b = float(j) - cy;
        return sqrtf(x2z2 + b * b) - r2;
    }
    __device__ double dist(double x, double y, double z, uint32_t) const {
        const double a = x - cx, b = y - cy, c = z - cz;
        const double q = sqrt(a * a + c * c) - double(r1);
        return sqrt(q * q + b * b) - double(r2);
    }
};
// union of spheres: every 128^3 block carries the list of spheres whose band can reach it
struct PrimSpheres {
    const float4* spheres;         // cx,cy,cz,r0 in voxel units (float, as the reference rasterises each sphere)
    const uint32_t* start;         // [nBlocks+1] into list, indexed by dense block id
    const uint32_t* list;
    __device__ float eval(int i, int j, int k, uint32_t blk) const {
        float m = 3.0e38f;
        for (uint32_t q = start[blk]; q < start[blk + 1]; ++q) {
            const float4 s = spheres[list[q]];
            const float a = float(i) - s.x, x2 = a * a;
            const float b = float(j) - s.y, x2y2 = b * b + x2;
            const float c = float(k) - s.z;
            m = fminf(m, sqrtf(x2y2 + c * c) - s.w);
        }
        return m;
    }
    __device__ double dist(double x, double y, double z, uint32_t blk) const {
        double m = 1e300;
        for (uint32_t q = start[blk]; q < start[blk + 1]; ++q) {
            const float4 s = spheres[list[q]];
            const double a = x - s.x, b = y - s.y, c = z - s.z;
            m = fmin(m, sqrt(a * a + b * b + c * c) - double(s.w));
        }
        return m;
    }
};

struct BlockGrid { int bx0, by0, bz0, nbx, nby, nbz; };   // dense box of 128^3 blocks covering the band
__host__ __device__ inline void blockCoord(const BlockGrid& g, uint32_t id, int& x, int& y, int& z)
{
    const int ix = int(id / uint32_t(g.nby * g.nbz)), r = int(id % uint32_t(g.nby * g.nbz));
    x = (g.bx0 + ix) * 128; y = (g.by0 + r / g.nbz) * 128; z = (g.bz0 + r % g.nbz) * 128;
}

// (1) conservative cull of 128^3 blocks
template<class P>
__global__ void k_cull_blocks(P prim, BlockGrid bg, uint32_t nBlocks, float hw, uint8_t* flag)
{
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nBlocks) return;
    int x, y, z; blockCoord(bg, id, x, y, z);
    const double d = prim.dist(x + 63.5, y + 63.5, z + 63.5, id);
    flag[id] = fabs(d) <= double(hw) + 110.0 + 1.0 ? 1 : 0;        // 63.5*sqrt(3) = 109.99
}

// (2) one CTA per candidate block: child mask of its 4096 leaf slots + exclusive prefix of popcounts per word
template<class P>
__global__ void __launch_bounds__(256)
k_classify_leaves(P prim, BlockGrid bg, const uint32_t* cand, float hw, unsigned long long* masks, uint16_t* prefix, uint32_t* leafCount)
{
    __shared__ unsigned long long smask[64];
    const uint32_t c = blockIdx.x, id = cand[c];
    if (threadIdx.x < 64) smask[threadIdx.x] = 0ull;
    __syncthreads();
    int ox, oy, oz; blockCoord(bg, id, ox, oy, oz);
    for (uint32_t m = threadIdx.x; m < 4096; m += 256) {
        const int x = ox + int((m >> 8) << 3), y = oy + int(((m >> 4) & 15u) << 3), z = oz + int((m & 15u) << 3);
        const double dc = prim.dist(x + 3.5, y + 3.5, z + 3.5, id);
        if (fabs(dc) > double(hw) + 6.07 + 0.05) continue;           // 3.5*sqrt(3) = 6.06
        bool any = false;
        for (int n = 0; n < 512 && !any; ++n) {
            const float v = prim.eval(x + (n >> 6), y + ((n >> 3) & 7), z + (n & 7), id);
            any = fabsf(v) < hw;
        }
        if (any) atomicOr(&smask[m >> 6], 1ull << (m & 63u));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int w = 0; w < 64; ++w) { prefix[size_t(c) * 64 + w] = uint16_t(run); run += __popcll(smask[w]); }
        leafCount[c] = run;
    }
    if (threadIdx.x < 64) masks[size_t(c) * 64 + threadIdx.x] = smask[threadIdx.x];
}

struct LowerDesc { uint32_t cand, blockId, leafBase, upper, slot; int ox, oy, oz; };
struct UpperDesc { int ox, oy, oz; };
struct Layout { uint64_t rootOff, upperOff, lowerOff, leafOff; float bg, dx; float hw; };

__device__ __forceinline__ float signedBg(float v, float bg) { return v < 0.f ? -bg : bg; }

// (4a) upper nodes: one thread per slot; childOfUpper[u*32768+n] = lower index or -1
template<class P>
__global__ void __launch_bounds__(256)
k_fill_upper(P prim, Layout L, const UpperDesc* uppers, const int* childOfUpper, const uint32_t* blockOfSlot, uint8_t* base)
{
    const uint32_t u = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    uint8_t* node = base + L.upperOff + uint64_t(u) * kUpperBytes;
    const UpperDesc ud = uppers[u];
    const int child = childOfUpper[size_t(u) * 32768 + n];
    const unsigned bit = __ballot_sync(0xffffffffu, child >= 0);
    if ((threadIdx.x & 31) == 0) {
        reinterpret_cast<uint32_t*>(node + kUpperCMask)[n >> 5] = bit;
        reinterpret_cast<uint32_t*>(node + kUpperVMask)[n >> 5] = 0u;
    }
    unsigned long long entry;
    if (child >= 0) entry = (unsigned long long)((long long)(L.lowerOff + uint64_t(child) * kLowerBytes) - (long long)(L.upperOff + uint64_t(u) * kUpperBytes));
    else {
        const int x = ud.ox + int((n >> 10) << 7), y = ud.oy + int(((n >> 5) & 31u) << 7), z = ud.oz + int((n & 31u) << 7);
        entry = __float_as_uint(signedBg(prim.eval(x, y, z, blockOfSlot[size_t(u) * 32768 + n]), L.bg));
    }
    reinterpret_cast<unsigned long long*>(node + kUpperTable)[n] = entry;
    if (n < 8) {   // header: bbox (node aligned), flags, then min/max/avg/dev = 0
        if (n == 0) { int* b = reinterpret_cast<int*>(node); b[0] = ud.ox; b[1] = ud.oy; b[2] = ud.oz; b[3] = ud.ox + 4095; b[4] = ud.oy + 4095; b[5] = ud.oz + 4095; b[6] = 0; b[7] = 0; }
        if (n < 4) reinterpret_cast<float*>(node + kUpperCMask + 4096)[n] = 0.f;
    }
}

// (4b) lower nodes: one thread per slot; also records the origin of every leaf
template<class P>
__global__ void __launch_bounds__(256)
k_fill_lower(P prim, Layout L, const LowerDesc* lowers, const unsigned long long* masks, const uint16_t* prefix, uint8_t* base, int4* leafOrigin)
{
    const uint32_t l = blockIdx.y, m = blockIdx.x * 256 + threadIdx.x;
    const LowerDesc ld = lowers[l];
    uint8_t* node = base + L.lowerOff + uint64_t(l) * kLowerBytes;
    const unsigned long long word = masks[size_t(ld.cand) * 64 + (m >> 6)];
    const bool child = (word >> (m & 63u)) & 1ull;
    if ((m & 63u) == 0) {
        reinterpret_cast<unsigned long long*>(node + kLowerCMask)[m >> 6] = word;
        reinterpret_cast<unsigned long long*>(node + kLowerVMask)[m >> 6] = 0ull;
    }
    const int x = ld.ox + int((m >> 8) << 3), y = ld.oy + int(((m >> 4) & 15u) << 3), z = ld.oz + int((m & 15u) << 3);
    unsigned long long entry;
    if (child) {
        const uint32_t leaf = ld.leafBase + prefix[size_t(ld.cand) * 64 + (m >> 6)] + __popcll(word & ((1ull << (m & 63u)) - 1ull));
        entry = (unsigned long long)((long long)(L.leafOff + uint64_t(leaf) * kLeafBytes) - (long long)(L.lowerOff + uint64_t(l) * kLowerBytes));
        leafOrigin[leaf] = make_int4(x, y, z, int(ld.blockId));
    } else {
        entry = __float_as_uint(signedBg(prim.eval(x, y, z, ld.blockId), L.bg));
    }
    reinterpret_cast<unsigned long long*>(node + kLowerTable)[m] = entry;
    if (m == 0) { int* b = reinterpret_cast<int*>(node); b[0] = ld.ox; b[1] = ld.oy; b[2] = ld.oz; b[3] = ld.ox + 127; b[4] = ld.oy + 127; b[5] = ld.oz + 127; b[6] = 0; b[7] = 0; }
    if (m < 4) reinterpret_cast<float*>(node + kLowerCMask + 512)[m] = 0.f;
}

// (4c) leaves: one CTA of 512 threads per leaf, thread n <-> voxel n (z fastest)
// stats[0..5] index bbox (atomic min/max), stats64[0] active voxel count
template<class P>
__global__ void __launch_bounds__(512)
k_fill_leaves(P prim, Layout L, const int4* leafOrigin, uint8_t* base, int* bbox, unsigned long long* voxelCount)
{
    __shared__ int smin[3], smax[3];
    __shared__ unsigned int scount;
    const uint32_t leaf = blockIdx.x, n = threadIdx.x;
    const int4 o = leafOrigin[leaf];
    uint8_t* node = base + L.leafOff + uint64_t(leaf) * kLeafBytes;
    if (n < 3) { smin[n] = 7; smax[n] = 0; }
    if (n == 0) scount = 0;
    __syncthreads();
    const int li = int(n >> 6), lj = int((n >> 3) & 7u), lk = int(n & 7u);
    const float v = prim.eval(o.x + li, o.y + lj, o.z + lk, uint32_t(o.w));
    const bool active = fabsf(v) < L.hw;
    reinterpret_cast<float*>(node + kLeafValues)[n] = active ? L.dx * v : signedBg(v, L.bg);
    const unsigned bits = __ballot_sync(0xffffffffu, active);
    if ((n & 31u) == 0) {
        reinterpret_cast<uint32_t*>(node + kLeafVMask)[n >> 5] = bits;
        if (bits) atomicAdd(&scount, __popc(bits));
    }
    if (active) {
        atomicMin(&smin[0], li); atomicMin(&smin[1], lj); atomicMin(&smin[2], lk);
        atomicMax(&smax[0], li); atomicMax(&smax[1], lj); atomicMax(&smax[2], lk);
    }
    __syncthreads();
    if (n == 0) {
        int* h = reinterpret_cast<int*>(node);
        h[0] = o.x + smin[0]; h[1] = o.y + smin[1]; h[2] = o.z + smin[2];                  // mBBoxMin
        node[12] = uint8_t(smax[0] - smin[0]); node[13] = uint8_t(smax[1] - smin[1]); node[14] = uint8_t(smax[2] - smin[2]);
        node[15] = 2;                                                                     // has bbox
        float* st = reinterpret_cast<float*>(node + 80); st[0] = st[1] = st[2] = st[3] = 0.f;
        atomicMin(bbox + 0, o.x + smin[0]); atomicMin(bbox + 1, o.y + smin[1]); atomicMin(bbox + 2, o.z + smin[2]);
        atomicMax(bbox + 3, o.x + smax[0]); atomicMax(bbox + 4, o.y + smax[1]); atomicMax(bbox + 5, o.z + smax[2]);
        atomicAdd(voxelCount, (unsigned long long)scount);
    }
}

// ---- header blocks written on the host -------------------------------------------------------------------------
template<typename T> void wr(uint8_t* p, T v) { std::memcpy(p, &v, sizeof(T)); }

void writeGridHeader(uint8_t* h, uint64_t gridSize, const char* name, double dx, const double t[3], uint32_t gridClass,
                     const int bbox[6], uint64_t rootOff, uint64_t upperOff, uint64_t lowerOff, uint64_t leafOff,
                     uint32_t nLeaf, uint32_t nLower, uint32_t nUpper, const uint32_t tileCount[3], uint64_t voxels)
{
    std::memset(h, 0, kGridBytes + kTreeBytes);
    wr<uint64_t>(h + 0, 0x314244566f6e614eULL);            // NANOVDB_MAGIC_GRID
    wr<uint64_t>(h + 8, ~uint64_t(0));                     // checksum disabled
    wr<uint32_t>(h + 16, (32u << 21) | (9u << 10) | 2u);   // version 32.9.2
    wr<uint32_t>(h + 20, (1u << 5) | (1u << 1));           // IsBreadthFirst | HasBBox
    wr<uint32_t>(h + 24, 0); wr<uint32_t>(h + 28, 1);
    wr<uint64_t>(h + 32, gridSize);
    std::strncpy(reinterpret_cast<char*>(h + 40), name, 255);
    uint8_t* map = h + 296;                                // nanovdb::Map(s, t) (NanoVDB.h:1443-1453)
    const float sf = float(dx);
    for (int i = 0; i < 3; ++i) {
        wr<float>(map + 4 * (4 * i), sf); wr<float>(map + 36 + 4 * (4 * i), 1.0f / sf); wr<float>(map + 72 + 4 * i, float(t[i]));
        wr<double>(map + 88 + 8 * (4 * i), dx); wr<double>(map + 160 + 8 * (4 * i), 1.0 / dx); wr<double>(map + 232 + 8 * i, t[i]);
    }
    wr<float>(map + 84, 1.0f); wr<double>(map + 256, 1.0);
    for (int i = 0; i < 3; ++i) {
        wr<double>(h + 560 + 8 * i, bbox[i] * dx + t[i]); wr<double>(h + 584 + 8 * i, (bbox[3 + i] + 1) * dx + t[i]);
        wr<double>(h + 608 + 8 * i, dx);
    }
    wr<uint32_t>(h + 632, gridClass); wr<uint32_t>(h + 636, 1);   // GridType::Float
    wr<int64_t>(h + 640, int64_t(gridSize)); wr<uint32_t>(h + 648, 0);
    uint8_t* tr = h + kGridBytes;                          // TreeData
    wr<int64_t>(tr + 0, int64_t(leafOff - kGridBytes)); wr<int64_t>(tr + 8, int64_t(lowerOff - kGridBytes));
    wr<int64_t>(tr + 16, int64_t(upperOff - kGridBytes)); wr<int64_t>(tr + 24, int64_t(rootOff - kGridBytes));
    wr<uint32_t>(tr + 32, nLeaf); wr<uint32_t>(tr + 36, nLower); wr<uint32_t>(tr + 40, nUpper);
    wr<uint32_t>(tr + 44, tileCount[0]); wr<uint32_t>(tr + 48, tileCount[1]); wr<uint32_t>(tr + 52, tileCount[2]);
    wr<uint64_t>(tr + 56, voxels);
}

uint64_t rootKeyHost(int x, int y, int z)
{
    return uint64_t(uint32_t(z) >> 12) | (uint64_t(uint32_t(y) >> 12) << 21) | (uint64_t(uint32_t(x) >> 12) << 42);
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template<typename T> T* as() { return static_cast<T*>(p); }
};

template<class P>
int buildLevelSet(vdbrt_ctx* ctx, const P& prim, const int lo[3], const int hi[3], double dx, double hwVox, const char* name,
                  const uint32_t* blockStart /*host copy for PrimSpheres or null*/, BlockGrid bg, vdbrt_grid** out)
{
    cudaStream_t st = ctx->stream;
    const uint32_t nBlocks = uint32_t(bg.nbx) * bg.nby * bg.nbz;
    const float hw = float(hwVox), bgVal = float(hwVox * dx);
    // (1) cull blocks
    DevBuf dFlag; CUDA_TRY(cudaMalloc(&dFlag.p, nBlocks));
    k_cull_blocks<<<(nBlocks + 255) / 256, 256, 0, st>>>(prim, bg, nBlocks, hw, dFlag.as<uint8_t>());
    CUDA_TRY(cudaGetLastError());
    std::vector<uint8_t> hFlag(nBlocks);
    CUDA_TRY(cudaMemcpyAsync(hFlag.data(), dFlag.p, nBlocks, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<uint32_t> cand;
    for (uint32_t i = 0; i < nBlocks; ++i) if (hFlag[i] && (!blockStart || blockStart[i + 1] > blockStart[i])) cand.push_back(i);
    if (cand.empty()) return setError(VDBRT_ERR_EMPTY_GRID, "primitive produces no narrow-band voxels");
    // (2) classify leaves
    const uint32_t nCand = uint32_t(cand.size());
    DevBuf dCand, dMasks, dPrefix, dCount;
    CUDA_TRY(cudaMalloc(&dCand.p, nCand * 4)); CUDA_TRY(cudaMalloc(&dMasks.p, size_t(nCand) * 64 * 8));
    CUDA_TRY(cudaMalloc(&dPrefix.p, size_t(nCand) * 64 * 2)); CUDA_TRY(cudaMalloc(&dCount.p, nCand * 4));
    CUDA_TRY(cudaMemcpyAsync(dCand.p, cand.data(), nCand * 4, cudaMemcpyHostToDevice, st));
    k_classify_leaves<<<nCand, 256, 0, st>>>(prim, bg, dCand.as<uint32_t>(), hw, dMasks.as<unsigned long long>(), dPrefix.as<uint16_t>(), dCount.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> hCount(nCand);
    CUDA_TRY(cudaMemcpyAsync(hCount.data(), dCount.p, nCand * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // (3) host scan: lowers ordered by (upper origin, slot), uppers by origin
    std::map<std::tuple<int, int, int>, uint32_t> upperIndex;
    std::vector<LowerDesc> lowers;
    for (uint32_t c = 0; c < nCand; ++c) {
        if (!hCount[c]) continue;
        LowerDesc d; d.cand = c; d.blockId = cand[c]; d.leafBase = 0; d.upper = 0;
        blockCoord(bg, cand[c], d.ox, d.oy, d.oz);
        d.slot = (uint32_t((d.ox & 4095) >> 7) << 10) | (uint32_t((d.oy & 4095) >> 7) << 5) | uint32_t((d.oz & 4095) >> 7);
        upperIndex[std::make_tuple(d.ox & ~4095, d.oy & ~4095, d.oz & ~4095)] = 0;
        lowers.push_back(d);
    }
    if (lowers.empty()) return setError(VDBRT_ERR_EMPTY_GRID, "primitive produces no narrow-band voxels");
    std::vector<UpperDesc> uppers;
    for (auto& kv : upperIndex) { kv.second = uint32_t(uppers.size()); uppers.push_back(UpperDesc{std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first)}); }
    for (auto& d : lowers) d.upper = upperIndex[std::make_tuple(d.ox & ~4095, d.oy & ~4095, d.oz & ~4095)];
    std::sort(lowers.begin(), lowers.end(), [](const LowerDesc& a, const LowerDesc& b) { return a.upper != b.upper ? a.upper < b.upper : a.slot < b.slot; });
    uint64_t nLeaf64 = 0;
    for (auto& d : lowers) { d.leafBase = uint32_t(nLeaf64); nLeaf64 += hCount[d.cand]; }
    if (nLeaf64 > 0xffffffffull) return setError(VDBRT_ERR_UNSUPPORTED, "too many leaf nodes");
    const uint32_t nLeaf = uint32_t(nLeaf64), nLower = uint32_t(lowers.size()), nUpper = uint32_t(uppers.size());
    std::vector<int> childOfUpper(size_t(nUpper) * 32768, -1);
    std::vector<uint32_t> blockOfSlot(size_t(nUpper) * 32768, 0);
    for (uint32_t l = 0; l < nLower; ++l) childOfUpper[size_t(lowers[l].upper) * 32768 + lowers[l].slot] = int(l);
    // block id (for per-block sphere lists) of every upper slot: slots outside the dense block box have no spheres -> use a
    // sentinel block with an empty list (last entry); single primitives ignore it
    for (uint32_t u = 0; u < nUpper; ++u) for (uint32_t n = 0; n < 32768; ++n) {
        const int bx = ((uppers[u].ox + int((n >> 10) << 7)) >> 7) - bg.bx0, by = ((uppers[u].oy + int(((n >> 5) & 31u) << 7)) >> 7) - bg.by0,
                  bz = ((uppers[u].oz + int((n & 31u) << 7)) >> 7) - bg.bz0;
        blockOfSlot[size_t(u) * 32768 + n] = (bx >= 0 && by >= 0 && bz >= 0 && bx < bg.nbx && by < bg.nby && bz < bg.nbz)
                                                 ? uint32_t((bx * bg.nby + by) * bg.nbz + bz) : nBlocks;
    }
    Layout L;
    L.rootOff = kGridBytes + kTreeBytes;
    L.upperOff = (L.rootOff + kRootBytes + 32ull * nUpper + 31) & ~31ull;
    L.lowerOff = L.upperOff + kUpperBytes * nUpper;
    L.leafOff = L.lowerOff + kLowerBytes * nLower;
    L.bg = bgVal; L.dx = float(dx); L.hw = hw;
    const uint64_t total = L.leafOff + kLeafBytes * nLeaf;
    // (4) allocate + fill
    auto* g = new vdbrt_grid;
    g->bytes = total; g->device = ctx->device;
    cudaError_t e = cudaMalloc(&g->dev, total);
    if (e != cudaSuccess) { delete g; return cudaFail(e, "cudaMalloc(grid)"); }
    auto failGrid = [&](int rc) { destroyGrid(g); return rc; };
    cudaMemsetAsync(g->dev, 0, total, st);                    // padding bytes are defined: the buffer is reproducible byte for byte
    DevBuf dLowers, dUppers, dChild, dBlockOf, dLeafOrg, dStats;
    if (cudaMalloc(&dLowers.p, nLower * sizeof(LowerDesc)) != cudaSuccess || cudaMalloc(&dUppers.p, nUpper * sizeof(UpperDesc)) != cudaSuccess ||
        cudaMalloc(&dChild.p, childOfUpper.size() * 4) != cudaSuccess || cudaMalloc(&dBlockOf.p, blockOfSlot.size() * 4) != cudaSuccess ||
        cudaMalloc(&dLeafOrg.p, size_t(nLeaf) * 16) != cudaSuccess || cudaMalloc(&dStats.p, 64) != cudaSuccess)
        return failGrid(setError(VDBRT_ERR_NOMEM, "out of device memory while building the grid"));
    cudaMemcpyAsync(dLowers.p, lowers.data(), nLower * sizeof(LowerDesc), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dUppers.p, uppers.data(), nUpper * sizeof(UpperDesc), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dChild.p, childOfUpper.data(), childOfUpper.size() * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dBlockOf.p, blockOfSlot.data(), blockOfSlot.size() * 4, cudaMemcpyHostToDevice, st);
    int initBox[8] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0, 0};
    cudaMemcpyAsync(dStats.p, initBox, 32, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(static_cast<uint8_t*>(dStats.p) + 32, 0, 32, st);
    k_fill_upper<<<dim3(128, nUpper), 256, 0, st>>>(prim, L, dUppers.as<UpperDesc>(), dChild.as<int>(), dBlockOf.as<uint32_t>(), g->dev);
    k_fill_lower<<<dim3(16, nLower), 256, 0, st>>>(prim, L, dLowers.as<LowerDesc>(), dMasks.as<unsigned long long>(), dPrefix.as<uint16_t>(), g->dev, dLeafOrg.as<int4>());
    k_fill_leaves<<<nLeaf, 512, 0, st>>>(prim, L, dLeafOrg.as<int4>(), g->dev, dStats.as<int>(), reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(dStats.p) + 32));
    e = cudaGetLastError();
    if (e != cudaSuccess) return failGrid(cudaFail(e, "grid fill kernels"));
    int hBox[8]; unsigned long long voxels = 0;
    cudaMemcpyAsync(hBox, dStats.p, 32, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&voxels, static_cast<uint8_t*>(dStats.p) + 32, 8, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return failGrid(cudaFail(e, "grid fill kernels"));
    // header + root on the host
    std::vector<uint8_t> head(L.upperOff, 0);
    const double t0[3] = {0, 0, 0};
    const uint32_t tiles[3] = {0, 0, 0};
    writeGridHeader(head.data(), total, name, dx, t0, VDBRT_GRID_CLASS_LEVEL_SET, hBox, L.rootOff, L.upperOff, L.lowerOff, L.leafOff, nLeaf, nLower, nUpper, tiles, voxels);
    uint8_t* root = head.data() + L.rootOff;
    for (int i = 0; i < 6; ++i) wr<int32_t>(root + 4 * i, hBox[i]);
    wr<uint32_t>(root + 24, nUpper); wr<float>(root + 28, bgVal);
    // root tiles sorted by key, as NanoVDB stores them
    std::vector<std::pair<uint64_t, uint32_t>> keys;
    for (uint32_t u = 0; u < nUpper; ++u) keys.push_back({rootKeyHost(uppers[u].ox, uppers[u].oy, uppers[u].oz), u});
    std::sort(keys.begin(), keys.end());
    for (uint32_t i = 0; i < nUpper; ++i) {
        uint8_t* t = root + kRootBytes + 32 * i;
        wr<uint64_t>(t, keys[i].first);
        wr<int64_t>(t + 8, int64_t(L.upperOff + kUpperBytes * keys[i].second) - int64_t(L.rootOff));
        wr<uint32_t>(t + 16, 0); wr<float>(t + 20, bgVal);
    }
    e = cudaMemcpyAsync(g->dev, head.data(), head.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return failGrid(cudaFail(e, "cudaMemcpyAsync(header)"));
    const int rc = finishGrid(ctx, g);
    if (rc != VDBRT_OK) return failGrid(rc);
    *out = g;
    return VDBRT_OK;
}

BlockGrid blockGridFor(const int lo[3], const int hi[3])
{
    BlockGrid bg;
    bg.bx0 = lo[0] >> 7; bg.by0 = lo[1] >> 7; bg.bz0 = lo[2] >> 7;
    bg.nbx = (hi[0] >> 7) - bg.bx0 + 1; bg.nby = (hi[1] >> 7) - bg.by0 + 1; bg.nbz = (hi[2] >> 7) - bg.bz0 + 1;
    return bg;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

} // namespace

extern "C" {

int vdbrt_build_levelset_sphere(vdbrt_ctx* ctx, double radius, const double center[3], double voxelSize, double halfWidth, vdbrt_grid** out)
{
    if (!ctx || !center || !out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (!(radius > 0) || !(voxelSize > 0) || !(halfWidth > 1)) return setError(VDBRT_ERR_INVALID_ARG, "radius and voxel size must be positive, half-width > 1");
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    // float set-up exactly as LevelSetSphere::rasterSphere (tools/LevelSetSphere.h:133-145)
    const float dx = float(voxelSize), w = float(halfWidth);
    PrimSphere p;
    p.r0 = float(radius) / dx;
    if (p.r0 < 1.5f) return setError(VDBRT_ERR_EMPTY_GRID, "radius below the Nyquist frequency");
    p.cx = float(center[0]) / dx; p.cy = float(center[1]) / dx; p.cz = float(center[2]) / dx;
    const float rmax = p.r0 + w;
    const int lo[3] = {int(std::floor(p.cx - rmax)), int(std::floor(p.cy - rmax)), int(std::floor(p.cz - rmax))};
    const int hi[3] = {int(std::ceil(p.cx + rmax)), int(std::ceil(p.cy + rmax)), int(std::ceil(p.cz + rmax))};
    return buildLevelSet(ctx, p, lo, hi, double(dx), double(w), "sphere_ls", nullptr, blockGridFor(lo, hi), out);
}

int vdbrt_build_levelset_torus(vdbrt_ctx* ctx, double majorRadius, double minorRadius, const double center[3], double voxelSize, double halfWidth, vdbrt_grid** out)
{
    if (!ctx || !center || !out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (!(minorRadius > 0) || !(majorRadius > minorRadius) || !(voxelSize > 0) || !(halfWidth > 0)) return setError(VDBRT_ERR_INVALID_ARG, "bad torus parameters");
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    // nanovdb initTorus (nanovdb/tools/CreatePrimitives.h:689-705)
    PrimTorus p;
    p.r1 = float(majorRadius / voxelSize); p.r2 = float(minorRadius / voxelSize);
    if (p.r2 < 1.5f) return setError(VDBRT_ERR_EMPTY_GRID, "radius below the Nyquist frequency");
    p.cx = float(center[0]) / float(voxelSize); p.cy = float(center[1]) / float(voxelSize); p.cz = float(center[2]) / float(voxelSize);
    const float rmax1 = p.r1 + p.r2 + float(halfWidth), rmax2 = p.r2 + float(halfWidth);
    const int lo[3] = {int(std::floor(p.cx - rmax1)), int(std::floor(p.cy - rmax2)), int(std::floor(p.cz - rmax1))};
    const int hi[3] = {int(std::ceil(p.cx + rmax1)), int(std::ceil(p.cy + rmax2)), int(std::ceil(p.cz + rmax1))};
    return buildLevelSet(ctx, p, lo, hi, voxelSize, halfWidth, "torus_ls", nullptr, blockGridFor(lo, hi), out);
}

int vdbrt_build_levelset_spheres(vdbrt_ctx* ctx, const double* spheres, uint32_t n, double voxelSize, double halfWidth, vdbrt_grid** out)
{
    if (!ctx || !spheres || !out || n == 0) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (!(voxelSize > 0) || !(halfWidth > 1)) return setError(VDBRT_ERR_INVALID_ARG, "voxel size must be positive, half-width > 1");
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    const float dx = float(voxelSize), w = float(halfWidth);
    std::vector<float4> s;
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (uint32_t i = 0; i < n; ++i) {
        float4 q;
        q.w = float(spheres[4 * i + 3]) / dx;
        if (q.w < 1.5f) continue;                        // rasterSphere returns an empty grid; a union with it is a no-op
        q.x = float(spheres[4 * i]) / dx; q.y = float(spheres[4 * i + 1]) / dx; q.z = float(spheres[4 * i + 2]) / dx;
        const float rmax = q.w + w, c[3] = {q.x, q.y, q.z};
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], int(std::floor(c[a] - rmax))); hi[a] = std::max(hi[a], int(std::ceil(c[a] + rmax))); }
        s.push_back(q);
    }
    if (s.empty()) return setError(VDBRT_ERR_EMPTY_GRID, "all spheres are below the Nyquist frequency");
    const BlockGrid bg = blockGridFor(lo, hi);
    const uint64_t nBlocks64 = uint64_t(bg.nbx) * bg.nby * bg.nbz;
    if (nBlocks64 > (1ull << 26)) return setError(VDBRT_ERR_UNSUPPORTED, "sphere set spans too large a domain");
    const uint32_t nBlocks = uint32_t(nBlocks64);
    // per-block sphere lists (CSR, host): a sphere is listed in every block its band (r + w + 1 voxel) can reach
    std::vector<uint32_t> start(nBlocks + 2, 0);
    auto forBlocks = [&](const float4& q, auto f) {
        const float rr = q.w + w + 1.0f;
        const int b0[3] = {int(std::floor(q.x - rr)) >> 7, int(std::floor(q.y - rr)) >> 7, int(std::floor(q.z - rr)) >> 7};
        const int b1[3] = {int(std::ceil(q.x + rr)) >> 7, int(std::ceil(q.y + rr)) >> 7, int(std::ceil(q.z + rr)) >> 7};
        for (int x = b0[0]; x <= b1[0]; ++x) for (int y = b0[1]; y <= b1[1]; ++y) for (int z = b0[2]; z <= b1[2]; ++z) {
            const int ix = x - bg.bx0, iy = y - bg.by0, iz = z - bg.bz0;
            if (ix < 0 || iy < 0 || iz < 0 || ix >= bg.nbx || iy >= bg.nby || iz >= bg.nbz) continue;
            f(uint32_t((ix * bg.nby + iy) * bg.nbz + iz));
        }
    };
    for (const float4& q : s) forBlocks(q, [&](uint32_t b) { ++start[b + 1]; });
    for (uint32_t b = 0; b <= nBlocks; ++b) start[b + 1] += start[b];     // start[nBlocks+1] == start[nBlocks]: sentinel block is empty
    std::vector<uint32_t> list(start[nBlocks] ? start[nBlocks] : 1), fill(start.begin(), start.end() - 1);
    for (uint32_t i = 0; i < s.size(); ++i) forBlocks(s[i], [&](uint32_t b) { list[fill[b]++] = i; });
    DevBuf dS, dStart, dList;
    CUDA_TRY(cudaMalloc(&dS.p, s.size() * 16)); CUDA_TRY(cudaMalloc(&dStart.p, start.size() * 4)); CUDA_TRY(cudaMalloc(&dList.p, list.size() * 4));
    CUDA_TRY(cudaMemcpyAsync(dS.p, s.data(), s.size() * 16, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(dStart.p, start.data(), start.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(dList.p, list.data(), list.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    PrimSpheres p{dS.as<float4>(), dStart.as<uint32_t>(), dList.as<uint32_t>()};
    return buildLevelSet(ctx, p, lo, hi, double(dx), double(w), "spheres_ls", start.data(), bg, out);
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// sdfToFogVolume on an existing device grid: topology transform (drop leaves without negative voxels, drop emptied
// internal nodes) + value remap.  cutoff = background (the reference's default clamps to the minimum SDF value = -bg).
// ---------------------------------------------------------------------------------------------------------------
namespace {

// per lower node: which leaf children survive (any value < 0) -> new child mask + count
__global__ void __launch_bounds__(256)
k_fog_classify(const uint8_t* __restrict__ base, uint64_t lowerOff, float weight, unsigned long long* masks, uint16_t* prefix, uint32_t* leafCount)
{
    __shared__ unsigned long long smask[64];
    const uint32_t l = blockIdx.x;
    const uint8_t* node = base + lowerOff + uint64_t(l) * kLowerBytes;
    if (threadIdx.x < 64) smask[threadIdx.x] = 0ull;
    __syncthreads();
    for (uint32_t m = threadIdx.x; m < 4096; m += 256) {
        if (!maskBit(node + kLowerCMask, m)) continue;
        const uint8_t* leaf = node + ldgs64(node + kLowerTable + 8u * m);
        bool any = false;
        for (int n = 0; n < 512 && !any; ++n) {
            const float v = ldgf(leaf + kLeafValues + 4 * n);
            any = (v > 0.f ? 0.f : v * weight) > 0.f;             // SDFVoxelsToFogVolume (tools/LevelSetUtil.h:494-496)
        }
        if (any) atomicOr(&smask[m >> 6], 1ull << (m & 63u));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int w = 0; w < 64; ++w) { prefix[size_t(l) * 64 + w] = uint16_t(run); run += __popcll(smask[w]); }
        leafCount[l] = run;
    }
    if (threadIdx.x < 64) masks[size_t(l) * 64 + threadIdx.x] = smask[threadIdx.x];
}

struct FogLower { uint32_t src, leafBase; };   // src = index of the lower node in the level-set grid

__global__ void __launch_bounds__(256)
k_fog_fill_lower(const uint8_t* __restrict__ src, uint64_t srcLowerOff, Layout L, const FogLower* lowers, const unsigned long long* masks,
                 const uint16_t* prefix, uint8_t* dst, const uint8_t** leafSrc, unsigned int* tileCount)
{
    const uint32_t l = blockIdx.y, m = blockIdx.x * 256 + threadIdx.x;
    const FogLower fl = lowers[l];
    const uint8_t* sn = src + srcLowerOff + uint64_t(fl.src) * kLowerBytes;
    uint8_t* node = dst + L.lowerOff + uint64_t(l) * kLowerBytes;
    const unsigned long long word = masks[size_t(fl.src) * 64 + (m >> 6)];
    const bool child = (word >> (m & 63u)) & 1ull;
    unsigned long long entry; bool on = false;
    if (child) {
        const uint32_t leaf = fl.leafBase + prefix[size_t(fl.src) * 64 + (m >> 6)] + __popcll(word & ((1ull << (m & 63u)) - 1ull));
        entry = (unsigned long long)((long long)(L.leafOff + uint64_t(leaf) * kLeafBytes) - (long long)(L.lowerOff + uint64_t(l) * kLowerBytes));
        leafSrc[leaf] = sn + ldgs64(sn + kLowerTable + 8u * m);
    } else {
        // a dropped leaf had no negative voxel -> tile 0/off; an SDF tile < 0 becomes 1/on (SDFTilesToFogVolume + ValueAllIter pass)
        float v = 1.f;
        if (!maskBit(sn + kLowerCMask, m)) v = ldgf(sn + kLowerTable + 8u * m);
        on = v < 0.f;
        entry = __float_as_uint(on ? 1.0f : 0.0f);
    }
    reinterpret_cast<unsigned long long*>(node + kLowerTable)[m] = entry;
    const unsigned onBits = __ballot_sync(0xffffffffu, on);
    if ((m & 31u) == 0) {
        reinterpret_cast<uint32_t*>(node + kLowerVMask)[m >> 5] = onBits;
        if (onBits) atomicAdd(tileCount, __popc(onBits));
    }
    if ((m & 63u) == 0) reinterpret_cast<unsigned long long*>(node + kLowerCMask)[m >> 6] = word;
    if (m < 8) reinterpret_cast<int*>(node)[m] = m < 6 ? reinterpret_cast<const int*>(sn)[m] : 0;
    if (m < 4) reinterpret_cast<float*>(node + kLowerCMask + 512)[m] = 0.f;
}

__global__ void __launch_bounds__(256)
k_fog_fill_upper(const uint8_t* __restrict__ src, uint64_t srcUpperOff, uint64_t srcLowerOff, Layout L, const uint32_t* upperSrc, const int* newLowerOfOld,
                 uint8_t* dst, unsigned int* tileCount)
{
    const uint32_t u = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    const uint8_t* sn = src + srcUpperOff + uint64_t(upperSrc[u]) * kUpperBytes;
    uint8_t* node = dst + L.upperOff + uint64_t(u) * kUpperBytes;
    int child = -1; bool on = false;
    unsigned long long entry;
    if (maskBit(sn + kUpperCMask, n)) {
        const uint64_t oldLower = (uint64_t(sn - src) + uint64_t(ldgs64(sn + kUpperTable + 8u * n)) - srcLowerOff) / kLowerBytes;
        child = newLowerOfOld[oldLower];
        if (child >= 0) entry = (unsigned long long)((long long)(L.lowerOff + uint64_t(child) * kLowerBytes) - (long long)(L.upperOff + uint64_t(u) * kUpperBytes));
        else entry = __float_as_uint(0.0f);            // the whole lower node lost its leaves: outside shell, value 0 / off
    } else {
        on = ldgf(sn + kUpperTable + 8u * n) < 0.f;
        entry = __float_as_uint(on ? 1.0f : 0.0f);
    }
    reinterpret_cast<unsigned long long*>(node + kUpperTable)[n] = entry;
    const unsigned cBits = __ballot_sync(0xffffffffu, child >= 0), vBits = __ballot_sync(0xffffffffu, on);
    if ((n & 31u) == 0) {
        reinterpret_cast<uint32_t*>(node + kUpperCMask)[n >> 5] = cBits;
        reinterpret_cast<uint32_t*>(node + kUpperVMask)[n >> 5] = vBits;
        if (vBits) atomicAdd(tileCount + 1, __popc(vBits));
    }
    if (n < 8) reinterpret_cast<int*>(node)[n] = n < 6 ? reinterpret_cast<const int*>(sn)[n] : 0;
    if (n < 4) reinterpret_cast<float*>(node + kUpperCMask + 4096)[n] = 0.f;
}

__global__ void __launch_bounds__(512)
k_fog_fill_leaves(Layout L, float weight, const uint8_t* const* leafSrc, uint8_t* dst, int* bbox, unsigned long long* voxelCount)
{
    __shared__ int smin[3], smax[3];
    __shared__ unsigned int scount;
    const uint32_t leaf = blockIdx.x, n = threadIdx.x;
    const uint8_t* sn = leafSrc[leaf];
    uint8_t* node = dst + L.leafOff + uint64_t(leaf) * kLeafBytes;
    if (n < 3) { smin[n] = 7; smax[n] = 0; }
    if (n == 0) scount = 0;
    __syncthreads();
    const float s = ldgf(sn + kLeafValues + 4u * n);
    const float v = s > 0.f ? 0.f : s * weight;
    const bool active = v > 0.f;
    reinterpret_cast<float*>(node + kLeafValues)[n] = v;
    const unsigned bits = __ballot_sync(0xffffffffu, active);
    if ((n & 31u) == 0) { reinterpret_cast<uint32_t*>(node + kLeafVMask)[n >> 5] = bits; if (bits) atomicAdd(&scount, __popc(bits)); }
    const int li = int(n >> 6), lj = int((n >> 3) & 7u), lk = int(n & 7u);
    if (active) {
        atomicMin(&smin[0], li); atomicMin(&smin[1], lj); atomicMin(&smin[2], lk);
        atomicMax(&smax[0], li); atomicMax(&smax[1], lj); atomicMax(&smax[2], lk);
    }
    __syncthreads();
    if (n == 0) {
        const int* sh = reinterpret_cast<const int*>(sn);
        const int ox = sh[0] & ~7, oy = sh[1] & ~7, oz = sh[2] & ~7;
        int* h = reinterpret_cast<int*>(node);
        h[0] = ox + smin[0]; h[1] = oy + smin[1]; h[2] = oz + smin[2];
        node[12] = uint8_t(smax[0] - smin[0]); node[13] = uint8_t(smax[1] - smin[1]); node[14] = uint8_t(smax[2] - smin[2]); node[15] = 2;
        float* st = reinterpret_cast<float*>(node + 80); st[0] = st[1] = st[2] = st[3] = 0.f;
        atomicMin(bbox + 0, ox + smin[0]); atomicMin(bbox + 1, oy + smin[1]); atomicMin(bbox + 2, oz + smin[2]);
        atomicMax(bbox + 3, ox + smax[0]); atomicMax(bbox + 4, oy + smax[1]); atomicMax(bbox + 5, oz + smax[2]);
        atomicAdd(voxelCount, (unsigned long long)scount);
    }
}

} // namespace

extern "C" int vdbrt_build_fog_from_levelset(vdbrt_ctx* ctx, const vdbrt_grid* ls, vdbrt_grid** out)
{
    if (!ctx || !ls || !out) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (ls->info.grid_class != VDBRT_GRID_CLASS_LEVEL_SET) return setError(VDBRT_ERR_NOT_LEVELSET, "sdfToFogVolume needs a level set");
    if (ls->leaf_kind != 0) return setError(VDBRT_ERR_UNSUPPORTED, "sdfToFogVolume needs float leaves: upload the quantised grid with quant_native = 0");
    std::lock_guard<std::mutex> lock(ctx->mx);
    DeviceGuard guard(ctx->device);
    cudaStream_t st = ctx->stream;
    // source layout
    uint8_t head[kGridBytes + kTreeBytes];
    CUDA_TRY(cudaMemcpyAsync(head, ls->dev, sizeof(head), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    auto rd64 = [&](size_t off) { int64_t v; std::memcpy(&v, head + off, 8); return v; };
    const uint64_t srcLeafOff = kGridBytes + uint64_t(rd64(kGridBytes + 0)), srcLowerOff = kGridBytes + uint64_t(rd64(kGridBytes + 8)),
                   srcUpperOff = kGridBytes + uint64_t(rd64(kGridBytes + 16)), srcRootOff = kGridBytes + uint64_t(rd64(kGridBytes + 24));
    (void)srcLeafOff;
    const uint32_t oldLower = ls->info.lower_count, oldUpper = ls->info.upper_count, oldTiles = ls->info.root_tiles;
    const float bgVal = ls->info.background;
    const float cutoff = -std::fabs(bgVal);                 // clamped to the minimum SDF value (tools/LevelSetUtil.h:2238-2239)
    const float weight = 1.0f / cutoff;                     // SDFVoxelsToFogVolume: mWeight = ValueType(1.0) / cutoffDistance
    DevBuf dMasks, dPrefix, dCount;
    CUDA_TRY(cudaMalloc(&dMasks.p, size_t(oldLower) * 64 * 8)); CUDA_TRY(cudaMalloc(&dPrefix.p, size_t(oldLower) * 64 * 2)); CUDA_TRY(cudaMalloc(&dCount.p, size_t(oldLower) * 4));
    k_fog_classify<<<oldLower, 256, 0, st>>>(ls->dev, srcLowerOff, weight, dMasks.as<unsigned long long>(), dPrefix.as<uint16_t>(), dCount.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> hCount(oldLower);
    CUDA_TRY(cudaMemcpyAsync(hCount.data(), dCount.p, size_t(oldLower) * 4, cudaMemcpyDeviceToHost, st));
    // parents of the lower nodes: read the upper child tables on the host (masks + offsets are small relative to the grid)
    std::vector<uint8_t> rootBuf(kRootBytes + 32ull * oldTiles);
    CUDA_TRY(cudaMemcpyAsync(rootBuf.data(), ls->dev + srcRootOff, rootBuf.size(), cudaMemcpyDeviceToHost, st));
    std::vector<uint8_t> upperBuf(size_t(oldUpper) * kUpperBytes);
    CUDA_TRY(cudaMemcpyAsync(upperBuf.data(), ls->dev + srcUpperOff, upperBuf.size(), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<int> newLowerOfOld(oldLower, -1);
    std::vector<FogLower> lowers;
    std::vector<uint32_t> upperSrc;
    uint64_t nLeaf64 = 0;
    // keep source order: uppers in source order, lowers in (upper, slot) order
    for (uint32_t u = 0; u < oldUpper; ++u) {
        const uint8_t* un = upperBuf.data() + size_t(u) * kUpperBytes;
        bool keep = false;
        std::vector<uint32_t> kids;
        for (uint32_t n = 0; n < 32768; ++n) {
            uint64_t w; std::memcpy(&w, un + kUpperCMask + 8 * (n >> 6), 8);
            if (!((w >> (n & 63)) & 1)) continue;
            int64_t off; std::memcpy(&off, un + kUpperTable + 8 * n, 8);
            const uint64_t ol = (srcUpperOff + uint64_t(u) * kUpperBytes + uint64_t(off) - srcLowerOff) / kLowerBytes;
            if (hCount[ol]) { kids.push_back(uint32_t(ol)); keep = true; }
        }
        if (!keep) continue;
        upperSrc.push_back(u);
        for (uint32_t ol : kids) { newLowerOfOld[ol] = int(lowers.size()); lowers.push_back(FogLower{ol, uint32_t(nLeaf64)}); nLeaf64 += hCount[ol]; }
    }
    if (lowers.empty()) return setError(VDBRT_ERR_EMPTY_GRID, "level set has no interior voxels");
    const uint32_t nLeaf = uint32_t(nLeaf64), nLower = uint32_t(lowers.size()), nUpper = uint32_t(upperSrc.size());
    Layout L;
    L.rootOff = kGridBytes + kTreeBytes;
    L.upperOff = (L.rootOff + kRootBytes + 32ull * nUpper + 31) & ~31ull;
    L.lowerOff = L.upperOff + kUpperBytes * nUpper;
    L.leafOff = L.lowerOff + kLowerBytes * nLower;
    L.bg = 0.f; L.dx = float(ls->info.voxel_size[0]); L.hw = 0.f;
    const uint64_t total = L.leafOff + kLeafBytes * nLeaf;
    auto* g = new vdbrt_grid;
    g->bytes = total; g->device = ctx->device;
    cudaError_t e = cudaMalloc(&g->dev, total);
    if (e != cudaSuccess) { delete g; return cudaFail(e, "cudaMalloc(grid)"); }
    auto failGrid = [&](int rc) { destroyGrid(g); return rc; };
    cudaMemsetAsync(g->dev, 0, total, st);                    // padding bytes are defined: the buffer is reproducible byte for byte
    DevBuf dLowers, dUpperSrc, dMap, dLeafSrc, dStats;
    if (cudaMalloc(&dLowers.p, nLower * sizeof(FogLower)) != cudaSuccess || cudaMalloc(&dUpperSrc.p, nUpper * 4) != cudaSuccess ||
        cudaMalloc(&dMap.p, size_t(oldLower) * 4) != cudaSuccess || cudaMalloc(&dLeafSrc.p, size_t(nLeaf) * 8) != cudaSuccess || cudaMalloc(&dStats.p, 64) != cudaSuccess)
        return failGrid(setError(VDBRT_ERR_NOMEM, "out of device memory while building the grid"));
    cudaMemcpyAsync(dLowers.p, lowers.data(), nLower * sizeof(FogLower), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dUpperSrc.p, upperSrc.data(), nUpper * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dMap.p, newLowerOfOld.data(), size_t(oldLower) * 4, cudaMemcpyHostToDevice, st);
    int initBox[8] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0, 0};
    cudaMemcpyAsync(dStats.p, initBox, 32, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(static_cast<uint8_t*>(dStats.p) + 32, 0, 32, st);
    unsigned int* dTiles = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(dStats.p) + 48);
    k_fog_fill_upper<<<dim3(128, nUpper), 256, 0, st>>>(ls->dev, srcUpperOff, srcLowerOff, L, dUpperSrc.as<uint32_t>(), dMap.as<int>(), g->dev, dTiles);
    k_fog_fill_lower<<<dim3(16, nLower), 256, 0, st>>>(ls->dev, srcLowerOff, L, dLowers.as<FogLower>(), dMasks.as<unsigned long long>(), dPrefix.as<uint16_t>(), g->dev,
                                                       dLeafSrc.as<const uint8_t*>(), dTiles);
    k_fog_fill_leaves<<<nLeaf, 512, 0, st>>>(L, weight, dLeafSrc.as<const uint8_t*>(), g->dev, dStats.as<int>(),
                                              reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(dStats.p) + 32));
    e = cudaGetLastError();
    if (e != cudaSuccess) return failGrid(cudaFail(e, "fog fill kernels"));
    int hBox[8]; unsigned long long voxels = 0; unsigned int hTiles[2] = {0, 0};
    cudaMemcpyAsync(hBox, dStats.p, 32, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&voxels, static_cast<uint8_t*>(dStats.p) + 32, 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(hTiles, dTiles, 8, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return failGrid(cudaFail(e, "fog fill kernels"));
    std::vector<uint8_t> hd(L.upperOff, 0);
    const uint32_t tiles[3] = {hTiles[0], hTiles[1], 0};
    // NOTE: the voxel-tight bbox here covers active voxels only; active tiles widen it in the reference's stats pass.  The
    // renderer never reads it (it uses the node-granular bbox computed by finishGrid).
    writeGridHeader(hd.data(), total, "fog", ls->info.voxel_size[0], ls->info.translation, VDBRT_GRID_CLASS_FOG_VOLUME, hBox, L.rootOff, L.upperOff, L.lowerOff,
                    L.leafOff, nLeaf, nLower, nUpper, tiles, voxels);
    uint8_t* root = hd.data() + L.rootOff;
    for (int i = 0; i < 6; ++i) wr<int32_t>(root + 4 * i, hBox[i]);
    wr<uint32_t>(root + 24, nUpper); wr<float>(root + 28, 0.0f);
    std::vector<std::pair<uint64_t, uint32_t>> keys;
    for (uint32_t u = 0; u < nUpper; ++u) {
        const int* b = reinterpret_cast<const int*>(upperBuf.data() + size_t(upperSrc[u]) * kUpperBytes);
        keys.push_back({rootKeyHost(b[0] & ~4095, b[1] & ~4095, b[2] & ~4095), u});
    }
    std::sort(keys.begin(), keys.end());
    for (uint32_t i = 0; i < nUpper; ++i) {
        uint8_t* t = root + kRootBytes + 32 * i;
        wr<uint64_t>(t, keys[i].first);
        wr<int64_t>(t + 8, int64_t(L.upperOff + kUpperBytes * keys[i].second) - int64_t(L.rootOff));
        wr<uint32_t>(t + 16, 0); wr<float>(t + 20, 0.0f);
    }
    e = cudaMemcpyAsync(g->dev, hd.data(), hd.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return failGrid(cudaFail(e, "cudaMemcpyAsync(header)"));
    const int rc = finishGrid(ctx, g);
    if (rc != VDBRT_OK) return failGrid(rc);
    *out = g;
    return VDBRT_OK;
}
