// vdbrt_device.cuh -- device-side building blocks of the sm_100a ray-tracing kernels.
//
// Data layout is NanoVDB's (NanoGrid<float>, nanovdb/nanovdb/NanoVDB.h:67-122); the ALGORITHM and PRECISION are
// OpenVDB's CPU RayTracer: double rays / DDA / hit times, float voxel values and stencil maths,
// no FMA contraction anywhere (this translation unit is compiled with --fmad=false; division and sqrt are IEEE).
// Paths in comments are relative to the OpenVDB tree (openvdb/openvdb/...).
//
// Design (B200-first, not a translation of the reference's nested C++ templates):
//   * TreeCursor  -- register-resident root/upper/lower/leaf path cache.  One cached coordinate + three 32-bit
//                    node handles (byte offset >> 5, nodes are 32 B aligned); a probe XORs the query with the
//                    cached coordinate and re-descends only from the first level that differs.  The root tile
//                    table lives in shared memory.  All grid reads are ld.global.nc.
//   * HDDA        -- the reference nests four fixed-stride DDAs as recursive template calls.  Here ONE DDA lives
//                    in registers and the suspended parent levels are parked in explicit save slots, so every
//                    lane of a warp executes the same "probe / step / descend / ascend" loop body whatever level
//                    it is on (no per-level code replication, no dynamic register indexing).  Per-level deltas
//                    are recomputed as double(+-DIM)*invDir, the reference's own expression, instead of stored.
//   * the same cursor drives the fog path as a resumable span generator (VolumeHDDA::hits without the std::vector).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace vdbrt {

// ---- NanoVDB offsets (GridData :1944-2135, TreeData :2393-2423, RootData :2621-2695, InternalData :3181-3291,
//      LeafData<float> :3671-3746)
constexpr uint32_t kRootTableSize = 24, kRootBackground = 28, kRootTiles = 64, kTileSize = 32;
constexpr uint32_t kUpperVMask = 32, kUpperCMask = 32 + 4096, kUpperTable = 8256;
constexpr uint32_t kLowerVMask = 32, kLowerCMask = 32 + 512, kLowerTable = 1088;
constexpr uint32_t kLeafVMask = 16, kLeafValues = 96;
constexpr int kMaxSmemTiles = 96;

struct DevGrid {
    const uint8_t* base;      // device copy of the serialised grid (GridData at offset 0)
    const uint8_t* tiles;     // RootData::Tile array (device)
    uint64_t root_off;        // byte offset of RootData from base (root tile children are relative to it)
    uint32_t table_size;
    float    background;
    double   scale[3], inv[3], trans[3];
    int32_t  bbox_min[3], bbox_max[3];   // node-granular bbox, max as CoordBBox stores it (no +1)
    uint32_t has_translation;
    uint32_t grid_class;
    double   voxel_size0;
    // a map with off-diagonal terms (rotation / shear: OpenVDB's AffineMap, math/Maps.h:411-445): nanovdb::Map::mMatD / mInvMatD as
    // stored (NanoVDB.h:1418-1554).  TOLERANCE path: bit-exactness is only claimed for scale(+translate) maps.
    uint32_t general, pad0;
    double   mat[9], imat[9];
    // Halo blocks (null = none): for leaf i, the 9x9x9 values getValue() returns at origin + (0..8)^3 -- the
    // leaf's own 512 values plus the +x/+y/+z faces, edges and corner taken from whatever lies there (neighbour leaf, tile,
    // background) -- at halo + 736 * i floats, index 81 x + 9 y + z.  Built once when the grid is registered (k_build_halo).
    // A BoxStencil cell (tools: math/Stencils.h:414-423) whose base voxel lies in a leaf reads its 8 corners from that
    // leaf's block with 8 loads off one pointer: no walk over the up to 8 leaves the cell touches.
    const float* halo;        // float leaves: 736 floats per block; quantised leaves: the same pointer holds the code blocks (LeafKind::block bytes each)
    uint32_t leaf0;           // handle (byte offset >> 5) of leaf 0; leaves are LeafKind::handles apart (2144 B = 67 handles for float leaves)
    uint32_t leaf_count;
    // "leaf or active tile" masks of the lower nodes (null = none): lowmask[64 i + w] = child-mask word w | value-mask word w of
    // lower node i (512 B per node, built by k_build_lowmask).  The volume walk only needs this one bit per 8^3 cell: at the
    // leaf level VolumeHDDA counts any existing leaf as active, otherwise the tile's state (math/DDA.h:308-309,326-327).
    const unsigned long long* lowmask;
    uint32_t lower0;          // handle of lower node 0; lower nodes are 33856 B = 1058 handles apart
    uint32_t lower_count;
};
constexpr uint32_t kHaloStride = 736;   // floats per halo block (729 used; 2944 B keeps blocks 32-byte aligned)

// ---- leaf kinds.  NanoGrid<Fp8> / NanoGrid<Fp16> (nanovdb/NanoVDB.h:3752-3930) can be rendered AS THEY ARE (vdbrt_upload_grid with
// the context's quant_native knob): the tree above the leaves is the float build's, a leaf is LeafFnBase (96 B: origin, flags, value
// mask, float mMinimum @80, float mQuantum @84, statistics) + 512 codes of 1 or 2 bytes, and
//     LeafData<FpX>::getValue(i) = float(code_i) * mQuantum + mMinimum        (:3876, :3906; product and sum rounded separately)
// is evaluated where the value is fetched.  The halo block of such a leaf keeps CODES too: 8 {minimum, quantum} pairs -- own leaf,
// +z, +y, +yz, +x, +xz, +xy, +xyz neighbour block; a block without a leaf is {its tile / background value, 0} with code 0 -- then
// 729 codes at 81 x + 9 y + z: 800 B (Fp8) or 1 536 B (Fp16) per leaf instead of 2 944.  Every kind is its own instantiation of
// the kernels (LEAF template parameter): the float kernels carry none of this (the branch inside one kernel cost them 32 %, round 1).
enum { kLeafFloat = 0, kLeafFp8 = 1, kLeafFp16 = 2 };
template<int LEAF> struct LeafKind;
template<> struct LeafKind<kLeafFloat> { static constexpr uint32_t handles = 67u, bytes = 2144u, block = 2944u; };
template<> struct LeafKind<kLeafFp8> { static constexpr uint32_t handles = 19u, bytes = 608u, block = 800u; };
template<> struct LeafKind<kLeafFp16> { static constexpr uint32_t handles = 35u, bytes = 1120u, block = 1536u; };
constexpr uint32_t kLeafMinimum = 80, kLeafQuantum = 84, kQHaloCodes = 64;

struct RootSmem {
    unsigned long long key[kMaxSmemTiles];
    uint32_t child[kMaxSmemTiles];   // node handle (offset>>5 from base), 0 = value tile
    uint32_t state[kMaxSmemTiles];
    float    value[kMaxSmemTiles];
    uint32_t count;                  // tiles staged; == table_size unless the table is too large
    uint32_t staged;
};

__device__ __forceinline__ unsigned long long ldg64(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned long long*>(p)); }
__device__ __forceinline__ long long ldgs64(const uint8_t* p) { return __ldg(reinterpret_cast<const long long*>(p)); }
__device__ __forceinline__ uint32_t ldg32(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ float ldgf(const uint8_t* p) { return __ldg(reinterpret_cast<const float*>(p)); }
// bit n of a nanovdb::Mask (little-endian 64-bit words): read through the 32-bit half that holds it, a 32-bit shift is one instruction
__device__ __forceinline__ bool maskBit(const uint8_t* mask, uint32_t n) { return (ldg32(mask + 4u * (n >> 5)) >> (n & 31u)) & 1u; }

// value n of a leaf (LeafData<float>::getValue / LeafData<FpX>::getValue)
template<int LEAF>
__device__ __forceinline__ float leafValue(const uint8_t* leaf, uint32_t n)
{
    if (LEAF == kLeafFloat) return ldgf(leaf + kLeafValues + 4u * n);
    const float code = LEAF == kLeafFp8 ? float(__ldg(leaf + kLeafValues + n)) : float(__ldg(reinterpret_cast<const unsigned short*>(leaf + kLeafValues) + n));
    return __fadd_rn(__fmul_rn(code, ldgf(leaf + kLeafQuantum)), ldgf(leaf + kLeafMinimum));
}
// value (x,y,z) in 0..8 of halo block `block`
template<int LEAF>
__device__ __forceinline__ float haloValue(const float* halo, uint32_t block, uint32_t x, uint32_t y, uint32_t z)
{
    const uint32_t i = x * 81u + y * 9u + z;
    if (LEAF == kLeafFloat) return __ldg(halo + size_t(block) * kHaloStride + i);
    const uint8_t* b = reinterpret_cast<const uint8_t*>(halo) + size_t(block) * LeafKind<LEAF>::block;
    const float2 mq = __ldg(reinterpret_cast<const float2*>(b) + (((x >> 3) << 2) | ((y >> 3) << 1) | (z >> 3)));      // {minimum, quantum} of the block the value came from
    const float code = LEAF == kLeafFp8 ? float(__ldg(b + kQHaloCodes + i)) : float(__ldg(reinterpret_cast<const unsigned short*>(b + kQHaloCodes) + i));
    return __fadd_rn(__fmul_rn(code, mq.y), mq.x);
}

// RootData::CoordToKey with NANOVDB_USE_SINGLE_ROOT_KEY (NanoVDB.h:2630-2640)
__device__ __forceinline__ unsigned long long rootKey(int x, int y, int z) {
    return (unsigned long long)(uint32_t(z) >> 12) | ((unsigned long long)(uint32_t(y) >> 12) << 21) | ((unsigned long long)(uint32_t(x) >> 12) << 42);
}
__device__ __forceinline__ uint32_t upperOffset(int x, int y, int z) { return (((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7); }
__device__ __forceinline__ uint32_t lowerOffset(int x, int y, int z) { return (((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3); }
__device__ __forceinline__ uint32_t leafOffset(int x, int y, int z) { return ((x & 7) << 6) | ((y & 7) << 3) | (z & 7); }

// cooperative staging of the root table into shared memory (call from all threads of the CTA, then __syncthreads)
__device__ __forceinline__ void stageRoot(const DevGrid& g, RootSmem& s)
{
    const uint32_t n = g.table_size <= (uint32_t)kMaxSmemTiles ? g.table_size : 0u;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint8_t* t = g.tiles + kTileSize * i;
        s.key[i] = ldg64(t);
        const long long child = ldgs64(t + 8);
        s.child[i] = child ? uint32_t((g.root_off + (unsigned long long)child) >> 5) : 0u;
        s.state[i] = ldg32(t + 16);
        s.value[i] = ldgf(t + 20);
    }
    if (threadIdx.x == 0) { s.count = n; s.staged = (n == g.table_size); }
}

struct Counters {   // per-thread work counters of the instrumented (never timed) launches
    uint32_t root, upper, lower, voxel, refills, psamples, ssamples, srays, hits, rays;
    // diagnostics of the warp-synchronous level-set loop (VDBRT_DEBUG_TILES): lanes that ran the node probe / voxel probe / stencil
    // evaluation / DDA step of an iteration [0..3], iterations in which the warp ran that phase at all [4..7] (counted by lane 0)
    uint32_t diag[8];
};

// ------------------------------------------------------------------------------------------------------------
// TreeCursor: semantics of tree::ValueAccessor::probeConstNode / probeValue / getValue / isValueOn
// (tree/ValueAccessor.h:455-510,805-830,937-953) on the NanoVDB node arrays.
// ------------------------------------------------------------------------------------------------------------
struct TreeCursor {
    int kx, ky, kz;          // coordinate of the last descent
    uint32_t n2, n1, n0;     // handles of the upper / lower / leaf node containing (kx,ky,kz); 0 = none

    __device__ __forceinline__ void reset() { kx = ky = kz = 0; n2 = n1 = n0 = 0u; }
    __device__ __forceinline__ static const uint8_t* node(const DevGrid& g, uint32_t h) { return g.base + ((unsigned long long)h << 5); }

    // root table search (RootNode::probeChild / findTile, NanoVDB.h:2785-2799): index of the tile or -1
    __device__ __forceinline__ static int findTile(const DevGrid& g, const RootSmem& s, int x, int y, int z)
    {
        const unsigned long long key = rootKey(x, y, z);
        if (s.staged) {
            for (uint32_t i = 0; i < s.count; ++i) if (s.key[i] == key) return int(i);
            return -1;
        }
        for (uint32_t i = 0; i < g.table_size; ++i) if (ldg64(g.tiles + kTileSize * i) == key) return int(i);
        return -1;
    }
    __device__ __forceinline__ static uint32_t tileChild(const DevGrid& g, const RootSmem& s, int i)
    {
        if (s.staged) return s.child[i];
        const long long child = ldgs64(g.tiles + kTileSize * i + 8);
        return child ? uint32_t((g.root_off + (unsigned long long)child) >> 5) : 0u;
    }

    // Re-descend from the first level whose node does not contain (x,y,z).  Returns the depth reached:
    // 0 = leaf, 1 = lower node (no leaf), 2 = upper node (no lower), 3 = nothing below the root.
    __device__ __forceinline__ int descend(const DevGrid& g, const RootSmem& s, int x, int y, int z)
    {
        const uint32_t d = uint32_t((x ^ kx) | (y ^ ky) | (z ^ kz));
        if (n0 && !(d & ~7u)) return 0;
        if (!(n1 && !(d & ~127u))) {
            if (!(n2 && !(d & ~4095u))) {
                const int t = findTile(g, s, x, y, z);
                n2 = t >= 0 ? tileChild(g, s, t) : 0u;
            }
            n1 = 0u;
            if (n2) {
                // the child-mask word and the table entry are fetched together (two independent loads, one latency); the
                // entry is only a child offset -- relative to this InternalData (:3190-3199) -- when the mask bit is set
                const uint8_t* u = node(g, n2);
                const uint32_t n = upperOffset(x, y, z);
                const uint32_t word = ldg32(u + kUpperCMask + 4u * (n >> 5));
                const long long ent = ldgs64(u + kUpperTable + 8u * n);
                if ((word >> (n & 31u)) & 1u) n1 = n2 + uint32_t(int(ent >> 5));        // offsets are multiples of 32
            }
        }
        n0 = 0u;
        if (n1) {
            const uint8_t* l = node(g, n1);
            const uint32_t n = lowerOffset(x, y, z);
            const uint32_t word = ldg32(l + kLowerCMask + 4u * (n >> 5));
            // (loading the entry only when the bit is set -- most cells of a lower node are empty -- was measured: no gain, +1 % on C2)
            const long long ent = ldgs64(l + kLowerTable + 8u * n);
            if ((word >> (n & 31u)) & 1u) n0 = n1 + uint32_t(int(ent >> 5));
        }
        kx = x; ky = y; kz = z;
        return n0 ? 0 : (n1 ? 1 : (n2 ? 2 : 3));
    }

    // value and active state at the deepest node containing the coordinate (after descend())
    template<int LEAF = kLeafFloat>
    __device__ __forceinline__ bool valueAt(const DevGrid& g, const RootSmem& s, int depth, int x, int y, int z, float& v) const
    {
        if (depth == 0) { const uint8_t* p = node(g, n0); const uint32_t n = leafOffset(x, y, z); v = leafValue<LEAF>(p, n); return maskBit(p + kLeafVMask, n); }
        if (depth == 1) { const uint8_t* p = node(g, n1); const uint32_t n = lowerOffset(x, y, z); v = ldgf(p + kLowerTable + 8u * n); return maskBit(p + kLowerVMask, n); }
        if (depth == 2) { const uint8_t* p = node(g, n2); const uint32_t n = upperOffset(x, y, z); v = ldgf(p + kUpperTable + 8u * n); return maskBit(p + kUpperVMask, n); }
        const int t = findTile(g, s, x, y, z);
        if (t < 0) { v = g.background; return false; }
        if (s.staged) { v = s.value[t]; return s.state[t] != 0u; }
        v = ldgf(g.tiles + kTileSize * t + 20); return ldg32(g.tiles + kTileSize * t + 16) != 0u;
    }
    __device__ __forceinline__ bool activeAt(const DevGrid& g, const RootSmem& s, int depth, int x, int y, int z) const
    {
        if (depth == 0) return maskBit(node(g, n0) + kLeafVMask, leafOffset(x, y, z));
        if (depth == 1) return maskBit(node(g, n1) + kLowerVMask, lowerOffset(x, y, z));
        if (depth == 2) return maskBit(node(g, n2) + kUpperVMask, upperOffset(x, y, z));
        const int t = findTile(g, s, x, y, z);
        if (t < 0) return false;
        return (s.staged ? s.state[t] : ldg32(g.tiles + kTileSize * t + 16)) != 0u;
    }
    template<int LEAF = kLeafFloat>
    __device__ __forceinline__ bool probeValue(const DevGrid& g, const RootSmem& s, int x, int y, int z, float& v)
    {
        const int depth = descend(g, s, x, y, z);
        return valueAt<LEAF>(g, s, depth, x, y, z, v);
    }
    template<int LEAF = kLeafFloat>
    __device__ __forceinline__ float getValue(const DevGrid& g, const RootSmem& s, int x, int y, int z)
    {
        float v; probeValue<LEAF>(g, s, x, y, z, v); return v;
    }

    // the 8 corners of the cell (x..x+1, y..y+1, z..z+1) in BoxStencil slot order 000,001,011,010,100,101,111,110
    // (math/Stencils.h:285-293,414-423; same order as BoxSampler::probeValues, tools/Interpolation.h:663-689).
    // A cell touches 2^(number of axes on which it straddles a leaf face) leaves.  The loop enumerates exactly those
    // leaves (the sub-masks of fmask in ascending order), so the k-th trip of every lane of a warp is a leaf visit --
    // a warp runs max(2^straddles) trips, not one per distinct neighbour code -- and pulls all corners that live in the
    // visited leaf with predicated loads off one base pointer.  KEEP = false: the fetch works on a COPY of the cursor, so
    // the caller's path cache keeps pointing at the node its DDA is walking (no re-descent after a stencil that straddled
    // a face); KEEP = true: the cursor follows the fetch (the fog sampler's own cursor, tools/Interpolation.h:420-425).
    template<bool KEEP, int LEAF = kLeafFloat>
    __device__ __forceinline__ void fetchCell(const DevGrid& g, const RootSmem& s, int x, int y, int z, float v[8])
    {
        if (g.halo) {
            // base voxel inside a leaf (92 % of a level set's stencil moves) -> that leaf's halo block
            TreeCursor copy = *this;
            TreeCursor& probe = KEEP ? *this : copy;
            const int depth = probe.descend(g, s, x, y, z);
            if (depth == 0) {
                const uint32_t i = (probe.n0 - g.leaf0) / LeafKind<LEAF>::handles;
                if (i < g.leaf_count) {
                    if (LEAF == kLeafFloat) {
                        const float* b = g.halo + size_t(i) * kHaloStride + (uint32_t(x & 7) * 81u + uint32_t(y & 7) * 9u + uint32_t(z & 7));
                        v[0] = __ldg(b); v[1] = __ldg(b + 1); v[2] = __ldg(b + 10); v[3] = __ldg(b + 9);
                        v[4] = __ldg(b + 81); v[5] = __ldg(b + 82); v[6] = __ldg(b + 91); v[7] = __ldg(b + 90);
                    } else {
                        const uint32_t bx = uint32_t(x & 7), by = uint32_t(y & 7), bz = uint32_t(z & 7);
                        v[0] = haloValue<LEAF>(g.halo, i, bx, by, bz);         v[1] = haloValue<LEAF>(g.halo, i, bx, by, bz + 1);
                        v[2] = haloValue<LEAF>(g.halo, i, bx, by + 1, bz + 1); v[3] = haloValue<LEAF>(g.halo, i, bx, by + 1, bz);
                        v[4] = haloValue<LEAF>(g.halo, i, bx + 1, by, bz);     v[5] = haloValue<LEAF>(g.halo, i, bx + 1, by, bz + 1);
                        v[6] = haloValue<LEAF>(g.halo, i, bx + 1, by + 1, bz + 1); v[7] = haloValue<LEAF>(g.halo, i, bx + 1, by + 1, bz);
                    }
                    return;
                }
            } else if (((x & 7) != 7) && ((y & 7) != 7) && ((z & 7) != 7)) {
                // no leaf at the base voxel and the cell stays inside its 8^3 block (a ray entering the band through a high
                // face: tester.init's position lies in the empty block it came from): one tile / the background covers all 8
                float tile;
                probe.template valueAt<LEAF>(g, s, depth, x, y, z, tile);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = tile;
                return;
            }
        }
        fetchCellAcrossLeaves<KEEP, LEAF>(g, s, x, y, z, v);
    }
    // the general case of fetchCell: no halo blocks, or a cell whose base voxel has no leaf and that reaches into the next 8^3 block.
    // (As a __noinline__ function -- it is the rare case of the render loops and a third of their stencil code -- the stencil's values move
    // to local memory: measured 7 % slower.)
    template<bool KEEP, int LEAF = kLeafFloat>
    __device__ __forceinline__ void fetchCellAcrossLeaves(const DevGrid& g, const RootSmem& s, int x, int y, int z, float v[8])
    {
        const int fmask = (((x & 7) == 7) ? 4 : 0) | (((y & 7) == 7) ? 2 : 0) | (((z & 7) == 7) ? 1 : 0);
        const uint32_t ox0 = uint32_t(x & 7) << 6, ox1 = uint32_t((x + 1) & 7) << 6, oy0 = uint32_t(y & 7) << 3, oy1 = uint32_t((y + 1) & 7) << 3;
        const uint32_t oz0 = uint32_t(z & 7), oz1 = uint32_t((z + 1) & 7);
        TreeCursor local = *this;
        TreeCursor& t = KEEP ? *this : local;
        int cmb = 0;
#pragma unroll 1
        do {
            const int rx = x + (cmb >> 2), ry = y + ((cmb >> 1) & 1), rz = z + (cmb & 1);   // a corner inside that leaf
            const int depth = t.descend(g, s, rx, ry, rz);
            float tile = 0.f;
            const uint8_t* lv = nullptr;
            if (depth == 0) lv = node(g, t.n0);
            else t.template valueAt<LEAF>(g, s, depth, rx, ry, rz, tile);     // one tile (or the background) covers the whole 8^3 block
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int dx = q >> 2, dy = (q >> 1) & 1, dz = (q ^ (q >> 1)) & 1;            // slot q -> corner offset
                if ((((dx << 2) | (dy << 1) | dz) & fmask) == cmb)
                    v[q] = lv ? leafValue<LEAF>(lv, (dx ? ox1 : ox0) | (dy ? oy1 : oy0) | (dz ? oz1 : oz0)) : tile;
            }
            cmb = (cmb - fmask) & fmask;                        // next sub-mask of fmask; wraps to 0 after fmask itself
        } while (cmb != 0);
    }
};

// ------------------------------------------------------------------------------------------------------------
// math::Ray<double> (math/Ray.h:26-295) and the grid's ScaleMap / ScaleTranslateMap (math/Maps.h:726-771,1255-1290)
// ------------------------------------------------------------------------------------------------------------
struct Ray {
    double ex, ey, ez, dx, dy, dz, ix, iy, iz, t0, t1;
    __device__ __forceinline__ void setDir(double x, double y, double z) { dx = x; dy = y; dz = z; ix = 1 / x; iy = 1 / y; iz = 1 / z; } // Ray.h:67-71
};

__device__ __forceinline__ double vlength(double x, double y, double z) { return sqrt(x * x + y * y + z * z); }  // math/Vec3.h:201-207
// Vec3::normalize (math/Vec3.h:363-371): isApproxEqual(d, 0, 1e-7) == !(|d-0| > 1e-7) (math/Math.h:428-431)
__device__ __forceinline__ void vnormalize(double& x, double& y, double& z)
{
    const double d = vlength(x, y, z);
    if (!(fabs(d - 0.0) > 1.0e-7)) return;
    const double s = 1.0 / d;
    x *= s; y *= s; z *= s;
}

// Ray::intersects(bbox,t0,t1) + clip (math/Ray.h:233-267).  pad = 0 for level sets (CoordBBox max as is),
// pad = 1 for VolumeRayIntersector (mBBox.max().offset(1), tools/RayIntersector.h:318).
__device__ __forceinline__ bool clipRay(Ray& r, const DevGrid& g, int pad)
{
    double a0 = r.t0, a1 = r.t1;
    {
        double a = (g.bbox_min[0] - r.ex) * r.ix, b = ((g.bbox_max[0] + pad) - r.ex) * r.ix;
        if (a > b) { const double t = a; a = b; b = t; }
        if (a > a0) a0 = a;
        if (b < a1) a1 = b;
        if (a0 > a1) return false;
    }
    {
        double a = (g.bbox_min[1] - r.ey) * r.iy, b = ((g.bbox_max[1] + pad) - r.ey) * r.iy;
        if (a > b) { const double t = a; a = b; b = t; }
        if (a > a0) a0 = a;
        if (b < a1) a1 = b;
        if (a0 > a1) return false;
    }
    {
        double a = (g.bbox_min[2] - r.ez) * r.iz, b = ((g.bbox_max[2] + pad) - r.ez) * r.iz;
        if (a > b) { const double t = a; a = b; b = t; }
        if (a > a0) a0 = a;
        if (b < a1) a1 = b;
        if (a0 > a1) return false;
    }
    r.t0 = a0; r.t1 = a1;
    return true;
}

// nanovdb::Map::applyMap / applyInverseMap / applyJacobian / applyInverseJacobian / applyIJT (NanoVDB.h:1473-1548) for general maps
__device__ __forceinline__ void mul3(const double* m, double& x, double& y, double& z)
{
    const double a = x * m[0] + y * m[1] + z * m[2], b = x * m[3] + y * m[4] + z * m[5], c = x * m[6] + y * m[7] + z * m[8];
    x = a; y = b; z = c;
}
__device__ __forceinline__ void mul3T(const double* m, double& x, double& y, double& z)
{
    const double a = x * m[0] + y * m[3] + z * m[6], b = x * m[1] + y * m[4] + z * m[7], c = x * m[2] + y * m[5] + z * m[8];
    x = a; y = b; z = c;
}

// Ray::worldToIndex == applyInverseMap (math/Ray.h:150-159)
__device__ __forceinline__ void worldToIndex(const DevGrid& g, Ray& r)
{
    double x, y, z, jx, jy, jz;
    if (g.general) {
        x = r.ex - g.trans[0]; y = r.ey - g.trans[1]; z = r.ez - g.trans[2];
        mul3(g.imat, x, y, z);
        jx = r.dx; jy = r.dy; jz = r.dz;
        mul3(g.imat, jx, jy, jz);
    } else {
        if (g.has_translation) { x = (r.ex - g.trans[0]) * g.inv[0]; y = (r.ey - g.trans[1]) * g.inv[1]; z = (r.ez - g.trans[2]) * g.inv[2]; }
        else { x = r.ex * g.inv[0]; y = r.ey * g.inv[1]; z = r.ez * g.inv[2]; }
        jx = r.dx * g.inv[0]; jy = r.dy * g.inv[1]; jz = r.dz * g.inv[2];
    }
    r.ex = x; r.ey = y; r.ez = z;
    const double len = vlength(jx, jy, jz);
    r.setDir(jx / len, jy / len, jz / len);
    r.t0 = len * r.t0; r.t1 = len * r.t1;
}
__device__ __forceinline__ void indexToWorldPos(const DevGrid& g, double& x, double& y, double& z)
{
    if (g.general) { mul3(g.mat, x, y, z); x += g.trans[0]; y += g.trans[1]; z += g.trans[2]; return; }
    if (g.has_translation) { x = x * g.scale[0] + g.trans[0]; y = y * g.scale[1] + g.trans[1]; z = z * g.scale[2] + g.trans[2]; }
    else { x = x * g.scale[0]; y = y * g.scale[1]; z = z * g.scale[2]; }
}
__device__ __forceinline__ void worldToIndexPos(const DevGrid& g, double& x, double& y, double& z)
{
    if (g.general) { x -= g.trans[0]; y -= g.trans[1]; z -= g.trans[2]; mul3(g.imat, x, y, z); return; }
    if (g.has_translation) { x = (x - g.trans[0]) * g.inv[0]; y = (y - g.trans[1]) * g.inv[1]; z = (z - g.trans[2]) * g.inv[2]; }
    else { x = x * g.inv[0]; y = y * g.inv[1]; z = z * g.inv[2]; }
}
// |J dir|: the factor between index-space and world-space times (getWorldTime, tools/RayIntersector.h:588-591)
__device__ __forceinline__ double jacobianLength(const DevGrid& g, double dx, double dy, double dz)
{
    if (g.general) { mul3(g.mat, dx, dy, dz); return vlength(dx, dy, dz); }
    return vlength(dx * g.scale[0], dy * g.scale[1], dz * g.scale[2]);
}

__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }   // std::min

// ------------------------------------------------------------------------------------------------------------
// One math::DDA<Ray,Log2Dim> (math/DDA.h:34-127) with the stride as a run-time shift.
// ------------------------------------------------------------------------------------------------------------
struct Dda {
    double t0, t1, nx, ny, nz;
    int vx, vy, vz;

    // double(+-2^shift) built from its bit pattern (exact; avoids an int->double conversion per step)
    __device__ __forceinline__ static double signedDim(int shift, bool positive) { return __hiloint2double(((1023 + shift) << 20) | (positive ? 0 : int(0x80000000u)), 0); }

    // DDA::init(ray, startTime, maxTime) (DDA.h:52-75)
    __device__ __forceinline__ void init(const Ray& r, double start, double maxT, int shift)
    {
        const int dim = 1 << shift;
        t0 = start; t1 = maxT;
        const double px = r.ex + r.dx * t0, py = r.ey + r.dy * t0, pz = r.ez + r.dz * t0;
        vx = int(floor(px)) & ~(dim - 1); vy = int(floor(py)) & ~(dim - 1); vz = int(floor(pz)) & ~(dim - 1);
        const double ax = (double(r.ix > 0 ? vx + dim : vx) - px) * r.ix, ay = (double(r.iy > 0 ? vy + dim : vy) - py) * r.iy;
        const double az = (double(r.iz > 0 ? vz + dim : vz) - pz) * r.iz;
        nx = r.dx == 0.0 ? DBL_MAX : t0 + ax;
        ny = r.dy == 0.0 ? DBL_MAX : t0 + ay;
        nz = r.dz == 0.0 ? DBL_MAX : t0 + az;
    }
    // DDA::step (DDA.h:83-90); MinIndex ties go to the largest index (math/Math.h:999-1007).
    // step/delta are rebuilt from the ray: step = +-DIM with the sign of 1/dir, delta = double(step) * (1/dir) = DIM * |1/dir| (DDA.h:61-73;
    // the same product, IEEE multiplication is symmetric in the signs).  An axis with dir == 0 has next == DBL_MAX (init) and can only
    // be the minimum when the other two are DBL_MAX as well, i.e. when this step fails (t0 = DBL_MAX > t1) and the DDA is dropped by
    // the caller -- so the reference's special case for it (step 0, delta DBL_MAX) needs no code here.
    // Branch-free: the three axes are handled by selects so that the lanes of a warp do not split on the axis.
    __device__ __forceinline__ bool step(const Ray& r, int shift)
    {
        const bool b1 = ny <= nx;
        double m = b1 ? ny : nx;
        const bool a2 = nz <= m;
        m = a2 ? nz : m;
        const bool a1 = b1 && !a2, a0 = !b1 && !a2;
        t0 = m;
        const double iv = a2 ? r.iz : (a1 ? r.iy : r.ix);
        const int dim = 1 << shift;
        const int st = __double2hiint(iv) < 0 ? -dim : dim;
        const double nn = m + __hiloint2double((1023 + shift) << 20, 0) * fabs(iv);     // mNext[axis] += mDelta[axis]
        nx = a0 ? nn : nx; ny = a1 ? nn : ny; nz = a2 ? nn : nz;
        vx += a0 ? st : 0; vy += a1 ? st : 0; vz += a2 ? st : 0;
        return t0 <= t1;
    }
    // DDA::next = math::Min(mT1, mNext[0], mNext[1], mNext[2]) (DDA.h:112, Math.h:734-738)
    __device__ __forceinline__ double next() const { return dmin(dmin(t1, nx), dmin(ny, nz)); }
};

struct DdaSave { double t1, nx, ny, nz; int vx, vy, vz; };   // a suspended parent level (its t0 is rewritten by step())
__device__ __forceinline__ void park(DdaSave& s, const Dda& d) { s.t1 = d.t1; s.nx = d.nx; s.ny = d.ny; s.nz = d.nz; s.vx = d.vx; s.vy = d.vy; s.vz = d.vz; }
__device__ __forceinline__ void unpark(Dda& d, const DdaSave& s) { d.t1 = s.t1; d.nx = s.nx; d.ny = s.ny; d.nz = s.nz; d.vx = s.vx; d.vy = s.vy; d.vz = s.vz; }

// ------------------------------------------------------------------------------------------------------------
// math::BoxStencil<FloatGrid> (math/Stencils.h:85-90,335-423)
// ------------------------------------------------------------------------------------------------------------
struct Stencil {
    int cx, cy, cz;
    float v[8];
    __device__ __forceinline__ void reset() { cx = cy = cz = 0x7fffffff; }    // BaseStencil: mCenter(Coord::max()) (:212)

    template<bool COUNT, int LEAF = kLeafFloat>
    __device__ __forceinline__ void moveTo(const DevGrid& g, const RootSmem& s, TreeCursor& acc, double x, double y, double z, Counters& c)
    {
        const int i = int(floor(x)), j = int(floor(y)), k = int(floor(z));
        if (i == cx && j == cy && k == cz) return;
        cx = i; cy = j; cz = k;
        if (COUNT) ++c.refills;
        acc.template fetchCell<false, LEAF>(g, s, i, j, k, v);
    }
    // interpolation(Vec3<float>) (:335-360): position converted to float first; every lerp in float
    __device__ __forceinline__ float interpolation(double x, double y, double z) const
    {
        const float u = float(x) - float(cx), vv = float(y) - float(cy), w = float(z) - float(cz);
        float V = v[0];
        float A = V + (v[1] - V) * w;
        V = v[3];
        float B = V + (v[2] - V) * w;
        const float C = A + (B - A) * vv;
        V = v[4];
        A = V + (v[5] - V) * w;
        V = v[7];
        B = V + (v[6] - V) * w;
        const float D = A + (B - A) * vv;
        return C + (D - C) * u;
    }
    // gradient(Vec3<float>) (:369-411) then applyIJT = * 1/scale in double, rounded back to float (Maps.h:767-771)
    __device__ __forceinline__ void gradient(const DevGrid& g, double x, double y, double z, float& gx, float& gy, float& gz) const
    {
        const float u = float(x) - float(cx), vv = float(y) - float(cy), w = float(z) - float(cz);
        float D0 = v[1] - v[0], D1 = v[2] - v[3], D2 = v[5] - v[4], D3 = v[6] - v[7];
        float A = D0 + (D1 - D0) * vv;
        float B = D2 + (D3 - D2) * vv;
        const float z_ = A + (B - A) * u;
        D0 = v[0] + D0 * w; D1 = v[3] + D1 * w; D2 = v[4] + D2 * w; D3 = v[7] + D3 * w;
        A = D0 + (D1 - D0) * vv;
        B = D2 + (D3 - D2) * vv;
        const float x_ = B - A;
        A = D1 - D0;
        B = D3 - D2;
        const float y_ = A + (B - A) * u;
        if (g.general) { double a = x_, b = y_, c = z_; mul3T(g.imat, a, b, c); gx = float(a); gy = float(b); gz = float(c); return; }   // applyIJT
        gx = float(double(x_) * g.inv[0]); gy = float(double(y_) * g.inv[1]); gz = float(double(z_) * g.inv[2]);
    }
};

// ------------------------------------------------------------------------------------------------------------
// LevelSetRayIntersector::intersects* = LinearSearchImpl (tools/RayIntersector.h:514-668) driven by
// LevelSetHDDA<Tree,2> (math/DDA.h:144-177), flattened into one loop over an explicit level variable.
// shift 12: root-level DDA probing upper nodes; 7: inside an upper node probing lower nodes; 3: inside a lower node
// probing leaves; 0: voxel DDA running the tester.
// ------------------------------------------------------------------------------------------------------------
struct LsHit { double time; int ix, iy, iz; double px, py, pz; float gx, gy, gz; };

// Suspended parent DDAs live in shared memory (structure of arrays, one column per thread): parking / resuming a level
// is a handful of ST.S / LD.S indexed by the level, instead of a three-way branch over 33 registers.
template<int THREADS>
struct WalkSmem {
    double t0[3][THREADS], t1[3][THREADS], nx[3][THREADS], ny[3][THREADS], nz[3][THREADS];
    int vx[3][THREADS], vy[3][THREADS], vz[3][THREADS];
    // t0 (the time the parent entered the cell it descends into) is only read when a ray is suspended inside a leaf: step() rewrites it
    __device__ __forceinline__ void park(int lvl, const Dda& d)
    {
        const int t = threadIdx.x;
        t0[lvl][t] = d.t0; t1[lvl][t] = d.t1; nx[lvl][t] = d.nx; ny[lvl][t] = d.ny; nz[lvl][t] = d.nz; vx[lvl][t] = d.vx; vy[lvl][t] = d.vy; vz[lvl][t] = d.vz;
    }
    __device__ __forceinline__ void unpark(int lvl, Dda& d) const
    {
        const int t = threadIdx.x;
        d.t1 = t1[lvl][t]; d.nx = nx[lvl][t]; d.ny = ny[lvl][t]; d.nz = nz[lvl][t]; d.vx = vx[lvl][t]; d.vy = vy[lvl][t]; d.vz = vz[lvl][t];
    }
};

// Traversal state of one ray.  lsAdvance() performs at most ONE cell probe, ONE level set-up, ONE stencil evaluation (two for the
// first one of a leaf visit, see lazyInit) and ONE DDA step, each of which exists exactly once in the instruction stream: all lanes
// of a warp run the same short loop body whatever level they are on (the warp-synchronous render loop reconverges after every
// phase), and the hot loop stays inside the instruction cache.  What is carried from one call to the next is kept small -- the
// child range of a descent and the time of a voxel's evaluation live inside one call only, the parents' DDAs in shared memory.
struct LsWalk {
    enum : uint32_t {
        kSkip = 1u,       // the current cell was already handled (we just came back up): only step
        kStep = 2u,       // the current cell is done: step
        kLazy = 4u,       // tester.init(T0) of this leaf visit has not been evaluated yet (V0 is not valid)
        kIdle = 8u        // no ray: the lane of a warp-synchronous loop that has nothing to advance (every phase tests it with its own flag)
    };
    Dda cur;
    double T0;            // LinearSearchImpl::mT[0]
    float V0;             // LinearSearchImpl::mV[0]
    int lvl;              // 0: root-level DDA (4096^3 cells) 1: inside an upper node (128^3) 2: inside a lower node (8^3) 3: voxels
    int shift;            // shiftOf(lvl): every DDA step wants it
    uint32_t f;           // the flags above in ONE register (bools cost a register and a mask each in the hot loop)

    __device__ __forceinline__ static int shiftOf(int lvl) { return (0x0003070C >> (8 * lvl)) & 0xff; }   // 12, 7, 3, 0
    __device__ __forceinline__ void setLevel(int l) { lvl = l; shift = shiftOf(l); }
    __device__ __forceinline__ void reset() { lvl = 0; shift = 12; f = 0u; T0 = 0.0; V0 = 0.f; }
    // `ray` must be the index-space ray already clipped to the node bbox (setIndexRay/setWorldRay, :548-562); sets up the root-level
    // DDA over the whole ray: LevelSetHDDA<TreeT, RootLevel>::test's math::DDA<Ray,Log2Dim> dda(tester.ray()) (DDA.h:150)
    __device__ __forceinline__ void begin(const Ray& ray) { reset(); cur.init(ray, ray.t0, ray.t1, 12); }
};

enum { kWalkContinue = 0, kWalkHit = 1, kWalkMiss = 2 };

// SYNC = true: the caller runs a warp-synchronous loop in which ALL 32 lanes call lsAdvance every iteration (lanes
// without a ray have LsWalk::kIdle set).  __syncwarp() between the phases makes the warp reconverge after each phase and
// stops the compiler from cloning the later phases per control-flow path.  SYNC = false: plain per-thread use.
// REFINE = true: LinearSearchImpl<GridT, Iterations> with Iterations = `iters` > 0 -- after the zero crossing the hit time is refined by
// `iters` secant steps, each one stencil evaluation at the current estimate (tools/RayIntersector.h:630-636).  A separate instantiation:
// the default kernels (Iterations = 0, what vdb_render and tools::rayTrace use) do not carry the loop.
//
// tester.init(dda.time()) (:597-601) evaluates the stencil at the leaf's entry time only to fill mT[0], mV[0], which nothing reads
// before the first voxel of the visit that passes the value gate (:622-624).  That evaluation is therefore made together with the
// first gated one (kLazy) and never for the leaf visits in which no voxel passes: 1.4 of 8.6 evaluations per ray on C2.  Same
// values in the same order -- the stencil's cache only makes moveTo cheaper, never changes what it returns.
template<bool COUNT, bool SYNC, int THREADS, bool REFINE = false, int LEAF = kLeafFloat>
__device__ __forceinline__ int lsAdvance(const DevGrid& g, const RootSmem& s, WalkSmem<THREADS>& sm,
                                         TreeCursor& acc, Stencil& st, const Ray& ray, float iso, float vmin, float vmax,
                                         LsWalk& w, LsHit& out, Counters& c, int iters = 0)
{
    Dda& cur = w.cur;
    int status = kWalkContinue;
    bool gate = false;                   // tester(ijk, t) of the voxel probed in this call is due, at time tq
    const uint32_t fIn = w.f; const int lvlIn = w.lvl;       // (diagnostics only)
    double tq;                           // only read under `gate`
    // ---- phase B: probe the current cell
    if (!(w.f & (LsWalk::kStep | LsWalk::kIdle))) {
        if (w.f & LsWalk::kSkip) w.f ^= LsWalk::kSkip | LsWalk::kStep;
        else {
            const int depth = acc.descend(g, s, cur.vx, cur.vy, cur.vz);
            if (w.lvl != 3) {
                // tester.hasNode<NodeT>(dda.voxel()) (:609-613): level 0 wants an upper node (depth <= 2), 1 a lower, 2 a leaf
                if (COUNT) { if (w.lvl == 0) ++c.root; else if (w.lvl == 1) ++c.upper; else ++c.lower; }
                if (depth <= 2 - w.lvl) {
                    // tester.setRange(dda.time(), dda.next()); recurse one level down (DDA.h:154-156)
                    const double c0 = cur.t0, c1 = cur.next();
                    sm.park(w.lvl, cur);
                    w.setLevel(w.lvl + 1);
                    // level set-up: math::DDA<Ray,Log2Dim> dda(tester.ray()) (DDA.h:150,172).  Done right here: the lanes that would run it
                    // as a phase of its own after a reconvergence are exactly the lanes in this branch (measured: 2.3 % faster than the phase)
                    cur.init(ray, c0, c1, w.shift);
                    if (w.lvl == 3) { w.f |= LsWalk::kLazy; w.T0 = c0; }   // tester.init(dda.time()) (:597-601), evaluated on demand
                } else w.f |= LsWalk::kStep;
            } else {
                // LinearSearchImpl::operator()(ijk, time) with time = dda.next() (:620-644)
                if (COUNT) ++c.voxel;
                float V;
                bool on;
                // a voxel of a leaf: the value comes from the leaf's halo block when there is one (the same values; the stencil
                // reads them from there too, so the leaf's own value array stays out of the caches), the active bit from the leaf
                uint32_t hi = 0xffffffffu;
                if (depth == 0 && g.halo) hi = (acc.n0 - g.leaf0) / LeafKind<LEAF>::handles;
                if (hi < g.leaf_count) {
                    V = haloValue<LEAF>(g.halo, hi, uint32_t(cur.vx & 7), uint32_t(cur.vy & 7), uint32_t(cur.vz & 7));
                    on = maskBit(TreeCursor::node(g, acc.n0) + kLeafVMask, leafOffset(cur.vx, cur.vy, cur.vz));
                } else on = acc.template valueAt<LEAF>(g, s, depth, cur.vx, cur.vy, cur.vz, V);
                if (on && V > vmin && V < vmax) { gate = true; tq = cur.next(); }
                w.f |= LsWalk::kStep;
            }
        }
    }
    if (COUNT && SYNC) {
        const bool probed = !(fIn & (LsWalk::kStep | LsWalk::kSkip | LsWalk::kIdle));
        const bool node = probed && lvlIn != 3, vox = probed && lvlIn == 3;
        c.diag[0] += node; c.diag[1] += vox;
        const unsigned bn = __ballot_sync(0xffffffffu, node), bv = __ballot_sync(0xffffffffu, vox);
        if ((threadIdx.x & 31) == 0) { c.diag[4] += bn != 0u; c.diag[5] += bv != 0u; }
        c.diag[2] += gate;
        const unsigned b = __ballot_sync(0xffffffffu, gate);
        if ((threadIdx.x & 31) == 0) c.diag[6] += b != 0u;
    }
    if (SYNC) __syncwarp();
    // ---- phase C: stencil evaluation: interpValue(time) (:652-657): pos = ray(time); stencil.moveTo(pos); interpolation(pos) - iso
    if (gate) {
#pragma unroll 1
        for (;;) {
            const bool initPass = (w.f & LsWalk::kLazy) != 0u;
            const double te = initPass ? w.T0 : tq;
            const double px = ray.ex + ray.dx * te, py = ray.ey + ray.dy * te, pz = ray.ez + ray.dz * te;
            st.template moveTo<COUNT, LEAF>(g, s, acc, px, py, pz, c);
            const float V1 = st.interpolation(px, py, pz) - iso;
            if (initPass) { w.V0 = V1; w.f &= ~LsWalk::kLazy; continue; }          // mV[0] = interpValue(mT[0])
            if (w.V0 * V1 <= 0.0f) {                                              // math::ZeroCrossing (math/Math.h:821)
                // The crossing.  The walk ends here; what is left of operator() -- mTime = interpTime(), the refinements -- and
                // getWorldPosAndNml are lsFinishHit(), which the caller runs when it finishes the ray (the render kernel: after the tile's
                // loop, for all its pixels together).  ONE register carries the hit to it: mV[1] in out.gx; mT[0] / mV[0] are w.T0 / w.V0,
                // mT[1] is the next() and the hit voxel the position of the DDA, which stays where it is.
                out.gx = V1;
                w.f &= ~LsWalk::kStep;
                status = kWalkHit;
            } else { w.T0 = tq; w.V0 = V1; }                                      // no crossing: slide
            break;
        }
    }
    if (SYNC) __syncwarp();
    // ---- phase D: while (dda.step()) ... ; an exhausted level returns false to its parent (DDA.h:158-159,174-175)
    if (COUNT && SYNC) {
        const bool sp = (w.f & (LsWalk::kStep | LsWalk::kIdle)) == LsWalk::kStep;
        c.diag[3] += sp;
        const unsigned b = __ballot_sync(0xffffffffu, sp);
        if ((threadIdx.x & 31) == 0) c.diag[7] += b != 0u;
    }
    if ((w.f & (LsWalk::kStep | LsWalk::kIdle)) == LsWalk::kStep) {
        w.f &= ~LsWalk::kStep;
        if (!cur.step(ray, w.shift)) {
            if (w.lvl == 0) status = kWalkMiss;
            else {
                // the parent steps right away (one level per call; a second exhausted level waits for the next call)
                w.setLevel(w.lvl - 1); sm.unpark(w.lvl, cur);
                if (!cur.step(ray, w.shift)) {
                    if (w.lvl == 0) status = kWalkMiss;
                    else { w.setLevel(w.lvl - 1); sm.unpark(w.lvl, cur); w.f |= LsWalk::kSkip; }
                }
            }
        }
    }
    return status;
}

// The rest of a hit, after lsAdvance() has returned kWalkHit: LinearSearchImpl<GridT, Iterations>'s secant refinements of the hit time
// (tools/RayIntersector.h:630-636; REFINE instantiations only, `iters` of them, each one stencil evaluation at the current estimate) and
// getWorldPosAndNml (:575-582): position and stencil gradient at the hit time.
template<bool COUNT, bool REFINE, int LEAF = kLeafFloat>
__device__ __forceinline__ void lsFinishHit(const DevGrid& g, const RootSmem& s, TreeCursor& acc, Stencil& st, const Ray& ray, float iso,
                                            const LsWalk& w, LsHit& out, Counters& c, int iters = 0)
{
    // mT[0..1], mV[0..1] of the crossing: mT[1] is the voxel's exit time again (the DDA has not moved since), mV[1] was left in out.gx
    double rT0 = w.T0, rT1 = w.cur.next();
    float rV0 = w.V0, rV1 = out.gx;
    double tq = rT0 + (rT1 - rT0) * rV0 / (rV0 - rV1);                            // interpTime (:646-650): float difference promoted to double
    out.time = tq;
    out.ix = w.cur.vx; out.iy = w.cur.vy; out.iz = w.cur.vz;
    if (REFINE) {
#pragma unroll 1
        for (int n = 0; n < iters; ++n) {
            // V = interpValue(mTime); m = ZeroCrossing(mV[0], V); mV[m] = V; mT[m] = mTime; mTime = interpTime() (:631-635)
            const double px = ray.ex + ray.dx * tq, py = ray.ey + ray.dy * tq, pz = ray.ez + ray.dz * tq;
            st.template moveTo<COUNT, LEAF>(g, s, acc, px, py, pz, c);
            const float V = st.interpolation(px, py, pz) - iso;
            if (rV0 * V <= 0.0f) { rV1 = V; rT1 = tq; } else { rV0 = V; rT0 = tq; }
            tq = rT0 + (rT1 - rT0) * rV0 / (rV0 - rV1);
        }
        out.time = tq;
    }
    const double px = ray.ex + ray.dx * tq, py = ray.ey + ray.dy * tq, pz = ray.ez + ray.dz * tq;
    st.template moveTo<COUNT, LEAF>(g, s, acc, px, py, pz, c);
    out.px = px; out.py = py; out.pz = pz;
    st.gradient(g, px, py, pz, out.gx, out.gy, out.gz);
}

// plain per-thread form (arbitrary-ray batches)
template<bool COUNT, int THREADS, int LEAF = kLeafFloat>
__device__ __forceinline__ bool intersectLevelSet(const DevGrid& g, const RootSmem& s, WalkSmem<THREADS>& sm, TreeCursor& acc, Stencil& st, Ray& ray,
                                                  float iso, float vmin, float vmax, LsHit& out, Counters& c, int iters = 0)
{
    LsWalk w; w.begin(ray);
#pragma unroll 1
    for (;;) {
        const int r = lsAdvance<COUNT, false, THREADS, true, LEAF>(g, s, sm, acc, st, ray, iso, vmin, vmax, w, out, c, iters);
        if (r == kWalkHit) lsFinishHit<COUNT, true, LEAF>(g, s, acc, st, ray, iso, w, out, c, iters);
        if (r != kWalkContinue) return r == kWalkHit;
    }
}

// ------------------------------------------------------------------------------------------------------------
// VolumeRayIntersector::hits == VolumeHDDA<BoolTree,Ray,2>::hits (math/DDA.h:210-217,247-264,321-336) as a
// resumable generator: next() returns the spans in the order the reference pushes them into its std::vector.
// The bool topology copy the reference walks (tools/RayIntersector.h:299-319, dilation 0) has the float tree's
// child topology and active states, so the float tree itself is probed.
// ------------------------------------------------------------------------------------------------------------
// Parking area of the fog kernel (structure of arrays, one column per thread): slots 0,1 = suspended parents of the
// primary walk, 2,3 = suspended parents of the shadow walk, 4 = the primary walk's current DDA while a shadow ray runs.
template<int THREADS>
struct FogSmem {
    double t1[5][THREADS], nx[5][THREADS], ny[5][THREADS], nz[5][THREADS];
    int vx[5][THREADS], vy[5][THREADS], vz[5][THREADS];
    double ray[6][THREADS];           // eye, dir of the primary index-space ray while a shadow ray is active (1/dir is recomputed)
    double dt0[THREADS];              // primary Dda::t0
    double ts0[THREADS], topT1[THREADS], tcur[THREADS], tend[THREADS], bound[THREADS];
    int misc[THREADS];                // primary walk: lvl | needStep << 8
    double sbase[8];                  // shadow ray constants shared by the CTA: dir xyz, inv xyz, t0, t1 (index space)
    __device__ __forceinline__ void park(int slot, const Dda& d)
    {
        const int t = threadIdx.x;
        t1[slot][t] = d.t1; nx[slot][t] = d.nx; ny[slot][t] = d.ny; nz[slot][t] = d.nz; vx[slot][t] = d.vx; vy[slot][t] = d.vy; vz[slot][t] = d.vz;
    }
    __device__ __forceinline__ void unpark(int slot, Dda& d) const
    {
        const int t = threadIdx.x;
        d.t1 = t1[slot][t]; d.nx = nx[slot][t]; d.ny = ny[slot][t]; d.nz = nz[slot][t]; d.vx = vx[slot][t]; d.vy = vy[slot][t]; d.vz = vz[slot][t];
    }
};

enum { kSpanContinue = 0, kSpanEmit = 1, kSpanDone = 2 };

// One unit of VolumeHDDA::hits per call (at most one DDA set-up or step and one cell probe, each a single site).
struct SpanWalk {
    Dda cur;
    double ts0, topT1;     // open span start (<0: none), maxTime of the root-level DDA
    double bound;          // entry time of the last probed cell: while a span is open its end cannot be earlier than this
    int lvl;               // 0 root-level DDA (4096^3), 1 inside an upper node (128^3), 2 inside a lower node (8^3); -1 = finished
    bool needStep;

    __device__ __forceinline__ static int shiftOf(int lvl) { return (0x0003070C >> (8 * lvl)) & 0xff; }
    // the root-level DDA is set up here, a child's DDA in the call that finds the child (math::DDA<RayT,Log2Dim> dda(ray), DDA.h:252,326)
    __device__ __forceinline__ void begin(const Ray& ray)
    {
        lvl = 0; needStep = false; ts0 = -1.0; topT1 = ray.t1; bound = ray.t0;
        cur.init(ray, ray.t0, ray.t1, 12);
    }
    // slotBase: first parking slot of this walk's parents (0 primary, 2 shadow)
    template<bool COUNT, class SM>
    __device__ __forceinline__ int advance(const DevGrid& g, const RootSmem& s, SM& sm, int slotBase, TreeCursor& acc, const Ray& ray,
                                           double& a, double& b, Counters& c)
    {
        if (lvl < 0) return kSpanDone;
        if (needStep) {
            if (!cur.step(ray, shiftOf(lvl))) {
                // a level is exhausted: "if (t.t0>=0) t.t1 = mDDA.maxTime()" -- only the outermost assignment survives
                // because any later close overwrites t1 (DDA.h:263,335)
                if (lvl == 0) {
                    lvl = -1;
                    if (ts0 >= 0.0) { a = ts0; b = topT1; ts0 = -1.0; return (b - a) > 1e-9 ? kSpanEmit : kSpanDone; }   // TimeSpan::valid (Ray.h:48)
                    return kSpanDone;
                }
                --lvl; sm.unpark(slotBase + lvl, cur);
                return kSpanContinue;                   // the parent steps on the next call
            }
        }
        needStep = true;
        bound = cur.t0;
        bool active;
        if (lvl == 2 && g.lowmask && acc.n1 && !(uint32_t((cur.vx ^ acc.kx) | (cur.vy ^ acc.ky) | (cur.vz ^ acc.kz)) & ~127u)) {
            // a cell of the lower node the cursor is in (the usual case; rounding can push a cell outside): one mask word
            if (COUNT) ++c.lower;
            const uint32_t n = lowerOffset(cur.vx, cur.vy, cur.vz);
            const uint32_t i = min((acc.n1 - g.lower0) / 1058u, g.lower_count - 1u);
            active = (__ldg(g.lowmask + ((size_t(i) << 6) | (n >> 6))) >> (n & 63u)) & 1ull;
        } else {
            const int depth = acc.descend(g, s, cur.vx, cur.vy, cur.vz);
            if (COUNT) { if (lvl == 0) ++c.root; else if (lvl == 1) ++c.upper; else ++c.lower; }
            if (lvl < 2 && depth <= 2 - lvl) {              // child node: walk it
                const double c0 = cur.t0, c1 = cur.next();       // ray.setTimes(time(), next()) (DDA.h:252,326)
                sm.park(slotBase + lvl, cur);
                ++lvl; needStep = false;
                cur.init(ray, c0, c1, shiftOf(lvl));
                return kSpanContinue;
            }
            // leaf level: any existing leaf counts as active (DDA.h:308-309,326-327); otherwise the tile's state
            active = (lvl == 2 && depth == 0) ? true : acc.activeAt(g, s, depth, cur.vx, cur.vy, cur.vz);
        }
        if (active) {
            if (ts0 < 0.0) ts0 = cur.t0;
        } else if (ts0 >= 0.0) {
            a = ts0; b = cur.t0; ts0 = -1.0;
            if ((b - a) > 1e-9) return kSpanEmit;
        }
        return kSpanContinue;
    }
};

// tools::BoxSampler::sample through GridSampler::wsSample (tools/Interpolation.h:420-425,712-762): corner values are
// float, each lerp is a + float((b-a) * w) with a DOUBLE weight.
__device__ __forceinline__ float lerpBox(float a, float b, double w) { const double t = (b - a) * w; return a + float(t); }
template<int LEAF = kLeafFloat>
__device__ __forceinline__ float boxSampleWorld(const DevGrid& g, const RootSmem& s, TreeCursor& acc, double wx, double wy, double wz)
{
    worldToIndexPos(g, wx, wy, wz);
    const int i = int(floor(wx)), j = int(floor(wy)), k = int(floor(wz));
    const double u = wx - i, v = wy - j, w = wz - k;
    float d[8];   // 000,001,011,010,100,101,111,110
    acc.template fetchCell<true, LEAF>(g, s, i, j, k, d);
    return lerpBox(lerpBox(lerpBox(d[0], d[1], w), lerpBox(d[3], d[2], w), v),
                   lerpBox(lerpBox(d[4], d[5], w), lerpBox(d[7], d[6], w), v), u);
}

// ------------------------------------------------------------------------------------------------------------
// cameras (tools/RayTracer.h:351-513) on the flattened POD, shaders (:565-753)
// ------------------------------------------------------------------------------------------------------------
struct DevCamera {
    uint32_t kind, width, height, pad;
    double m[16];
    double eye[3], dir[3];
    double scale_w, scale_h, t0, t1;
};
// A NanoGrid<Vec3f> that colours the surface (the GridT = Vec3SGrid forms of the shaders, tools/RayTracer.h:542-725).
// Node layout of the Vec3f build (nanovdb/NanoVDB.h, verified with a sizeof/offsetof probe): RootData 96 B + 32 B tiles
// (value at +20), internal nodes: same mask offsets as the float build, 16-byte table entries (Vec3f value | int64 child),
// leaves: value mask at +16, 512 x 12 B values at +128.
constexpr uint32_t kColRootTiles = 96, kColTileValue = 20, kColUpperTable = 8256, kColLowerTable = 1088, kColLeafValues = 128;
struct DevColor {
    const uint8_t* base;      // null: constant colour
    const uint8_t* tiles;
    uint64_t root_off;
    uint32_t table_size, has_translation;
    float    background[3];
    double   inv[3], trans[3];
};
struct DevShader { uint32_t kind; float r, g, b, a; double bmin[3], inv[3]; DevColor col; };

// tools::PointSampler::sample(acc, xform.worldToIndex(xyz), v) (tools/Interpolation.h:600-617): probeValue at the nearest
// voxel, ::round half away from zero; the value of a tile or the background where the tree has no voxel
__device__ __forceinline__ void colorAt(const DevColor& c, double wx, double wy, double wz, float v[3])
{
    if (c.has_translation) { wx = (wx - c.trans[0]) * c.inv[0]; wy = (wy - c.trans[1]) * c.inv[1]; wz = (wz - c.trans[2]) * c.inv[2]; }
    else { wx = wx * c.inv[0]; wy = wy * c.inv[1]; wz = wz * c.inv[2]; }
    const int x = int(round(wx)), y = int(round(wy)), z = int(round(wz));
    const uint8_t* p = nullptr;                                     // where the three floats are
    const unsigned long long key = rootKey(x, y, z);
    for (uint32_t i = 0; i < c.table_size; ++i) {
        const uint8_t* t = c.tiles + kTileSize * i;
        if (ldg64(t) != key) continue;
        const long long child = ldgs64(t + 8);
        if (!child) { p = t + kColTileValue; break; }
        const uint8_t* u = c.base + c.root_off + child;
        uint32_t n = upperOffset(x, y, z);
        if (!maskBit(u + kUpperCMask, n)) { p = u + kColUpperTable + 16u * n; break; }
        const uint8_t* l = u + ldgs64(u + kColUpperTable + 16u * n);
        n = lowerOffset(x, y, z);
        if (!maskBit(l + kLowerCMask, n)) { p = l + kColLowerTable + 16u * n; break; }
        const uint8_t* f = l + ldgs64(l + kColLowerTable + 16u * n);
        p = f + kColLeafValues + 12u * leafOffset(x, y, z);
        break;
    }
    if (p) { v[0] = ldgf(p); v[1] = ldgf(p + 4); v[2] = ldgf(p + 8); }
    else { v[0] = c.background[0]; v[1] = c.background[1]; v[2] = c.background[2]; }
}

__device__ __forceinline__ void cameraRay(const DevCamera& c, uint32_t i, uint32_t j, double io, double jo, Ray& ray)
{
    ray.ex = c.eye[0]; ray.ey = c.eye[1]; ray.ez = c.eye[2];
    ray.t0 = c.t0; ray.t1 = c.t1;
    // rasterToScreen (:391-395)
    const double sx = (2 * (double(i) + io) / double(c.width) - 1) * c.scale_w;
    const double sy = (1 - 2 * (double(j) + jo) / double(c.height)) * c.scale_h;
    if (c.kind == 0) {                                            // PerspectiveCamera::getRay (:452-462)
        const double sz = -1.0;
        double x = sx * c.m[0] + sy * c.m[4] + sz * c.m[8];       // Mat4::transform3x3 (math/Mat4.h:1070-1076)
        double y = sx * c.m[1] + sy * c.m[5] + sz * c.m[9];
        double z = sx * c.m[2] + sy * c.m[6] + sz * c.m[10];
        vnormalize(x, y, z);
        const double sc = 1.0 / (x * c.dir[0] + y * c.dir[1] + z * c.dir[2]);
        ray.t0 *= sc; ray.t1 *= sc;                               // scaleTimes
        ray.setDir(x, y, z);
    } else {                                                      // OrthographicCamera::getRay (:505-512)
        const double sz = 0.0;
        ray.ex = sx * c.m[0] + sy * c.m[4] + sz * c.m[8] + c.m[12];   // Vec3 * Mat4 (math/Mat4.h:1180-1188)
        ray.ey = sx * c.m[1] + sy * c.m[5] + sz * c.m[9] + c.m[13];
        ray.ez = sx * c.m[2] + sy * c.m[6] + sz * c.m[10] + c.m[14];
        ray.setDir(c.dir[0], c.dir[1], c.dir[2]);
    }
}

__device__ __forceinline__ float4 shade(const DevShader& s, double wx, double wy, double wz, double nx, double ny, double nz,
                                        double dx, double dy, double dz)
{
    if (s.col.base) {
        // colour from a Vec3f grid sampled at the hit position (GridT = Vec3SGrid, SamplerType = PointSampler)
        float v[3];
        colorAt(s.col, wx, wy, wz, v);
        switch (s.kind) {
        case 0: return make_float4(v[0], v[1], v[2], 1.0f);                                         // MatteShader<GridT> (:542-562)
        case 1: return make_float4(float(v[0] * (nx + 1.0)), float(v[1] * (ny + 1.0)), float(v[2] * (nz + 1.0)), 1.0f);   // NormalShader<GridT> (:591-611)
        case 2: {                                                                                    // PositionShader<GridT> (:640-668)
            const double rx = (wx - s.bmin[0]) * s.inv[0], ry = (wy - s.bmin[1]) * s.inv[1], rz = (wz - s.bmin[2]) * s.inv[2];
            return make_float4(v[0] * float(rx), v[1] * float(ry), v[2] * float(rz), 1.0f);
        }
        default: {                                                                                   // DiffuseShader<GridT> (:702-725)
            const float f = float(fabs(nx * dx + ny * dy + nz * dz));
            return make_float4(v[0] * f, v[1] * f, v[2] * f, 1.0f);
        }
        }
    }
    switch (s.kind) {
    case 0: return make_float4(s.r, s.g, s.b, s.a);                                              // MatteShader (:565-581)
    case 1: {                                                                                    // NormalShader (:614-630)
        const float r = s.r * 0.5f, g = s.g * 0.5f, b = s.b * 0.5f;
        return make_float4(r * float(nx + 1.0), g * float(ny + 1.0), b * float(nz + 1.0), 1.0f);
    }
    case 2: {                                                                                    // PositionShader (:671-690)
        const double rx = (wx - s.bmin[0]) * s.inv[0], ry = (wy - s.bmin[1]) * s.inv[1], rz = (wz - s.bmin[2]) * s.inv[2];
        return make_float4(s.r * float(rx), s.g * float(ry), s.b * float(rz), 1.0f);
    }
    default: {                                                                                   // DiffuseShader (:728-753)
        const float f = float(fabs(nx * dx + ny * dy + nz * dz));
        return make_float4(s.r * f, s.g * f, s.b * f, 1.0f);
    }
    }
}

} // namespace vdbrt
