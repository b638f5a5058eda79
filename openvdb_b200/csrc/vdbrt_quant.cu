// vdbrt_quant.cu -- quantised grids (NanoGrid<Fp4|Fp8|Fp16|FpN>) become NanoGrid<float> while they are
// uploaded.
//
// The render kernels read leaf values as plain floats.  Dequantising inside the voxel fetch was built and measured: the
// launch-uniform branch and its eight inlined copies in the stencil fetch cost 32 % on FLOAT grids (C2 3.35 -> 4.43 ms,
// profiles/r01_summary.md), so the leaves are expanded ONCE, here, with the arithmetic of the reference's accessors
//     LeafData<FpX>::getValue(i) = float(code_i) * mQuantum + mMinimum      (nanovdb/nanovdb/NanoVDB.h:3843,3876,3906,3961)
// (a float product rounded, then a float sum rounded: compiled --fmad=false, written with __fmul_rn/__fadd_rn anyway).
// That is also what the reference itself does before it ray-traces such a grid: nanovdb::tools::nanoToOpenVDB reads every
// voxel through getValue into a FloatGrid (nanovdb/nanovdb/tools/NanoToOpenVDB.h:511-518), so the frames are bit-identical.
//
// Only the leaves differ between NanoGrid<FpX> and NanoGrid<float>: ValueType is float for all of them, so GridData,
// TreeData, the root and both internal levels have the same layout, and LeafFnBase (NanoVDB.h:3752-3811) has the size of the
// float leaf's header (96 B).  The expanded buffer is therefore
//     [0, leafOffset)  copied, with GridData retyped and the lower nodes' child offsets re-pointed
//     leaf i at leafOffset + 2144 * i, i = rank of the source leaf's address among all leaves
// FpN leaves have different sizes (96 + 64 * bitWidth bytes, NanoVDB.h:3934), so the source address of leaf i is not
// arithmetic: the addresses are collected from the lower nodes' tables and sorted (cub radix sort), which also makes no
// assumption on the order the leaves were serialised in.
#include "vdbrt_host.h"
#include <cub/device/device_radix_sort.cuh>
#include <cstring>

namespace vdbrt {
namespace {

constexpr uint32_t kLowerSize = 33856, kLeafSize = 2144, kLeafHeader = 96;
constexpr uint32_t kTypeFp4 = 13, kTypeFp8 = 14, kTypeFp16 = 15, kTypeFpN = 16;          // nanovdb::GridType, NanoVDB.h:232-235

// status word written by the kernels: bit 0 a child offset points outside the leaf area, bit 1 leaves overlap / leave gaps,
// bit 2 a bit width this library does not know, bit 3 a leaf reaches past the end of the buffer
struct QuantCtl { unsigned int count, status; };

// one thread per lower-node slot: the source address (in 32-byte units from the first leaf) of every child leaf
__global__ void k_quant_collect(const uint8_t* __restrict__ src, uint64_t lowerOff, uint32_t lowerCount, uint64_t leafOff, uint64_t leafEnd,
                                uint32_t* __restrict__ keys, uint32_t capacity, QuantCtl* ctl)
{
    const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t node = uint32_t(t >> 12), slot = uint32_t(t & 4095);
    if (node >= lowerCount) return;
    const uint64_t nodeOff = lowerOff + uint64_t(node) * kLowerSize;
    const uint64_t word = *reinterpret_cast<const uint64_t*>(src + nodeOff + kLowerCMask + 8 * (slot >> 6));
    if (!((word >> (slot & 63)) & 1)) return;
    const int64_t child = *reinterpret_cast<const int64_t*>(src + nodeOff + kLowerTable + 8 * slot);
    const uint64_t addr = nodeOff + uint64_t(child);
    if (addr < leafOff || addr + kLeafHeader > leafEnd || ((addr - leafOff) & 31)) { atomicOr(&ctl->status, 1u); return; }
    const unsigned int pos = atomicAdd(&ctl->count, 1u);
    if (pos < capacity) keys[pos] = uint32_t((addr - leafOff) >> 5);
}

__device__ inline uint32_t rankOf(const uint32_t* __restrict__ keys, uint32_t n, uint32_t key)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// one thread per lower-node slot: child offsets of the expanded buffer (relative to the lower node, NanoVDB.h:3190-3199)
__global__ void k_quant_relink(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t lowerOff, uint32_t lowerCount, uint64_t leafOff,
                               const uint32_t* __restrict__ keys, uint32_t leafCount)
{
    const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t node = uint32_t(t >> 12), slot = uint32_t(t & 4095);
    if (node >= lowerCount) return;
    const uint64_t nodeOff = lowerOff + uint64_t(node) * kLowerSize;
    const uint64_t word = *reinterpret_cast<const uint64_t*>(src + nodeOff + kLowerCMask + 8 * (slot >> 6));
    if (!((word >> (slot & 63)) & 1)) return;
    const int64_t child = *reinterpret_cast<const int64_t*>(src + nodeOff + kLowerTable + 8 * slot);
    const uint32_t key = uint32_t((nodeOff + uint64_t(child) - leafOff) >> 5);
    const uint32_t idx = rankOf(keys, leafCount, key);
    *reinterpret_cast<int64_t*>(dst + nodeOff + kLowerTable + 8 * slot) = int64_t(leafOff + uint64_t(idx) * kLeafSize) - int64_t(nodeOff);
}

// one 128-thread block per leaf, four consecutive voxels per thread (one 16-byte store)
__global__ void __launch_bounds__(128) k_quant_expand(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t leafOff, uint64_t leafEnd,
                                                      const uint32_t* __restrict__ keys, uint32_t leafCount, uint32_t gridType, QuantCtl* ctl)
{
    const uint32_t leaf = blockIdx.x;
    const uint64_t from = leafOff + (uint64_t(keys[leaf]) << 5);
    const uint8_t* in = src + from;
    const uint32_t head3 = *reinterpret_cast<const uint32_t*>(in + 12);               // mBBoxDif[3], mFlags
    uint32_t log2w;                                                                   // bits per code = 1 << log2w
    if (gridType == kTypeFp4) log2w = 2; else if (gridType == kTypeFp8) log2w = 3; else if (gridType == kTypeFp16) log2w = 4;
    else log2w = head3 >> 29;                                                         // FpN: mFlags >> 5 (NanoVDB.h:3933)
    if (log2w > 4) { if (threadIdx.x == 0) atomicOr(&ctl->status, 4u); return; }
    const uint64_t size = kLeafHeader + (uint64_t(64) << log2w);
    if (from + size > leafEnd) { if (threadIdx.x == 0) atomicOr(&ctl->status, 8u); return; }
    if (threadIdx.x == 0) {                                                          // the leaves must tile the leaf area
        const uint64_t next = leaf + 1 < leafCount ? leafOff + (uint64_t(keys[leaf + 1]) << 5) : from + size;
        if (next != from + size || (leaf == 0 && from != leafOff)) atomicOr(&ctl->status, 2u);
    }
    uint8_t* out = dst + leafOff + uint64_t(leaf) * kLeafSize;
    const float minimum = *reinterpret_cast<const float*>(in + 80), quantum = *reinterpret_cast<const float*>(in + 84);
    if (threadIdx.x < 20) {                                                          // bbox min, dif + flags, value mask
        uint32_t w = reinterpret_cast<const uint32_t*>(in)[threadIdx.x];
        if (threadIdx.x == 3) w &= 0x1fffffffu;                                       // the float leaf has no bit width in its flags
        reinterpret_cast<uint32_t*>(out)[threadIdx.x] = w;
    } else if (threadIdx.x < 24) {                                                   // LeafFnBase::getMin/getMax/getAvg/getDev (:3789-3799)
        const uint32_t s = threadIdx.x - 20;
        const float q = float(*reinterpret_cast<const uint16_t*>(in + 88 + 2 * s));
        reinterpret_cast<float*>(out + 80)[s] = s < 3 ? __fadd_rn(__fmul_rn(q, quantum), minimum) : __fmul_rn(q, quantum);
    }
    // voxels 4t .. 4t+3: their codes are 4 << log2w consecutive bits, which never straddle a 64-bit word
    const uint32_t bit = (threadIdx.x * 4u) << log2w;
    const uint64_t word = *reinterpret_cast<const uint64_t*>(in + kLeafHeader + 8 * (bit >> 6));
    const uint32_t w = 1u << log2w;
    const uint64_t mask = (uint64_t(1) << w) - 1;
    const uint64_t bits = word >> (bit & 63);
    float4 v;
    v.x = __fadd_rn(__fmul_rn(float(uint32_t(bits & mask)), quantum), minimum);
    v.y = __fadd_rn(__fmul_rn(float(uint32_t((bits >> w) & mask)), quantum), minimum);
    v.z = __fadd_rn(__fmul_rn(float(uint32_t((bits >> (2 * w)) & mask)), quantum), minimum);
    v.w = __fadd_rn(__fmul_rn(float(uint32_t((bits >> (3 * w)) & mask)), quantum), minimum);
    reinterpret_cast<float4*>(out + kLeafHeader)[threadIdx.x] = v;
}

template<typename T> T rd(const uint8_t* p) { T v; std::memcpy(&v, p, sizeof(T)); return v; }
template<typename T> void wr(uint8_t* p, T v) { std::memcpy(p, &v, sizeof(T)); }

} // namespace

bool isQuantisedType(uint32_t gridType) { return gridType >= kTypeFp4 && gridType <= kTypeFpN; }

// src: the whole quantised grid in device memory; head: host copy of its GridData + TreeData (736 bytes).
// On success *outDev is a cudaMalloc'ed NanoGrid<float> of *outBytes bytes (complete when ctx->stream has drained).
int expandQuantised(vdbrt_ctx* ctx, const uint8_t* src, uint64_t srcBytes, const uint8_t* head, uint8_t** outDev, uint64_t* outBytes)
{
    constexpr uint64_t kGridSize = 672;
    const uint32_t gridType = rd<uint32_t>(head + 636);
    const uint64_t gridBytes = rd<uint64_t>(head + 32);
    if (gridBytes > srcBytes) return setError(VDBRT_ERR_BAD_GRID, "grid size exceeds the buffer");
    const uint8_t* tree = head + kGridSize;
    const uint64_t leafOff = kGridSize + uint64_t(rd<int64_t>(tree + 0)), lowerOff = kGridSize + uint64_t(rd<int64_t>(tree + 8));
    const uint32_t leafCount = rd<uint32_t>(tree + 32), lowerCount = rd<uint32_t>(tree + 36);
    const uint32_t blindCount = rd<uint32_t>(head + 648);
    const uint64_t blindOff = uint64_t(rd<int64_t>(head + 640));
    const uint64_t leafEnd = (blindCount && blindOff >= leafOff && blindOff <= gridBytes) ? blindOff : gridBytes;   // blind data follows the leaves
    if (leafOff > leafEnd || (leafOff & 31) || lowerOff + uint64_t(lowerCount) * kLowerSize > leafOff)
        return setError(VDBRT_ERR_BAD_GRID, "node offsets of the quantised grid are inconsistent");
    if (((leafEnd - leafOff) >> 5) > 0xffffffffull) return setError(VDBRT_ERR_UNSUPPORTED, "quantised grids above 128 GB are not supported");

    const uint64_t dstBytes = leafOff + uint64_t(leafCount) * kLeafSize;
    uint8_t* dst = nullptr;
    uint32_t* keys = nullptr;       // [0,n) collected, [n,2n) sorted
    QuantCtl* ctl = nullptr;
    void* tmp = nullptr;
    auto release = [&](bool all) { if (all) cudaFree(dst); cudaFree(keys); cudaFree(ctl); cudaFree(tmp); };
    auto cuda = [&](cudaError_t e, const char* what) { if (e == cudaSuccess) return false; release(true); cudaFail(e, what); return true; };
    if (cuda(cudaMalloc(&dst, dstBytes), "cudaMalloc(expanded grid)")) return VDBRT_ERR_CUDA;
    if (cuda(cudaMemcpyAsync(dst, src, leafOff, cudaMemcpyDeviceToDevice, ctx->stream), "cudaMemcpyAsync(grid head)")) return VDBRT_ERR_CUDA;

    // GridData of the expanded buffer (NanoVDB.h:1944-1966): a stand-alone NanoGrid<float> without blind data or checksum
    uint8_t gd[kGridSize];
    std::memcpy(gd, head, kGridSize);
    wr<uint64_t>(gd + 8, ~uint64_t(0));             // Checksum: none
    wr<uint32_t>(gd + 24, 0u); wr<uint32_t>(gd + 28, 1u);   // grid 0 of 1
    wr<uint64_t>(gd + 32, dstBytes);
    wr<uint32_t>(gd + 636, 1u);                     // GridType::Float
    wr<int64_t>(gd + 640, int64_t(dstBytes)); wr<uint32_t>(gd + 648, 0u);
    if (cuda(cudaMemcpyAsync(dst, gd, kGridSize, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(GridData)")) return VDBRT_ERR_CUDA;
    if (cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize")) return VDBRT_ERR_CUDA;   // gd is a stack buffer

    if (leafCount) {
        if (cuda(cudaMalloc(&keys, sizeof(uint32_t) * 2 * size_t(leafCount)), "cudaMalloc(leaf keys)")) return VDBRT_ERR_CUDA;
        if (cuda(cudaMalloc(&ctl, sizeof(QuantCtl)), "cudaMalloc(ctl)")) return VDBRT_ERR_CUDA;
        if (cuda(cudaMemsetAsync(ctl, 0, sizeof(QuantCtl), ctx->stream), "cudaMemsetAsync")) return VDBRT_ERR_CUDA;
        const uint64_t slots = uint64_t(lowerCount) << 12;
        const unsigned blocks = unsigned((slots + 255) / 256);
        if (blocks) k_quant_collect<<<blocks, 256, 0, ctx->stream>>>(src, lowerOff, lowerCount, leafOff, leafEnd, keys, leafCount, ctl);
        QuantCtl h{};
        if (cuda(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(ctl)")) return VDBRT_ERR_CUDA;
        if (cuda(cudaStreamSynchronize(ctx->stream), "k_quant_collect")) return VDBRT_ERR_CUDA;
        if (h.status || h.count != leafCount) {
            release(true);
            return setError(VDBRT_ERR_BAD_GRID, h.status ? "a lower node's child offset points outside the leaf nodes"
                                                         : "the child masks of the lower nodes do not add up to the leaf count");
        }
        size_t tmpBytes = 0;
        if (cuda(cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, keys, keys + leafCount, int(leafCount), 0, 32, ctx->stream), "cub::SortKeys(size)")) return VDBRT_ERR_CUDA;
        if (cuda(cudaMalloc(&tmp, tmpBytes ? tmpBytes : 1), "cudaMalloc(sort)")) return VDBRT_ERR_CUDA;
        if (cuda(cub::DeviceRadixSort::SortKeys(tmp, tmpBytes, keys, keys + leafCount, int(leafCount), 0, 32, ctx->stream), "cub::SortKeys")) return VDBRT_ERR_CUDA;
        const uint32_t* sorted = keys + leafCount;
        k_quant_relink<<<blocks, 256, 0, ctx->stream>>>(src, dst, lowerOff, lowerCount, leafOff, sorted, leafCount);
        k_quant_expand<<<leafCount, 128, 0, ctx->stream>>>(src, dst, leafOff, leafEnd, sorted, leafCount, gridType, ctl);
        if (cuda(cudaGetLastError(), "k_quant_expand launch")) return VDBRT_ERR_CUDA;
        if (cuda(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(ctl)")) return VDBRT_ERR_CUDA;
        if (cuda(cudaStreamSynchronize(ctx->stream), "k_quant_expand")) return VDBRT_ERR_CUDA;
        if (h.status) {
            release(true);
            return setError(h.status & 4u ? VDBRT_ERR_UNSUPPORTED : VDBRT_ERR_BAD_GRID,
                            h.status & 4u ? "FpN leaf with a bit width above 16"
                                          : (h.status & 8u ? "a quantised leaf reaches past the end of the grid" : "the quantised leaves do not tile the leaf area"));
        }
    }
    release(false);
    *outDev = dst; *outBytes = dstBytes;
    return VDBRT_OK;
}

} // namespace vdbrt
