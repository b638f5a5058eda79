// nanovdb_bar.cu -- TEST / BENCH INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
//
// The prior-art bar on the same GPU: NanoVDB's OWN level-set ray tracing kernels, i.e. the reference's
// nanovdb::math::zeroCrossing (nanovdb/nanovdb/math/HDDA.h:190-214, float rays, one variable-stride HDDA, sign change of raw voxel
// values) driven exactly as the reference's examples drive it:
//   mode 0  one thread per pixel, 512-thread blocks: ex_raytrace_level_set (examples/ex_raytrace_level_set/nanovdb.cu:46-71 through
//           ex_util/ComputePrimitives.h:94-103,127-139)
//   mode 1  persistent blocks (256 threads x 4 per SM) that pull 32 pixels per warp from a global counter:
//           renderIsoSurfacePersistentKernel (examples/ex_raytrace_iso_surface/common.h:88-103,166-189)
// Only the camera differs from the examples: vdb_render's perspective camera (the POD of include/vdbrt.h, evaluated in float) so that
// the frame is the bench's frame.  It is a DIFFERENT algorithm from the OpenVDB CPU ray tracer this repository reproduces bit for bit
// (SURVEY.md 0.2): no parity is expected, only the hit counts are compared.  Built from the reference headers where they lie
// (oracle/Makefile: nvcc -use_fast_math like the reference's own CUDA build, nanovdb/nanovdb/CMakeLists.txt:98) into oracle/_ref/.
#include <cuda_runtime.h>
#include <nanovdb/NanoVDB.h>
#include <nanovdb/math/Ray.h>
#include <nanovdb/math/HDDA.h>
#include "../include/vdbrt.h"

namespace {

using GridT = nanovdb::FloatGrid;
using Vec3T = nanovdb::math::Vec3<float>;
using RayT = nanovdb::math::Ray<float>;

struct Cam { float m[9], eye[3], sw, sh; int w, h; };

__device__ inline void renderPixel(const Cam& c, int i, const GridT* grid, float* image)
{
    const int x = i % c.w, y = i / c.w;
    // BaseCamera::rasterToScreen + PerspectiveCamera::getRay (openvdb/tools/RayTracer.h:391-395,452-462) in float
    const float sx = (2.f * (float(x) + 0.5f) / float(c.w) - 1.f) * c.sw, sy = (1.f - 2.f * (float(y) + 0.5f) / float(c.h)) * c.sh;
    Vec3T dir(sx * c.m[0] + sy * c.m[3] - c.m[6], sx * c.m[1] + sy * c.m[4] - c.m[7], sx * c.m[2] + sy * c.m[5] - c.m[8]);
    dir.normalize();
    RayT wRay(Vec3T(c.eye[0], c.eye[1], c.eye[2]), dir);
    RayT iRay = wRay.worldToIndexF(*grid);
    auto acc = grid->tree().getAccessor();
    float t0, v;
    nanovdb::Coord ijk;
    image[i] = nanovdb::math::zeroCrossing(iRay, acc, ijk, v, t0) ? t0 * float(grid->voxelSize()[0]) : 0.f;
}

__global__ void k_thread_per_pixel(Cam c, const GridT* grid, float* image, int numPixels)
{
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < numPixels) renderPixel(c, i, grid, image);
}

__global__ void k_persistent(Cam c, const GridT* grid, float* image, int numPixels, int* nextPixel)
{
    const unsigned lane = threadIdx.x & 31u;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(nextPixel, 32);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        const int i = base + int(lane);
        if (i >= numPixels) break;
        renderPixel(c, i, grid, image);
    }
}

__global__ void k_count(const float* image, int n, unsigned long long* out)
{
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    const bool hit = i < n && image[i] > 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

} // namespace

extern "C" int nvbar_render(const void* deviceGrid, const vdbrt_camera* cam, int mode, int warmup, int iters, float* msPerFrame, unsigned long long* hits)
{
    if (!deviceGrid || !cam || cam->kind != VDBRT_CAMERA_PERSPECTIVE || iters < 1) return 1;
    Cam c;
    // the POD's 4x4 is row-major with the 3x3 in rows 0..2 (Mat4::transform3x3: v0*m0 + v1*m4 + v2*m8)
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c.m[3 * r + k] = float(cam->m[4 * r + k]);
    for (int k = 0; k < 3; ++k) c.eye[k] = float(cam->eye[k]);
    c.sw = float(cam->scale_w); c.sh = float(cam->scale_h); c.w = int(cam->width); c.h = int(cam->height);
    const int n = c.w * c.h;
    float* image = nullptr; int* next = nullptr; unsigned long long* dHits = nullptr;
    if (cudaMalloc(&image, size_t(n) * 4) != cudaSuccess || cudaMalloc(&next, 4) != cudaSuccess || cudaMalloc(&dHits, 8) != cudaSuccess) return 2;
    cudaDeviceProp prop; int dev = 0;
    cudaGetDevice(&dev); cudaGetDeviceProperties(&prop, dev);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const GridT* grid = static_cast<const GridT*>(deviceGrid);
    auto frame = [&]() {
        if (mode == 0) k_thread_per_pixel<<<(n + 511) / 512, 512>>>(c, grid, image, n);
        else { cudaMemsetAsync(next, 0, 4); k_persistent<<<prop.multiProcessorCount * 4, 256>>>(c, grid, image, n, next); }
    };
    for (int i = 0; i < warmup; ++i) frame();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) frame();
    cudaEventRecord(e1);
    int rc = cudaEventSynchronize(e1) == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : 3;
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    if (msPerFrame) *msPerFrame = ms / float(iters);
    cudaMemset(dHits, 0, 8);
    k_count<<<(n + 255) / 256, 256>>>(image, n, dHits);
    if (hits) cudaMemcpy(hits, dHits, 8, cudaMemcpyDeviceToHost);
    cudaFree(image); cudaFree(next); cudaFree(dHits); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}
