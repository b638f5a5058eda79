// ref_driver.cc -- TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or shipped with the product.
//
// A small C API over the UNMODIFIED reference (OpenVDB 13.0.1 + NanoVDB headers under /root/reference, compiled
// by oracle/Makefile into oracle/_ref/libvdbref.so).  It lets tests/ and bench.py's reference arm
//   * build the synthetic grids of BASELINE.json with the reference's own generators and serialise them with
//     nanovdb::tools::createNanoGrid (so the SAME grid feeds reference, oracle port and GPU),
//   * run tools::rayTrace / LevelSetRayIntersector / VolumeRender / VolumeRayIntersector::hits on them,
//   * read the per-pixel records parity is defined on (hit mask, first-hit voxel, t, xyz, normal, RGBA).
//
// No reference arithmetic is restated here.  The first-hit voxel is not an output of the reference API, so the
// stock LinearSearchImpl is wrapped by a forwarding tester (PeekSearch) that records the Coord handed to the
// operator() call that returned true; `#define private public` gives the wrapper access to the tester methods
// that the reference only exposes to its friend LevelSetHDDA.  The wrapper forwards every call unchanged.
#include <algorithm>
#include <any>
#include <deque>
#include <fstream>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>
#include <tbb_shim.h>
#include <openvdb/openvdb.h>
#include <openvdb/tools/Interpolation.h>
#include <openvdb/tools/Morphology.h>
#define private public
#define protected public
#include <openvdb/math/Stencils.h>
#include <openvdb/tools/RayIntersector.h>
#include <openvdb/tools/RayTracer.h>
#undef private
#undef protected
#include <openvdb/tools/LevelSetSphere.h>
#include <openvdb/tools/LevelSetUtil.h>
#include <openvdb/tools/Composite.h>
#include <nanovdb/NanoVDB.h>
#include <nanovdb/tools/CreateNanoGrid.h>
#include <nanovdb/tools/CreatePrimitives.h>
#include <nanovdb/tools/NanoToOpenVDB.h>

#include "../include/vdbrt.h"

#include <chrono>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

using namespace openvdb;

namespace {

thread_local std::string g_err;

struct RefGrid {
    FloatGrid::Ptr grid;
    nanovdb::GridHandle<nanovdb::HostBuffer> nano; // lazily created serialisation
    nanovdb::GridHandle<nanovdb::HostBuffer> quant; // last quantised serialisation (vdbref_grid_nanovdb_quantized)
};

// Forwarding tester: identical behaviour to the stock LinearSearchImpl, plus bookkeeping.
template<typename GridT>
class PeekSearch
{
public:
    using ImplT = tools::LinearSearchImpl<GridT, 0, double>;
    using RayT = typename ImplT::RayT;
    using VecT = typename ImplT::VecT;
    using ValueT = typename ImplT::ValueT;
    PeekSearch(const GridT& grid, const ValueT& iso = zeroVal<ValueT>()) : mImpl(grid, iso) {}
    const ValueT& getIsoValue() const { return mImpl.getIsoValue(); }
    bool setIndexRay(const RayT& r) { return mImpl.setIndexRay(r); }
    bool setWorldRay(const RayT& r) { return mImpl.setWorldRay(r); }
    void getIndexPos(VecT& xyz) const { mImpl.getIndexPos(xyz); }
    void getWorldPos(VecT& xyz) const { mImpl.getWorldPos(xyz); }
    void getWorldPosAndNml(VecT& xyz, VecT& nml) { trackStencil([&] { mImpl.getWorldPosAndNml(xyz, nml); }); }
    double getIndexTime() const { return mImpl.getIndexTime(); }
    double getWorldTime() const { return mImpl.getWorldTime(); }
    // what LevelSetHDDA calls
    void init(double t0) { trackStencil([&] { mImpl.init(t0); }); }
    void setRange(double t0, double t1) { mImpl.setRange(t0, t1); }
    const RayT& ray() const { return mImpl.ray(); }
    template<typename NodeT> bool hasNode(const Coord& ijk)
    {
        ++probes[NodeT::LEVEL];
        return mImpl.template hasNode<NodeT>(ijk);
    }
    bool operator()(const Coord& ijk, double time)
    {
        ++voxelProbes;
        bool hit = false;
        trackStencil([&] { hit = mImpl(ijk, time); });
        if (hit) hitIjk = ijk;
        return hit;
    }
    template<typename F> void trackStencil(F f)
    {
        const Coord before = mImpl.mStencil.mCenter;
        f();
        if (mImpl.mStencil.mCenter != before) ++stencilRefills;
    }
    ImplT mImpl;
    Coord hitIjk;
    uint64_t probes[4] = {0, 0, 0, 0}; // by NodeT::LEVEL: [0]=leaf probes (inside lower), [1]=lower, [2]=upper
    uint64_t voxelProbes = 0, stencilRefills = 0;
};

using PeekIntersector = tools::LevelSetRayIntersector<FloatGrid, PeekSearch<FloatGrid>>;
using StockIntersector = tools::LevelSetRayIntersector<FloatGrid>;
using VolIntersector = tools::VolumeRayIntersector<FloatGrid>;

} // namespace

extern "C" {

// constructor arguments of the reference cameras (tools/RayTracer.h:436-445,494-502) + optional lookAt
struct vdbref_camera_desc {
    uint32_t kind, width, height, use_lookat;
    double rotation[3], translation[3];
    double focal_or_frame; // focal length (perspective) or frame width (orthographic)
    double aperture;
    double near_plane, far_plane;
    double target[3], up[3];
};

const char* vdbref_last_error() { return g_err.c_str(); }

void vdbref_set_threads(int n) { tbb_shim::set_num_threads(n); }
int vdbref_hardware_threads() { return int(std::thread::hardware_concurrency()); }

static void ensureInit()
{
    static bool done = false;
    if (!done) { openvdb::initialize(); done = true; }
}

void* vdbref_grid_sphere(double radius, double cx, double cy, double cz, double voxel, double halfWidth)
{
    try {
        ensureInit();
        auto* g = new RefGrid;
        g->grid = tools::createLevelSetSphere<FloatGrid>(float(radius), Vec3f(float(cx), float(cy), float(cz)),
                                                         float(voxel), float(halfWidth));
        return g;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

// NanoVDB-only primitive: build with nanovdb, convert with nanoToOpenVDB for the reference tracer
void* vdbref_grid_torus(double R, double r, double cx, double cy, double cz, double voxel, double halfWidth)
{
    try {
        ensureInit();
        auto* g = new RefGrid;
        g->nano = nanovdb::tools::createLevelSetTorus<float>(R, r, nanovdb::Vec3d(cx, cy, cz), voxel, halfWidth);
        g->grid = nanovdb::tools::nanoToOpenVDB(*g->nano.grid<float>());
        return g;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

void* vdbref_grid_nano_sphere(double radius, double cx, double cy, double cz, double voxel, double halfWidth)
{
    try {
        ensureInit();
        auto* g = new RefGrid;
        g->nano = nanovdb::tools::createLevelSetSphere<float>(radius, nanovdb::Vec3d(cx, cy, cz), voxel, halfWidth);
        g->grid = nanovdb::tools::nanoToOpenVDB(*g->nano.grid<float>());
        return g;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

// deep copy of a level set turned into a fog volume with openvdb::tools::sdfToFogVolume (default cutoff)
void* vdbref_grid_fog_from_levelset(void* ls)
{
    try {
        auto* g = new RefGrid;
        g->grid = static_cast<RefGrid*>(ls)->grid->deepCopy();
        tools::sdfToFogVolume(*g->grid);
        return g;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

// csgUnion of n spheres {cx,cy,cz,r} (world units)
void* vdbref_grid_spheres_union(const double* spheres, uint32_t n, double voxel, double halfWidth)
{
    try {
        ensureInit();
        auto* g = new RefGrid;
        for (uint32_t s = 0; s < n; ++s) {
            const double* p = spheres + 4 * s;
            auto sph = tools::createLevelSetSphere<FloatGrid>(float(p[3]), Vec3f(float(p[0]), float(p[1]), float(p[2])),
                                                              float(voxel), float(halfWidth));
            if (!g->grid) g->grid = sph; else tools::csgUnion(*g->grid, *sph);
        }
        return g;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

// The same union folded by `threads` workers: worker w unions the spheres w, w+threads, ... into its own grid, then the
// partial grids are unioned in worker order.  csgUnion keeps min(a,b) with the minimum's active state (ties keep a's; equal
// values of these spheres have equal states), so the fold order does not change the result -- asserted against the
// sequential fold in tests/test_oracle_vs_reference.py.  Only used to build BASELINE config 4's grid for the timed
// reference arm in seconds instead of minutes; the render that is timed is the stock one.
void* vdbref_grid_spheres_union_mt(const double* spheres, uint32_t n, double voxel, double halfWidth, int threads)
{
    try {
        ensureInit();
        if (threads < 1) threads = 1;
        if (uint32_t(threads) > n) threads = int(n ? n : 1);
        std::vector<FloatGrid::Ptr> part(threads);
        std::vector<std::string> errs(threads);
        const int saved = tbb_shim::num_threads();
        tbb_shim::set_num_threads(1);                       // the library's own parallel_for stays serial inside the workers
        auto work = [&](int w) {
            try {
                for (uint32_t s = uint32_t(w); s < n; s += uint32_t(threads)) {
                    const double* p = spheres + 4 * s;
                    auto sph = tools::createLevelSetSphere<FloatGrid>(float(p[3]), Vec3f(float(p[0]), float(p[1]), float(p[2])),
                                                                      float(voxel), float(halfWidth));
                    if (!part[w]) part[w] = sph; else tools::csgUnion(*part[w], *sph);
                }
            } catch (std::exception& e) { errs[w] = e.what(); }
        };
        std::vector<std::thread> pool;
        for (int w = 1; w < threads; ++w) pool.emplace_back(work, w);
        work(0);
        for (auto& t : pool) t.join();
        tbb_shim::set_num_threads(saved);
        for (auto& e : errs) if (!e.empty()) { g_err = e; return nullptr; }
        auto* g = new RefGrid;
        for (int w = 0; w < threads; ++w) {
            if (!part[w]) continue;
            if (!g->grid) g->grid = part[w]; else tools::csgUnion(*g->grid, *part[w]);
            part[w].reset();
        }
        return g;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

// hand-made grid for the known-answer tests: setValue(ijk, v) for n voxels, then Grid::fill(bbox, value, active) for
// nb boxes {min xyz, max xyz} (a node-aligned active box becomes an active tile)
void* vdbref_grid_custom(float background, uint32_t gridClass, double voxelSize, const double* translation,
                         const int32_t* ijk, const float* values, uint64_t n,
                         const int32_t* boxes, const float* boxValues, const uint8_t* boxActive, uint64_t nb)
{
    try {
        ensureInit();
        auto* g = new RefGrid;
        g->grid = FloatGrid::create(background);
        auto xf = math::Transform::createLinearTransform(voxelSize);
        if (translation && (translation[0] != 0 || translation[1] != 0 || translation[2] != 0))
            xf->postTranslate(Vec3d(translation[0], translation[1], translation[2]));
        g->grid->setTransform(xf);
        g->grid->setGridClass(gridClass == VDBRT_GRID_CLASS_LEVEL_SET ? GRID_LEVEL_SET : (gridClass == VDBRT_GRID_CLASS_FOG_VOLUME ? GRID_FOG_VOLUME : GRID_UNKNOWN));
        for (uint64_t i = 0; i < n; ++i) g->grid->tree().setValue(Coord(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]), values[i]);
        for (uint64_t b = 0; b < nb; ++b)
            g->grid->fill(CoordBBox(Coord(boxes[6 * b], boxes[6 * b + 1], boxes[6 * b + 2]), Coord(boxes[6 * b + 3], boxes[6 * b + 4], boxes[6 * b + 5])),
                          boxValues[b], boxActive[b] != 0);
        return g;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

// any NanoGrid<float> buffer (e.g. one produced by the product's GPU builder) -> OpenVDB grid
void* vdbref_grid_from_nanovdb(const void* buf, uint64_t bytes)
{
    try {
        ensureInit();
        auto* g = new RefGrid;
        auto hb = nanovdb::HostBuffer::create(bytes);
        std::memcpy(hb.data(), buf, bytes);
        g->nano = nanovdb::GridHandle<nanovdb::HostBuffer>(std::move(hb));
        // float and quantised (Fp4/Fp8/Fp16/FpN) grids: nanoToOpenVDB reads every voxel through the NanoVDB accessors
        // (LeafData<FpX>::getValue) into a FloatGrid -- the only way the reference ray tracer can be given such a grid
        const nanovdb::GridType t = g->nano.gridType(0);
        if (t != nanovdb::GridType::Float && t != nanovdb::GridType::Fp4 && t != nanovdb::GridType::Fp8 &&
            t != nanovdb::GridType::Fp16 && t != nanovdb::GridType::FpN) { g_err = "not a float-valued NanoGrid"; delete g; return nullptr; }
        g->grid = gridPtrCast<FloatGrid>(nanovdb::tools::nanoToOpenVDB(g->nano, 0));
        if (!g->grid) { g_err = "nanoToOpenVDB did not return a FloatGrid"; delete g; return nullptr; }
        return g;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

void vdbref_grid_free(void* h) { delete static_cast<RefGrid*>(h); }

// nanovdb::tools::createNanoGrid(openvdbGrid) serialisation; returns size, *out points at grid-owned memory
uint64_t vdbref_grid_nanovdb(void* h, const void** out)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        if (!g->nano) g->nano = nanovdb::tools::createNanoGrid(*g->grid);
        if (out) *out = g->nano.data();
        return g->nano.bufferSize();
    } catch (std::exception& e) { g_err = e.what(); return 0; }
}
// nanovdb::tools::createNanoGrid<FloatGrid, Fp4|Fp8|Fp16|FpN>: the quantised serialisations (gridType 13..16 = nanovdb::GridType);
// tolerance < 0 keeps FpN's default oracle (AbsDiff picks its tolerance from the grid class).  The handle is kept by the
// grid (one quantised serialisation at a time) and *out points into it.
uint64_t vdbref_grid_nanovdb_quantized(void* h, uint32_t gridType, int dither, float tolerance, const void** out)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        namespace nt = nanovdb::tools;
        using SrcT = openvdb::FloatGrid;
        const auto sm = nt::StatsMode::Default; const auto cm = nanovdb::CheckMode::Default;
        const bool d = dither != 0;
        switch (gridType) {
        case 13: g->quant = nt::createNanoGrid<SrcT, nanovdb::Fp4>(*g->grid, sm, cm, d); break;
        case 14: g->quant = nt::createNanoGrid<SrcT, nanovdb::Fp8>(*g->grid, sm, cm, d); break;
        case 15: g->quant = nt::createNanoGrid<SrcT, nanovdb::Fp16>(*g->grid, sm, cm, d); break;
        case 16: g->quant = tolerance < 0 ? nt::createNanoGrid<SrcT, nanovdb::FpN>(*g->grid, sm, cm, d)
                                          : nt::createNanoGrid<SrcT, nanovdb::FpN>(*g->grid, sm, cm, d, 0, nt::AbsDiff(tolerance)); break;
        default: g_err = "gridType must be 13 (Fp4), 14 (Fp8), 15 (Fp16) or 16 (FpN)"; return 0;
        }
        if (out) *out = g->quant.data();
        return g->quant.bufferSize();
    } catch (std::exception& e) { g_err = e.what(); return 0; }
}
// drop the cached serialisation so the next vdbref_grid_nanovdb re-converts from the OpenVDB grid
void vdbref_grid_reserialize(void* h) { static_cast<RefGrid*>(h)->nano = nanovdb::GridHandle<nanovdb::HostBuffer>(); }

// stats[0]=active voxels, [1]=leaf count, [2]=active tiles ; bbox = evalActiveBoundingBox(false) (6 ints), background
void vdbref_grid_stats(void* h, uint64_t* stats, int32_t* nodeBBox, int32_t* voxelBBox, float* background)
{
    auto* g = static_cast<RefGrid*>(h);
    stats[0] = g->grid->activeVoxelCount();
    stats[1] = g->grid->tree().leafCount();
    stats[2] = g->grid->tree().activeTileCount();
    CoordBBox b;
    g->grid->tree().root().evalActiveBoundingBox(b, false);
    for (int i = 0; i < 3; ++i) { nodeBBox[i] = b.min()[i]; nodeBBox[3 + i] = b.max()[i]; }
    CoordBBox v = g->grid->evalActiveVoxelBoundingBox();
    for (int i = 0; i < 3; ++i) { voxelBBox[i] = v.min()[i]; voxelBBox[3 + i] = v.max()[i]; }
    *background = g->grid->background();
}

// random access, for checking a foreign grid voxel by voxel: values[i], active[i] = probeValue(ijk[i])
void vdbref_grid_probe(void* h, const int32_t* ijk, uint64_t n, float* values, uint8_t* active)
{
    auto acc = static_cast<RefGrid*>(h)->grid->getConstAccessor();
    for (uint64_t i = 0; i < n; ++i) {
        float v;
        active[i] = acc.probeValue(Coord(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]), v);
        values[i] = v;
    }
}

} // extern "C"

namespace {

// expose the protected camera state without touching the reference
std::unique_ptr<tools::BaseCamera> makeCamera(tools::Film& film, const vdbref_camera_desc& d)
{
    const Vec3R rot(d.rotation[0], d.rotation[1], d.rotation[2]), tr(d.translation[0], d.translation[1], d.translation[2]);
    std::unique_ptr<tools::BaseCamera> cam;
    if (d.kind == VDBRT_CAMERA_PERSPECTIVE)
        cam.reset(new tools::PerspectiveCamera(film, rot, tr, d.focal_or_frame, d.aperture, d.near_plane, d.far_plane));
    else
        cam.reset(new tools::OrthographicCamera(film, rot, tr, d.focal_or_frame, d.near_plane, d.far_plane));
    if (d.use_lookat) cam->lookAt(Vec3R(d.target[0], d.target[1], d.target[2]), Vec3R(d.up[0], d.up[1], d.up[2]));
    return cam;
}

std::unique_ptr<tools::BaseShader> makeShader(const vdbrt_shader& s)
{
    const tools::Film::RGBA c(s.rgba[0], s.rgba[1], s.rgba[2], s.rgba[3]);
    switch (s.kind) {
    case VDBRT_SHADER_MATTE: return std::unique_ptr<tools::BaseShader>(new tools::MatteShader<>(c));
    case VDBRT_SHADER_NORMAL: return std::unique_ptr<tools::BaseShader>(new tools::NormalShader<>(c));
    case VDBRT_SHADER_POSITION: {
        // PositionShader stores min and 1/extents; rebuild a bbox with exactly those (min, min + 1/inv) is lossy,
        // so construct from a bbox and then overwrite the two const members with the caller's values.
        math::BBox<Vec3R> bb(Vec3R(0.0), Vec3R(1.0));
        auto* p = new tools::PositionShader<>(bb, c);
        const_cast<Vec3R&>(p->mMin) = Vec3R(s.bbox_min[0], s.bbox_min[1], s.bbox_min[2]);
        const_cast<Vec3R&>(p->mInvDim) = Vec3R(s.inv_dim[0], s.inv_dim[1], s.inv_dim[2]);
        return std::unique_ptr<tools::BaseShader>(p);
    }
    default: return std::unique_ptr<tools::BaseShader>(new tools::DiffuseShader<>(c));
    }
}

struct RefColor {
    Vec3SGrid::Ptr grid;
    nanovdb::GridHandle<nanovdb::HostBuffer> nano;
};

// the GridT = Vec3SGrid forms of the shaders (tools/RayTracer.h:542-725), default PointSampler
std::unique_ptr<tools::BaseShader> makeColorShader(const vdbrt_shader& s, const Vec3SGrid& cg)
{
    switch (s.kind) {
    case VDBRT_SHADER_MATTE: return std::unique_ptr<tools::BaseShader>(new tools::MatteShader<Vec3SGrid>(cg));
    case VDBRT_SHADER_NORMAL: return std::unique_ptr<tools::BaseShader>(new tools::NormalShader<Vec3SGrid>(cg));
    case VDBRT_SHADER_POSITION: {
        math::BBox<Vec3R> bb(Vec3R(0.0), Vec3R(1.0));
        auto* p = new tools::PositionShader<Vec3SGrid>(bb, cg);
        const_cast<Vec3R&>(p->mMin) = Vec3R(s.bbox_min[0], s.bbox_min[1], s.bbox_min[2]);
        const_cast<Vec3R&>(p->mInvDim) = Vec3R(s.inv_dim[0], s.inv_dim[1], s.inv_dim[2]);
        return std::unique_ptr<tools::BaseShader>(p);
    }
    default: return std::unique_ptr<tools::BaseShader>(new tools::DiffuseShader<Vec3SGrid>(cg));
    }
}

} // namespace

extern "C" {

// A Vec3SGrid to colour the level set `h` with: its own transform (voxel size `voxel`, translation `t`), one voxel for every
// active voxel of the level set (nearest colour voxel of its world position) with a colour derived from the coordinate, one
// active 8^3 tile, background (0.2, 0.4, 0.6).
void* vdbref_color_grid(void* h, double voxel, const double* t)
{
    try {
        ensureInit();
        auto* g = static_cast<RefGrid*>(h);
        auto* c = new RefColor;
        c->grid = Vec3SGrid::create(Vec3s(0.2f, 0.4f, 0.6f));
        math::Transform::Ptr xf = math::Transform::createLinearTransform(voxel);
        xf->postTranslate(Vec3d(t[0], t[1], t[2]));
        c->grid->setTransform(xf);
        auto acc = c->grid->getAccessor();
        Coord first; bool haveFirst = false;
        for (auto it = g->grid->cbeginValueOn(); it; ++it) {
            if (!it.isVoxelValue()) continue;
            const Vec3d w = g->grid->indexToWorld(it.getCoord());
            const Vec3d p = xf->worldToIndex(w);
            const Coord ijk(int(::round(p[0])), int(::round(p[1])), int(::round(p[2])));
            const uint32_t hsh = uint32_t(ijk[0]) * 73856093u ^ uint32_t(ijk[1]) * 19349663u ^ uint32_t(ijk[2]) * 83492791u;
            acc.setValue(ijk, Vec3s(0.1f + 0.8f * float(hsh & 255u) / 255.0f, 0.1f + 0.8f * float((hsh >> 8) & 255u) / 255.0f,
                                    0.1f + 0.8f * float((hsh >> 16) & 255u) / 255.0f));
            if (!haveFirst) { first = ijk; haveFirst = true; }
        }
        if (haveFirst) c->grid->tree().addTile(/*level=*/1, first, Vec3s(0.9f, 0.1f, 0.3f), /*active=*/true);   // replaces one leaf by a tile
        return c;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}
void vdbref_color_free(void* h) { delete static_cast<RefColor*>(h); }
uint64_t vdbref_color_nanovdb(void* h, const void** out)
{
    try {
        auto* c = static_cast<RefColor*>(h);
        if (!c->nano) c->nano = nanovdb::tools::createNanoGrid(*c->grid);
        if (out) *out = c->nano.data();
        return c->nano.bufferSize();
    } catch (std::exception& e) { g_err = e.what(); return 0; }
}

// tools::rayTrace with one of the colour-grid shaders
double vdbref_render_levelset_color(void* h, void* color, const vdbref_camera_desc* d, const vdbrt_shader* sh, float iso, uint32_t spp,
                                    unsigned int seed, int threaded, float* filmRGBA)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        auto* c = static_cast<RefColor*>(color);
        tools::Film film(d->width, d->height);
        const size_t npx = size_t(d->width) * d->height;
        std::memcpy(const_cast<tools::Film::RGBA*>(film.pixels()), filmRGBA, npx * 16);
        auto cam = makeCamera(film, *d);
        auto shader = makeColorShader(*sh, *c->grid);
        const auto t0 = std::chrono::steady_clock::now();
        StockIntersector inter(*g->grid, iso);
        tools::rayTrace(*g->grid, inter, *shader, *cam, spp, seed, threaded != 0);
        const auto t1 = std::chrono::steady_clock::now();
        std::memcpy(filmRGBA, film.pixels(), npx * 16);
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (openvdb::ValueError& e) { g_err = std::string("ValueError: ") + e.what(); return -1.0;
    } catch (openvdb::RuntimeError& e) { g_err = std::string("RuntimeError: ") + e.what(); return -1.0;
    } catch (std::exception& e) { g_err = e.what(); return -1.0; }
}

// flatten the reference camera into the product's POD, for checking vdbrt_camera_* bit for bit
int vdbref_camera_pod(const vdbref_camera_desc* d, vdbrt_camera* out)
{
    try {
        tools::Film film(d->width, d->height);
        auto cam = makeCamera(film, *d);
        std::memset(out, 0, sizeof(*out));
        out->kind = d->kind; out->width = d->width; out->height = d->height;
        const Mat4d m = cam->mScreenToWorld.getMat4();
        std::memcpy(out->m, m.asPointer(), 16 * sizeof(double));
        for (int i = 0; i < 3; ++i) { out->eye[i] = cam->mRay.eye()[i]; out->dir[i] = cam->mRay.dir()[i]; }
        out->scale_w = cam->mScaleWidth; out->scale_h = cam->mScaleHeight;
        out->t0 = cam->mRay.t0(); out->t1 = cam->mRay.t1();
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// BaseCamera::getRay for a list of pixels (i,j,io,jo) -> world rays
int vdbref_camera_rays(const vdbref_camera_desc* d, const uint32_t* ij, const double* offsets, uint64_t n, vdbrt_ray* rays)
{
    try {
        tools::Film film(d->width, d->height);
        auto cam = makeCamera(film, *d);
        for (uint64_t k = 0; k < n; ++k) {
            const double io = offsets ? offsets[2 * k] : 0.5, jo = offsets ? offsets[2 * k + 1] : 0.5;
            const math::Ray<double> r = cam->getRay(ij[2 * k], ij[2 * k + 1], io, jo);
            for (int a = 0; a < 3; ++a) { rays[k].eye[a] = r.eye()[a]; rays[k].dir[a] = r.dir()[a]; }
            rays[k].t0 = r.t0(); rays[k].t1 = r.t1();
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

void vdbref_jitter_table(unsigned int seed, double* out16)
{
    math::Rand01<double> rand(seed);
    for (int i = 0; i < 16; ++i) out16[i] = rand();
}

// tools::rayTrace(grid, LevelSetRayIntersector(grid, iso), shader, camera, spp, seed, threaded) into `film`
// (in/out, RGBA float4).  Returns elapsed seconds of intersector construction + render (the vdb_render -v region,
// openvdb_cmd/vdb_render/main.cc:475-503) or a negative value on error (message via vdbref_last_error, with the
// exception class as prefix).
double vdbref_render_levelset(void* h, const vdbref_camera_desc* d, const vdbrt_shader* sh, float iso, uint32_t spp,
                              unsigned int seed, int threaded, float* filmRGBA)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        tools::Film film(d->width, d->height);
        const size_t npx = size_t(d->width) * d->height;
        std::memcpy(const_cast<tools::Film::RGBA*>(film.pixels()), filmRGBA, npx * 16);
        auto cam = makeCamera(film, *d);
        auto shader = makeShader(*sh);
        const auto t0 = std::chrono::steady_clock::now();
        StockIntersector inter(*g->grid, iso);
        tools::rayTrace(*g->grid, inter, *shader, *cam, spp, seed, threaded != 0);
        const auto t1 = std::chrono::steady_clock::now();
        std::memcpy(filmRGBA, film.pixels(), npx * 16);
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (openvdb::ValueError& e) { g_err = std::string("ValueError: ") + e.what(); return -1.0;
    } catch (openvdb::RuntimeError& e) { g_err = std::string("RuntimeError: ") + e.what(); return -1.0;
    } catch (std::exception& e) { g_err = e.what(); return -1.0; }
}

// primary-ray records per pixel using LevelSetRayIntersector<FloatGrid, PeekSearch> (forwarding tester).
// Also cross-checks the forwarding tester against the stock intersector: returns the number of pixels whose
// (hit, world xyz, normal, world time) differ bitwise between the two -- must be 0.
int64_t vdbref_levelset_records(void* h, const vdbref_camera_desc* d, float iso, vdbrt_aux* aux, vdbrt_counters* ctr)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        tools::Film film(d->width, d->height);
        auto cam = makeCamera(film, *d);
        PeekIntersector peek(*g->grid, iso);
        StockIntersector stock(*g->grid, iso);
        int64_t mismatches = 0;
        if (ctr) std::memset(ctr, 0, sizeof(*ctr));
        for (uint32_t j = 0; j < d->height; ++j) {
            for (uint32_t i = 0; i < d->width; ++i) {
                const size_t p = size_t(j) * d->width + i;
                const math::Ray<double> ray = cam->getRay(i, j);
                Vec3R xyz(0.0), nml(0.0), xyz2(0.0), nml2(0.0);
                double tw = 0.0, tw2 = 0.0;
                const bool hit = peek.intersectsWS(ray, xyz, nml, tw);
                const bool hit2 = stock.intersectsWS(ray, xyz2, nml2, tw2);
                if (hit != hit2 || std::memcmp(&xyz, &xyz2, 24) || std::memcmp(&nml, &nml2, 24) || std::memcmp(&tw, &tw2, 8)) ++mismatches;
                if (aux->hit) aux->hit[p] = hit;
                if (hit) {
                    const Coord c = peek.mTester.hitIjk;
                    if (aux->ijk) { aux->ijk[3 * p] = c[0]; aux->ijk[3 * p + 1] = c[1]; aux->ijk[3 * p + 2] = c[2]; }
                    if (aux->t_index) aux->t_index[p] = peek.mTester.getIndexTime();
                    if (aux->t_world) aux->t_world[p] = tw;
                    for (int a = 0; a < 3; ++a) {
                        if (aux->xyz) aux->xyz[3 * p + a] = xyz[a];
                        if (aux->nml) aux->nml[3 * p + a] = nml[a];
                    }
                }
            }
        }
        if (ctr) {
            auto& t = peek.mTester;
            ctr->rays = uint64_t(d->width) * d->height;
            ctr->root_probes = t.probes[2]; ctr->upper_probes = t.probes[1]; ctr->lower_probes = t.probes[0];
            ctr->voxel_probes = t.voxelProbes; ctr->stencil_refills = t.stencilRefills;
        }
        return mismatches;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// LevelSetRayIntersector::intersectsWS / intersectsIS on arbitrary rays
int vdbref_intersect_levelset(void* h, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso, vdbrt_hit* hits)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        PeekIntersector peek(*g->grid, iso);
        for (uint64_t k = 0; k < n; ++k) {
            const vdbrt_ray& r = rays[k];
            const math::Ray<double> ray(Vec3R(r.eye[0], r.eye[1], r.eye[2]), Vec3R(r.dir[0], r.dir[1], r.dir[2]), r.t0, r.t1);
            vdbrt_hit& o = hits[k];
            std::memset(&o, 0, sizeof(o));
            Vec3R w(0.0), nm(0.0);
            double tw = 0.0;
            bool hit;
            if (space == VDBRT_SPACE_WORLD) {
                hit = peek.intersectsWS(ray, w, nm, tw);
            } else {
                // same sequence with an index-space ray: setIndexRay + HDDA + world outputs
                hit = peek.mTester.setIndexRay(ray) && math::LevelSetHDDA<FloatTree, 2>::test(peek.mTester);
                if (hit) { peek.mTester.getWorldPosAndNml(w, nm); tw = peek.mTester.getWorldTime(); }
            }
            o.hit = hit;
            if (hit) {
                const Coord c = peek.mTester.hitIjk;
                o.ijk[0] = c[0]; o.ijk[1] = c[1]; o.ijk[2] = c[2];
                o.t_index = peek.mTester.getIndexTime(); o.t_world = tw;
                Vec3R xi; peek.mTester.getIndexPos(xi);
                for (int a = 0; a < 3; ++a) { o.xyz_index[a] = xi[a]; o.xyz_world[a] = w[a]; o.nml[a] = nm[a]; }
            }
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// The stock intersector with LinearSearchImpl<FloatGrid, ITER, double> as its search (tools/RayIntersector.h:79-82,630-636): what the
// reference's own accuracy sweep instantiates (unittest/TestLevelSetRayIntersector.cc:236-309 uses <FloatGrid, 2>).  No first-hit voxel
// here (the stock search does not expose it): ijk stays zero.
} // extern "C"
template<int ITER>
static void intersectIter(FloatGrid& grid, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso, vdbrt_hit* hits)
{
    using SearchT = tools::LinearSearchImpl<FloatGrid, ITER, double>;
    tools::LevelSetRayIntersector<FloatGrid, SearchT> inter(grid, iso);
    for (uint64_t k = 0; k < n; ++k) {
        const vdbrt_ray& r = rays[k];
        const math::Ray<double> ray(Vec3R(r.eye[0], r.eye[1], r.eye[2]), Vec3R(r.dir[0], r.dir[1], r.dir[2]), r.t0, r.t1);
        vdbrt_hit& o = hits[k];
        std::memset(&o, 0, sizeof(o));
        Vec3R w(0.0), nm(0.0), xi(0.0);
        double tw = 0.0, ti = 0.0;
        bool hit;
        if (space == VDBRT_SPACE_WORLD) {
            hit = inter.intersectsWS(ray, w, nm, tw);                                   // world position, normal, world time
            if (hit) inter.intersectsIS(ray.worldToIndex(grid), xi, ti);                // the same hit in index space (setWorldRay does this map)
        } else {
            hit = inter.intersectsIS(ray, xi, ti);                                      // index-space rays: index position and time only
        }
        o.hit = hit;
        if (hit) {
            o.t_index = ti; o.t_world = tw;
            for (int a = 0; a < 3; ++a) { o.xyz_index[a] = xi[a]; o.xyz_world[a] = w[a]; o.nml[a] = nm[a]; }
        }
    }
}

extern "C" int vdbref_intersect_levelset_iter(void* h, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso, uint32_t iterations, vdbrt_hit* hits)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        switch (iterations) {
        case 0: intersectIter<0>(*g->grid, rays, n, space, iso, hits); break;
        case 1: intersectIter<1>(*g->grid, rays, n, space, iso, hits); break;
        case 2: intersectIter<2>(*g->grid, rays, n, space, iso, hits); break;
        case 3: intersectIter<3>(*g->grid, rays, n, space, iso, hits); break;
        default: g_err = "iterations must be 0..3 in the reference driver"; return 1;
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// tools::rayTrace(grid, intersector, shader, camera, ...) with that intersector (tools/RayTracer.h:56-64)
template<int ITER>
static double renderIter(FloatGrid& grid, const vdbref_camera_desc* d, const vdbrt_shader* sh, float iso, uint32_t spp, unsigned seed, int threaded, float* filmRGBA)
{
    using SearchT = tools::LinearSearchImpl<FloatGrid, ITER, double>;
    using InterT = tools::LevelSetRayIntersector<FloatGrid, SearchT>;
    tools::Film film(d->width, d->height);
    std::memcpy(reinterpret_cast<void*>(const_cast<tools::Film::RGBA*>(film.pixels())), filmRGBA, size_t(d->width) * d->height * 16);
    auto cam = makeCamera(film, *d);
    auto shader = makeShader(*sh);
    const auto t0 = std::chrono::steady_clock::now();
    InterT inter(grid, iso);
    tools::rayTrace<FloatGrid, InterT>(grid, inter, *shader, *cam, spp, seed, threaded != 0);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::memcpy(filmRGBA, film.pixels(), size_t(d->width) * d->height * 16);
    return s;
}

extern "C" double vdbref_render_levelset_iter(void* h, const vdbref_camera_desc* d, const vdbrt_shader* sh, float iso, uint32_t spp, unsigned seed,
                                   int threaded, uint32_t iterations, float* filmRGBA)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        switch (iterations) {
        case 0: return renderIter<0>(*g->grid, d, sh, iso, spp, seed, threaded, filmRGBA);
        case 1: return renderIter<1>(*g->grid, d, sh, iso, spp, seed, threaded, filmRGBA);
        case 2: return renderIter<2>(*g->grid, d, sh, iso, spp, seed, threaded, filmRGBA);
        case 3: return renderIter<3>(*g->grid, d, sh, iso, spp, seed, threaded, filmRGBA);
        default: g_err = "iterations must be 0..3 in the reference driver"; return -1.0;
        }
    } catch (std::exception& e) { g_err = e.what(); return -1.0; }
}

extern "C" {
// VolumeRayIntersector::setWorldRay/setIndexRay + hits()
int vdbref_volume_spans(void* h, const vdbrt_ray* rays, uint64_t n, uint32_t space, uint32_t maxSpans, double* spans, int32_t* counts)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        VolIntersector inter(*g->grid);
        std::vector<math::Ray<double>::TimeSpan> list;
        for (uint64_t k = 0; k < n; ++k) {
            const vdbrt_ray& r = rays[k];
            const math::Ray<double> ray(Vec3R(r.eye[0], r.eye[1], r.eye[2]), Vec3R(r.dir[0], r.dir[1], r.dir[2]), r.t0, r.t1);
            const bool ok = space == VDBRT_SPACE_WORLD ? inter.setWorldRay(ray) : inter.setIndexRay(ray);
            if (!ok) { counts[k] = -1; continue; }
            inter.hits(list);
            counts[k] = int32_t(list.size());
            for (size_t s = 0; s < list.size() && s < maxSpans; ++s) {
                spans[(k * maxSpans + s) * 2] = list[s].t0;
                spans[(k * maxSpans + s) * 2 + 1] = list[s].t1;
            }
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// VolumeRender<VolumeRayIntersector<FloatGrid>>::render; returns seconds (intersector construction + render)
double vdbref_render_volume(void* h, const vdbref_camera_desc* d, const vdbrt_vol_opts* o, int threaded, float* filmRGBA)
{
    try {
        auto* g = static_cast<RefGrid*>(h);
        tools::Film film(d->width, d->height);
        const size_t npx = size_t(d->width) * d->height;
        std::memcpy(const_cast<tools::Film::RGBA*>(film.pixels()), filmRGBA, npx * 16);
        auto cam = makeCamera(film, *d);
        const auto t0 = std::chrono::steady_clock::now();
        VolIntersector inter(*g->grid);
        tools::VolumeRender<VolIntersector> r(inter, *cam);
        // setLightDir normalises with unit(); the POD already holds a unit vector, so assign the member directly
        r.mLightDir = Vec3R(o->light_dir[0], o->light_dir[1], o->light_dir[2]);
        r.setLightColor(o->light_color[0], o->light_color[1], o->light_color[2]);
        r.setPrimaryStep(o->primary_step);
        r.setShadowStep(o->shadow_step);
        r.setScattering(o->scattering[0], o->scattering[1], o->scattering[2]);
        r.setAbsorption(o->absorption[0], o->absorption[1], o->absorption[2]);
        r.setLightGain(o->light_gain);
        r.setCutOff(o->cutoff);
        r.render(threaded != 0);
        const auto t1 = std::chrono::steady_clock::now();
        std::memcpy(filmRGBA, film.pixels(), npx * 16);
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (openvdb::RuntimeError& e) { g_err = std::string("RuntimeError: ") + e.what(); return -1.0;
    } catch (std::exception& e) { g_err = e.what(); return -1.0; }
}

// VolumeRender defaults as the reference constructs them (RayTracer.h:929-936)
void vdbref_vol_defaults(vdbrt_vol_opts* o)
{
    ensureInit();
    auto grid = FloatGrid::create(0.0f);
    grid->setTransform(math::Transform::createLinearTransform(1.0));
    grid->tree().setValue(Coord(0, 0, 0), 1.0f);
    tools::Film film(2, 2);
    tools::PerspectiveCamera cam(film);
    VolIntersector inter(*grid);
    tools::VolumeRender<VolIntersector> r(inter, cam);
    std::memset(o, 0, sizeof(*o));
    o->primary_step = r.mPrimaryStep; o->shadow_step = r.mShadowStep; o->cutoff = r.mCutOff; o->light_gain = r.mLightGain;
    for (int a = 0; a < 3; ++a) {
        o->light_dir[a] = r.mLightDir[a]; o->light_color[a] = r.mLightColor[a];
        o->absorption[a] = r.mAbsorption[a]; o->scattering[a] = r.mScattering[a];
    }
}

} // extern "C"
