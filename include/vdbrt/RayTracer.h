// vdbrt/RayTracer.h -- C++ facade over the C ABI (include/vdbrt.h) with the reference's class names and call shapes.
//
// Mirrors openvdb/openvdb/tools/RayTracer.h and tools/RayIntersector.h for the one path this library accelerates:
//   tools::Film, BaseCamera / PerspectiveCamera / OrthographicCamera, BaseShader + the four constant-colour shaders,
//   LevelSetRayIntersector, VolumeRayIntersector, LevelSetRayTracer, VolumeRender, rayTrace().
// Host code written against the reference changes its grid type (a NanoVDB-serialised grid uploaded to the GPU) and its
// namespace; the per-pixel loops (RayTracer.h:899-918, 991-1070) run in the CUDA kernels of libvdbrt.so.
//
// What cannot be a drop-in: user subclasses of BaseCamera / BaseShader (virtual getRay / operator()) cannot run on the
// device, so only the two cameras and the four const-colour shaders are accepted; anything else throws RuntimeError
// (there is no CPU fallback, by design).  Errors the reference throws at construction are thrown here at the same
// points (intersector / tracer construction), with the same exception kinds.
#pragma once
#include "../vdbrt.h"

#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>
#include <sstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace vdbrt {

struct RuntimeError : std::runtime_error { using std::runtime_error::runtime_error; };   // openvdb::RuntimeError
struct ValueError : std::runtime_error { using std::runtime_error::runtime_error; };     // openvdb::ValueError
struct IoError : std::runtime_error { using std::runtime_error::runtime_error; };        // openvdb::IoError

inline void check(int code)
{
    if (code == VDBRT_OK) return;
    const std::string msg = vdbrt_last_error();
    if (code == VDBRT_ERR_ISO_RANGE || code == VDBRT_ERR_SPP_ZERO) throw ValueError(msg);
    throw RuntimeError(msg);
}

struct Vec3R { double x = 0, y = 0, z = 0; Vec3R() = default; Vec3R(double a, double b, double c) : x(a), y(b), z(c) {} explicit Vec3R(double v) : x(v), y(v), z(v) {} };

/// one GPU
class Context
{
public:
    explicit Context(int device = 0) { check(vdbrt_create(device, &mCtx)); }
    ~Context() { vdbrt_destroy(mCtx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    vdbrt_ctx* get() const { return mCtx; }
private:
    vdbrt_ctx* mCtx = nullptr;
};

/// A NanoGrid<float> resident on the GPU (replaces the openvdb::FloatGrid argument of the reference's classes).
class FloatGrid
{
public:
    using Ptr = std::shared_ptr<FloatGrid>;
    /// upload a serialised grid (nanovdb::GridHandle<HostBuffer>::data(), size()): NanoGrid<float>, or a quantised
    /// NanoGrid<Fp4|Fp8|Fp16|FpN> (createNanoGrid<openvdb::FloatGrid, nanovdb::Fp8>(grid) ...), whose leaves are expanded on the device
    static Ptr upload(Context& ctx, const void* nanovdbBuffer, uint64_t bytes)
    {
        vdbrt_grid* g = nullptr;
        check(vdbrt_upload_grid(ctx.get(), nanovdbBuffer, bytes, VDBRT_MEM_HOST, &g));
        return Ptr(new FloatGrid(ctx, g));
    }
    static Ptr createLevelSetSphere(Context& ctx, double radius, const Vec3R& center, double voxelSize, double halfWidth = 3.0)
    {
        vdbrt_grid* g = nullptr;
        const double c[3] = {center.x, center.y, center.z};
        check(vdbrt_build_levelset_sphere(ctx.get(), radius, c, voxelSize, halfWidth, &g));
        return Ptr(new FloatGrid(ctx, g));
    }
    static Ptr createLevelSetTorus(Context& ctx, double majorRadius, double minorRadius, const Vec3R& center, double voxelSize, double halfWidth = 3.0)
    {
        vdbrt_grid* g = nullptr;
        const double c[3] = {center.x, center.y, center.z};
        check(vdbrt_build_levelset_torus(ctx.get(), majorRadius, minorRadius, c, voxelSize, halfWidth, &g));
        return Ptr(new FloatGrid(ctx, g));
    }
    /// nanovdb::io::readGrid + deviceUpload (nanovdb/io/IO.h, GridHandle.h:243-327): the named grid of a .nvdb file, or the
    /// first floating-point grid when no name is given (vdb_render's rule, openvdb_cmd/vdb_render/main.cc:771-786)
    static Ptr read(Context& ctx, const std::string& fileName, const std::string& gridName = "")
    {
        void* buf = nullptr; uint64_t bytes = 0;
        check(vdbrt_nvdb_read(fileName.c_str(), gridName.empty() ? nullptr : gridName.c_str(), &buf, &bytes));
        vdbrt_grid* g = nullptr;
        const int rc = vdbrt_upload_grid(ctx.get(), buf, bytes, VDBRT_MEM_HOST, &g);
        vdbrt_buffer_free(buf);
        check(rc);
        return Ptr(new FloatGrid(ctx, g));
    }
    /// nanovdb::io::writeGrid of the grid as it lives on the device
    void write(const std::string& fileName, uint32_t codec = VDBRT_CODEC_NONE) const
    {
        const vdbrt_grid_info i = info();
        std::vector<unsigned char> host(i.bytes);
        check(vdbrt_grid_download(mCtx->get(), mGrid, host.data(), i.bytes));
        check(vdbrt_nvdb_write(fileName.c_str(), host.data(), i.bytes, codec));
    }
    /// openvdb::tools::sdfToFogVolume
    Ptr sdfToFogVolume() const
    {
        vdbrt_grid* g = nullptr;
        check(vdbrt_build_fog_from_levelset(mCtx->get(), mGrid, &g));
        return Ptr(new FloatGrid(*mCtx, g));
    }
    ~FloatGrid() { vdbrt_free_grid(mCtx->get(), mGrid); }
    vdbrt_grid_info info() const { vdbrt_grid_info i; check(vdbrt_grid_get_info(mGrid, &i)); return i; }
    Context& context() const { return *mCtx; }
    const vdbrt_grid* get() const { return mGrid; }
private:
    FloatGrid(Context& ctx, vdbrt_grid* g) : mCtx(&ctx), mGrid(g) {}
    Context* mCtx; vdbrt_grid* mGrid;
};

/// A NanoGrid<Vec3f> resident on the GPU: the colour grid of the GridT = Vec3SGrid shaders (openvdb::Vec3SGrid serialised with
/// nanovdb::tools::createNanoGrid).  The shader objects keep a reference to it, like the reference's shaders keep an accessor.
class Vec3SGrid
{
public:
    using Ptr = std::shared_ptr<Vec3SGrid>;
    static Ptr upload(Context& ctx, const void* nanovdbBuffer, uint64_t bytes)
    {
        vdbrt_grid* g = nullptr;
        check(vdbrt_upload_color_grid(ctx.get(), nanovdbBuffer, bytes, VDBRT_MEM_HOST, &g));
        return Ptr(new Vec3SGrid(ctx, g));
    }
    /// the named Vec3f grid of a .nvdb file (vdb_render -color, openvdb_cmd/vdb_render/main.cc:788-795)
    static Ptr read(Context& ctx, const std::string& fileName, const std::string& gridName)
    {
        void* buf = nullptr; uint64_t bytes = 0;
        check(vdbrt_nvdb_read_typed(fileName.c_str(), gridName.c_str(), 6u, &buf, &bytes));
        vdbrt_grid* g = nullptr;
        const int rc = vdbrt_upload_color_grid(ctx.get(), buf, bytes, VDBRT_MEM_HOST, &g);
        vdbrt_buffer_free(buf);
        check(rc);
        return Ptr(new Vec3SGrid(ctx, g));
    }
    ~Vec3SGrid() { vdbrt_free_grid(mCtx->get(), mGrid); }
    const vdbrt_grid* get() const { return mGrid; }
private:
    Vec3SGrid(Context& ctx, vdbrt_grid* g) : mCtx(&ctx), mGrid(g) {}
    Context* mCtx; vdbrt_grid* mGrid;
};

namespace tools {

/// tools::Film (RayTracer.h:226-345): pixels live in pinned host memory so the copies to/from the GPU run at full speed.
class Film
{
public:
    struct RGBA {
        using ValueT = float;
        RGBA() : r(0), g(0), b(0), a(1) {}
        explicit RGBA(ValueT i) : r(i), g(i), b(i), a(1) {}
        RGBA(ValueT _r, ValueT _g, ValueT _b, ValueT _a = 1.0f) : r(_r), g(_g), b(_b), a(_a) {}
        RGBA(double _r, double _g, double _b, double _a = 1.0) : r(ValueT(_r)), g(ValueT(_g)), b(ValueT(_b)), a(ValueT(_a)) {}
        RGBA operator*(ValueT s) const { return RGBA(r * s, g * s, b * s); }
        RGBA operator+(const RGBA& o) const { return RGBA(r + o.r, g + o.g, b + o.b); }
        RGBA operator*(const RGBA& o) const { return RGBA(r * o.r, g * o.g, b * o.b); }
        RGBA& operator+=(const RGBA& o) { r += o.r; g += o.g; b += o.b; a += o.a; return *this; }
        void over(const RGBA& rhs) { const float s = rhs.a * (1.0f - a); r = a * r + s * rhs.r; g = a * g + s * rhs.g; b = a * b + s * rhs.b; a = a + s; }
        ValueT r, g, b, a;
    };
    Film(size_t width, size_t height) : mWidth(width), mHeight(height), mSize(width * height) { alloc(); fill(RGBA()); }
    Film(size_t width, size_t height, const RGBA& bg) : mWidth(width), mHeight(height), mSize(width * height) { alloc(); fill(bg); }
    ~Film() { vdbrt_host_free(mPixels); }
    Film(const Film&) = delete;
    Film& operator=(const Film&) = delete;
    const RGBA& pixel(size_t w, size_t h) const { return mPixels[w + h * mWidth]; }
    RGBA& pixel(size_t w, size_t h) { mUniform = false; return mPixels[w + h * mWidth]; }
    void fill(const RGBA& rgb = RGBA(0)) { for (size_t i = 0; i < mSize; ++i) mPixels[i] = rgb; mUniform = true; mFill = rgb; }
    void checkerboard(const RGBA& c1 = RGBA(0.3f), const RGBA& c2 = RGBA(0.6f), size_t size = 32)
    {
        RGBA* p = mPixels;
        for (size_t j = 0; j < mHeight; ++j) for (size_t i = 0; i < mWidth; ++i, ++p) *p = ((i & size) ^ (j & size)) ? c1 : c2;
        mUniform = false;
    }
    template<typename Type = unsigned char>
    std::unique_ptr<Type[]> convertToBitBuffer(const bool alpha = true) const
    {
        const size_t totalSize = mSize * (alpha ? 4 : 3);
        std::unique_ptr<Type[]> buffer(new Type[totalSize]);
        Type* q = buffer.get();
        const RGBA* p = mPixels;
        size_t n = mSize;
        while (n--) { *q++ = Type(255.0f * p->r); *q++ = Type(255.0f * p->g); *q++ = Type(255.0f * p->b); if (alpha) *q++ = Type(255.0f * p->a); ++p; }
        return buffer;
    }
    void savePPM(const std::string& fileName);   // defined below (same P6 layout as the reference, RayTracer.h:319-335)
    size_t width() const { return mWidth; }
    size_t height() const { return mHeight; }
    size_t numPixels() const { return mSize; }
    const RGBA* pixels() const { return mPixels; }
    // facade internals: a film that was only ever fill()ed lets the library skip the host->device copy
    bool uniform(RGBA& c) const { c = mFill; return mUniform; }
    RGBA* data() { return mPixels; }
    void markRendered() { mUniform = false; }
private:
    void alloc() { void* p = nullptr; check(vdbrt_host_alloc(mSize * sizeof(RGBA) + 16, &p)); mPixels = static_cast<RGBA*>(p); }
    size_t mWidth, mHeight, mSize;
    RGBA* mPixels = nullptr;
    bool mUniform = true;
    RGBA mFill;
};

/// math::Ray<double> as the facade's callers hand it over: eye, direction, times (math/Ray.h:57-63: t0 = 1e-9, t1 = max by default)
struct Ray : vdbrt_ray
{
    Ray(const Vec3R& e = Vec3R(0.0), const Vec3R& d = Vec3R(1.0, 0.0, 0.0), double t0_ = 1e-9, double t1_ = std::numeric_limits<double>::max())
    { eye[0] = e.x; eye[1] = e.y; eye[2] = e.z; dir[0] = d.x; dir[1] = d.y; dir[2] = d.z; t0 = t0_; t1 = t1_; }
    explicit Ray(const vdbrt_ray& r) : vdbrt_ray(r) {}
    Vec3R operator()(double t) const { return Vec3R(eye[0] + dir[0] * t, eye[1] + dir[1] * t, eye[2] + dir[2] * t); }      // math/Ray.h:109
};

class BaseCamera
{
public:
    virtual ~BaseCamera() = default;
    Film::RGBA& pixel(size_t i, size_t j) { return mFilm->pixel(i, j); }
    size_t width() const { return mFilm->width(); }
    size_t height() const { return mFilm->height(); }
    /// BaseCamera::lookAt (RayTracer.h:379-389)
    void lookAt(const Vec3R& xyz, const Vec3R& up = Vec3R(0.0, 1.0, 0.0))
    {
        const double t[3] = {xyz.x, xyz.y, xyz.z}, u[3] = {up.x, up.y, up.z};
        check(vdbrt_camera_look_at(&mPod, t, u));
    }
    /// BaseCamera::rasterToScreen (RayTracer.h:391-395)
    Vec3R rasterToScreen(double i, double j, double z) const
    {
        return Vec3R((2 * i / double(mFilm->width()) - 1) * mPod.scale_w, (1 - 2 * j / double(mFilm->height())) * mPod.scale_h, z);
    }
    /// getRay (RayTracer.h:452-462, 505-512): the world-space ray through pixel (i, j); the offsets in [0, 1], 0.5 = the pixel's centre
    Ray getRay(size_t i, size_t j, double iOffset = 0.5, double jOffset = 0.5) const
    {
        const uint32_t ij[2] = {uint32_t(i), uint32_t(j)};
        const double off[2] = {iOffset, jOffset};
        vdbrt_ray r;
        check(vdbrt_camera_get_rays(&mPod, ij, off, 1, &r));
        return Ray(r);
    }
    const vdbrt_camera& pod() const { return mPod; }
    Film& film() const { return *mFilm; }
protected:
    explicit BaseCamera(Film& film) : mFilm(&film) {}
    Film* mFilm;
    vdbrt_camera mPod;
};

class PerspectiveCamera : public BaseCamera
{
public:
    PerspectiveCamera(Film& film, const Vec3R& rotation = Vec3R(0.0), const Vec3R& translation = Vec3R(0.0), double focalLength = 50.0,
                      double aperture = 41.2136, double nearPlane = 1e-3, double farPlane = std::numeric_limits<double>::max())
        : BaseCamera(film)
    {
        const double r[3] = {rotation.x, rotation.y, rotation.z}, t[3] = {translation.x, translation.y, translation.z};
        check(vdbrt_camera_perspective(&mPod, uint32_t(film.width()), uint32_t(film.height()), r, t, focalLength, aperture, nearPlane, farPlane));
    }
    /// horizontal field of view in degrees from a focal length and an aperture in mm, and back (RayTracer.h:466-475)
    static double focalLengthToFieldOfView(double length, double aperture) { return 360.0 / 3.14159265358979323846 * std::atan(aperture / (2.0 * length)); }
    static double fieldOfViewToFocalLength(double fov, double aperture) { return aperture / (2.0 * (std::tan(fov * 3.14159265358979323846 / 360.0))); }
};

class OrthographicCamera : public BaseCamera
{
public:
    OrthographicCamera(Film& film, const Vec3R& rotation = Vec3R(0.0), const Vec3R& translation = Vec3R(0.0), double frameWidth = 1.0,
                       double nearPlane = 1e-3, double farPlane = std::numeric_limits<double>::max())
        : BaseCamera(film)
    {
        const double r[3] = {rotation.x, rotation.y, rotation.z}, t[3] = {translation.x, translation.y, translation.z};
        check(vdbrt_camera_orthographic(&mPod, uint32_t(film.width()), uint32_t(film.height()), r, t, frameWidth, nearPlane, farPlane));
    }
};

/// BaseShader: the device runs the four constant-colour shaders; the object only carries their parameters.
class BaseShader
{
public:
    virtual ~BaseShader() = default;
    virtual BaseShader* copy() const = 0;
    const vdbrt_shader& pod() const { return mPod; }
protected:
    BaseShader(uint32_t kind, const Film::RGBA& c) { std::memset(&mPod, 0, sizeof(mPod)); mPod.kind = kind; mPod.rgba[0] = c.r; mPod.rgba[1] = c.g; mPod.rgba[2] = c.b; mPod.rgba[3] = c.a; }
    vdbrt_shader mPod;
};
// the reference's default template argument GridT = Film::RGBA selects the constant-colour specialisation (RayTracer.h:565,614,671,728)
template<typename GridT = Film::RGBA> class MatteShader;
template<typename GridT = Film::RGBA> class NormalShader;
template<typename GridT = Film::RGBA> class PositionShader;
template<typename GridT = Film::RGBA> class DiffuseShader;
template<> class MatteShader<Film::RGBA> : public BaseShader { public:
    MatteShader(const Film::RGBA& c = Film::RGBA(1.0f)) : BaseShader(VDBRT_SHADER_MATTE, c) {}
    BaseShader* copy() const override { return new MatteShader(*this); } };
template<> class NormalShader<Film::RGBA> : public BaseShader { public:
    NormalShader(const Film::RGBA& c = Film::RGBA(1.0f)) : BaseShader(VDBRT_SHADER_NORMAL, c) {}
    BaseShader* copy() const override { return new NormalShader(*this); } };
template<> class PositionShader<Film::RGBA> : public BaseShader { public:
    /// bbox in world space: min and max corners (math::BBox<Vec3R>)
    PositionShader(const Vec3R& bboxMin, const Vec3R& bboxMax, const Film::RGBA& c = Film::RGBA(1.0f)) : BaseShader(VDBRT_SHADER_POSITION, c)
    {
        mPod.bbox_min[0] = bboxMin.x; mPod.bbox_min[1] = bboxMin.y; mPod.bbox_min[2] = bboxMin.z;
        mPod.inv_dim[0] = 1.0 / (bboxMax.x - bboxMin.x); mPod.inv_dim[1] = 1.0 / (bboxMax.y - bboxMin.y); mPod.inv_dim[2] = 1.0 / (bboxMax.z - bboxMin.z);
    }
    BaseShader* copy() const override { return new PositionShader(*this); } };
template<> class DiffuseShader<Film::RGBA> : public BaseShader { public:
    DiffuseShader(const Film::RGBA& d = Film::RGBA(1.0f)) : BaseShader(VDBRT_SHADER_DIFFUSE, d) {}
    BaseShader* copy() const override { return new DiffuseShader(*this); } };

// the colour-grid forms (RayTracer.h:542-562, 591-611, 640-668, 702-725; default PointSampler): colour = the grid's value at the
// voxel nearest to the hit position.  The grid must outlive the shader (the reference's shaders hold an accessor the same way).
template<> class MatteShader<Vec3SGrid> : public BaseShader { public:
    MatteShader(const Vec3SGrid& grid) : BaseShader(VDBRT_SHADER_MATTE, Film::RGBA(1.0f)) { mPod.color_grid = grid.get(); }
    BaseShader* copy() const override { return new MatteShader(*this); } };
template<> class NormalShader<Vec3SGrid> : public BaseShader { public:
    NormalShader(const Vec3SGrid& grid) : BaseShader(VDBRT_SHADER_NORMAL, Film::RGBA(1.0f)) { mPod.color_grid = grid.get(); }
    BaseShader* copy() const override { return new NormalShader(*this); } };
template<> class PositionShader<Vec3SGrid> : public BaseShader { public:
    PositionShader(const Vec3R& bboxMin, const Vec3R& bboxMax, const Vec3SGrid& grid) : BaseShader(VDBRT_SHADER_POSITION, Film::RGBA(1.0f))
    {
        mPod.bbox_min[0] = bboxMin.x; mPod.bbox_min[1] = bboxMin.y; mPod.bbox_min[2] = bboxMin.z;
        mPod.inv_dim[0] = 1.0 / (bboxMax.x - bboxMin.x); mPod.inv_dim[1] = 1.0 / (bboxMax.y - bboxMin.y); mPod.inv_dim[2] = 1.0 / (bboxMax.z - bboxMin.z);
        mPod.color_grid = grid.get();
    }
    BaseShader* copy() const override { return new PositionShader(*this); } };
template<> class DiffuseShader<Vec3SGrid> : public BaseShader { public:
    DiffuseShader(const Vec3SGrid& grid) : BaseShader(VDBRT_SHADER_DIFFUSE, Film::RGBA(1.0f)) { mPod.color_grid = grid.get(); }
    BaseShader* copy() const override { return new DiffuseShader(*this); } };

/// tools::LinearSearchImpl<GridT, Iterations> (RayIntersector.h:514-668) as a tag: the facade's intersector only needs the iteration count
template<typename GridT, int Iterations = 0>
struct LinearSearchImpl { static constexpr uint32_t iterations = Iterations; };

/// tools::LevelSetRayIntersector (RayIntersector.h:79-246): validates at construction; the twelve intersectsIS / intersectsWS overloads
/// of the reference for one ray (one device call each), and the same for batches of rays (the shape to use on a GPU)
template<typename GridT = FloatGrid, typename SearchImplT = LinearSearchImpl<GridT, 0>>
class LevelSetRayIntersector
{
public:
    using RayType = Ray;
    LevelSetRayIntersector(const GridT& grid, float isoValue = 0.0f) : mGrid(&grid), mIso(isoValue)
    {
        // same construction-time checks as the reference, evaluated by the library (empty batch = validation only)
        check(vdbrt_intersect_levelset(grid.context().get(), grid.get(), nullptr, 0, VDBRT_SPACE_WORLD, isoValue, nullptr, VDBRT_MEM_HOST));
    }
    const float& getIsoValue() const { return mIso; }
    static constexpr uint32_t iterations() { return SearchImplT::iterations; }
    /// intersectsWS / intersectsIS for n rays at once; hits[i].hit == 0 leaves the record zeroed (outputs untouched)
    void intersectsWS(const vdbrt_ray* rays, size_t n, vdbrt_hit* hits) const { batch(rays, n, VDBRT_SPACE_WORLD, hits); }
    void intersectsIS(const vdbrt_ray* rays, size_t n, vdbrt_hit* hits) const { batch(rays, n, VDBRT_SPACE_INDEX, hits); }
    // ---- index-space rays (RayIntersector.h:119-160); outputs are untouched on a miss, as in the reference
    bool intersectsIS(const vdbrt_ray& iRay) const { vdbrt_hit h; batch(&iRay, 1, VDBRT_SPACE_INDEX, &h); return h.hit != 0; }
    bool intersectsIS(const vdbrt_ray& iRay, double& iTime) const
    { vdbrt_hit h; batch(&iRay, 1, VDBRT_SPACE_INDEX, &h); if (!h.hit) return false; iTime = h.t_index; return true; }
    bool intersectsIS(const vdbrt_ray& iRay, Vec3R& xyz) const
    { vdbrt_hit h; batch(&iRay, 1, VDBRT_SPACE_INDEX, &h); if (!h.hit) return false; xyz = vec(h.xyz_index); return true; }
    bool intersectsIS(const vdbrt_ray& iRay, Vec3R& xyz, double& iTime) const
    { vdbrt_hit h; batch(&iRay, 1, VDBRT_SPACE_INDEX, &h); if (!h.hit) return false; xyz = vec(h.xyz_index); iTime = h.t_index; return true; }
    // ---- world-space rays (RayIntersector.h:162-240)
    bool intersectsWS(const vdbrt_ray& wRay) const { vdbrt_hit h; batch(&wRay, 1, VDBRT_SPACE_WORLD, &h); return h.hit != 0; }
    bool intersectsWS(const vdbrt_ray& wRay, double& wTime) const
    { vdbrt_hit h; batch(&wRay, 1, VDBRT_SPACE_WORLD, &h); if (!h.hit) return false; wTime = h.t_world; return true; }
    bool intersectsWS(const vdbrt_ray& wRay, Vec3R& world) const
    { vdbrt_hit h; batch(&wRay, 1, VDBRT_SPACE_WORLD, &h); if (!h.hit) return false; world = vec(h.xyz_world); return true; }
    bool intersectsWS(const vdbrt_ray& wRay, Vec3R& world, double& wTime) const
    { vdbrt_hit h; batch(&wRay, 1, VDBRT_SPACE_WORLD, &h); if (!h.hit) return false; world = vec(h.xyz_world); wTime = h.t_world; return true; }
    bool intersectsWS(const vdbrt_ray& wRay, Vec3R& world, Vec3R& normal) const
    { vdbrt_hit h; batch(&wRay, 1, VDBRT_SPACE_WORLD, &h); if (!h.hit) return false; world = vec(h.xyz_world); normal = vec(h.nml); return true; }
    bool intersectsWS(const vdbrt_ray& wRay, Vec3R& world, Vec3R& normal, double& wTime) const
    {
        vdbrt_hit h; batch(&wRay, 1, VDBRT_SPACE_WORLD, &h);
        if (!h.hit) return false;
        world = vec(h.xyz_world); normal = vec(h.nml); wTime = h.t_world;
        return true;
    }
    const GridT& grid() const { return *mGrid; }
private:
    static Vec3R vec(const double* v) { return Vec3R(v[0], v[1], v[2]); }
    void batch(const vdbrt_ray* rays, size_t n, uint32_t space, vdbrt_hit* hits) const
    { check(vdbrt_intersect_levelset_ex(mGrid->context().get(), mGrid->get(), rays, n, space, mIso, SearchImplT::iterations, hits, VDBRT_MEM_HOST)); }
    const GridT* mGrid; float mIso;
};

/// tools::VolumeRayIntersector (RayIntersector.h:277-485): setIndexRay / setWorldRay + march() / hits() for one ray (each march is one
/// device call), hits() for batches of rays.  The bbox is the node-granular one, max padded by one (RayIntersector.h:318).
template<typename GridT = FloatGrid>
class VolumeRayIntersector
{
public:
    struct TimeSpan {                                                   // math::Ray::TimeSpan (math/Ray.h:38-55)
        double t0, t1;
        TimeSpan(double a = -1.0, double b = -1.0) : t0(a), t1(b) {}
        bool valid(double eps = 1e-9) const { return (t1 - t0) > eps; }
        void get(double& a, double& b) const { a = t0; b = t1; }
    };
    explicit VolumeRayIntersector(const GridT& grid) : mGrid(&grid)
    { check(vdbrt_volume_spans(grid.context().get(), grid.get(), nullptr, 0, VDBRT_SPACE_WORLD, 0, nullptr, nullptr, VDBRT_MEM_HOST)); }
    /// spans[i*maxSpans + k] = {t0,t1}; counts[i] = -1 when the ray misses the bbox
    void hits(const vdbrt_ray* rays, size_t n, bool indexSpace, uint32_t maxSpans, double* spans, int32_t* counts) const
    { check(vdbrt_volume_spans(mGrid->context().get(), mGrid->get(), rays, n, indexSpace ? VDBRT_SPACE_INDEX : VDBRT_SPACE_WORLD, maxSpans, spans, counts, VDBRT_MEM_HOST)); }
    /// setIndexRay (:368-374): false if the ray misses the bbox; the ray is clipped to it and mTmax = its exit time
    bool setIndexRay(const vdbrt_ray& iRay) { return this->start(iRay, VDBRT_SPACE_INDEX); }
    /// setWorldRay (:387-390) = setIndexRay(wRay.worldToIndex(grid))
    bool setWorldRay(const vdbrt_ray& wRay) { return this->start(wRay, VDBRT_SPACE_WORLD); }
    /// march (:392-397): the next span of active values along the current ray (index-space times); invalid when there is none.
    /// Afterwards the ray starts Delta<double> behind the span, exactly as the reference re-arms mRay.
    TimeSpan march()
    {
        TimeSpan t(-1.0, -1.0);
        if (!mArmed || !((mRay.t1 - mRay.t0) > 1e-5f)) return t;       // VolumeHDDA::march: if (ray.valid()) (Ray::valid, eps = Delta<float>)
        double span[2] = {-1.0, -1.0}; int32_t count = 0;
        check(vdbrt_volume_spans(mGrid->context().get(), mGrid->get(), &mRay, 1, VDBRT_SPACE_INDEX, 1, span, &count, VDBRT_MEM_HOST));
        if (count > 0) { t.t0 = span[0]; t.t1 = span[1]; }
        if (t.t1 > 0) { mRay.t0 = t.t1 + 1e-9; mRay.t1 = mTmax; }      // mRay.setTimes(t.t1 + Delta<RealType>, mTmax)
        else mArmed = false;
        return t;
    }
    bool march(double& t0, double& t1) { const TimeSpan t = this->march(); t.get(t0, t1); return t.valid(); }
    /// hits (:428-432): all spans of the current ray
    template<typename ListType> void hits(ListType& list)
    {
        list.clear();
        if (!mArmed) return;
        uint32_t cap = 16;
        for (;;) {
            std::vector<double> spans(2 * size_t(cap)); int32_t count = 0;
            check(vdbrt_volume_spans(mGrid->context().get(), mGrid->get(), &mRay, 1, VDBRT_SPACE_INDEX, cap, spans.data(), &count, VDBRT_MEM_HOST));
            if (count > int32_t(cap)) { cap = uint32_t(count); continue; }
            for (int32_t k = 0; k < count; ++k) list.push_back(TimeSpan(spans[2 * k], spans[2 * k + 1]));
            return;
        }
    }
    Vec3R getIndexPos(double time) const { return Vec3R(mRay.eye[0] + mRay.dir[0] * time, mRay.eye[1] + mRay.dir[1] * time, mRay.eye[2] + mRay.dir[2] * time); }
    /// getWorldPos (:440) = grid.indexToWorld(ray(time)): ScaleMap / ScaleTranslateMap, one multiply (and one add) per component
    Vec3R getWorldPos(double time) const
    {
        const Vec3R p = getIndexPos(time); const vdbrt_grid_info i = mGrid->info();
        const bool tr = i.translation[0] != 0 || i.translation[1] != 0 || i.translation[2] != 0;
        return tr ? Vec3R(p.x * mScale[0] + i.translation[0], p.y * mScale[1] + i.translation[1], p.z * mScale[2] + i.translation[2])
                  : Vec3R(p.x * mScale[0], p.y * mScale[1], p.z * mScale[2]);
    }
    /// getWorldTime (:442-445): time * |J dir| of the current index ray, whose direction has unit length (diagonal maps, like getWorldPos above; a grid with a general affine
    /// map -- the tolerance path -- reports its hit positions and times through the batch calls of the C ABI instead)
    double getWorldTime(double time) const
    {
        const double x = mRay.dir[0] * mScale[0], y = mRay.dir[1] * mScale[1], z = mRay.dir[2] * mScale[2];
        return time * std::sqrt(x * x + y * y + z * z);
    }
    /// print (:459-469): "BBox: [min] -> [max]" (levels 2 and 3 of the reference add statistics of its bool tree, which does not exist here)
    void print(std::ostream& os = std::cout, int verboseLevel = 1) const
    {
        if (verboseLevel > 0) {
            const vdbrt_grid_info i = mGrid->info();
            os << "BBox: [" << i.node_bbox[0] << ", " << i.node_bbox[1] << ", " << i.node_bbox[2] << "] -> [" << i.node_bbox[3] + 1 << ", "
               << i.node_bbox[4] + 1 << ", " << i.node_bbox[5] + 1 << "]" << std::endl;
        }
    }
    const GridT& grid() const { return *mGrid; }
private:
    bool start(const vdbrt_ray& ray, uint32_t space)
    {
        // the library clips the ray and returns its clipped index-space form: one span query with room for none
        mArmed = false;
        vdbrt_ray clipped;
        int hit = 0;
        check(vdbrt_volume_clip(mGrid->context().get(), mGrid->get(), &ray, space, &clipped, &hit, mScale));
        if (!hit) return false;
        mRay = clipped; mTmax = clipped.t1; mArmed = true;
        return true;
    }
    const GridT* mGrid;
    vdbrt_ray mRay{}; double mTmax = 0.0; bool mArmed = false; double mScale[3] = {1.0, 1.0, 1.0};
};

/// tools::LevelSetRayTracer (RayTracer.h:72-140,789-918)
template<typename GridT = FloatGrid, typename IntersectorT = LevelSetRayIntersector<GridT>>
class LevelSetRayTracer
{
public:
    LevelSetRayTracer(const GridT& grid, const BaseShader& shader, BaseCamera& camera, size_t pixelSamples = 1, unsigned int seed = 0)
        : mInter(grid), mShader(shader.copy()), mCamera(&camera) { setPixelSamples(pixelSamples, seed); }
    LevelSetRayTracer(const IntersectorT& inter, const BaseShader& shader, BaseCamera& camera, size_t pixelSamples = 1, unsigned int seed = 0)
        : mInter(inter), mShader(shader.copy()), mCamera(&camera) { setPixelSamples(pixelSamples, seed); }
    void setGrid(const GridT& grid) { mInter = IntersectorT(grid); }
    void setIntersector(const IntersectorT& inter) { mInter = inter; }
    void setShader(const BaseShader& shader) { mShader.reset(shader.copy()); }
    void setCamera(BaseCamera& camera) { mCamera = &camera; }
    void setPixelSamples(size_t pixelSamples, unsigned int seed = 0)
    {
        if (pixelSamples == 0) throw ValueError("pixelSamples must be larger than zero!");       // RayTracer.h:877-879
        std::memset(&mOpts, 0, sizeof(mOpts));
        mOpts.spp = uint32_t(pixelSamples);
        if (pixelSamples > 1) check(vdbrt_jitter_table(seed, mOpts.jitter));
    }
    /// the `threaded` flag of the reference is accepted and ignored: the frame is rendered by the GPU either way
    void render(bool /*threaded*/ = true) const
    {
        Film& film = mCamera->film();
        vdbrt_ls_opts o = mOpts;
        o.iso = mInter.getIsoValue();
        vdbrt_film f; std::memset(&f, 0, sizeof(f));
        f.pixels = reinterpret_cast<float*>(film.data()); f.width = uint32_t(film.width()); f.height = uint32_t(film.height()); f.memspace = VDBRT_MEM_HOST;
        Film::RGBA bg;
        if (film.uniform(bg)) { o.flags |= VDBRT_LS_UNIFORM_BG; f.bg_rgba[0] = bg.r; f.bg_rgba[1] = bg.g; f.bg_rgba[2] = bg.b; f.bg_rgba[3] = bg.a; }
        check(vdbrt_render_levelset(mInter.grid().context().get(), mInter.grid().get(), &mCamera->pod(), &mShader->pod(), &o, &f, nullptr));
        film.markRendered();
    }
private:
    IntersectorT mInter;
    std::unique_ptr<const BaseShader> mShader;
    BaseCamera* mCamera;
    vdbrt_ls_opts mOpts;
};

/// tools::rayTrace (RayTracer.h:48-64,758-783)
template<typename GridT>
inline void rayTrace(const GridT& grid, const BaseShader& shader, BaseCamera& camera, size_t pixelSamples = 1, unsigned int seed = 0, bool threaded = true)
{
    LevelSetRayTracer<GridT, LevelSetRayIntersector<GridT>> tracer(grid, shader, camera, pixelSamples, seed);
    tracer.render(threaded);
}
template<typename GridT, typename IntersectorT>
inline void rayTrace(const GridT&, const IntersectorT& inter, const BaseShader& shader, BaseCamera& camera, size_t pixelSamples = 1, unsigned int seed = 0, bool threaded = true)
{
    LevelSetRayTracer<GridT, IntersectorT> tracer(inter, shader, camera, pixelSamples, seed);
    tracer.render(threaded);
}

/// tools::VolumeRender (RayTracer.h:148-220,922-1070)
template<typename IntersectorT = VolumeRayIntersector<FloatGrid>>
class VolumeRender
{
public:
    VolumeRender(const IntersectorT& inter, BaseCamera& camera) : mInter(inter), mCamera(&camera) { check(vdbrt_vol_opts_default(&mOpts)); }
    void setCamera(BaseCamera& camera) { mCamera = &camera; }
    void setIntersector(const IntersectorT& inter) { mInter = inter; }
    /// throws (like Vec3::unit) if the vector is null
    void setLightDir(double x, double y, double z)
    {
        const double len = std::sqrt(x * x + y * y + z * z);
        if (!(len > 0.0)) throw RuntimeError("Normalizing null 3-vector");
        mOpts.light_dir[0] = x / len; mOpts.light_dir[1] = y / len; mOpts.light_dir[2] = z / len;
    }
    void setLightColor(double r, double g, double b) { mOpts.light_color[0] = r; mOpts.light_color[1] = g; mOpts.light_color[2] = b; }
    void setPrimaryStep(double s) { mOpts.primary_step = s; }
    void setShadowStep(double s) { mOpts.shadow_step = s; }
    void setScattering(double x, double y, double z) { mOpts.scattering[0] = x; mOpts.scattering[1] = y; mOpts.scattering[2] = z; }
    void setAbsorption(double x, double y, double z) { mOpts.absorption[0] = x; mOpts.absorption[1] = y; mOpts.absorption[2] = z; }
    void setLightGain(double g) { mOpts.light_gain = g; }
    void setCutOff(double c) { mOpts.cutoff = c; }
    /// print (RayTracer.h:958-973)
    void print(std::ostream& os = std::cout, int verboseLevel = 1)
    {
        auto v3 = [](const double* v) { std::ostringstream b; b << "[" << v[0] << ", " << v[1] << ", " << v[2] << "]"; return b.str(); };
        if (verboseLevel > 0) {
            os << "\nPrimary step: " << mOpts.primary_step << "\nShadow step: " << mOpts.shadow_step << "\nCutoff: " << mOpts.cutoff
               << "\nLightGain: " << mOpts.light_gain << "\nLightDir: " << v3(mOpts.light_dir) << "\nLightColor: " << v3(mOpts.light_color)
               << "\nAbsorption: " << v3(mOpts.absorption) << "\nScattering: " << v3(mOpts.scattering) << std::endl;
        }
        mInter.print(os, verboseLevel);
    }
    void render(bool /*threaded*/ = true) const
    {
        Film& film = mCamera->film();
        vdbrt_film f; std::memset(&f, 0, sizeof(f));
        f.pixels = reinterpret_cast<float*>(film.data()); f.width = uint32_t(film.width()); f.height = uint32_t(film.height()); f.memspace = VDBRT_MEM_HOST;
        check(vdbrt_render_volume(mInter.grid().context().get(), mInter.grid().get(), &mCamera->pod(), &mOpts, &f));
        film.markRendered();
    }
private:
    IntersectorT mInter;
    BaseCamera* mCamera;
    vdbrt_vol_opts mOpts;
};

} // namespace tools
} // namespace vdbrt

#include <cmath>
#include <fstream>
#include <iostream>
inline void vdbrt::tools::Film::savePPM(const std::string& fileName)
{
    std::string name(fileName);
    if (name.find_last_of(".") == std::string::npos) name.append(".ppm");
    std::ofstream os(name.c_str(), std::ios_base::binary);
    if (!os.is_open()) { std::cerr << "Error opening PPM file \"" << name << "\"" << std::endl; return; }
    auto buf = this->convertToBitBuffer<unsigned char>(/*alpha=*/false);
    os << "P6\n" << mWidth << " " << mHeight << "\n255\n";
    os.write(reinterpret_cast<const char*>(buf.get()), 3 * mSize * sizeof(unsigned char));
}
