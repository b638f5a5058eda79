// vdbrt_oracle.cc -- TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or shipped with the product.
//
// CPU restatement of OpenVDB 13.0.1's ray-tracing hot path over a NanoVDB-serialised NanoGrid<float>:
// scalar C++, no reference headers, every function citing the reference file:line it follows (paths relative to
// /root/reference/openvdb/openvdb unless they start with nanovdb/).  It is the checker the CUDA kernels are
// compared with in tests/, __graft_entry__.smoke() and the "port" leg of bench.py; it is itself pinned against the
// unmodified reference (oracle/_ref/libvdbref.so) by tests/test_oracle_vs_reference.py and against the
// reference's own known-answer tests by tests/test_reference_kats.py.
//
// Arithmetic contract: rays, DDA and hit times are double; voxel values, the BoxStencil
// interpolation and gradient are float; nothing may be contracted into FMA (build: plain -O2 -ffp-contract=off,
// no -march/-mfma/-ffast-math), operand order follows the reference expression by expression.
#include "vdbrt_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

// ------------------------------------------------------------------------------------------------------------
// NanoVDB layout (nanovdb/nanovdb/NanoVDB.h:67-122 diagram; GridData :1944-2135, TreeData :2393-2423,
// RootData/Tile :2621-2695, InternalData :3181-3291, LeafData<float> :3671-3746).  Only offsets are restated.
// ------------------------------------------------------------------------------------------------------------
constexpr uint64_t MAGIC_NUMB = 0x304244566f6e614eULL, MAGIC_GRID = 0x314244566f6e614eULL; // NanoVDB.h:139-140
constexpr size_t GRID_SIZE = 672, TREE_SIZE = 64;
constexpr size_t OFF_VERSION = 16, OFF_GRIDCOUNT = 28, OFF_GRIDSIZE = 32, OFF_MATD = 384, OFF_VECD = 528,
                 OFF_VOXELSIZE = 608, OFF_CLASS = 632, OFF_TYPE = 636;
constexpr size_t ROOT_TABLESIZE = 24, ROOT_BACKGROUND = 28, ROOT_TILES = 64, TILE_SIZE = 32;
constexpr size_t UPPER_VMASK = 32, UPPER_CMASK = 32 + 4096, UPPER_TABLE = 8256, UPPER_SIZE = 270400;
constexpr size_t LOWER_VMASK = 32, LOWER_CMASK = 32 + 512, LOWER_TABLE = 1088, LOWER_SIZE = 33856;
constexpr size_t LEAF_VMASK = 16, LEAF_VALUES = 96, LEAF_SIZE = 2144;

template<typename T> inline T rd(const uint8_t* p) { T v; std::memcpy(&v, p, sizeof(T)); return v; }
inline bool maskBit(const uint8_t* mask, uint32_t n) { return (rd<uint64_t>(mask + 8 * (n >> 6)) >> (n & 63)) & 1; } // NanoVDB.h:1238

struct Coord { int32_t x, y, z;
    int32_t& operator[](int i) { return (&x)[i]; }
    int32_t operator[](int i) const { return (&x)[i]; }
    bool operator!=(const Coord& o) const { return x != o.x || y != o.y || z != o.z; } };

// RootData::CoordToKey, NANOVDB_USE_SINGLE_ROOT_KEY (NanoVDB.h:2630-2640)
inline uint64_t rootKey(const Coord& c) {
    return (uint64_t(uint32_t(c.z) >> 12)) | (uint64_t(uint32_t(c.y) >> 12) << 21) | (uint64_t(uint32_t(c.x) >> 12) << 42);
}
inline uint32_t upperOffset(const Coord& c) { return (((c.x & 4095) >> 7) << 10) | (((c.y & 4095) >> 7) << 5) | ((c.z & 4095) >> 7); } // :3560-3565
inline uint32_t lowerOffset(const Coord& c) { return (((c.x & 127) >> 3) << 8) | (((c.y & 127) >> 3) << 4) | ((c.z & 127) >> 3); }
inline uint32_t leafOffset(const Coord& c) { return ((c.x & 7) << 6) | ((c.y & 7) << 3) | (c.z & 7); } // :4516-4519

// Leaf values.  NanoGrid<float>: mValues[n] at +96 (NanoVDB.h:3671-3746).  Quantised grids (GridType 13..16 = Fp4, Fp8, Fp16,
// FpN) share the 96-byte header LeafFnBase {.. float mMinimum @80, float mQuantum @84 ..} followed by packed codes, and
// LeafData<FpX>::getValue(n) = float(code_n) * mQuantum + mMinimum (NanoVDB.h:3843, 3876, 3906, 3961); FpN keeps log2 of its bit
// width in mFlags >> 5 (:3933).  nanovdb::tools::nanoToOpenVDB builds the FloatGrid the reference ray-traces from exactly these
// values (tools/NanoToOpenVDB.h:511-518).  Compiled with -ffp-contract=off: product and sum round separately.
inline float leafValue(const uint8_t* lf, uint32_t n, uint32_t gridType) {
    if (gridType == 1) return rd<float>(lf + LEAF_VALUES + 4 * n);
    const uint32_t log2w = gridType == 13 ? 2u : gridType == 14 ? 3u : gridType == 15 ? 4u : uint32_t(lf[15] >> 5);
    const uint8_t* codes = lf + LEAF_VALUES;
    uint32_t code;
    switch (log2w) {
    case 0: code = (codes[n >> 3] >> (n & 7)) & 1u; break;
    case 1: code = (codes[n >> 2] >> ((n & 3) << 1)) & 3u; break;
    case 2: code = (codes[n >> 1] >> ((n & 1) << 2)) & 15u; break;
    case 3: code = codes[n]; break;
    default: code = rd<uint16_t>(codes + 2 * n); break;
    }
    const float product = float(code) * rd<float>(lf + 84);
    return product + rd<float>(lf + 80);
}

struct CoordBBox { Coord mn{INT_MAX, INT_MAX, INT_MAX}, mx{INT_MIN, INT_MIN, INT_MIN};
    void expand(const Coord& lo, int dim) { // math::CoordBBox::expand(min, dim)
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], lo[a]); mx[a] = std::max(mx[a], lo[a] + dim - 1); } }
    bool empty() const { return mn.x > mx.x || mn.y > mx.y || mn.z > mx.z; } };

} // namespace

struct oracle_grid {
    const uint8_t* base = nullptr; uint64_t bytes = 0;
    const uint8_t* tree = nullptr; const uint8_t* root = nullptr; const uint8_t* tiles = nullptr;
    const uint8_t* firstLeaf = nullptr; const uint8_t* firstLower = nullptr; const uint8_t* firstUpper = nullptr;
    uint32_t tableSize = 0, leafCount = 0, lowerCount = 0, upperCount = 0, gridClass = 0, gridType = 1;
    uint64_t activeVoxels = 0;
    float background = 0.f;
    double scale[3], inv[3], trans[3], voxelSize[3];
    bool hasTranslation = false;
    // a map with off-diagonal terms (rotation / shear: OpenVDB's AffineMap, math/Maps.h:411-445): NanoVDB's Map::mMatD / mInvMatD as
    // stored (NanoVDB.h:1418-1554).  TOLERANCE path: the reference multiplies through 4x4 matrices in another order of operations.
    bool general = false;
    double mat[9], imat[9];
    CoordBBox nodeBBox;   // RootNode::evalActiveBoundingBox(bbox, /*visitVoxels=*/false)  (tree/RootNode.h:1532-1541)
    int32_t indexBBox[6];
};

namespace {

// ---- tree access: semantics of tree::ValueAccessor probeConstNode / probeValue / getValue / isValueOn
// (tree/ValueAccessor.h:455-510,805-830,937-953) evaluated on the NanoVDB twin (NanoVDB.h:5252-5469).  The accessor's
// node cache is a pure cache, so every query may descend from the root.
struct Access {
    const oracle_grid& g;
    // small path cache, results identical with or without it
    Coord k2{INT_MAX, 0, 0}, k1{INT_MAX, 0, 0}, k0{INT_MAX, 0, 0};
    const uint8_t *n2 = nullptr, *n1 = nullptr, *n0 = nullptr;
    explicit Access(const oracle_grid& grid) : g(grid) {}

    const uint8_t* findTile(const Coord& c) const { // RootNode linear tile search (NanoVDB.h:2785-2799)
        const uint64_t key = rootKey(c);
        for (uint32_t i = 0; i < g.tableSize; ++i) { const uint8_t* t = g.tiles + TILE_SIZE * i; if (rd<uint64_t>(t) == key) return t; }
        return nullptr;
    }
    const uint8_t* upper(const Coord& c) { // node containing c at level 2 or null
        if (n2 && ((c.x & ~4095) == k2.x) && ((c.y & ~4095) == k2.y) && ((c.z & ~4095) == k2.z)) return n2;
        const uint8_t* t = findTile(c);
        if (!t) return nullptr;
        const int64_t child = rd<int64_t>(t + 8);
        if (child == 0) return nullptr;
        n2 = g.root + child; k2 = Coord{c.x & ~4095, c.y & ~4095, c.z & ~4095};   // child offset relative to RootData (:2692)
        return n2;
    }
    const uint8_t* lower(const Coord& c) {
        if (n1 && ((c.x & ~127) == k1.x) && ((c.y & ~127) == k1.y) && ((c.z & ~127) == k1.z)) return n1;
        const uint8_t* u = upper(c);
        if (!u) return nullptr;
        const uint32_t n = upperOffset(c);
        if (!maskBit(u + UPPER_CMASK, n)) return nullptr;
        n1 = u + rd<int64_t>(u + UPPER_TABLE + 8 * n); k1 = Coord{c.x & ~127, c.y & ~127, c.z & ~127}; // relative to this node (:3190-3199)
        return n1;
    }
    const uint8_t* leaf(const Coord& c) {
        if (n0 && ((c.x & ~7) == k0.x) && ((c.y & ~7) == k0.y) && ((c.z & ~7) == k0.z)) return n0;
        const uint8_t* l = lower(c);
        if (!l) return nullptr;
        const uint32_t n = lowerOffset(c);
        if (!maskBit(l + LOWER_CMASK, n)) return nullptr;
        n0 = l + rd<int64_t>(l + LOWER_TABLE + 8 * n); k0 = Coord{c.x & ~7, c.y & ~7, c.z & ~7};
        return n0;
    }
    // ValueAccessor::probeValue: value + active state of voxel or covering tile, background/inactive outside the root table
    bool probeValue(const Coord& c, float& v) {
        if (const uint8_t* lf = leaf(c)) { const uint32_t n = leafOffset(c); v = leafValue(lf, n, g.gridType); return maskBit(lf + LEAF_VMASK, n); }
        if (const uint8_t* l = lower(c)) { const uint32_t n = lowerOffset(c); v = rd<float>(l + LOWER_TABLE + 8 * n); return maskBit(l + LOWER_VMASK, n); }
        if (const uint8_t* u = upper(c)) { const uint32_t n = upperOffset(c); v = rd<float>(u + UPPER_TABLE + 8 * n); return maskBit(u + UPPER_VMASK, n); }
        if (const uint8_t* t = findTile(c)) { v = rd<float>(t + 20); return rd<uint32_t>(t + 16) != 0; }
        v = g.background; return false;
    }
    float getValue(const Coord& c) { float v; probeValue(c, v); return v; }
    bool isValueOn(const Coord& c) { float v; return probeValue(c, v); }
};

// RootNode/InternalNode/LeafNode::evalActiveBoundingBox(bbox, false): union of the node boxes of every leaf that has
// an active voxel and of every active tile (tree/LeafNode.h:1505-1517, InternalNode.h:1246-1256, RootNode.h:1532-1541).
CoordBBox evalNodeBBox(const oracle_grid& g)
{
    CoordBBox bb;
    for (uint32_t i = 0; i < g.tableSize; ++i) {
        const uint8_t* t = g.tiles + TILE_SIZE * i;
        const uint64_t key = rd<uint64_t>(t);
        const Coord org{int32_t(uint32_t((key >> 42) & 0x1FFFFF) << 12), int32_t(uint32_t((key >> 21) & 0x1FFFFF) << 12), int32_t(uint32_t(key & 0x1FFFFF) << 12)};
        const int64_t child = rd<int64_t>(t + 8);
        if (child == 0) { if (rd<uint32_t>(t + 16)) bb.expand(org, 4096); continue; }
        const uint8_t* u = g.root + child;
        for (uint32_t n = 0; n < 32768; ++n) {
            const Coord uo{org.x + int32_t((n >> 10) << 7), org.y + int32_t(((n >> 5) & 31) << 7), org.z + int32_t((n & 31) << 7)};
            if (!maskBit(u + UPPER_CMASK, n)) { if (maskBit(u + UPPER_VMASK, n)) bb.expand(uo, 128); continue; }
            const uint8_t* l = u + rd<int64_t>(u + UPPER_TABLE + 8 * n);
            for (uint32_t m = 0; m < 4096; ++m) {
                const Coord lo{uo.x + int32_t((m >> 8) << 3), uo.y + int32_t(((m >> 4) & 15) << 3), uo.z + int32_t((m & 15) << 3)};
                if (!maskBit(l + LOWER_CMASK, m)) { if (maskBit(l + LOWER_VMASK, m)) bb.expand(lo, 8); continue; }
                const uint8_t* lf = l + rd<int64_t>(l + LOWER_TABLE + 8 * m);
                bool any = false;
                for (int w = 0; w < 8 && !any; ++w) any = rd<uint64_t>(lf + LEAF_VMASK + 8 * w) != 0;
                if (any) bb.expand(lo, 8);
            }
        }
    }
    return bb;
}

// ------------------------------------------------------------------------------------------------------------
// math::Ray<double> (math/Ray.h:26-295)
// ------------------------------------------------------------------------------------------------------------
struct Vec3 { double x, y, z;
    double& operator[](int i) { return (&x)[i]; }
    double operator[](int i) const { return (&x)[i]; } };
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       // math/Vec3.h:192-198
inline double length(const Vec3& a) { return std::sqrt(double(a.x * a.x + a.y * a.y + a.z * a.z)); } // math/Vec3.h:201-207
inline bool normalize(Vec3& a) {                                                                    // math/Vec3.h:363-371
    const double d = length(a);
    if (!(std::fabs(d - 0.0) > 1.0e-7)) return false;   // isApproxEqual(d, 0, eps) == !(Abs(a-b) > tol)  (math/Math.h:428-431)
    const double s = 1.0 / d;
    a.x *= s; a.y *= s; a.z *= s;
    return true;
}

struct Ray {
    Vec3 eye, dir, inv; double t0, t1;
    void setDir(const Vec3& d) { dir = d; inv = Vec3{1 / d.x, 1 / d.y, 1 / d.z}; }                   // Ray.h:67-71
    Vec3 operator()(double t) const { return Vec3{eye.x + dir.x * t, eye.y + dir.y * t, eye.z + dir.z * t}; } // Ray.h:109
    // Ray::intersects(bbox,t0,t1) + clip (Ray.h:233-267); bbox max used as given (no +1 for level sets)
    bool clip(const Coord& mn, const Coord& mx) {
        double a0 = t0, a1 = t1;
        for (int i = 0; i < 3; ++i) {
            double a = (mn[i] - eye[i]) * inv[i];
            double b = (mx[i] - eye[i]) * inv[i];
            if (a > b) std::swap(a, b);
            if (a > a0) a0 = a;
            if (b < a1) a1 = b;
            if (a0 > a1) return false;
        }
        t0 = a0; t1 = a1;
        return true;
    }
};
Ray makeRay(const vdbrt_ray& r) { Ray o; o.eye = Vec3{r.eye[0], r.eye[1], r.eye[2]}; o.setDir(Vec3{r.dir[0], r.dir[1], r.dir[2]}); o.t0 = r.t0; o.t1 = r.t1; return o; }

// ScaleMap / ScaleTranslateMap (math/Maps.h:726-771,1255-1290): one multiply per component with the stored inverse.
// General maps: nanovdb::Map::applyMap / applyInverseMap / applyJacobian / applyInverseJacobian / applyIJT (NanoVDB.h:1473-1548) without fma.
inline Vec3 mul3(const double* m, const Vec3& p) { return Vec3{p.x * m[0] + p.y * m[1] + p.z * m[2], p.x * m[3] + p.y * m[4] + p.z * m[5], p.x * m[6] + p.y * m[7] + p.z * m[8]}; }
inline Vec3 mul3T(const double* m, const Vec3& p) { return Vec3{p.x * m[0] + p.y * m[3] + p.z * m[6], p.x * m[1] + p.y * m[4] + p.z * m[7], p.x * m[2] + p.y * m[5] + p.z * m[8]}; }
inline Vec3 applyMap(const oracle_grid& g, const Vec3& p) {
    if (g.general) { const Vec3 q = mul3(g.mat, p); return Vec3{q.x + g.trans[0], q.y + g.trans[1], q.z + g.trans[2]}; }
    if (g.hasTranslation) return Vec3{p.x * g.scale[0] + g.trans[0], p.y * g.scale[1] + g.trans[1], p.z * g.scale[2] + g.trans[2]};
    return Vec3{p.x * g.scale[0], p.y * g.scale[1], p.z * g.scale[2]};
}
inline Vec3 applyInverseMap(const oracle_grid& g, const Vec3& p) {
    if (g.general) return mul3(g.imat, Vec3{p.x - g.trans[0], p.y - g.trans[1], p.z - g.trans[2]});
    if (g.hasTranslation) return Vec3{(p.x - g.trans[0]) * g.inv[0], (p.y - g.trans[1]) * g.inv[1], (p.z - g.trans[2]) * g.inv[2]};
    return Vec3{p.x * g.inv[0], p.y * g.inv[1], p.z * g.inv[2]};
}
inline Vec3 applyJacobian(const oracle_grid& g, const Vec3& d) { return g.general ? mul3(g.mat, d) : Vec3{d.x * g.scale[0], d.y * g.scale[1], d.z * g.scale[2]}; }
inline Vec3 applyInverseJacobian(const oracle_grid& g, const Vec3& d) { return g.general ? mul3(g.imat, d) : Vec3{d.x * g.inv[0], d.y * g.inv[1], d.z * g.inv[2]}; }

// Ray::worldToIndex == applyInverseMap (Ray.h:150-159): new eye, renormalised direction, times scaled by |J^-1 dir|
Ray worldToIndex(const oracle_grid& g, const Ray& w) {
    Ray r;
    r.eye = applyInverseMap(g, w.eye);
    const Vec3 d = applyInverseJacobian(g, w.dir);
    const double len = length(d);
    r.setDir(Vec3{d.x / len, d.y / len, d.z / len});
    r.t0 = len * w.t0; r.t1 = len * w.t1;
    return r;
}

// ------------------------------------------------------------------------------------------------------------
// math::DDA<Ray,Log2Dim> (math/DDA.h:34-127)
// ------------------------------------------------------------------------------------------------------------
inline int floorToInt(double x) { return int(std::floor(x)); }   // math::Floor (math/Math.h:917), Coord::floor (math/Coord.h:57-60)

template<int LOG2DIM>
struct DDA {
    static constexpr int DIM = 1 << LOG2DIM;
    double t0, t1; Coord vox, stp; double delta[3], nxt[3];
    void init(const Ray& ray, double start, double maxT) {                     // DDA.h:52-75
        t0 = start; t1 = maxT;
        const Vec3 pos = ray(t0);
        vox = Coord{floorToInt(pos.x) & ~(DIM - 1), floorToInt(pos.y) & ~(DIM - 1), floorToInt(pos.z) & ~(DIM - 1)};
        for (int a = 0; a < 3; ++a) {
            if (ray.dir[a] == 0.0) {                                           // math::isZero handles +/-0 (DDA.h:61-64)
                stp[a] = 0; nxt[a] = DBL_MAX; delta[a] = DBL_MAX;
            } else if (ray.inv[a] > 0) {
                stp[a] = DIM; nxt[a] = t0 + (vox[a] + DIM - pos[a]) * ray.inv[a]; delta[a] = stp[a] * ray.inv[a];
            } else {
                stp[a] = -DIM; nxt[a] = t0 + (vox[a] - pos[a]) * ray.inv[a]; delta[a] = stp[a] * ray.inv[a];
            }
        }
    }
    void init(const Ray& ray) { init(ray, ray.t0, ray.t1); }
    bool step() {                                                              // DDA.h:83-90, MinIndex ties -> largest index (Math.h:999-1007)
        int axis = 0;
        for (int i = 1; i < 3; ++i) if (nxt[i] <= nxt[axis]) axis = i;
        t0 = nxt[axis]; nxt[axis] += delta[axis]; vox[axis] += stp[axis];
        return t0 <= t1;
    }
    double time() const { return t0; }
    double maxTime() const { return t1; }
    double next() const { return std::min(std::min(t1, nxt[0]), std::min(nxt[1], nxt[2])); } // math::Min of four (Math.h:734-738)
};

// ------------------------------------------------------------------------------------------------------------
// math::BoxStencil<FloatGrid> (math/Stencils.h:34-221,285-428)
// ------------------------------------------------------------------------------------------------------------
struct BoxStencil {
    Coord center{INT_MAX, INT_MAX, INT_MAX};     // BaseStencil ctor: mCenter(Coord::max()) (Stencils.h:212)
    float v[8];                                  // slots: 000,001,011,010,100,101,111,110 (Stencils.h:285-293)
    uint64_t refills = 0;
    void moveTo(Access& acc, const Vec3& xyz) {  // Stencils.h:85-90 -> :47-52 -> init :414-423
        const Coord ijk{floorToInt(xyz.x), floorToInt(xyz.y), floorToInt(xyz.z)};
        if (!(ijk != center)) return;
        center = ijk; ++refills;
        v[0] = acc.getValue(ijk);
        v[1] = acc.getValue(Coord{ijk.x, ijk.y, ijk.z + 1});
        v[2] = acc.getValue(Coord{ijk.x, ijk.y + 1, ijk.z + 1});
        v[3] = acc.getValue(Coord{ijk.x, ijk.y + 1, ijk.z});
        v[4] = acc.getValue(Coord{ijk.x + 1, ijk.y, ijk.z});
        v[5] = acc.getValue(Coord{ijk.x + 1, ijk.y, ijk.z + 1});
        v[6] = acc.getValue(Coord{ijk.x + 1, ijk.y + 1, ijk.z + 1});
        v[7] = acc.getValue(Coord{ijk.x + 1, ijk.y + 1, ijk.z});
    }
    // interpolation(Vec3<float>) (Stencils.h:335-360): position converted to float FIRST, all lerps in float
    float interpolation(const Vec3& xyzd) const {
        const float xf = float(xyzd.x), yf = float(xyzd.y), zf = float(xyzd.z);
        const float u = xf - float(center.x), vv = yf - float(center.y), w = zf - float(center.z);
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
    // gradient(Vec3<float>) (Stencils.h:369-411) followed by applyIJT (= * 1/scale, Maps.h:767-771), result float
    void gradient(const oracle_grid& g, const Vec3& xyzd, float out[3]) const {
        const float xf = float(xyzd.x), yf = float(xyzd.y), zf = float(xyzd.z);
        const float u = xf - float(center.x), vv = yf - float(center.y), w = zf - float(center.z);
        float D[4] = {v[1] - v[0], v[2] - v[3], v[5] - v[4], v[6] - v[7]};
        float A = D[0] + (D[1] - D[0]) * vv;
        float B = D[2] + (D[3] - D[2]) * vv;
        const float gz = A + (B - A) * u;
        D[0] = v[0] + D[0] * w;
        D[1] = v[3] + D[1] * w;
        D[2] = v[4] + D[2] * w;
        D[3] = v[7] + D[3] * w;
        A = D[0] + (D[1] - D[0]) * vv;
        B = D[2] + (D[3] - D[2]) * vv;
        const float gx = B - A;
        A = D[1] - D[0];
        B = D[3] - D[2];
        const float gy = A + (B - A) * u;
        if (g.general) { const Vec3 t = mul3T(g.imat, Vec3{double(gx), double(gy), double(gz)}); out[0] = float(t.x); out[1] = float(t.y); out[2] = float(t.z); return; }   // applyIJT
        out[0] = float(double(gx) * g.inv[0]); out[1] = float(double(gy) * g.inv[1]); out[2] = float(double(gz) * g.inv[2]);
    }
};

// ------------------------------------------------------------------------------------------------------------
// tools::LinearSearchImpl<FloatGrid,0,double> (tools/RayIntersector.h:514-668) + math::LevelSetHDDA (math/DDA.h:144-177)
// ------------------------------------------------------------------------------------------------------------
struct Counters { uint64_t probes[3] = {0, 0, 0}, voxel = 0, refills = 0, pSamples = 0, sSamples = 0, sRays = 0, hits = 0, rays = 0; };

// oracle_set_lazy_init (test knob): evaluate tester.init's mV[0] only when the first voxel of the leaf visit passes the value gate --
// the evaluation order of the CUDA kernels (lsAdvance, vdbrt_device.cuh).  mV[0] is read nowhere else (RayIntersector.h:620-644), so the
// results must be identical and only the stencil refill count drops; tests/test_oracle_vs_reference.py checks exactly that.
static int g_lazyInit = 0;

struct Tester {
    const oracle_grid& g; Access acc; BoxStencil st; Ray ray;
    double time = 0; float V[2]; double T[2]; float iso, vmin, vmax; Coord hitIjk{0, 0, 0};
    int iterations = 0;                                         // LinearSearchImpl<GridT, Iterations, RealT>
    bool lazy = g_lazyInit != 0, pendingInit = false;
    Counters* ctr;
    Tester(const oracle_grid& grid, float isoValue, Counters* c) : g(grid), acc(grid), iso(isoValue), ctr(c) {
        vmin = isoValue - float(2 * grid.voxelSize[0]);         // RayIntersector.h:530-531 (as float)
        vmax = isoValue + float(2 * grid.voxelSize[0]);
    }
    bool setIndexRay(const Ray& r) { ray = r; return ray.clip(g.nodeBBox.mn, g.nodeBBox.mx); }        // :548-552
    bool setWorldRay(const Ray& r) { ray = worldToIndex(g, r); return ray.clip(g.nodeBBox.mn, g.nodeBBox.mx); } // :558-562
    double interpValue(double t) { const Vec3 pos = ray(t); st.moveTo(acc, pos); return st.interpolation(pos) - iso; } // :652-657
    void init(double t0) { T[0] = t0; if (lazy) pendingInit = true; else V[0] = float(interpValue(t0)); }   // :597-601
    void setRange(double a, double b) { ray.t0 = a; ray.t1 = b; }
    template<int LEVEL> bool hasNode(const Coord& c) {                                                  // :609-613
        if (ctr) ++ctr->probes[LEVEL];
        if (LEVEL == 2) return acc.upper(c) != nullptr;
        if (LEVEL == 1) return acc.lower(c) != nullptr;
        return acc.leaf(c) != nullptr;
    }
    bool operator()(const Coord& ijk, double t) {                                                       // :620-644
        if (ctr) ++ctr->voxel;
        float v;
        if (acc.probeValue(ijk, v) && v > vmin && v < vmax) {
            if (pendingInit) { V[0] = float(interpValue(T[0])); pendingInit = false; }
            T[1] = t; V[1] = float(interpValue(t));
            if (V[0] * V[1] <= 0.0f) {                                                                  // math::ZeroCrossing (Math.h:821)
                time = T[0] + (T[1] - T[0]) * V[0] / (V[0] - V[1]);                                    // interpTime :646-650
                for (int n = 0; n < iterations; ++n) {                                                  // secant refinements :630-636
                    const float W = float(interpValue(time));
                    const int m = (V[0] * W <= 0.0f) ? 1 : 0;
                    V[m] = W; T[m] = time;
                    time = T[0] + (T[1] - T[0]) * V[0] / (V[0] - V[1]);
                }
                hitIjk = ijk;
                return true;
            }
            T[0] = T[1]; V[0] = V[1];
        }
        return false;
    }
    void getWorldPosAndNml(Vec3& xyzIndex, Vec3& xyzWorld, Vec3& nml) {                                 // :575-582
        xyzIndex = ray(time);
        st.moveTo(acc, xyzIndex);
        float gf[3]; st.gradient(g, xyzIndex, gf);
        nml = Vec3{double(gf[0]), double(gf[1]), double(gf[2])};
        normalize(nml);
        xyzWorld = applyMap(g, xyzIndex);
    }
    double getWorldTime() const { return time * length(applyJacobian(g, ray.dir)); }                    // :588-591
};

template<int LEVEL> struct LevelLog2 { };
template<> struct LevelLog2<2> { static constexpr int value = 12; };
template<> struct LevelLog2<1> { static constexpr int value = 7; };
template<> struct LevelLog2<0> { static constexpr int value = 3; };

template<int LEVEL> struct LevelSetHDDA {                                                               // DDA.h:144-161
    static bool test(Tester& t) {
        DDA<LevelLog2<LEVEL>::value> dda; dda.init(t.ray);
        do {
            if (t.template hasNode<LEVEL>(dda.vox)) {
                t.setRange(dda.time(), dda.next());
                if (LevelSetHDDA<LEVEL - 1>::test(t)) return true;
            }
        } while (dda.step());
        return false;
    }
};
template<> struct LevelSetHDDA<-1> {                                                                    // DDA.h:165-177
    static bool test(Tester& t) {
        DDA<0> dda; dda.init(t.ray);
        t.init(dda.time());
        do { if (t(dda.vox, dda.next())) return true; } while (dda.step());
        return false;
    }
};

// ------------------------------------------------------------------------------------------------------------
// tools::VolumeRayIntersector (tools/RayIntersector.h:277-485) + math::VolumeHDDA::hits (math/DDA.h:187-338)
// ------------------------------------------------------------------------------------------------------------
struct TimeSpan { double t0, t1; bool valid() const { return (t1 - t0) > 1e-9; } };                    // Ray.h:48 (eps = Delta<double>)

struct VolumeIntersector {
    const oracle_grid& g; Access acc; Ray ray; double tmax = 0; Coord bmin, bmax; Counters* ctr;
    VolumeIntersector(const oracle_grid& grid, Counters* c) : g(grid), acc(grid), ctr(c) {
        bmin = grid.nodeBBox.mn;
        bmax = Coord{grid.nodeBBox.mx.x + 1, grid.nodeBBox.mx.y + 1, grid.nodeBBox.mx.z + 1};          // mBBox.max().offset(1) (:318)
    }
    bool setIndexRay(const Ray& r) { ray = r; const bool hit = ray.clip(bmin, bmax); if (hit) tmax = ray.t1; return hit; } // :368-374
    bool setWorldRay(const Ray& r) { return setIndexRay(worldToIndex(g, r)); }                           // :391-394
    Vec3 getWorldPos(double t) const { return applyMap(g, ray(t)); }                                     // :440

    template<int LEVEL> void hitsLevel(std::vector<TimeSpan>& times, TimeSpan& t) {                     // DDA.h:247-264 / :321-336
        DDA<LevelLog2<LEVEL>::value> dda; dda.init(ray);
        do {
            if (ctr) ++ctr->probes[LEVEL];
            bool child = false;
            if (LEVEL == 2) child = acc.upper(dda.vox) != nullptr;
            else if (LEVEL == 1) child = acc.lower(dda.vox) != nullptr;
            if (LEVEL > 0 && child) {
                ray.t0 = dda.time(); ray.t1 = dda.next();
                hitsLevel<(LEVEL > 0 ? LEVEL - 1 : 0)>(times, t);
            } else if (LEVEL == 0 ? (acc.leaf(dda.vox) != nullptr || acc.isValueOn(dda.vox)) : acc.isValueOn(dda.vox)) {
                if (t.t0 < 0) t.t0 = dda.time();
            } else if (t.t0 >= 0) {
                t.t1 = dda.time();
                if (t.valid()) times.push_back(t);
                t = TimeSpan{-1, -1};
            }
        } while (dda.step());
        if (t.t0 >= 0) t.t1 = dda.maxTime();
    }
    void hits(std::vector<TimeSpan>& times) {                                                            // DDA.h:210-217
        TimeSpan t{-1, -1};
        times.clear();
        hitsLevel<2>(times, t);
        if (t.valid()) times.push_back(t);
    }
};

// tools::BoxSampler::sample via GridSampler::wsSample (tools/Interpolation.h:420-425,658-688,712-762):
// float corner values, (b-a) in float times a DOUBLE weight, rounded to float, added to a in float.
inline float lerpBox(float a, float b, double w) { const double temp = (b - a) * w; return a + float(temp); }
float boxSampleWorld(const oracle_grid& g, Access& acc, const Vec3& ws)
{
    const Vec3 p = applyInverseMap(g, ws);                                       // Transform::worldToIndex
    const int i = floorToInt(p.x), j = floorToInt(p.y), k = floorToInt(p.z);     // local_util::floorVec3 (:586-589)
    const double u = p.x - i, v = p.y - j, w = p.z - k;
    float d000, d001, d011, d010, d100, d101, d111, d110;                        // probeValues order (:663-689)
    acc.probeValue(Coord{i, j, k}, d000);
    acc.probeValue(Coord{i, j, k + 1}, d001);
    acc.probeValue(Coord{i, j + 1, k + 1}, d011);
    acc.probeValue(Coord{i, j + 1, k}, d010);
    acc.probeValue(Coord{i + 1, j, k}, d100);
    acc.probeValue(Coord{i + 1, j, k + 1}, d101);
    acc.probeValue(Coord{i + 1, j + 1, k + 1}, d111);
    acc.probeValue(Coord{i + 1, j + 1, k}, d110);
    return lerpBox(lerpBox(lerpBox(d000, d001, w), lerpBox(d010, d011, w), v),
                   lerpBox(lerpBox(d100, d101, w), lerpBox(d110, d111, w), v), u);
}

// ------------------------------------------------------------------------------------------------------------
// cameras (tools/RayTracer.h:351-513) on the flattened POD
// ------------------------------------------------------------------------------------------------------------
inline Vec3 rasterToScreen(const vdbrt_camera& c, double i, double j, double z) {                       // :391-395
    return Vec3{(2 * i / double(c.width) - 1) * c.scale_w, (1 - 2 * j / double(c.height)) * c.scale_h, z};
}
inline Vec3 transform3x3(const double* m, const Vec3& v) {                                              // math/Mat4.h:1070-1076
    return Vec3{v.x * m[0] + v.y * m[4] + v.z * m[8], v.x * m[1] + v.y * m[5] + v.z * m[9], v.x * m[2] + v.y * m[6] + v.z * m[10]};
}
inline Vec3 transformPoint(const double* m, const Vec3& v) {                                            // Vec3 * Mat4 (math/Mat4.h:1180-1188)
    return Vec3{v.x * m[0] + v.y * m[4] + v.z * m[8] + m[12], v.x * m[1] + v.y * m[5] + v.z * m[9] + m[13], v.x * m[2] + v.y * m[6] + v.z * m[10] + m[14]};
}
Ray cameraRay(const vdbrt_camera& c, uint32_t i, uint32_t j, double io, double jo)
{
    Ray ray;
    ray.eye = Vec3{c.eye[0], c.eye[1], c.eye[2]};
    ray.setDir(Vec3{c.dir[0], c.dir[1], c.dir[2]});
    ray.t0 = c.t0; ray.t1 = c.t1;
    if (c.kind == VDBRT_CAMERA_PERSPECTIVE) {                                                           // :452-462
        Vec3 dir = rasterToScreen(c, double(i) + io, double(j) + jo, -1.0);
        dir = transform3x3(c.m, dir);
        normalize(dir);
        const double s = 1.0 / dot(dir, ray.dir);
        ray.t0 *= s; ray.t1 *= s;                                                                        // scaleTimes
        ray.setDir(dir);
    } else {                                                                                            // :505-512
        const Vec3 eye = rasterToScreen(c, double(i) + io, double(j) + jo, 0.0);
        ray.eye = transformPoint(c.m, eye);
    }
    return ray;
}

// ------------------------------------------------------------------------------------------------------------
// shaders (tools/RayTracer.h:565-581,614-630,671-690,728-753) and Film::RGBA arithmetic (:231-262)
// ------------------------------------------------------------------------------------------------------------
struct RGBA { float r, g, b, a; };
} // namespace

// A NanoGrid<Vec3f> for the GridT = Vec3SGrid forms of the shaders (tools/RayTracer.h:542-725).  Vec3f node layout
// (nanovdb/NanoVDB.h; sizeof/offsetof probe): RootData 96 B + 32 B tiles with the value at +20; internal nodes keep the float
// build's mask offsets and have 16-byte table entries; leaves: values (12 B each) at +128.
struct oracle_color {
    const uint8_t* base = nullptr; const uint8_t* root = nullptr; const uint8_t* tiles = nullptr;
    uint32_t tableSize = 0; float background[3] = {0.f, 0.f, 0.f};
    double inv[3], trans[3]; bool hasTranslation = false;
};

namespace {

// tools::PointSampler::sample(acc, xform.worldToIndex(xyz), v) (tools/Interpolation.h:600-617): probeValue at the voxel
// ::round()ed from the index position; tile value or background where there is no voxel
void colorAt(const oracle_color& c, const Vec3& w, float v[3])
{
    double p[3];
    for (int a = 0; a < 3; ++a) p[a] = c.hasTranslation ? (w[a] - c.trans[a]) * c.inv[a] : w[a] * c.inv[a];
    const int x = int(::round(p[0])), y = int(::round(p[1])), z = int(::round(p[2]));
    const uint64_t key = uint64_t(uint32_t(z) >> 12) | (uint64_t(uint32_t(y) >> 12) << 21) | (uint64_t(uint32_t(x) >> 12) << 42);
    const uint8_t* src = nullptr;
    for (uint32_t i = 0; i < c.tableSize && !src; ++i) {
        const uint8_t* t = c.tiles + 32 * i;
        if (rd<uint64_t>(t) != key) continue;
        const int64_t child = rd<int64_t>(t + 8);
        if (!child) { src = t + 20; break; }
        const uint8_t* u = c.root + child;
        uint32_t n = (((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7);
        if (!((rd<uint64_t>(u + 32 + 4096 + 8 * (n >> 6)) >> (n & 63)) & 1)) { src = u + 8256 + 16 * n; break; }
        const uint8_t* l = u + rd<int64_t>(u + 8256 + 16 * n);
        n = (((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3);
        if (!((rd<uint64_t>(l + 32 + 512 + 8 * (n >> 6)) >> (n & 63)) & 1)) { src = l + 1088 + 16 * n; break; }
        const uint8_t* f = l + rd<int64_t>(l + 1088 + 16 * n);
        src = f + 128 + 12 * (((x & 7) << 6) | ((y & 7) << 3) | (z & 7));
    }
    for (int a = 0; a < 3; ++a) v[a] = src ? rd<float>(src + 4 * a) : c.background[a];
}

RGBA shade(const vdbrt_shader& s, const oracle_color* col, const Vec3& xyz, const Vec3& nml, const Vec3& dir)
{
    if (col) {
        float v[3];
        colorAt(*col, xyz, v);
        switch (s.kind) {
        case VDBRT_SHADER_MATTE: return RGBA{v[0], v[1], v[2], 1.0f};                                              // :549-554
        case VDBRT_SHADER_NORMAL: return RGBA{float(v[0] * (nml.x + 1.0)), float(v[1] * (nml.y + 1.0)), float(v[2] * (nml.z + 1.0)), 1.0f};   // :598-603
        case VDBRT_SHADER_POSITION: {                                                                                // :655-661
            const double rx = (xyz.x - s.bbox_min[0]) * s.inv_dim[0], ry = (xyz.y - s.bbox_min[1]) * s.inv_dim[1], rz = (xyz.z - s.bbox_min[2]) * s.inv_dim[2];
            return RGBA{v[0] * float(rx), v[1] * float(ry), v[2] * float(rz), 1.0f};
        }
        default: { const float f = float(std::fabs(dot(nml, dir))); return RGBA{v[0] * f, v[1] * f, v[2] * f, 1.0f}; }   // :709-717
        }
    }
    switch (s.kind) {
    case VDBRT_SHADER_MATTE: return RGBA{s.rgba[0], s.rgba[1], s.rgba[2], s.rgba[3]};
    case VDBRT_SHADER_NORMAL: {   // mRGBA = c*0.5f (alpha -> 1); mRGBA * RGBA(n+1.0) with the doubles cast to float
        const float r = s.rgba[0] * 0.5f, g = s.rgba[1] * 0.5f, b = s.rgba[2] * 0.5f;
        return RGBA{r * float(nml.x + 1.0), g * float(nml.y + 1.0), b * float(nml.z + 1.0), 1.0f};
    }
    case VDBRT_SHADER_POSITION: {
        const double rx = (xyz.x - s.bbox_min[0]) * s.inv_dim[0], ry = (xyz.y - s.bbox_min[1]) * s.inv_dim[1], rz = (xyz.z - s.bbox_min[2]) * s.inv_dim[2];
        return RGBA{s.rgba[0] * float(rx), s.rgba[1] * float(ry), s.rgba[2] * float(rz), 1.0f};
    }
    default: {                    // Diffuse: mRGBA * float(|n . rayDir|)
        const float f = float(std::fabs(dot(nml, dir)));
        return RGBA{s.rgba[0] * f, s.rgba[1] * f, s.rgba[2] * f, 1.0f};
    }
    }
}

bool ownsPixel(const vdbrt_partition& p, uint32_t i, uint32_t j, uint32_t width)
{
    if (p.count <= 1) return true;
    const uint32_t tw = p.tile_w ? p.tile_w : 64, th = p.tile_h ? p.tile_h : 64;
    const uint32_t tilesX = (width + tw - 1) / tw;
    const uint32_t tile = (j / th) * tilesX + (i / tw);
    return tile % p.count == p.rank;
}

template<typename F> void parallelRows(uint32_t height, int threads, F f)
{
    if (threads <= 1) { f(0u, height, 0); return; }
    std::atomic<uint32_t> next{0};
    const uint32_t chunk = 4;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back([&, t]() {
        for (;;) { const uint32_t b = next.fetch_add(chunk); if (b >= height) break; f(b, std::min(height, b + chunk), t); } });
    for (auto& th : pool) th.join();
}

void addCounters(vdbrt_counters* out, const std::vector<Counters>& cs)
{
    if (!out) return;
    std::memset(out, 0, sizeof(*out));
    for (const Counters& c : cs) {
        out->rays += c.rays; out->root_probes += c.probes[2]; out->upper_probes += c.probes[1]; out->lower_probes += c.probes[0];
        out->voxel_probes += c.voxel; out->stencil_refills += c.refills; out->primary_samples += c.pSamples;
        out->shadow_samples += c.sSamples; out->shadow_rays += c.sRays; out->hits += c.hits;
    }
}

template<int L> int ddaTrace(const Ray& ray, int maxSteps, double* out)
{
    DDA<L> dda; dda.init(ray);
    int n = 0;
    do {
        if (n >= maxSteps) break;
        double* o = out + 5 * n++;
        o[0] = dda.time(); o[1] = dda.next(); o[2] = dda.vox.x; o[3] = dda.vox.y; o[4] = dda.vox.z;
    } while (dda.step());
    return n;
}
} // namespace

extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

int oracle_grid_open(const void* buf, uint64_t bytes, oracle_grid** out)
{
    if (!buf || !out) return fail(VDBRT_ERR_INVALID_ARG, "null argument");
    const uint8_t* b = static_cast<const uint8_t*>(buf);
    if (bytes < GRID_SIZE + TREE_SIZE) return fail(VDBRT_ERR_BAD_GRID, "buffer smaller than GridData+TreeData");
    const uint64_t magic = rd<uint64_t>(b);
    if (magic != MAGIC_NUMB && magic != MAGIC_GRID) return fail(VDBRT_ERR_BAD_GRID, "bad magic number");
    if ((rd<uint32_t>(b + OFF_VERSION) >> 21) != 32) return fail(VDBRT_ERR_BAD_GRID, "incompatible NanoVDB major version");
    if (rd<uint64_t>(b + OFF_GRIDSIZE) > bytes) return fail(VDBRT_ERR_BAD_GRID, "grid size exceeds buffer");
    const uint32_t gridType = rd<uint32_t>(b + OFF_TYPE);
    if (gridType != 1 && !(gridType >= 13 && gridType <= 16)) return fail(VDBRT_ERR_NOT_FLOAT, "grid type is not Float");
    auto* g = new oracle_grid;
    g->gridType = gridType;
    g->base = b; g->bytes = bytes; g->tree = b + GRID_SIZE;
    g->firstLeaf = g->tree + rd<int64_t>(g->tree + 0); g->firstLower = g->tree + rd<int64_t>(g->tree + 8);
    g->firstUpper = g->tree + rd<int64_t>(g->tree + 16); g->root = g->tree + rd<int64_t>(g->tree + 24);
    g->leafCount = rd<uint32_t>(g->tree + 32); g->lowerCount = rd<uint32_t>(g->tree + 36); g->upperCount = rd<uint32_t>(g->tree + 40);
    g->activeVoxels = rd<uint64_t>(g->tree + 56);
    g->tableSize = rd<uint32_t>(g->root + ROOT_TABLESIZE);
    g->background = rd<float>(g->root + ROOT_BACKGROUND);
    g->tiles = g->root + ROOT_TILES;
    g->gridClass = rd<uint32_t>(b + OFF_CLASS);
    for (int i = 0; i < 6; ++i) g->indexBBox[i] = rd<int32_t>(g->root + 4 * i);
    double m[9]; for (int i = 0; i < 9; ++i) m[i] = rd<double>(b + OFF_MATD + 8 * i);
    g->general = m[1] != 0 || m[2] != 0 || m[3] != 0 || m[5] != 0 || m[6] != 0 || m[7] != 0;
    for (int i = 0; i < 9; ++i) { g->mat[i] = m[i]; g->imat[i] = rd<double>(b + OFF_MATD + 72 + 8 * i); }
    for (int a = 0; a < 3; ++a) {
        g->scale[a] = m[4 * a]; g->inv[a] = 1.0 / g->scale[a];                  // ScaleMap ctor: mScaleValuesInverse = 1.0/scale (Maps.h:674)
        g->trans[a] = rd<double>(b + OFF_VECD + 8 * a);
        // mVoxelSize: |scale|, or for an affine map the length of the image of the unit vector (Maps.h:630-633,667)
        g->voxelSize[a] = g->general ? std::sqrt(m[a] * m[a] + m[3 + a] * m[3 + a] + m[6 + a] * m[6 + a]) : std::fabs(g->scale[a]);
    }
    g->hasTranslation = g->trans[0] != 0 || g->trans[1] != 0 || g->trans[2] != 0;
    g->nodeBBox = evalNodeBBox(*g);
    *out = g;
    return VDBRT_OK;
}

void oracle_grid_close(oracle_grid* g) { delete g; }

int oracle_grid_get_info(const oracle_grid* g, vdbrt_grid_info* info)
{
    std::memset(info, 0, sizeof(*info));
    info->bytes = g->bytes; info->active_voxels = g->activeVoxels; info->leaf_count = g->leafCount; info->lower_count = g->lowerCount;
    info->upper_count = g->upperCount; info->root_tiles = g->tableSize;
    for (int i = 0; i < 6; ++i) info->index_bbox[i] = g->indexBBox[i];
    for (int a = 0; a < 3; ++a) { info->node_bbox[a] = g->nodeBBox.mn[a]; info->node_bbox[3 + a] = g->nodeBBox.mx[a];
        info->voxel_size[a] = g->voxelSize[a]; info->translation[a] = g->trans[a]; }
    info->background = g->background; info->grid_class = g->gridClass; info->source_type = g->gridType;
    return VDBRT_OK;
}

int oracle_grid_probe(const oracle_grid* g, const int32_t* ijk, uint64_t n, float* values, uint8_t* active)
{
    Access acc(*g);
    for (uint64_t i = 0; i < n; ++i) active[i] = acc.probeValue(Coord{ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]}, values[i]);
    return VDBRT_OK;
}

// Validation performed by the reference at construction (RayIntersector.h:100-112,527-541)
static int checkLevelSet(const oracle_grid* g, float iso)
{
    // member mTester (LinearSearchImpl) is constructed first: empty / iso checks precede the intersector's own checks
    if (g->tableSize == 0) return fail(VDBRT_ERR_EMPTY_GRID, "LinearSearchImpl does not supports empty grids");
    if (iso <= -g->background || iso >= g->background) return fail(VDBRT_ERR_ISO_RANGE, "The iso-value must be inside the narrow-band!");
    const double s0 = g->voxelSize[0];
    if (std::fabs(s0 - g->voxelSize[1]) > 5e-7 || std::fabs(s0 - g->voxelSize[2]) > 5e-7)
        return fail(VDBRT_ERR_NONUNIFORM, "LevelSetRayIntersector only supports uniform voxels!");
    if (g->gridClass != VDBRT_GRID_CLASS_LEVEL_SET) return fail(VDBRT_ERR_NOT_LEVELSET, "LevelSetRayIntersector only supports level sets!");
    return VDBRT_OK;
}
static int checkVolume(const oracle_grid* g)
{
    const double s0 = g->voxelSize[0];
    if (std::fabs(s0 - g->voxelSize[1]) > 5e-7 || std::fabs(s0 - g->voxelSize[2]) > 5e-7)
        return fail(VDBRT_ERR_NONUNIFORM, "VolumeRayIntersector only supports uniform voxels!");
    if (g->tableSize == 0) return fail(VDBRT_ERR_EMPTY_GRID, "LinearSearchImpl does not supports empty grids");
    return VDBRT_OK;
}

// LevelSetRayTracer::operator() (tools/RayTracer.h:899-918); jitter index n(i,j) = 2*(spp-1)*(j*W+i) (threaded=false order)
int oracle_color_open(const void* buf, uint64_t bytes, oracle_color** out)
{
    if (!buf || !out) return fail(VDBRT_ERR_INVALID_ARG, "null argument");
    const uint8_t* b = static_cast<const uint8_t*>(buf);
    if (bytes < GRID_SIZE + TREE_SIZE) return fail(VDBRT_ERR_BAD_GRID, "buffer smaller than GridData+TreeData");
    const uint64_t magic = rd<uint64_t>(b);
    if (magic != MAGIC_NUMB && magic != MAGIC_GRID) return fail(VDBRT_ERR_BAD_GRID, "bad magic number");
    if (rd<uint32_t>(b + OFF_TYPE) != 6) return fail(VDBRT_ERR_NOT_FLOAT, "grid type is not Vec3f");
    auto* c = new oracle_color;
    c->base = b; c->root = b + GRID_SIZE + rd<int64_t>(b + GRID_SIZE + 24); c->tiles = c->root + 96;
    c->tableSize = rd<uint32_t>(c->root + ROOT_TABLESIZE);
    for (int a = 0; a < 3; ++a) {
        c->background[a] = rd<float>(c->root + ROOT_BACKGROUND + 4 * a);
        c->inv[a] = 1.0 / rd<double>(b + OFF_MATD + 8 * 4 * a);
        c->trans[a] = rd<double>(b + OFF_VECD + 8 * a);
    }
    c->hasTranslation = c->trans[0] != 0 || c->trans[1] != 0 || c->trans[2] != 0;
    *out = c;
    return VDBRT_OK;
}
void oracle_color_close(oracle_color* c) { delete c; }

int oracle_render_levelset(const oracle_grid* g, const vdbrt_camera* cam, const vdbrt_shader* shader, const vdbrt_ls_opts* opts,
                           vdbrt_film* film, vdbrt_aux* aux, vdbrt_counters* ctr, int threads)
{
    return oracle_render_levelset_color(g, nullptr, cam, shader, opts, film, aux, ctr, threads);
}

// the same with a colour grid feeding the shader (shader->color_grid is the PRODUCT's handle and is ignored here)
int oracle_render_levelset_color(const oracle_grid* g, const oracle_color* color, const vdbrt_camera* cam, const vdbrt_shader* shader,
                                 const vdbrt_ls_opts* opts, vdbrt_film* film, vdbrt_aux* aux, vdbrt_counters* ctr, int threads)
{
    if (!g || !cam || !shader || !opts || !film || !film->pixels) return fail(VDBRT_ERR_INVALID_ARG, "null argument");
    if (opts->spp == 0) return fail(VDBRT_ERR_SPP_ZERO, "pixelSamples must be larger than zero!");
    if (int e = checkLevelSet(g, opts->iso)) return e;
    const uint32_t W = film->width, H = film->height;
    const uint32_t sub = opts->spp - 1;
    const float frac = 1.0f / (1.0f + float(sub));
    std::vector<Counters> counters(std::max(1, threads));
    parallelRows(H, threads, [&](uint32_t j0, uint32_t j1, int tid) {
        Counters& c = counters[tid];
        Tester tester(*g, opts->iso, ctr ? &c : nullptr);
        tester.iterations = int(opts->iterations);
        for (uint32_t j = j0; j < j1; ++j) for (uint32_t i = 0; i < W; ++i) {
            if (!ownsPixel(opts->part, i, j, W)) continue;
            const size_t p = size_t(j) * W + i;
            float* px = film->pixels + 4 * p;
            const RGBA bg{px[0], px[1], px[2], px[3]};
            RGBA col;
            uint64_t n = uint64_t(2) * sub * p;
            for (uint32_t k = 0; k <= sub; ++k) {
                const Ray ray = k == 0 ? cameraRay(*cam, i, j, 0.5, 0.5)
                                       : cameraRay(*cam, i, j, opts->jitter[n & 15], opts->jitter[(n + 1) & 15]);
                if (k > 0) n += 2;
                ++c.rays;
                Vec3 xi, xw, nml;
                bool hit = tester.setWorldRay(ray) && LevelSetHDDA<2>::test(tester);
                RGBA s = bg;
                if (hit) { tester.getWorldPosAndNml(xi, xw, nml); s = shade(*shader, color, xw, nml, ray.dir); ++c.hits; }
                if (k == 0) {
                    col = s;
                    if (aux) {
                        if (aux->hit) aux->hit[p] = hit;
                        if (hit) {
                            if (aux->ijk) { aux->ijk[3 * p] = tester.hitIjk.x; aux->ijk[3 * p + 1] = tester.hitIjk.y; aux->ijk[3 * p + 2] = tester.hitIjk.z; }
                            if (aux->t_index) aux->t_index[p] = tester.time;
                            if (aux->t_world) aux->t_world[p] = tester.getWorldTime();
                            for (int a = 0; a < 3; ++a) { if (aux->xyz) aux->xyz[3 * p + a] = xw[a]; if (aux->nml) aux->nml[3 * p + a] = nml[a]; }
                        }
                    }
                } else { col.r += s.r; col.g += s.g; col.b += s.b; col.a += s.a; }                       // RGBA::operator+= (:250)
            }
            px[0] = col.r * frac; px[1] = col.g * frac; px[2] = col.b * frac; px[3] = 1.0f;             // bg = c*frac, alpha rebuilt as 1 (:247)
        }
        c.refills += tester.st.refills;
    });
    addCounters(ctr, counters);
    return VDBRT_OK;
}

int oracle_intersect_levelset(const oracle_grid* g, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso, vdbrt_hit* hits)
{
    return oracle_intersect_levelset_iter(g, rays, n, space, iso, 0u, hits);
}

void oracle_set_lazy_init(int on) { g_lazyInit = on; }

int oracle_intersect_levelset_iter(const oracle_grid* g, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso, uint32_t iterations, vdbrt_hit* hits)
{
    if (int e = checkLevelSet(g, iso)) return e;
    Tester tester(*g, iso, nullptr);
    tester.iterations = int(iterations);
    for (uint64_t k = 0; k < n; ++k) {
        const Ray ray = makeRay(rays[k]);
        vdbrt_hit& o = hits[k];
        std::memset(&o, 0, sizeof(o));
        const bool ok = space == VDBRT_SPACE_WORLD ? tester.setWorldRay(ray) : tester.setIndexRay(ray);
        if (!ok || !LevelSetHDDA<2>::test(tester)) continue;
        Vec3 xi, xw, nml;
        tester.getWorldPosAndNml(xi, xw, nml);
        o.hit = 1; o.ijk[0] = tester.hitIjk.x; o.ijk[1] = tester.hitIjk.y; o.ijk[2] = tester.hitIjk.z;
        o.t_index = tester.time; o.t_world = tester.getWorldTime();
        for (int a = 0; a < 3; ++a) { o.xyz_index[a] = xi[a]; o.xyz_world[a] = xw[a]; o.nml[a] = nml[a]; }
    }
    return VDBRT_OK;
}

int oracle_volume_spans(const oracle_grid* g, const vdbrt_ray* rays, uint64_t n, uint32_t space, uint32_t maxSpans, double* spans, int32_t* counts)
{
    if (int e = checkVolume(g)) return e;
    VolumeIntersector inter(*g, nullptr);
    std::vector<TimeSpan> list;
    for (uint64_t k = 0; k < n; ++k) {
        const Ray ray = makeRay(rays[k]);
        const bool ok = space == VDBRT_SPACE_WORLD ? inter.setWorldRay(ray) : inter.setIndexRay(ray);
        if (!ok) { counts[k] = -1; continue; }
        inter.hits(list);
        counts[k] = int32_t(list.size());
        for (size_t s = 0; s < list.size() && s < maxSpans; ++s) { spans[(k * maxSpans + s) * 2] = list[s].t0; spans[(k * maxSpans + s) * 2 + 1] = list[s].t1; }
    }
    return VDBRT_OK;
}

// VolumeRender::operator() (tools/RayTracer.h:991-1070)
int oracle_render_volume(const oracle_grid* g, const vdbrt_camera* cam, const vdbrt_vol_opts* o, vdbrt_film* film, vdbrt_counters* ctr, int threads)
{
    if (!g || !cam || !o || !film || !film->pixels) return fail(VDBRT_ERR_INVALID_ARG, "null argument");
    if (int e = checkVolume(g)) return e;
    const uint32_t W = film->width, H = film->height;
    Vec3 extinction, albedo;
    for (int a = 0; a < 3; ++a) {
        extinction[a] = -o->scattering[a] - o->absorption[a];                                           // :996
        albedo[a] = o->light_color[a] * o->scattering[a] / (o->scattering[a] + o->absorption[a]);       // :997
    }
    const double sGain = o->light_gain, pStep = o->primary_step, sStep = o->shadow_step, cutoff = o->cutoff;
    std::vector<Counters> counters(std::max(1, threads));
    parallelRows(H, threads, [&](uint32_t j0, uint32_t j1, int tid) {
        Counters& c = counters[tid];
        Counters* cp = ctr ? &c : nullptr;
        VolumeIntersector primary(*g, cp), shadow(*g, cp);
        Access sampler(*g);
        std::vector<TimeSpan> pTS, sTS;
        Ray sRay; sRay.eye = Vec3{0, 0, 0}; sRay.setDir(Vec3{o->light_dir[0], o->light_dir[1], o->light_dir[2]});
        sRay.t0 = 1e-9; sRay.t1 = DBL_MAX;                                                              // Ray ctor defaults (Ray.h:57-63)
        // EXTENSION (include/vdbrt.h, vdbrt_vol_opts::spp): sample 0 through the pixel centre, the others through the jittered
        // offsets of LevelSetRayTracer::operator() (tools/RayTracer.h:903-913); pixel = (sum of the samples' RGBA) * float(1/spp)
        const uint32_t sub = o->spp > 1 ? o->spp - 1 : 0u;
        const float frac = 1.0f / (1.0f + float(sub));
        for (uint32_t j = j0; j < j1; ++j) for (uint32_t i = 0; i < W; ++i) {
            if (!ownsPixel(o->part, i, j, W)) continue;
            const size_t pixel = size_t(j) * W + i;
            float* px = film->pixels + 4 * pixel;
            RGBA acc{0.f, 0.f, 0.f, 0.f};
            uint64_t n = uint64_t(2) * sub * pixel;
            for (uint32_t smp = 0; smp <= sub; ++smp) {
            RGBA out{0.f, 0.f, 0.f, 0.f};                                                               // :1020
            const Ray pRay = smp == 0 ? cameraRay(*cam, i, j, 0.5, 0.5) : cameraRay(*cam, i, j, o->jitter[n & 15], o->jitter[(n + 1) & 15]);
            if (smp > 0) n += 2;
            ++c.rays;
            if (primary.setWorldRay(pRay)) {
            Vec3 pTrans{1.0, 1.0, 1.0}, pLumi{0.0, 0.0, 0.0};
            primary.hits(pTS);
            bool done = false;
            for (size_t k = 0; k < pTS.size() && !done; ++k) {
                double pT = pStep * std::ceil(pTS[k].t0 / pStep); const double pT1 = pTS[k].t1;
                for (; pT <= pT1; pT += pStep) {
                    const Vec3 pPos = primary.getWorldPos(pT);
                    const double density = boxSampleWorld(*g, sampler, pPos);
                    ++c.pSamples;
                    if (density < cutoff) continue;
                    Vec3 dT;
                    for (int a = 0; a < 3; ++a) dT[a] = std::exp(extinction[a] * density * pStep);
                    Vec3 sTrans{1.0, 1.0, 1.0};
                    sRay.eye = pPos;
                    ++c.sRays;
                    if (!shadow.setWorldRay(sRay)) continue;
                    shadow.hits(sTS);
                    bool lit = false;
                    for (size_t l = 0; l < sTS.size() && !lit; ++l) {
                        double sT = sStep * std::ceil(sTS[l].t0 / sStep); const double sT1 = sTS[l].t1;
                        for (; sT <= sT1; sT += sStep) {
                            const double d = boxSampleWorld(*g, sampler, shadow.getWorldPos(sT));
                            ++c.sSamples;
                            if (d < cutoff) continue;
                            for (int a = 0; a < 3; ++a) sTrans[a] *= std::exp(extinction[a] * d * sStep / (1.0 + sT * sGain));
                            if (sTrans.x * sTrans.x + sTrans.y * sTrans.y + sTrans.z * sTrans.z < cutoff) { lit = true; break; } // goto Luminance
                        }
                    }
                    for (int a = 0; a < 3; ++a) { pLumi[a] += albedo[a] * sTrans[a] * pTrans[a] * (1.0 - dT[a]); pTrans[a] *= dT[a]; }
                    if (pTrans.x * pTrans.x + pTrans.y * pTrans.y + pTrans.z * pTrans.z < cutoff) { done = true; break; }     // goto Pixel
                }
            }
            out.r = float(pLumi.x); out.g = float(pLumi.y); out.b = float(pLumi.z);
            out.a = float(1.0f - (pTrans.x + pTrans.y + pTrans.z) / 3.0f);
            if (out.a > 0.f) ++c.hits;
            }
            if (smp == 0) acc = out;
            else { acc.r += out.r; acc.g += out.g; acc.b += out.b; acc.a += out.a; }
            }
            px[0] = acc.r * frac; px[1] = acc.g * frac; px[2] = acc.b * frac; px[3] = acc.a * frac;
        }
    });
    addCounters(ctr, counters);
    return VDBRT_OK;
}

// Film::RGBA::over (tools/RayTracer.h:252-259) per pixel: top = top.over(bottom)
int oracle_film_over(float* top, const float* bottom, uint64_t pixels)
{
    for (uint64_t i = 0; i < pixels; ++i) {
        float* t = top + 4 * i; const float* b = bottom + 4 * i;
        const float s = b[3] * (1.0f - t[3]);
        t[0] = t[3] * t[0] + s * b[0]; t[1] = t[3] * t[1] + s * b[1]; t[2] = t[3] * t[2] + s * b[2];
        t[3] = t[3] + s;
    }
    return VDBRT_OK;
}

int oracle_camera_rays(const vdbrt_camera* cam, const uint32_t* ij, const double* offsets, uint64_t n, vdbrt_ray* rays)
{
    for (uint64_t k = 0; k < n; ++k) {
        const Ray r = cameraRay(*cam, ij[2 * k], ij[2 * k + 1], offsets ? offsets[2 * k] : 0.5, offsets ? offsets[2 * k + 1] : 0.5);
        for (int a = 0; a < 3; ++a) { rays[k].eye[a] = r.eye[a]; rays[k].dir[a] = r.dir[a]; }
        rays[k].t0 = r.t0; rays[k].t1 = r.t1;
    }
    return VDBRT_OK;
}

int oracle_dda_trace(const vdbrt_ray* ray, int log2dim, int maxSteps, double* out)
{
    const Ray r = makeRay(*ray);
    switch (log2dim) {
    case 0: return ddaTrace<0>(r, maxSteps, out);
    case 3: return ddaTrace<3>(r, maxSteps, out);
    case 7: return ddaTrace<7>(r, maxSteps, out);
    case 12: return ddaTrace<12>(r, maxSteps, out);
    default: return -1;
    }
}
int oracle_ray_clip(const vdbrt_ray* ray, const int32_t bbox[6], double* t0, double* t1)
{
    Ray r = makeRay(*ray);
    const bool hit = r.clip(Coord{bbox[0], bbox[1], bbox[2]}, Coord{bbox[3], bbox[4], bbox[5]});
    if (hit) { *t0 = r.t0; *t1 = r.t1; }
    return hit ? 1 : 0;
}

} // extern "C"
