// vdbrt_camera.cc -- host helpers of the C ABI that flatten the reference's camera classes into vdbrt_camera.
// Pure set-up code (runs once per frame on the host, exactly where the reference runs it); the per-pixel ray
// generation itself happens on the device (vdbrt_device.cuh: cameraRay).  Built with -ffp-contract=off so the
// matrices are bit-identical to the reference's.  Paths below are relative to openvdb/openvdb/.
#include "../../include/vdbrt.h"

#include <cmath>
#include <cstring>
#include <random>
#include <string>

namespace vdbrt { int setError(int code, const std::string& msg); }

namespace {

struct Mat4 { double m[16]; };

Mat4 identity() { Mat4 r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0 : 0.0; return r; }

// Mat4::postRotate (math/Mat4.h:880-955): clockwise, columns of the row-vector matrix are mixed
void postRotate(Mat4& M, int axis, double angle)
{
    double* mm = M.m;
    const double c = std::cos(angle), s = -std::sin(angle);
    if (axis == 0) {
        const double a2 = c * mm[2] - s * mm[1], a6 = c * mm[6] - s * mm[5], a10 = c * mm[10] - s * mm[9], a14 = c * mm[14] - s * mm[13];
        mm[1] = c * mm[1] + s * mm[2]; mm[5] = c * mm[5] + s * mm[6]; mm[9] = c * mm[9] + s * mm[10]; mm[13] = c * mm[13] + s * mm[14];
        mm[2] = a2; mm[6] = a6; mm[10] = a10; mm[14] = a14;
    } else if (axis == 1) {
        const double a2 = c * mm[2] + s * mm[0], a6 = c * mm[6] + s * mm[4], a10 = c * mm[10] + s * mm[8], a14 = c * mm[14] + s * mm[12];
        mm[0] = c * mm[0] - s * mm[2]; mm[4] = c * mm[4] - s * mm[6]; mm[8] = c * mm[8] - s * mm[10]; mm[12] = c * mm[12] - s * mm[14];
        mm[2] = a2; mm[6] = a6; mm[10] = a10; mm[14] = a14;
    } else {
        const double a1 = c * mm[1] - s * mm[0], a5 = c * mm[5] - s * mm[4], a9 = c * mm[9] - s * mm[8], a13 = c * mm[13] - s * mm[12];
        mm[0] = c * mm[0] + s * mm[1]; mm[4] = c * mm[4] + s * mm[5]; mm[8] = c * mm[8] + s * mm[9]; mm[12] = c * mm[12] + s * mm[13];
        mm[1] = a1; mm[5] = a5; mm[9] = a9; mm[13] = a13;
    }
}

// Mat4::postTranslate (math/Mat4.h:714-721): *this = *this * translation(tr), full 4x4 product (operator*=, :439-469)
void postTranslate(Mat4& M, const double tr[3])
{
    const Mat4 m0 = M;
    const double s1[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, tr[0], tr[1], tr[2], 1};
    const double* s0 = m0.m;
    for (int i = 0; i < 4; ++i) {
        const int i4 = 4 * i;
        M.m[i4 + 0] = s0[i4 + 0] * s1[0] + s0[i4 + 1] * s1[4] + s0[i4 + 2] * s1[8] + s0[i4 + 3] * s1[12];
        M.m[i4 + 1] = s0[i4 + 0] * s1[1] + s0[i4 + 1] * s1[5] + s0[i4 + 2] * s1[9] + s0[i4 + 3] * s1[13];
        M.m[i4 + 2] = s0[i4 + 0] * s1[2] + s0[i4 + 1] * s1[6] + s0[i4 + 2] * s1[10] + s0[i4 + 3] * s1[14];
        M.m[i4 + 3] = s0[i4 + 0] * s1[3] + s0[i4 + 1] * s1[7] + s0[i4 + 2] * s1[11] + s0[i4 + 3] * s1[15];
    }
}

// BaseCamera::initRay (tools/RayTracer.h:404-409): eye = applyMap(0), dir = applyJacobian(0,0,-1)
void initRay(vdbrt_camera* c, double t0, double t1)
{
    const double* m = c->m;
    const double z0 = 0.0, zm = -1.0;
    c->t0 = t0; c->t1 = t1;
    c->eye[0] = z0 * m[0] + z0 * m[4] + z0 * m[8] + m[12];       // Vec3 * Mat4 (math/Mat4.h:1180-1188)
    c->eye[1] = z0 * m[1] + z0 * m[5] + z0 * m[9] + m[13];
    c->eye[2] = z0 * m[2] + z0 * m[6] + z0 * m[10] + m[14];
    c->dir[0] = z0 * m[0] + z0 * m[4] + zm * m[8];               // Mat4::transform3x3 (math/Mat4.h:1070-1076)
    c->dir[1] = z0 * m[1] + z0 * m[5] + zm * m[9];
    c->dir[2] = z0 * m[2] + z0 * m[6] + zm * m[10];
}

// BaseCamera ctor (tools/RayTracer.h:354-366)
int baseCamera(vdbrt_camera* c, uint32_t kind, uint32_t w, uint32_t h, const double rot[3], const double tr[3],
               double frameWidth, double nearPlane, double farPlane)
{
    if (!c || !rot || !tr) return vdbrt::setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (w == 0 || h == 0) return vdbrt::setError(VDBRT_ERR_INVALID_ARG, "empty film");
    std::memset(c, 0, sizeof(*c));
    c->kind = kind; c->width = w; c->height = h;
    c->scale_w = frameWidth;
    c->scale_h = frameWidth * double(h) / double(w);
    const double pi = 3.14159265358979323846;                    // math::pi<double>()
    Mat4 M = identity();
    postRotate(M, 0, rot[0] * pi / 180.0);
    postRotate(M, 1, rot[1] * pi / 180.0);
    postRotate(M, 2, rot[2] * pi / 180.0);
    postTranslate(M, tr);
    std::memcpy(c->m, M.m, sizeof(M.m));
    initRay(c, nearPlane, farPlane);
    return VDBRT_OK;
}

double length3(const double v[3]) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
// Vec3::unit(eps = 0) (math/Vec3.h:375-389): false <=> the reference throws ArithmeticError
bool unit3(const double v[3], double out[3])
{
    const double len = length3(v);
    if (!(std::fabs(len - 0.0) > 0.0)) return false;
    out[0] = v[0] / len; out[1] = v[1] / len; out[2] = v[2] / len;
    return true;
}
void cross3(const double a[3], const double b[3], double out[3])   // math/Vec3.h:221-226
{
    out[0] = a[1] * b[2] - a[2] * b[1]; out[1] = a[2] * b[0] - a[0] * b[2]; out[2] = a[0] * b[1] - a[1] * b[0];
}

} // namespace

extern "C" {

int vdbrt_camera_perspective(vdbrt_camera* cam, uint32_t width, uint32_t height, const double rotation[3], const double translation[3],
                             double focal_length, double aperture, double near_plane, double far_plane)
{
    // PerspectiveCamera ctor: frame width = 0.5*aperture/focalLength (tools/RayTracer.h:436-445)
    return baseCamera(cam, VDBRT_CAMERA_PERSPECTIVE, width, height, rotation, translation, 0.5 * aperture / focal_length, near_plane, far_plane);
}

int vdbrt_camera_orthographic(vdbrt_camera* cam, uint32_t width, uint32_t height, const double rotation[3], const double translation[3],
                              double frame_width, double near_plane, double far_plane)
{
    // OrthographicCamera ctor: 0.5*frameWidth (tools/RayTracer.h:494-502)
    return baseCamera(cam, VDBRT_CAMERA_ORTHOGRAPHIC, width, height, rotation, translation, 0.5 * frame_width, near_plane, far_plane);
}

// BaseCamera::lookAt (tools/RayTracer.h:379-389) with math::aim (math/Mat.h:726-745)
int vdbrt_camera_look_at(vdbrt_camera* cam, const double xyz[3], const double upIn[3])
{
    if (!cam || !xyz) return vdbrt::setError(VDBRT_ERR_INVALID_ARG, "null argument");
    const double defUp[3] = {0.0, 1.0, 0.0};
    const double* up = upIn ? upIn : defUp;
    const double* m = cam->m;
    const double z0 = 0.0;
    const double orig[3] = {z0 * m[0] + z0 * m[4] + z0 * m[8] + m[12], z0 * m[1] + z0 * m[5] + z0 * m[9] + m[13], z0 * m[2] + z0 * m[6] + z0 * m[10] + m[14]};
    const double dir[3] = {orig[0] - xyz[0], orig[1] - xyz[1], orig[2] - xyz[2]};
    double forward[3], upUnit[3], tmp[3], horizontal[3], up2[3];
    // any failing unit() throws in the reference and lookAt swallows it: the camera stays as it was
    if (!unit3(dir, forward)) return VDBRT_OK;
    if (!unit3(up, upUnit)) return VDBRT_OK;
    cross3(upUnit, forward, tmp);
    if (!unit3(tmp, horizontal)) return VDBRT_OK;
    cross3(forward, horizontal, tmp);
    if (!unit3(tmp, up2)) return VDBRT_OK;
    Mat4 M;
    M.m[0] = horizontal[0]; M.m[1] = horizontal[1]; M.m[2] = horizontal[2]; M.m[3] = 0.0;
    M.m[4] = up2[0]; M.m[5] = up2[1]; M.m[6] = up2[2]; M.m[7] = 0.0;
    M.m[8] = forward[0]; M.m[9] = forward[1]; M.m[10] = forward[2]; M.m[11] = 0.0;
    M.m[12] = 0.0; M.m[13] = 0.0; M.m[14] = 0.0; M.m[15] = 1.0;     // padMat4
    postTranslate(M, orig);
    std::memcpy(cam->m, M.m, sizeof(M.m));
    initRay(cam, cam->t0, cam->t1);
    return VDBRT_OK;
}

// PerspectiveCamera::getRay (tools/RayTracer.h:452-462) / OrthographicCamera::getRay (:505-512) on the host: the rays the render kernels
// build themselves (cameraRay, vdbrt_device.cuh), for callers that trace their own batches (vdbrt_intersect_levelset, vdbrt_volume_spans).
// pixels: n (i, j) pairs; offsets: n (iOffset, jOffset) pairs or NULL for the pixel centres (0.5, 0.5).
int vdbrt_camera_get_rays(const vdbrt_camera* c, const uint32_t* pixels, const double* offsets, uint64_t n, vdbrt_ray* rays)
{
    if (!c || !pixels || !rays) return vdbrt::setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (c->kind > VDBRT_CAMERA_ORTHOGRAPHIC) return vdbrt::setError(VDBRT_ERR_INVALID_ARG, "unknown camera kind");
    const double* m = c->m;
    for (uint64_t k = 0; k < n; ++k) {
        const double io = offsets ? offsets[2 * k] : 0.5, jo = offsets ? offsets[2 * k + 1] : 0.5;
        // rasterToScreen (:391-395)
        const double sx = (2 * (double(pixels[2 * k]) + io) / double(c->width) - 1) * c->scale_w;
        const double sy = (1 - 2 * (double(pixels[2 * k + 1]) + jo) / double(c->height)) * c->scale_h;
        vdbrt_ray& r = rays[k];
        r.t0 = c->t0; r.t1 = c->t1;
        if (c->kind == VDBRT_CAMERA_PERSPECTIVE) {
            const double sz = -1.0;
            double d[3] = {sx * m[0] + sy * m[4] + sz * m[8], sx * m[1] + sy * m[5] + sz * m[9], sx * m[2] + sy * m[6] + sz * m[10]};   // applyJacobian
            const double len = length3(d);
            if (std::fabs(len - 0.0) > 1.0e-7) { const double s = 1.0 / len; d[0] *= s; d[1] *= s; d[2] *= s; }    // Vec3::normalize (math/Vec3.h:363-371)
            const double sc = 1.0 / (d[0] * c->dir[0] + d[1] * c->dir[1] + d[2] * c->dir[2]);
            r.t0 *= sc; r.t1 *= sc;                                                                                  // scaleTimes
            for (int a = 0; a < 3; ++a) { r.eye[a] = c->eye[a]; r.dir[a] = d[a]; }
        } else {
            const double sz = 0.0;
            r.eye[0] = sx * m[0] + sy * m[4] + sz * m[8] + m[12];                                                    // applyMap
            r.eye[1] = sx * m[1] + sy * m[5] + sz * m[9] + m[13];
            r.eye[2] = sx * m[2] + sy * m[6] + sz * m[10] + m[14];
            for (int a = 0; a < 3; ++a) r.dir[a] = c->dir[a];
        }
    }
    return VDBRT_OK;
}

// LevelSetRayTracer::setPixelSamples (tools/RayTracer.h:883-885): math::Rand01<double>(seed) = std::mt19937 +
// std::uniform_real_distribution<double> (math/Math.h:176-206)
int vdbrt_jitter_table(unsigned int seed, double out[16])
{
    if (!out) return vdbrt::setError(VDBRT_ERR_INVALID_ARG, "null argument");
    std::mt19937 engine(static_cast<std::mt19937::result_type>(seed));
    std::uniform_real_distribution<double> dist;
    for (int i = 0; i < 16; ++i) out[i] = dist(engine);
    return VDBRT_OK;
}

// sphere set of BASELINE configs 4/5
int vdbrt_random_spheres(uint64_t seed, uint32_t n, double extent, double rmin, double rmax, double* out)
{
    if (!out) return vdbrt::setError(VDBRT_ERR_INVALID_ARG, "null argument");
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> pos(-extent, extent), rad(rmin, rmax);
    for (uint32_t s = 0; s < n; ++s) { out[4 * s] = pos(rng); out[4 * s + 1] = pos(rng); out[4 * s + 2] = pos(rng); out[4 * s + 3] = rad(rng); }
    return VDBRT_OK;
}

// VolumeRender ctor defaults (tools/RayTracer.h:929-936)
int vdbrt_vol_opts_default(vdbrt_vol_opts* o)
{
    if (!o) return vdbrt::setError(VDBRT_ERR_INVALID_ARG, "null argument");
    std::memset(o, 0, sizeof(*o));
    o->primary_step = 1.0; o->shadow_step = 3.0; o->cutoff = 0.005; o->light_gain = 0.2;
    const double l[3] = {0.3, 0.3, 0.0};
    unit3(l, o->light_dir);                                      // Vec3R(0.3, 0.3, 0).unit()
    for (int a = 0; a < 3; ++a) { o->light_color[a] = 0.7; o->absorption[a] = 0.1; o->scattering[a] = 1.5; }
    return VDBRT_OK;
}

} // extern "C"
