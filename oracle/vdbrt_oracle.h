/* vdbrt_oracle.h -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * C API of the CPU restatement of the ray-tracing hot path (vdbrt_oracle.cc -> libvdbrt_oracle.so).  It takes the
 * same PODs as the product (include/vdbrt.h) but runs on host NanoVDB buffers with plain scalar C++.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * PARITY STATUS: pinned -- tests/test_oracle_vs_reference.py checks every function here bit-for-bit against
 * the unmodified reference compiled into oracle/_ref/libvdbref.so, and tests/test_reference_kats.py replays the
 * reference's own known-answer tests (TestRay.cc, TestLevelSetRayIntersector.cc, TestVolumeRayIntersector.cc).
 */
#ifndef VDBRT_ORACLE_H_INCLUDED
#define VDBRT_ORACLE_H_INCLUDED
#include "../include/vdbrt.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_grid oracle_grid;

const char* oracle_last_error(void);
/* parse + validate a NanoGrid<float> buffer (kept by reference: the caller keeps `buf` alive) */
int  oracle_grid_open(const void* buf, uint64_t bytes, oracle_grid** out);
void oracle_grid_close(oracle_grid* g);
int  oracle_grid_get_info(const oracle_grid* g, vdbrt_grid_info* info);
/* probeValue on arbitrary coordinates */
int  oracle_grid_probe(const oracle_grid* g, const int32_t* ijk, uint64_t n, float* values, uint8_t* active);

int  oracle_render_levelset(const oracle_grid* g, const vdbrt_camera* cam, const vdbrt_shader* shader,
                            const vdbrt_ls_opts* opts, vdbrt_film* film, vdbrt_aux* aux, vdbrt_counters* ctr,
                            int threads);
int  oracle_render_volume(const oracle_grid* g, const vdbrt_camera* cam, const vdbrt_vol_opts* opts,
                          vdbrt_film* film, vdbrt_counters* ctr, int threads);
int  oracle_intersect_levelset(const oracle_grid* g, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso,
                               vdbrt_hit* hits);
/* test knob: testers created after this call evaluate tester.init's value on demand (the CUDA kernels' evaluation order; same results, fewer
 * stencil refills) */
void oracle_set_lazy_init(int on);
/* LinearSearchImpl<GridT, Iterations>: `iterations` secant refinements of the hit time (tools/RayIntersector.h:630-636) */
int  oracle_intersect_levelset_iter(const oracle_grid* g, const vdbrt_ray* rays, uint64_t n, uint32_t space, float iso,
                                    uint32_t iterations, vdbrt_hit* hits);
int  oracle_volume_spans(const oracle_grid* g, const vdbrt_ray* rays, uint64_t n, uint32_t space, uint32_t max_spans,
                         double* spans, int32_t* counts);
/* BaseCamera::getRay for pixels ij[2k],ij[2k+1] with offsets (NULL -> 0.5,0.5) */
typedef struct oracle_color oracle_color;   /* a NanoGrid<Vec3f> feeding the colour-grid shaders */
int  oracle_color_open(const void* nanovdb_buffer, uint64_t bytes, oracle_color** out);
void oracle_color_close(oracle_color* c);
int  oracle_render_levelset_color(const oracle_grid* g, const oracle_color* color, const vdbrt_camera* cam, const vdbrt_shader* shader,
                                  const vdbrt_ls_opts* opts, vdbrt_film* film, vdbrt_aux* aux, vdbrt_counters* ctr, int threads);
int  oracle_film_over(float* top, const float* bottom, uint64_t pixels);     /* Film::RGBA::over per pixel, top = top.over(bottom) */
int  oracle_camera_rays(const vdbrt_camera* cam, const uint32_t* ij, const double* offsets, uint64_t n, vdbrt_ray* rays);
/* math::DDA<Ray,Log2Dim> trace for the TestRay.testDDA known answers: writes up to max_steps records of
 * {time, next, voxel x,y,z} (5 doubles each) and returns the number of records                                 */
int  oracle_dda_trace(const vdbrt_ray* ray, int log2dim, int max_steps, double* out);
/* math::Ray::clip(CoordBBox) as LinearSearchImpl (pad=0) / VolumeRayIntersector (pad=1) use it: returns hit flag */
int  oracle_ray_clip(const vdbrt_ray* ray, const int32_t bbox[6], double* t0, double* t1);

#ifdef __cplusplus
}
#endif
#endif
