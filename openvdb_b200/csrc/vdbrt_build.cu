// vdbrt_build.cu -- GPU construction of NanoVDB level-set / fog grids (inputs for the benches).  Placeholder: filled in
// by a later milestone; the entry points exist so the ABI is complete.
#include "vdbrt_host.h"

extern "C" {
int vdbrt_build_levelset_sphere(vdbrt_ctx*, double, const double*, double, double, vdbrt_grid**) { return vdbrt::setError(VDBRT_ERR_UNSUPPORTED, "not built yet"); }
int vdbrt_build_levelset_torus(vdbrt_ctx*, double, double, const double*, double, double, vdbrt_grid**) { return vdbrt::setError(VDBRT_ERR_UNSUPPORTED, "not built yet"); }
int vdbrt_build_levelset_spheres(vdbrt_ctx*, const double*, uint32_t, double, double, vdbrt_grid**) { return vdbrt::setError(VDBRT_ERR_UNSUPPORTED, "not built yet"); }
int vdbrt_build_fog_from_levelset(vdbrt_ctx*, const vdbrt_grid*, vdbrt_grid**) { return vdbrt::setError(VDBRT_ERR_UNSUPPORTED, "not built yet"); }
}
