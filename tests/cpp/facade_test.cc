// facade_test.cc -- host code written against the reference's API shape (tools::Film / PerspectiveCamera / DiffuseShader /
// LevelSetRayIntersector / rayTrace / VolumeRender), compiled against include/vdbrt/RayTracer.h and run on the GPU.
// Mirrors the body of vdb_render's render<FloatGrid>() (openvdb_cmd/vdb_render/main.cc:415-497) for BASELINE config 1.
#include <vdbrt/RayTracer.h>

#include <cfloat>
#include <cstdio>
#include <cstdlib>

using namespace vdbrt;

#define EXPECT(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main(int argc, char** argv)
{
    const int res = argc > 1 ? std::atoi(argv[1]) : 1024;
    Context ctx(0);
    // createLevelSetSphere<FloatGrid>(100, (0,0,0), 1, 3)  
    FloatGrid::Ptr grid = FloatGrid::createLevelSetSphere(ctx, 100.0, Vec3R(0.0), 1.0, 3.0);
    EXPECT(grid->info().leaf_count == 4025 && grid->info().active_voxels == 753990);

    tools::Film film(res, res);
    // vdb_render stores its lens options as float (main.cc:62,83-87)
    tools::PerspectiveCamera camera(film, Vec3R(0.0), Vec3R(0.0, 0.0, 300.0), double(50.0f), double(41.2136f), double(1e-3f), double(FLT_MAX));
    camera.lookAt(Vec3R(0.0, 0.0, 0.0));
    tools::DiffuseShader<> shader;
    tools::LevelSetRayIntersector<FloatGrid> inter(*grid, 0.0f);
    tools::rayTrace(*grid, inter, shader, camera, /*samples=*/1, /*seed=*/0, /*threaded=*/true);
    size_t hits = 0; double sum = 0.0;
    for (size_t j = 0; j < film.height(); ++j) for (size_t i = 0; i < film.width(); ++i) { const auto& p = film.pixel(i, j); hits += p.r > 0.f; sum += p.r; EXPECT(p.a == 1.0f); }
    std::printf("level set: %zu hit pixels, sum(r) = %.3f\n", hits, sum);
    if (res == 1024) EXPECT(hits == 606028);
    if (res == 256) EXPECT(hits == 37896);

    // single rays through the intersector: centre ray hits (0,0,100) at t = 200
    vdbrt_ray r = {{0.0, 0.0, 300.0}, {0.0, 0.0, -1.0}, 1e-9, DBL_MAX};
    Vec3R xyz, nml; double t = 0.0;
    EXPECT(inter.intersectsWS(r, xyz, nml, t));
    EXPECT(std::fabs(xyz.z - 100.0) < 1e-3 && std::fabs(t - 200.0) < 1e-3 && nml.z > 0.999);

    // error behaviour: same exception kinds at the same construction points as the reference
    bool threw = false;
    try { tools::LevelSetRayIntersector<FloatGrid> bad(*grid, 3.0f); } catch (const ValueError&) { threw = true; }
    EXPECT(threw);
    threw = false;
    try { tools::LevelSetRayTracer<> tr(*grid, shader, camera, 0); } catch (const ValueError&) { threw = true; }
    EXPECT(threw);

    // fog: sdfToFogVolume + VolumeRender with vdb_render's settings (main.cc:486-494), primary step 0.5
    FloatGrid::Ptr fog = grid->sdfToFogVolume();
    threw = false;
    try { tools::LevelSetRayIntersector<FloatGrid> bad(*fog); } catch (const std::runtime_error&) { threw = true; }
    EXPECT(threw);
    tools::Film vfilm(res / 4, res / 4);
    tools::PerspectiveCamera vcam(vfilm, Vec3R(0.0), Vec3R(0.0, 0.0, 300.0), double(50.0f), double(41.2136f), double(1e-3f), double(FLT_MAX));
    vcam.lookAt(Vec3R(0.0));
    tools::VolumeRayIntersector<FloatGrid> vinter(*fog);
    tools::VolumeRender<tools::VolumeRayIntersector<FloatGrid>> renderer(vinter, vcam);
    renderer.setLightDir(0.3, 0.3, 0.0);
    renderer.setLightColor(0.7, 0.7, 0.7);
    renderer.setPrimaryStep(0.5);
    renderer.setShadowStep(3.0);
    renderer.setScattering(1.5, 1.5, 1.5);
    renderer.setAbsorption(0.1, 0.1, 0.1);
    renderer.setLightGain(0.2);
    renderer.setCutOff(0.005);
    renderer.render();
    double alpha = 0.0;
    for (size_t j = 0; j < vfilm.height(); ++j) for (size_t i = 0; i < vfilm.width(); ++i) alpha += vfilm.pixel(i, j).a;
    std::printf("fog: sum(alpha) = %.3f\n", alpha);
    EXPECT(alpha > 100.0);
    if (argc > 2) film.savePPM(argv[2]);
    std::printf("facade ok\n");
    return 0;
}
