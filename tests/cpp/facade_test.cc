// facade_test.cc -- host code written against the reference's API shape (tools::Film / PerspectiveCamera / DiffuseShader /
// LevelSetRayIntersector / rayTrace / VolumeRender), compiled against include/vdbrt/RayTracer.h and run on the GPU.
// Mirrors the body of vdb_render's render<FloatGrid>() (openvdb_cmd/vdb_render/main.cc:415-497) for BASELINE config 1.
#include <vdbrt/RayTracer.h>

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <vector>

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
    // ---- the twelve single-ray overloads of LevelSetRayIntersector (RayIntersector.h:119-240); a miss leaves the outputs alone
    {
        tools::Ray wr(Vec3R(0.0, 0.0, 300.0), Vec3R(0.0, 0.0, -1.0));
        Vec3R w(7.0), n(7.0), w2(7.0); double tw = -1.0, tw2 = -1.0;
        EXPECT(inter.intersectsWS(wr));
        EXPECT(inter.intersectsWS(wr, tw) && std::fabs(tw - 200.0) < 1e-3);
        EXPECT(inter.intersectsWS(wr, w) && std::fabs(w.z - 100.0) < 1e-3);
        EXPECT(inter.intersectsWS(wr, w2, tw2) && w2.z == w.z && tw2 == tw);
        EXPECT(inter.intersectsWS(wr, w2, n) && n.z > 0.999);
        Vec3R xi(7.0); double ti = -1.0;
        EXPECT(inter.intersectsIS(wr));                                    // voxel size 1: the same ray in index space
        EXPECT(inter.intersectsIS(wr, ti) && ti == tw);
        EXPECT(inter.intersectsIS(wr, xi) && xi.z == w.z);
        EXPECT(inter.intersectsIS(wr, xi, ti));
        tools::Ray miss(Vec3R(0.0, 500.0, 300.0), Vec3R(0.0, 0.0, -1.0));
        Vec3R keep(1.0, 2.0, 3.0); double tkeep = 42.0;
        EXPECT(!inter.intersectsWS(miss, keep, tkeep) && keep.x == 1.0 && keep.z == 3.0 && tkeep == 42.0);
        EXPECT(!inter.intersectsIS(miss, keep, tkeep) && keep.y == 2.0 && tkeep == 42.0);
        // LinearSearchImpl<GridT, 2>: two secant refinements move the hit closer to the analytic sphere
        tools::LevelSetRayIntersector<FloatGrid, tools::LinearSearchImpl<FloatGrid, 2>> inter2(*grid);
        tools::Ray slanted(Vec3R(30.0, 20.0, 300.0), Vec3R(0.0, 0.0, -1.0));
        Vec3R p0, p2;
        EXPECT(inter.intersectsWS(slanted, p0) && inter2.intersectsWS(slanted, p2));
        const double e0 = std::fabs(std::sqrt(p0.x * p0.x + p0.y * p0.y + p0.z * p0.z) - 100.0), e2 = std::fabs(std::sqrt(p2.x * p2.x + p2.y * p2.y + p2.z * p2.z) - 100.0);
        std::printf("|x| - r: %.3g with Iterations = 0, %.3g with Iterations = 2\n", e0, e2);
        EXPECT(e2 <= e0 && e2 < 1e-2);
    }
    // ---- VolumeRayIntersector: setWorldRay + march() until the spans are used up == hits() (RayIntersector.h:368-432)
    {
        tools::Ray wr(Vec3R(0.0, 0.0, 300.0), Vec3R(0.0, 0.0, -1.0));
        EXPECT(vinter.setWorldRay(wr));
        std::vector<tools::VolumeRayIntersector<FloatGrid>::TimeSpan> list;
        vinter.hits(list);
        EXPECT(list.size() == 1 && std::fabs(list[0].t0 - 196.0) < 8.5 && std::fabs(list[0].t1 - 404.0) < 8.5);
        double t0 = 0.0, t1 = 0.0; size_t n = 0;
        EXPECT(vinter.setWorldRay(wr));
        while (vinter.march(t0, t1)) { EXPECT(n < list.size() && t0 == list[n].t0 && t1 == list[n].t1); ++n; }
        EXPECT(n == list.size());
        EXPECT(std::fabs(vinter.getWorldPos(list[0].t0).z - (300.0 - list[0].t0)) < 1e-9);
        tools::Ray miss(Vec3R(0.0, 500.0, 300.0), Vec3R(0.0, 0.0, -1.0));
        EXPECT(!vinter.setWorldRay(miss) && !vinter.march(t0, t1));
        std::ostringstream os;
        renderer.print(os, 1);
        EXPECT(os.str().find("Primary step: 0.5") != std::string::npos && os.str().find("LightDir: [0.707107, 0.707107, 0]") != std::string::npos);
        EXPECT(os.str().find("BBox: [-104, -104, -104] -> [104, 104, 104]") != std::string::npos);
        std::printf("%s", os.str().c_str());
    }
    // ---- BaseCamera::getRay / rasterToScreen, the field-of-view helpers (RayTracer.h:391-395,452-475), getWorldTime (RayIntersector.h:442-445)
    {
        // the ray through a pixel that the render above has marked as hit must hit, one that it left alone must miss
        const size_t c = size_t(res) / 2;
        const tools::Ray centre = camera.getRay(c, c), corner = camera.getRay(0, 0, 0.25, 0.75);
        EXPECT(centre.eye[2] == 300.0 && centre.dir[2] < -0.999 && corner.dir[0] < 0.0 && corner.dir[1] > 0.0);
        EXPECT((film.pixel(c, c).r > 0.f) == inter.intersectsWS(centre));
        EXPECT((film.pixel(0, 0).r > 0.f) == inter.intersectsWS(camera.getRay(0, 0)));
        const Vec3R s = camera.rasterToScreen(double(res), 0.0, -1.0);
        EXPECT(s.x > 0.0 && s.y > 0.0 && s.z == -1.0 && std::fabs(s.x - 0.5 * double(41.2136f) / double(50.0f)) < 1e-12);
        const double fov = tools::PerspectiveCamera::focalLengthToFieldOfView(50.0, 41.2136);
        EXPECT(std::fabs(fov - 44.8) < 0.1 && std::fabs(tools::PerspectiveCamera::fieldOfViewToFocalLength(fov, 41.2136) - 50.0) < 1e-9);
        tools::Ray wr(Vec3R(0.0, 0.0, 300.0), Vec3R(0.0, 0.0, -2.0));
        // the index ray's direction is normalised and its times scaled (Ray::applyInverseMap, math/Ray.h:204-213): voxel size 1 -> |J dir| = 1
        EXPECT(vinter.setWorldRay(wr) && std::fabs(vinter.getWorldTime(3.0) - 3.0) < 1e-12);
    }
    if (argc > 2) film.savePPM(argv[2]);
    std::printf("facade ok\n");
    return 0;
}
