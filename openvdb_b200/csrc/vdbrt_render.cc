// vdbrt_render -- the command line of openvdb_cmd/vdb_render/main.cc over the GPU path.
//
// Same options, same defaults (main.cc:56-106), the same "look at the centre of the volume unless -rotate or -lookat is
// given" rule (main.cc:807-813), the same timed region behind -v (intersector construction + render, main.cc:475-503) and
// the same PPM writer.  Written against the C++ facade (include/vdbrt/RayTracer.h) the way main.cc is written against
// tools/RayTracer.h.  Differences, all forced by the scope of this library:
//   * the input is a NanoVDB file (.nvdb: segments, codecs NONE / ZIP, or a raw grid buffer) instead of a .vdb file, or one of
//     the built-in generators  sphere:R[,voxel[,halfwidth]]  torus:R,r  fogsphere:R  (the GPU box has no asset files);
//   * .ppm and .png output (the .png is the 8-bit RGB image of main.cc:335-390 -- Film::convertToBitBuffer<uint8_t>(alpha = false) --
//     written with zlib alone, no libpng: same pixels, not the same file bytes); .exr needs OpenEXR and is refused the way the
//     reference refuses it when built without (main.cc:248-253);
//   * -color NAME names a Vec3f grid of the same file (the colour-grid forms of the four shaders);
//   * -cpus is accepted and ignored, -gpu N picks the device.
#include <vdbrt/RayTracer.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include <zlib.h>

using namespace vdbrt;

namespace {

const char* gProgName = "vdbrt_render";

// PngWriter::write (main.cc:340-390): 8-bit RGB, no interlace, the bytes of Film::convertToBitBuffer<uint8_t>(alpha = false).
// One IDAT chunk, filter type 0 on every row, deflated by zlib.
void savePNG(const std::string& fname, tools::Film& film)
{
    const size_t w = film.width(), h = film.height();
    auto bits = film.convertToBitBuffer<unsigned char>(/*alpha=*/false);
    std::vector<unsigned char> rows((3 * w + 1) * h);
    for (size_t y = 0; y < h; ++y) {
        rows[(3 * w + 1) * y] = 0;
        std::memcpy(&rows[(3 * w + 1) * y + 1], bits.get() + 3 * w * y, 3 * w);
    }
    uLongf clen = compressBound(uLong(rows.size()));
    std::vector<unsigned char> z(clen);
    if (compress2(z.data(), &clen, rows.data(), uLong(rows.size()), Z_DEFAULT_COMPRESSION) != Z_OK) throw RuntimeError("Error writing PNG data buffers.");
    std::FILE* fp = std::fopen(fname.c_str(), "wb");
    if (!fp) throw IoError("Unable to open '" + fname + "' for writing");
    auto be32 = [](unsigned char* p, uint32_t v) { p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v; };
    auto chunk = [&](const char* type, const unsigned char* data, size_t n) {
        unsigned char head[8];
        be32(head, uint32_t(n)); std::memcpy(head + 4, type, 4);
        uLong crc = crc32(0L, head + 4, 4);
        if (n) crc = crc32(crc, data, uInt(n));
        unsigned char tail[4]; be32(tail, uint32_t(crc));
        std::fwrite(head, 1, 8, fp); if (n) std::fwrite(data, 1, n, fp); std::fwrite(tail, 1, 4, fp);
    };
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::fwrite(sig, 1, 8, fp);
    unsigned char ihdr[13];
    be32(ihdr, uint32_t(w)); be32(ihdr + 4, uint32_t(h));
    ihdr[8] = 8; ihdr[9] = 2 /* PNG_COLOR_TYPE_RGB */; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0 /* PNG_INTERLACE_NONE */;
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), size_t(clen));
    chunk("IEND", nullptr, 0);
    const bool ok = !std::ferror(fp);
    std::fclose(fp);
    if (!ok) throw IoError("Error writing PNG data buffers.");
}
const double LIGHT_DEFAULTS[] = {0.3, 0.3, 0.0, 0.7, 0.7, 0.7};

struct RenderOpts {                               // main.cc:56-106
    std::string shader = "diffuse", color, camera = "perspective";
    float aperture = 41.2136f, focal = 50.0f, frame = 1.0f, znear = 1.0e-3f, zfar = std::numeric_limits<float>::max();
    double isovalue = 0.0;
    Vec3R rotate{0.0}, translate{0.0}, target{0.0}, up{0.0, 1.0, 0.0};
    bool lookat = false;
    size_t samples = 1;
    Vec3R absorb{0.1};
    std::vector<double> light{LIGHT_DEFAULTS, LIGHT_DEFAULTS + 6};
    Vec3R scatter{1.5};
    double cutoff = 0.005, gain = 0.2;
    double step[2] = {1.0, 3.0};
    size_t width = 1920, height = 1080;
    int threads = 0, gpu = 0;
    bool verbose = false;

    std::string validate() const
    {
        if (shader != "diffuse" && shader != "matte" && shader != "normal" && shader != "position")
            return "expected diffuse, matte, normal or position shader, got \"" + shader + "\"";
        if (camera.rfind("ortho", 0) != 0 && camera.rfind("persp", 0) != 0)
            return "expected perspective or orthographic camera, got \"" + camera + "\"";
        if (width < 1 || height < 1) { std::ostringstream o; o << "expected width > 0 and height > 0, got " << width << "x" << height; return o.str(); }
        return "";
    }
};

std::ostream& operator<<(std::ostream& os, const RenderOpts& o)    // RenderOpts::put (main.cc:124-156)
{
    os << " -absorb " << o.absorb.x << "," << o.absorb.y << "," << o.absorb.z << " -aperture " << o.aperture << " -camera " << o.camera
       << " -cpus " << o.threads << " -cutoff " << o.cutoff << " -far " << o.zfar << " -focal " << o.focal << " -frame " << o.frame
       << " -gain " << o.gain << " -isovalue " << o.isovalue
       << " -light " << o.light[0] << "," << o.light[1] << "," << o.light[2] << "," << o.light[3] << "," << o.light[4] << "," << o.light[5];
    if (o.lookat) os << " -lookat " << o.target.x << "," << o.target.y << "," << o.target.z;
    os << " -near " << o.znear << " -res " << o.width << "x" << o.height;
    if (!o.lookat) os << " -rotate " << o.rotate.x << "," << o.rotate.y << "," << o.rotate.z;
    os << " -shader " << o.shader << " -samples " << o.samples << " -scatter " << o.scatter.x << "," << o.scatter.y << "," << o.scatter.z
       << " -shadowstep " << o.step[1] << " -step " << o.step[0] << " -translate " << o.translate.x << "," << o.translate.y << "," << o.translate.z;
    if (o.lookat) os << " -up " << o.up.x << "," << o.up.y << "," << o.up.z;
    if (o.verbose) os << " -v";
    return os;
}

[[noreturn]] void usage(int status = EXIT_FAILURE)
{
    RenderOpts o;
    const double fov = 360.0 / M_PI * std::atan(o.aperture / (2.0 * o.focal));      // focalLengthToFieldOfView (RayTracer.h:466-469)
    std::ostringstream s;
    s << std::setprecision(3) <<
"Usage: " << gProgName << " in.nvdb out.{ppm,png} [options]\n"
"Which: ray-traces NanoVDB volumes on the GPU (option set of OpenVDB's vdb_render)\n"
"       in.nvdb may also be sphere:R[,voxel[,halfwidth]], torus:R,r or fogsphere:R\n"
"Options:\n"
"    -aperture F       perspective camera aperture in mm (default: " << o.aperture << ")\n"
"    -camera S         camera type; either \"persp[ective]\" or \"ortho[graphic]\" (default: " << o.camera << ")\n"
"    -cpus N           accepted for compatibility, ignored\n"
"    -gpu N            CUDA device (default: 0)\n"
"    -far F            camera far plane depth (default: " << o.zfar << ")\n"
"    -focal F          perspective camera focal length in mm (default: " << o.focal << ")\n"
"    -fov F            perspective camera field of view in degrees (default: " << fov << ")\n"
"    -frame F          ortho camera frame width in world units (default: " << o.frame << ")\n"
"    -lookat X,Y,Z     rotate the camera to point to (X, Y, Z)\n"
"    -name S           name of the volume to be rendered (default: the first floating-point volume in the file)\n"
"    -near F           camera near plane depth (default: " << o.znear << ")\n"
"    -res WxH          image dimensions in pixels (default: " << o.width << "x" << o.height << ")\n"
"    -r, -rotate X,Y,Z camera rotation in degrees (default: look at the center of the volume)\n"
"    -t, -translate X,Y,Z  camera translation\n"
"    -up X,Y,Z         vector that should point up after rotation with -lookat (default: [0, 1, 0])\n"
"    -v                verbose (print timing and diagnostics)\n"
"    -h, -help         print this usage message and exit\n"
"Level set options:\n"
"    -color S          name of a vec3s volume to be used to set material colors\n"
"    -isovalue F       isovalue in world units for level set ray intersection (default: " << o.isovalue << ")\n"
"    -samples N        number of samples (rays) per pixel (default: " << o.samples << ")\n"
"    -shader S         shader name; either \"diffuse\", \"matte\", \"normal\" or \"position\" (default: " << o.shader << ")\n"
"Dense volume options:\n"
"    -absorb R,G,B     absorption coefficients (default: [0.1, 0.1, 0.1])\n"
"    -cutoff F         density and transmittance cutoff value (default: " << o.cutoff << ")\n"
"    -gain F           amount of scatter along the shadow ray (default: " << o.gain << ")\n"
"    -light X,Y,Z[,R,G,B]  light source direction and optional color (default: [0.3, 0.3, 0, 0.7, 0.7, 0.7])\n"
"    -scatter R,G,B    scattering coefficients (default: [1.5, 1.5, 1.5])\n"
"    -shadowstep F     step size in voxels for integration along the shadow ray (default: " << o.step[1] << ")\n"
"    -step F           step size in voxels for integration along the primary ray (default: " << o.step[0] << ")\n";
    std::cerr << s.str();
    std::exit(status);
}

std::vector<double> strToVec(const std::string& s)                 // main.cc:278-297
{
    std::vector<double> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(std::atof(tok.c_str()));
    return v;
}
Vec3R strToVec3d(const std::string& s)
{
    const std::vector<double> v = strToVec(s);
    return Vec3R(v.size() > 0 ? v[0] : 0.0, v.size() > 1 ? v[1] : 0.0, v.size() > 2 ? v[2] : 0.0);
}
bool endsWith(const std::string& s, const std::string& e) { return s.size() >= e.size() && s.compare(s.size() - e.size(), e.size(), e) == 0; }

FloatGrid::Ptr openGrid(Context& ctx, const std::string& name, const std::string& gridName)
{
    const size_t colon = name.find(':');
    if (colon != std::string::npos && name.find('/') == std::string::npos && !endsWith(name, ".nvdb")) {
        const std::string kind = name.substr(0, colon);
        const std::vector<double> a = strToVec(name.substr(colon + 1));
        if (kind == "sphere" && !a.empty()) return FloatGrid::createLevelSetSphere(ctx, a[0], Vec3R(0.0), a.size() > 1 ? a[1] : 1.0, a.size() > 2 ? a[2] : 3.0);
        if (kind == "torus" && a.size() >= 2) return FloatGrid::createLevelSetTorus(ctx, a[0], a[1], Vec3R(0.0), a.size() > 2 ? a[2] : 1.0, a.size() > 3 ? a[3] : 3.0);
        if (kind == "fogsphere" && !a.empty()) return FloatGrid::createLevelSetSphere(ctx, a[0], Vec3R(0.0), a.size() > 1 ? a[1] : 1.0, 3.0)->sdfToFogVolume();
        throw ValueError("unknown generator \"" + name + "\"");
    }
    return FloatGrid::read(ctx, name, gridName);
}

// render<GridType>() of main.cc:414-520
void render(FloatGrid& grid, const Vec3SGrid* colorgrid, const std::string& imgFilename, const RenderOpts& opts)
{
    const vdbrt_grid_info info = grid.info();
    const bool isLevelSet = info.grid_class == VDBRT_GRID_CLASS_LEVEL_SET;
    tools::Film film(opts.width, opts.height);
    std::unique_ptr<tools::BaseCamera> camera;
    if (opts.camera.rfind("persp", 0) == 0)
        camera.reset(new tools::PerspectiveCamera(film, opts.rotate, opts.translate, opts.focal, opts.aperture, opts.znear, opts.zfar));
    else
        camera.reset(new tools::OrthographicCamera(film, opts.rotate, opts.translate, opts.frame, opts.znear, opts.zfar));
    if (opts.lookat) camera->lookAt(opts.target, opts.up);

    std::unique_ptr<tools::BaseShader> shader;
    if (opts.shader == "matte") { if (colorgrid) shader.reset(new tools::MatteShader<Vec3SGrid>(*colorgrid)); else shader.reset(new tools::MatteShader<>()); }
    else if (opts.shader == "normal") { if (colorgrid) shader.reset(new tools::NormalShader<Vec3SGrid>(*colorgrid)); else shader.reset(new tools::NormalShader<>()); }
    else if (opts.shader == "position") {
        // bboxIndex(bbox.min().asVec3d(), bbox.max().asVec3d()).applyMap(map): scale(+translate) maps keep the corners (main.cc:452-455)
        double lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            const double p = info.index_bbox[a] * info.voxel_size[a] + info.translation[a], q = info.index_bbox[3 + a] * info.voxel_size[a] + info.translation[a];
            lo[a] = std::min(p, q); hi[a] = std::max(p, q);
        }
        if (colorgrid) shader.reset(new tools::PositionShader<Vec3SGrid>(Vec3R(lo[0], lo[1], lo[2]), Vec3R(hi[0], hi[1], hi[2]), *colorgrid));
        else shader.reset(new tools::PositionShader<>(Vec3R(lo[0], lo[1], lo[2]), Vec3R(hi[0], hi[1], hi[2])));
    } else if (colorgrid) shader.reset(new tools::DiffuseShader<Vec3SGrid>(*colorgrid));
    else shader.reset(new tools::DiffuseShader<>());

    if (opts.verbose) std::cout << gProgName << ": ray-tracing..." << std::endl;
    const auto start = std::chrono::steady_clock::now();
    if (isLevelSet) {
        tools::LevelSetRayIntersector<FloatGrid> intersector(grid, float(opts.isovalue));
        tools::rayTrace(grid, intersector, *shader, *camera, opts.samples, /*seed=*/0, opts.threads != 1);
    } else {
        using IntersectorType = tools::VolumeRayIntersector<FloatGrid>;
        IntersectorType intersector(grid);
        tools::VolumeRender<IntersectorType> renderer(intersector, *camera);
        renderer.setLightDir(opts.light[0], opts.light[1], opts.light[2]);
        renderer.setLightColor(opts.light[3], opts.light[4], opts.light[5]);
        renderer.setPrimaryStep(opts.step[0]);
        renderer.setShadowStep(opts.step[1]);
        renderer.setScattering(opts.scatter.x, opts.scatter.y, opts.scatter.z);
        renderer.setAbsorption(opts.absorb.x, opts.absorb.y, opts.absorb.z);
        renderer.setLightGain(opts.gain);
        renderer.setCutOff(opts.cutoff);
        renderer.render(opts.threads != 1);
    }
    if (opts.verbose) {
        std::ostringstream o;
        o << gProgName << ": ...completed in " << std::setprecision(3) << std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() << " sec";
        std::cout << o.str() << std::endl;
    }
    if (endsWith(imgFilename, ".png")) { savePNG(imgFilename, film); return; }
    if (!endsWith(imgFilename, ".ppm")) throw ValueError("unsupported image file format (" + imgFilename + ")");
    std::string filename = imgFilename;
    filename.erase(filename.size() - 4);          // strip .ppm extension; savePPM appends it again
    film.savePPM(filename);
}

} // namespace

int main(int argc, char* argv[])
{
    gProgName = argv[0];
    if (const char* p = std::strrchr(gProgName, '/')) gProgName = p + 1;
    if (argc == 1) usage();
    std::string vdbFilename, imgFilename, gridName;
    RenderOpts opts;
    bool hasFocal = false, hasFov = false, hasRotate = false, hasLookAt = false;
    float fov = 0.0f;
    auto need = [&](int i) { if (i + 1 >= argc) { std::cerr << gProgName << ": option " << argv[i] << " requires an argument\n"; usage(); } };
    for (int i = 1; i < argc; ++i) {
        const std::string arg = argv[i];
        if (arg[0] == '-' && arg.size() > 1 && !(std::isdigit(arg[1]))) {
            if (arg == "-absorb") { need(i); opts.absorb = strToVec3d(argv[++i]); }
            else if (arg == "-aperture") { need(i); opts.aperture = float(std::atof(argv[++i])); }
            else if (arg == "-camera") { need(i); opts.camera = argv[++i]; }
            else if (arg == "-color") { need(i); opts.color = argv[++i]; }
            else if (arg == "-compression") { need(i); ++i; }                    // EXR only
            else if (arg == "-cpus") { need(i); opts.threads = std::max(0, std::atoi(argv[++i])); }
            else if (arg == "-gpu") { need(i); opts.gpu = std::atoi(argv[++i]); }
            else if (arg == "-cutoff") { need(i); opts.cutoff = std::atof(argv[++i]); }
            else if (arg == "-isovalue") { need(i); opts.isovalue = std::atof(argv[++i]); }
            else if (arg == "-far") { need(i); opts.zfar = float(std::atof(argv[++i])); }
            else if (arg == "-focal") { need(i); opts.focal = float(std::atof(argv[++i])); hasFocal = true; }
            else if (arg == "-fov") { need(i); fov = float(std::atof(argv[++i])); hasFov = true; }
            else if (arg == "-frame") { need(i); opts.frame = float(std::atof(argv[++i])); }
            else if (arg == "-gain") { need(i); opts.gain = std::atof(argv[++i]); }
            else if (arg == "-light") { need(i); opts.light = strToVec(argv[++i]); opts.light.resize(6, 0.7); }
            else if (arg == "-lookat") { need(i); opts.lookat = true; opts.target = strToVec3d(argv[++i]); hasLookAt = true; }
            else if (arg == "-name") { need(i); gridName = argv[++i]; }
            else if (arg == "-near") { need(i); opts.znear = float(std::atof(argv[++i])); }
            else if (arg == "-r" || arg == "-rotate") { need(i); opts.rotate = strToVec3d(argv[++i]); hasRotate = true; }
            else if (arg == "-res") { need(i); unsigned w = 0, h = 0; if (std::sscanf(argv[++i], "%ux%u", &w, &h) == 2) { opts.width = w; opts.height = h; } }
            else if (arg == "-scatter") { need(i); opts.scatter = strToVec3d(argv[++i]); }
            else if (arg == "-shader") { need(i); opts.shader = argv[++i]; }
            else if (arg == "-shadowstep") { need(i); opts.step[1] = std::atof(argv[++i]); }
            else if (arg == "-samples") { need(i); opts.samples = size_t(std::max(0, std::atoi(argv[++i]))); }
            else if (arg == "-step") { need(i); opts.step[0] = std::atof(argv[++i]); }
            else if (arg == "-t" || arg == "-translate") { need(i); opts.translate = strToVec3d(argv[++i]); }
            else if (arg == "-up") { need(i); opts.up = strToVec3d(argv[++i]); }
            else if (arg == "-v") opts.verbose = true;
            else if (arg == "-h" || arg == "-help" || arg == "--help") usage(EXIT_SUCCESS);
            else { std::cerr << gProgName << ": \"" << arg << "\" is not a valid option\n"; usage(); }
        } else if (vdbFilename.empty()) vdbFilename = arg;
        else if (imgFilename.empty()) imgFilename = arg;
        else usage();
    }
    if (vdbFilename.empty() || imgFilename.empty()) usage();
    if (hasFov) {
        if (hasFocal) { std::cerr << gProgName << ": specify -focal or -fov, but not both\n"; usage(); }
        opts.focal = float(opts.aperture / (2.0 * std::tan(fov * M_PI / 360.0)));      // fieldOfViewToFocalLength (RayTracer.h:472-475)
    }
    if (hasLookAt && hasRotate) { std::cerr << gProgName << ": specify -lookat or -r[otate], but not both\n"; usage(); }
    { const std::string err = opts.validate(); if (!err.empty()) { std::cerr << gProgName << ": " << err << "\n"; usage(); } }

    int retcode = EXIT_SUCCESS;
    try {
        // isExtensionSupported (main.cc:244-264)
        if (endsWith(imgFilename, ".exr")) throw RuntimeError("vdbrt_render has not been compiled with .exr support.");
        if (!endsWith(imgFilename, ".ppm") && !endsWith(imgFilename, ".png")) throw ValueError("unsupported image file format (" + imgFilename + ")");
        const auto start = std::chrono::steady_clock::now();
        if (opts.verbose) {
            std::cout << gProgName << ": reading ";
            if (!gridName.empty()) std::cout << gridName << " from ";
            std::cout << vdbFilename << "..." << std::endl;
        }
        Context ctx(opts.gpu);
        FloatGrid::Ptr grid = openGrid(ctx, vdbFilename, gridName);
        Vec3SGrid::Ptr colorgrid;
        if (!opts.color.empty()) colorgrid = Vec3SGrid::read(ctx, vdbFilename, opts.color);     // "... is not a vec3s color volume" otherwise
        if (opts.verbose) {
            std::ostringstream o;
            o << gProgName << ": ...completed in " << std::setprecision(3) << std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() << " sec";
            std::cout << o.str() << std::endl;
        }
        if (!hasLookAt && !hasRotate) {
            // point the camera to the centre of the grid: indexToWorld(evalActiveVoxelBoundingBox().getCenter()) (main.cc:807-813,
            // math/Coord.h:378: 0.5 * (min + max))
            const vdbrt_grid_info info = grid->info();
            double c[3];
            for (int a = 0; a < 3; ++a) c[a] = 0.5 * double(info.index_bbox[a] + info.index_bbox[3 + a]) * info.voxel_size[a] + info.translation[a];
            opts.target = Vec3R(c[0], c[1], c[2]);
            opts.lookat = true;
        }
        if (opts.verbose) std::cout << opts << std::endl;
        render(*grid, colorgrid.get(), imgFilename, opts);
    } catch (const std::exception& e) {
        std::cerr << gProgName << ": " << e.what() << std::endl;
        retcode = EXIT_FAILURE;
    }
    return retcode;
}
