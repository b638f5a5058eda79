// vdbrt_io.cc -- NanoVDB file ingestion for the ray-tracing path and the PPM writer of tools::Film.
//
// File layout restated from nanovdb/nanovdb/io/IO.h and NanoVDB.h:5860-5929 (nothing is included from the reference):
//   a file is a sequence of SEGMENTS; a segment is
//     FileHeader   16 B  { uint64 magic ("NanoVDB0" or "NanoVDB2"), uint32 version, uint16 gridCount, uint16 codec }
//     gridCount x  FileMetaData 176 B (+ nameSize bytes of grid name)
//     gridCount x  grid payload: codec NONE = gridSize raw bytes; ZIP = uint64 compressedSize + one zlib stream
//                  (Internal::read, IO.h:276-292); BLOSC is not supported here (the reference needs libblosc for it too).
//   A file that starts with a GridData magic ("NanoVDB1", or "NanoVDB0" followed by a GridData checksum/version) and
//   no segment header is a raw grid buffer (io::writeUncompressedGrid with raw = true, NanoVDB.h:5952-5996).
// Host-only code: no CUDA here.  The buffers it returns are 32-byte aligned as NanoVDB requires.
#include "../../include/vdbrt.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include <sys/types.h>
#include <zlib.h>

namespace vdbrt {
int setError(int code, const std::string& msg);
}
using vdbrt::setError;

namespace {

constexpr uint64_t MAGIC_NUMB = 0x304244566f6e614eULL, MAGIC_GRID = 0x314244566f6e614eULL, MAGIC_FILE = 0x324244566f6e614eULL;
constexpr uint16_t CODEC_NONE = 0, CODEC_ZIP = 1, CODEC_BLOSC = 2;

#pragma pack(push, 1)
struct FileHeader { uint64_t magic; uint32_t version; uint16_t gridCount; uint16_t codec; };
struct FileMetaData {            // NanoVDB.h:5913-5929
    uint64_t gridSize, fileSize, nameKey, voxelCount;
    uint32_t gridType, gridClass;
    double   worldBBox[6];
    int32_t  indexBBox[6];
    double   voxelSize[3];
    uint32_t nameSize;
    uint32_t nodeCount[4];
    uint32_t tileCount[3];
    uint16_t codec;
    uint16_t blindDataCount;
    uint32_t version;
};
#pragma pack(pop)
static_assert(sizeof(FileHeader) == 16 && sizeof(FileMetaData) == 176, "NanoVDB file structs");

struct File {
    std::FILE* f = nullptr;
    int64_t size = 0;                                              // length of the file (input files)
    explicit File(const char* path, const char* mode) : f(std::fopen(path, mode))
    {
        if (f && mode[0] == 'r' && fseeko(f, 0, SEEK_END) == 0) { size = int64_t(ftello(f)); fseeko(f, 0, SEEK_SET); }
    }
    ~File() { if (f) std::fclose(f); }
    bool read(void* dst, size_t n) { return std::fread(dst, 1, n, f) == n; }
    bool write(const void* src, size_t n) { return std::fwrite(src, 1, n, f) == n; }
    int64_t tell() { return int64_t(ftello(f)); }
    bool seek(int64_t at) { return at >= 0 && at <= size && fseeko(f, off_t(at), SEEK_SET) == 0; }
};

struct Entry { FileMetaData meta; std::string name; int64_t payload; };   // payload: file offset of the grid's bytes

// Nothing a file says about sizes is trusted before it has been checked against the length of the file: a grid name is at most
// kMaxName bytes, a payload lies inside the file, an uncompressed grid is as long as its payload, a ZIP stream is shorter than it.
constexpr uint32_t kMaxName = 4096;
constexpr uint64_t kMaxGrid = 1ull << 40;                         // 1 TiB: more than any GPU holds; bounds the allocation for a ZIP payload

// every extern "C" body runs inside this: no C++ exception may cross the C ABI
template<class F> int guarded(F&& body)
{
    try { return body(); }
    catch (const std::bad_alloc&) { return setError(VDBRT_ERR_IO, "out of host memory"); }
    catch (const std::exception& e) { return setError(VDBRT_ERR_BAD_GRID, std::string("malformed NanoVDB file: ") + e.what()); }
}

// io::stringHash (nanovdb/io/IO.h:718-729): the name key stored in FileMetaData (readers compare it before the name)
uint64_t nameHash(const char* s)
{
    uint64_t hash = 0;
    if (!s) return hash;
    for (const unsigned char* p = reinterpret_cast<const unsigned char*>(s); *p; ++p) {
        const uint64_t overflow = hash >> (64 - 8);
        hash *= 67;
        hash += *p + overflow;
    }
    return hash;
}

// walks all segments and lists their grids
int scan(File& in, std::vector<Entry>& out)
{
    for (;;) {
        FileHeader h;
        const int64_t at = in.tell();
        const size_t got = std::fread(&h, 1, sizeof(h), in.f);
        if (got == 0) break;                                        // clean end of file
        if (got != sizeof(h)) return setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB file header");
        if (h.magic != MAGIC_NUMB && h.magic != MAGIC_FILE) {
            if (at == 0 && h.magic == MAGIC_GRID) return setError(VDBRT_ERR_BAD_GRID, "Expected a NanoVDB file, but read a raw NanoVDB grid!");
            if ((h.magic & 0xffffffffULL) == 0x56444220ULL) return setError(VDBRT_ERR_BAD_GRID, "Expected a NanoVDB file, but read an OpenVDB file!");
            return setError(VDBRT_ERR_BAD_GRID, "Expected a NanoVDB file, but read a file of unknown type!");
        }
        if ((h.version >> 21) != 32) return setError(VDBRT_ERR_BAD_GRID, "incompatible NanoVDB file version (need major 32)");
        std::vector<Entry> seg(h.gridCount);
        for (auto& e : seg) {
            if (!in.read(&e.meta, sizeof(FileMetaData))) return setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB grid meta data");
            if (e.meta.nameSize > kMaxName || int64_t(e.meta.nameSize) > in.size - in.tell())
                return setError(VDBRT_ERR_BAD_GRID, "NanoVDB grid name longer than the file allows");
            std::vector<char> name(size_t(e.meta.nameSize) + 1, '\0');
            if (e.meta.nameSize && !in.read(name.data(), e.meta.nameSize)) return setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB grid name");
            e.name = name.data();
        }
        int64_t pos = in.tell();
        for (auto& e : seg) {
            const uint64_t left = uint64_t(in.size - pos);
            if (e.meta.fileSize > left) return setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB file (a grid payload ends beyond the end of the file)");
            if (e.meta.gridSize > kMaxGrid || (e.meta.codec == CODEC_NONE && e.meta.gridSize != e.meta.fileSize) || (e.meta.codec == CODEC_ZIP && e.meta.fileSize < 8))
                return setError(VDBRT_ERR_BAD_GRID, "inconsistent sizes in the NanoVDB grid meta data");
            e.payload = pos; pos += int64_t(e.meta.fileSize); out.push_back(e);
        }
        if (!in.seek(pos)) return setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB file");
    }
    return VDBRT_OK;
}

void* alignedAlloc(uint64_t bytes) { void* p = nullptr; return posix_memalign(&p, 32, size_t((bytes + 31) & ~uint64_t(31))) == 0 ? p : nullptr; }

} // namespace

extern "C" {

int vdbrt_nvdb_list(const char* path, vdbrt_nvdb_meta* out, uint32_t capacity, uint32_t* count)
{
  return guarded([&]() -> int {
    if (!path || !count) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    File in(path, "rb");
    if (!in.f) return setError(VDBRT_ERR_IO, std::string("Unable to open file named \"") + path + "\" for input");
    uint64_t first = 0;
    if (in.read(&first, 8) && first == MAGIC_GRID) {               // a raw grid buffer: one grid, described by its own header
        *count = 1;
        if (out && capacity) {
            std::memset(out, 0, sizeof(*out));
            uint8_t head[736];
            std::rewind(in.f);
            if (!in.read(head, sizeof(head))) return setError(VDBRT_ERR_BAD_GRID, "truncated raw NanoVDB grid");
            std::memcpy(&out->grid_bytes, head + 32, 8); out->file_bytes = out->grid_bytes;
            std::memcpy(out->name, head + 40, 255);
            std::memcpy(&out->grid_class, head + 632, 4); std::memcpy(&out->grid_type, head + 636, 4);
            std::memcpy(&out->active_voxels, head + 672 + 56, 8);
            std::memcpy(out->voxel_size, head + 608, 24);
        }
        return VDBRT_OK;
    }
    std::rewind(in.f);
    std::vector<Entry> all;
    if (int rc = scan(in, all)) return rc;
    *count = uint32_t(all.size());
    for (uint32_t i = 0; out && i < capacity && i < all.size(); ++i) {
        const FileMetaData& m = all[i].meta;
        std::memset(&out[i], 0, sizeof(out[i]));
        std::strncpy(out[i].name, all[i].name.c_str(), sizeof(out[i].name) - 1);
        out[i].grid_bytes = m.gridSize; out[i].file_bytes = m.fileSize; out[i].active_voxels = m.voxelCount;
        out[i].grid_type = m.gridType; out[i].grid_class = m.gridClass; out[i].codec = m.codec;
        for (int k = 0; k < 6; ++k) { out[i].index_bbox[k] = m.indexBBox[k]; out[i].world_bbox[k] = m.worldBBox[k]; }
        for (int k = 0; k < 3; ++k) out[i].voxel_size[k] = m.voxelSize[k];
    }
    return VDBRT_OK;
  });
}

int vdbrt_nvdb_read(const char* path, const char* gridName, void** buffer, uint64_t* bytes)
{
    return vdbrt_nvdb_read_typed(path, gridName, 1u, buffer, bytes);
}

int vdbrt_nvdb_read_typed(const char* path, const char* gridName, uint32_t gridType, void** buffer, uint64_t* bytes)
{
  return guarded([&]() -> int {
    if (!path || !buffer || !bytes) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    *buffer = nullptr; *bytes = 0;
    File in(path, "rb");
    if (!in.f) return setError(VDBRT_ERR_IO, std::string("Unable to open file named \"") + path + "\" for input");
    uint64_t first = 0;
    // "float" includes the quantised float types Fp4/Fp8/Fp16/FpN (GridType 13..16): vdbrt_upload_grid expands them
    auto matches = [gridType](uint32_t t) { return gridType == 0u || t == gridType || (gridType == 1u && t >= 13u && t <= 16u); };
    const char* what = gridType == 1u ? "scalar, floating-point" : (gridType == 6u ? "vec3s color" : "matching");
    if (in.read(&first, 8) && first == MAGIC_GRID) {               // raw grid buffer: one grid, described by its own GridData
        const int64_t size = in.size;
        uint8_t head[736];
        std::rewind(in.f);
        if (size < int64_t(sizeof(head)) || !in.read(head, sizeof(head))) return setError(VDBRT_ERR_BAD_GRID, "truncated raw NanoVDB grid");
        uint64_t gridSize; uint32_t type;
        std::memcpy(&gridSize, head + 32, 8); std::memcpy(&type, head + 636, 4);
        if (gridSize > uint64_t(size)) return setError(VDBRT_ERR_BAD_GRID, "truncated raw NanoVDB grid (grid size exceeds the file)");
        char name[257]; std::memcpy(name, head + 40, 256); name[256] = 0;
        if (gridName && *gridName && std::strcmp(name, gridName) != 0) return setError(VDBRT_ERR_IO, std::string("no grid named \"") + gridName + "\" in file " + path);
        if (!matches(type)) return setError(VDBRT_ERR_NOT_FLOAT, std::string(gridName && *gridName ? gridName : "the grid") + " is not a " + what + " volume");
        void* p = alignedAlloc(uint64_t(size));
        if (!p) return setError(VDBRT_ERR_IO, "out of host memory");
        std::rewind(in.f);
        if (!in.read(p, size_t(size))) { std::free(p); return setError(VDBRT_ERR_BAD_GRID, "truncated raw NanoVDB grid"); }
        *buffer = p; *bytes = uint64_t(size);
        return VDBRT_OK;
    }
    std::rewind(in.f);
    std::vector<Entry> all;
    if (int rc = scan(in, all)) return rc;
    const Entry* pick = nullptr;
    for (const Entry& e : all) {
        if (gridName && *gridName) { if (e.name == gridName) { pick = &e; break; } }
        else if (matches(e.meta.gridType)) { pick = &e; break; }   // vdb_render: the first floating-point volume (main.cc:771-786)
    }
    if (!pick) {
        if (gridName && *gridName) return setError(VDBRT_ERR_IO, std::string("no grid named \"") + gridName + "\" in file " + path);
        return setError(VDBRT_ERR_NOT_FLOAT, std::string("no ") + what + " volumes in file " + path);
    }
    if (!matches(pick->meta.gridType))
        return setError(VDBRT_ERR_NOT_FLOAT, std::string(gridName ? gridName : "") + " is not a " + what + " volume");   // main.cc:766-769,790-794
    void* p = alignedAlloc(pick->meta.gridSize);
    if (!p) return setError(VDBRT_ERR_IO, "out of host memory");
    if (!in.seek(pick->payload)) { std::free(p); return setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB file"); }
    int rc = VDBRT_OK;
    if (pick->meta.codec == CODEC_NONE) {
        if (!in.read(p, size_t(pick->meta.gridSize))) rc = setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB grid payload");
    } else if (pick->meta.codec == CODEC_ZIP) {
        uint64_t csize = 0;
        if (!in.read(&csize, 8)) rc = setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB grid payload");
        else if (csize > pick->meta.fileSize - 8) rc = setError(VDBRT_ERR_BAD_GRID, "ZIP stream longer than its payload");
        else {
            std::vector<unsigned char> tmp(csize);
            if (!in.read(tmp.data(), size_t(csize))) rc = setError(VDBRT_ERR_BAD_GRID, "truncated NanoVDB grid payload");
            else {
                uLongf n = uLongf(pick->meta.gridSize);
                if (uncompress(static_cast<Bytef*>(p), &n, tmp.data(), uLong(csize)) != Z_OK) rc = setError(VDBRT_ERR_BAD_GRID, "Internal read error in ZIP");
                else if (uint64_t(n) != pick->meta.gridSize) rc = setError(VDBRT_ERR_BAD_GRID, "UNZIP failed on byte size");
            }
        }
    } else rc = setError(VDBRT_ERR_UNSUPPORTED, pick->meta.codec == CODEC_BLOSC ? "BLOSC compression codec was disabled during build" : "unknown compression codec");
    if (rc != VDBRT_OK) { std::free(p); return rc; }
    *buffer = p; *bytes = pick->meta.gridSize;
    return VDBRT_OK;
  });
}

int vdbrt_nvdb_write(const char* path, const void* buffer, uint64_t bytes, uint32_t codec)
{
  return guarded([&]() -> int {
    if (!path || !buffer) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    if (bytes < 736) return setError(VDBRT_ERR_BAD_GRID, "buffer smaller than GridData+TreeData");
    if (codec != CODEC_NONE && codec != CODEC_ZIP) return setError(VDBRT_ERR_UNSUPPORTED, "only the NONE and ZIP codecs are built in");
    const uint8_t* g = static_cast<const uint8_t*>(buffer);
    uint64_t magic, gridSize; uint32_t version;
    std::memcpy(&magic, g, 8); std::memcpy(&version, g + 16, 4); std::memcpy(&gridSize, g + 32, 8);
    if (magic != MAGIC_GRID && magic != MAGIC_NUMB) return setError(VDBRT_ERR_BAD_GRID, "not a NanoVDB grid (bad magic number)");
    if (gridSize > bytes) return setError(VDBRT_ERR_BAD_GRID, "grid size exceeds the buffer");
    // Segment::write + FileGridMetaData(size, codec, gridData) (IO.h:318-345, 391-399)
    FileHeader head = {MAGIC_FILE, version, 1, uint16_t(codec)};
    FileMetaData m;
    std::memset(&m, 0, sizeof(m));
    const char* name = reinterpret_cast<const char*>(g + 40);
    const uint8_t* tree = g + 672;
    m.gridSize = gridSize; m.fileSize = gridSize; m.nameKey = nameHash(name);
    std::memcpy(&m.voxelCount, tree + 56, 8);
    std::memcpy(&m.gridClass, g + 632, 4); std::memcpy(&m.gridType, g + 636, 4);
    std::memcpy(m.worldBBox, g + 560, 48);
    uint64_t rootOff; std::memcpy(&rootOff, tree + 24, 8);
    if (672 + rootOff + 24 <= bytes) std::memcpy(m.indexBBox, g + 672 + rootOff, 24);      // RootData::mBBox == GridData::indexBBox()
    std::memcpy(m.voxelSize, g + 608, 24);
    m.nameSize = uint32_t(std::strlen(name) + 1);
    std::memcpy(m.nodeCount, tree + 32, 12); m.nodeCount[3] = 1;
    std::memcpy(m.tileCount, tree + 44, 12);
    m.codec = uint16_t(codec);
    uint32_t blind; std::memcpy(&blind, g + 648, 4); m.blindDataCount = uint16_t(blind);
    m.version = version;
    std::vector<unsigned char> packed;
    if (codec == CODEC_ZIP) {
        uLongf n = compressBound(uLong(gridSize));
        packed.resize(n);
        if (compress(packed.data(), &n, g, uLong(gridSize)) != Z_OK) return setError(VDBRT_ERR_IO, "Internal write error in ZIP");
        packed.resize(n);
        m.fileSize = 8 + uint64_t(n);
    }
    File out(path, "wb");
    if (!out.f) return setError(VDBRT_ERR_IO, std::string("Unable to open file named \"") + path + "\" for output");
    bool ok = out.write(&head, sizeof(head)) && out.write(&m, sizeof(m)) && out.write(name, m.nameSize);
    if (codec == CODEC_ZIP) { const uint64_t n = packed.size(); ok = ok && out.write(&n, 8) && out.write(packed.data(), packed.size()); }
    else ok = ok && out.write(g, size_t(gridSize));
    if (!ok) return setError(VDBRT_ERR_IO, "Failed writing NanoVDB file");
    return VDBRT_OK;
  });
}

int vdbrt_buffer_free(void* buffer) { std::free(buffer); return VDBRT_OK; }

// tools::Film::savePPM (tools/RayTracer.h:319-335) with convertToBitBuffer<unsigned char>(alpha = false) (:300-317):
// every channel is static_cast<unsigned char>(255.0f * value)
int vdbrt_film_save_ppm(const char* fileName, const float* rgba, uint32_t width, uint32_t height)
{
  return guarded([&]() -> int {
    if (!fileName || !rgba) return setError(VDBRT_ERR_INVALID_ARG, "null argument");
    std::string name(fileName);
    if (name.find_last_of(".") == std::string::npos) name.append(".ppm");
    File out(name.c_str(), "wb");
    if (!out.f) return setError(VDBRT_ERR_IO, "Error opening PPM file \"" + name + "\"");
    const size_t n = size_t(width) * height;
    std::vector<unsigned char> buf(3 * n);
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) buf[3 * i + c] = static_cast<unsigned char>(255.0f * rgba[4 * i + c]);
    std::fprintf(out.f, "P6\n%u %u\n255\n", width, height);
    if (!out.write(buf.data(), buf.size())) return setError(VDBRT_ERR_IO, "Failed writing PPM file");
    return VDBRT_OK;
  });
}

} // extern "C"
