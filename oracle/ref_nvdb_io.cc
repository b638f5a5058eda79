// TEST INFRASTRUCTURE: the reference's own NanoVDB file reader / writer (nanovdb/io/IO.h, header-only, compiled from
// /root/reference where it lies) behind a two-verb command line, used to pin openvdb_b200/csrc/vdbrt_io.cc:
//   ref_nvdb_io write OUT.nvdb CODEC RAW_GRID [RAW_GRID ...]   nanovdb::io::writeGrid(s) of raw grid buffers (CODEC: none | zip)
//   ref_nvdb_io read IN.nvdb NAME RAW_OUT                      nanovdb::io::readGrid by name ("-" = first grid), dumps the grid bytes
#include <nanovdb/NanoVDB.h>
#include <nanovdb/io/IO.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

static nanovdb::GridHandle<nanovdb::HostBuffer> loadRaw(const char* path)
{
    std::ifstream is(path, std::ios::binary | std::ios::ate);
    if (!is) throw std::runtime_error(std::string("cannot open ") + path);
    const size_t n = size_t(is.tellg());
    is.seekg(0);
    auto buffer = nanovdb::HostBuffer::create(n);
    is.read(reinterpret_cast<char*>(buffer.data()), n);
    return nanovdb::GridHandle<nanovdb::HostBuffer>(std::move(buffer));
}

int main(int argc, char** argv)
{
    try {
        if (argc >= 5 && !std::strcmp(argv[1], "write")) {
            const nanovdb::io::Codec codec = !std::strcmp(argv[3], "zip") ? nanovdb::io::Codec::ZIP : nanovdb::io::Codec::NONE;
            std::vector<nanovdb::GridHandle<nanovdb::HostBuffer>> handles;
            for (int i = 4; i < argc; ++i) handles.push_back(loadRaw(argv[i]));
            if (handles.size() == 1) nanovdb::io::writeGrid(argv[2], handles[0], codec);
            else nanovdb::io::writeGrids(argv[2], handles, codec);
            return 0;
        }
        if (argc == 5 && !std::strcmp(argv[1], "read")) {
            auto h = !std::strcmp(argv[3], "-") ? nanovdb::io::readGrid(argv[2]) : nanovdb::io::readGrid(argv[2], std::string(argv[3]));
            if (!h) throw std::runtime_error("grid not found");
            std::ofstream os(argv[4], std::ios::binary);
            os.write(reinterpret_cast<const char*>(h.data()), std::streamsize(h.size()));
            return 0;
        }
        std::cerr << "usage: ref_nvdb_io write OUT.nvdb none|zip RAW... | read IN.nvdb NAME|- RAW_OUT\n";
        return 2;
    } catch (const std::exception& e) {
        std::cerr << "ref_nvdb_io: " << e.what() << "\n";
        return 1;
    }
}
