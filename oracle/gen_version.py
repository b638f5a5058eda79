#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle/): generate openvdb/version.h for the reference build.

Reads the reference's own template (openvdb/openvdb/version.h.in) and performs the
substitutions its CMake would (version 13.0.1, ABI 13; only OPENVDB_USE_ZLIB defined;
no explicit instantiation), writing <out>/openvdb/version.h.  Nothing is copied into
the repository: the output lives under oracle/_ref/ (git-ignored).
"""
import re, sys, os

def main(ref_root, out_dir):
    src = open(os.path.join(ref_root, "openvdb/openvdb/version.h.in")).read()
    major, minor, patch, abi = 13, 0, 1, 13
    subs = {
        "OpenVDB_MAJOR_VERSION": str(major), "OpenVDB_MINOR_VERSION": str(minor),
        "OpenVDB_PATCH_VERSION": str(patch), "OPENVDB_ABI_VERSION_NUMBER": str(abi),
        "OPENVDB_PACKED_VERSION": "0x%02x%02x%04x" % (major, minor, patch),
        "OPENVDB_NAMESPACE_SUFFIX": "", "OPENVDB_X86_INSTRSET": "0",
    }
    src = re.sub(r"\$\{(\w+)\}", lambda m: subs.get(m.group(1), ""), src)
    src = re.sub(r"@(\w+)@", "", src)
    defined = {"OPENVDB_USE_ZLIB"}
    def cmakedefine(m):
        name = m.group(1)
        return ("#define %s" % name) if name in defined else ("/* #undef %s */" % name)
    src = re.sub(r"#cmakedefine\s+(\w+).*", cmakedefine, src)
    d = os.path.join(out_dir, "openvdb")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "version.h"), "w").write(src)

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
