// ORACLE SCAFFOLDING (test infrastructure, not product code).
// Minimal stand-in for the reference's mgsr.hpp: the placement/seeding translation units only
// use mgsr::FastqFile (a FILE* over a plain file or a `gzip -dc` pipe; reference mgsr.hpp:153-186).
// The real header drags in Eigen/TBB/panman, none of which exist in this image.
#pragma once
#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>
namespace mgsr {
inline bool isGzipped(const std::string& p) {
    if (p.size() >= 3 && p.compare(p.size() - 3, 3, ".gz") == 0) return true;
    unsigned char b[2] = {0, 0};
    std::ifstream f(p, std::ios::binary);
    if (f.read(reinterpret_cast<char*>(b), 2)) return b[0] == 0x1f && b[1] == 0x8b;
    return false;
}
struct FastqFile {
    FILE* fp = nullptr;
    bool piped = false;
    explicit FastqFile(const std::string& path) {
        if (isGzipped(path)) { fp = popen(("gzip -dc '" + path + "'").c_str(), "r"); piped = true; }
        else fp = fopen(path.c_str(), "r");
        if (!fp) throw std::runtime_error("Failed to open FASTQ file: " + path);
    }
    ~FastqFile() { if (fp) { if (piped) pclose(fp); else fclose(fp); } }
    FastqFile(const FastqFile&) = delete;
    FastqFile& operator=(const FastqFile&) = delete;
};
}  // namespace mgsr
