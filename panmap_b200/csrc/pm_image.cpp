// pm_image.cpp -- the flattened index (FlatIndex, pm_host.h) as a file: the "cached device image" of SURVEY.md section 8(f2).
//
// The reference caches its index as the `.idx` file and re-validates it on every run (main.cpp:371-396: reusable only if it is not
// older than its source and was built with the same parameters).  Here the `.idx` itself is the source and what is cached is the
// result of flattenIndex() -- packed delta words, segment masks, seed dictionary, tile chains, BFS order -- i.e. exactly the arrays
// pm_index_create uploads, so that opening an index is one file read plus the host-to-device copies (host side only, no CUDA).
//
// File: 8-byte magic, u32 format version, u32 sizeof(FlatIndex) of the writer, the stamp of the source (size, mtime, its 32 header
// bytes, shard / n_shards), then one section per field in the fixed order of visit() below -- u64 byte count + the bytes, padded to
// 8 -- and a 64-bit checksum of everything before it.  A reader accepts an image only if magic, version, struct size, stamp and
// checksum all match; anything else is a miss and the caller re-flattens.
#include "pm_host.h"

#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace pm {
namespace {

constexpr uint64_t kMagic = 0x0154414C464D50ull;   // "PMFLAT\x01"
constexpr uint32_t kImageVersion = 2;   // 2: four-lane checksum

inline uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v; h *= 0x9E3779B97F4A7C15ull; return h ^ (h >> 29);
}
// four independent lanes over 32-byte strides (one multiply chain per lane: the serial chain of a single lane ran at ~1.5 GB/s, a third of
// the time of opening a cached image), folded at the end together with the length; the tail goes through lane 0
uint64_t checksum(const uint8_t* p, size_t n, uint64_t h) {
    uint64_t a = h, b = h ^ 0x9E3779B97F4A7C15ull, c = h ^ 0xC2B2AE3D27D4EB4Full, d = h ^ 0x165667B19E3779F9ull;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        uint64_t v[4]; std::memcpy(v, p + i, 32);
        a = mix(a, v[0]); b = mix(b, v[1]); c = mix(c, v[2]); d = mix(d, v[3]);
    }
    for (; i + 8 <= n; i += 8) { uint64_t v; std::memcpy(&v, p + i, 8); a = mix(a, v); }
    if (i < n) { uint64_t v = 0; std::memcpy(&v, p + i, n - i); a = mix(a, v); }
    return mix(mix(mix(mix(a, b), c), d), (uint64_t)n);
}

struct Sink {   // serialises into one buffer
    std::vector<uint8_t> b;
    void raw(const void* p, size_t n) { const size_t at = b.size(); b.resize(at + ((n + 7) & ~size_t(7)), 0); if (n) std::memcpy(b.data() + at, p, n); }
    template <class T> void scalar(T& v) { raw(&v, sizeof(T)); }
    template <class T> void vec(std::vector<T>& v) { uint64_t n = v.size() * sizeof(T); raw(&n, 8); raw(v.data(), (size_t)n); }
};
struct Source {   // reads the same sequence back, bounds-checked
    const uint8_t* p; size_t n, at = 0;
    void need(size_t k) { if (k > n - at) throw std::runtime_error("index image truncated"); }
    void raw(void* dst, size_t k) { const size_t padded = (k + 7) & ~size_t(7); need(padded); if (k) std::memcpy(dst, p + at, k); at += padded; }
    template <class T> void scalar(T& v) { raw(&v, sizeof(T)); }
    template <class T> void vec(std::vector<T>& v) {
        uint64_t bytes = 0; raw(&bytes, 8);
        if (bytes % sizeof(T) != 0) throw std::runtime_error("index image: section size is not a multiple of its element size");
        need((size_t)bytes);
        v.resize((size_t)(bytes / sizeof(T)));
        raw(v.data(), (size_t)bytes);
    }
};

// every field of FlatIndex, in file order.  A field added to the struct changes sizeof(FlatIndex) and with it the image header, so
// stale images are rejected; the assert below is the reminder to list the field here.
template <class IO> void visit(FlatIndex& F, IO& io) {
    io.scalar(F.N); io.scalar(F.D); io.scalar(F.S); io.scalar(F.sp);
    io.scalar(F.nodeBegin); io.scalar(F.nodeEnd); io.scalar(F.nAnc); io.scalar(F.nLocal); io.scalar(F.nLocalDeltas);
    io.vec(F.parent); io.vec(F.depth); io.vec(F.subEnd); io.vec(F.bfsRank); io.vec(F.isLeaf);
    io.vec(F.gMagSq); io.vec(F.gMag); io.vec(F.gUnique);
    io.vec(F.dictHash); io.vec(F.dictKeys); io.vec(F.dictVals); io.scalar(F.dictMask);
    io.vec(F.lNode); io.vec(F.dw); io.vec(F.endMask);
    io.scalar(F.nFast); io.scalar(F.nDeltaChunks); io.scalar(F.nSeg);
    io.vec(F.chunkSeg); io.vec(F.nodeSeg); io.vec(F.boundarySegs);
    io.vec(F.genSlot); io.vec(F.genId); io.vec(F.genPc); io.scalar(F.nGenNodes);
    io.vec(F.evSlot); io.vec(F.evIdx);
    io.vec(F.rootId); io.vec(F.rootChild);
    io.vec(F.carrySlot); io.vec(F.chainOff); io.vec(F.chainNodes); io.scalar(F.nK2Tiles);
    io.vec(F.bfsNodes); io.vec(F.bfsRanks);
    io.raw(F.homo, sizeof(F.homo));
}
static_assert(sizeof(FlatIndex) == 848, "FlatIndex changed: list the new field in visit() and bump kImageVersion");

struct Header {
    uint64_t magic; uint32_t version, structBytes;
    ImageStamp stamp;
};

}  // namespace

ImageStamp stampOfFile(const std::string& path, uint32_t shard, uint32_t nShards) {
    ImageStamp s{};
    struct stat st;
    if (::stat(path.c_str(), &st) != 0) throw IoError("cannot stat index file: " + path);
    s.srcSize = (uint64_t)st.st_size;
    s.srcMtimeNs = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw IoError("cannot open index file: " + path);
    const size_t got = std::fread(s.srcHeader, 1, sizeof(s.srcHeader), f);
    std::fclose(f);
    if (got < sizeof(s.srcHeader)) std::memset(s.srcHeader + got, 0, sizeof(s.srcHeader) - got);
    s.shard = shard; s.nShards = nShards;
    return s;
}

uint64_t writeFlatImage(const std::string& path, FlatIndex& F, const std::vector<std::string>& nodeIds, const ImageStamp& stamp) {
    Sink out;
    Header h{kMagic, kImageVersion, (uint32_t)sizeof(FlatIndex), stamp};
    out.raw(&h, sizeof(h));
    visit(F, out);
    {   // node ids: count, then the lengths, then the characters back to back
        uint64_t n = nodeIds.size(); out.raw(&n, 8);
        std::vector<uint32_t> len(nodeIds.size()); size_t total = 0;
        for (size_t i = 0; i < nodeIds.size(); ++i) { len[i] = (uint32_t)nodeIds[i].size(); total += len[i]; }
        out.vec(len);
        std::vector<char> chars; chars.reserve(total);
        for (const auto& s : nodeIds) chars.insert(chars.end(), s.begin(), s.end());
        out.vec(chars);
    }
    uint64_t sum = checksum(out.b.data(), out.b.size(), 0x504D464C4154ull);
    out.raw(&sum, 8);
    // written beside the target and renamed into place: a reader never sees a half-written image, concurrent writers do not interleave
    const std::string tmp = path + ".tmp." + std::to_string((long)::getpid());
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) throw IoError("cannot create index image: " + tmp);
    const bool ok = std::fwrite(out.b.data(), 1, out.b.size(), f) == out.b.size();
    if (std::fclose(f) != 0 || !ok) { std::remove(tmp.c_str()); throw IoError("short write: " + tmp); }
    if (std::rename(tmp.c_str(), path.c_str()) != 0) { std::remove(tmp.c_str()); throw IoError("cannot move index image into place: " + path); }
    return out.b.size();
}

bool readFlatImage(const std::string& path, FlatIndex& F, std::vector<std::string>& nodeIds, const ImageStamp* expect, std::string* why) {
    auto miss = [&](const char* w) { if (why) *why = w; return false; };
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return miss("no image file");
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (sz < (long)(sizeof(Header) + 8)) { std::fclose(f); return miss("image too small"); }
    Header h;
    if (std::fread(&h, 1, sizeof(h), f) != sizeof(h)) { std::fclose(f); return miss("short read"); }
    if (h.magic != kMagic || h.version != kImageVersion || h.structBytes != sizeof(FlatIndex)) { std::fclose(f); return miss("image written by another format version"); }
    if (expect && std::memcmp(&h.stamp, expect, sizeof(ImageStamp)) != 0) { std::fclose(f); return miss("source index changed since the image was written (size / mtime / header / shard)"); }
    std::vector<uint8_t> buf((size_t)sz);
    std::memcpy(buf.data(), &h, sizeof(h));
    const bool ok = std::fread(buf.data() + sizeof(h), 1, buf.size() - sizeof(h), f) == buf.size() - sizeof(h);
    std::fclose(f);
    if (!ok) return miss("short read");
    uint64_t want; std::memcpy(&want, buf.data() + buf.size() - 8, 8);
    if (checksum(buf.data(), buf.size() - 8, 0x504D464C4154ull) != want) return miss("image checksum mismatch (truncated or corrupt)");
    try {
        Source in{buf.data(), buf.size() - 8};
        Header again; in.raw(&again, sizeof(again));
        visit(F, in);
        uint64_t n = 0; in.raw(&n, 8);
        std::vector<uint32_t> len; in.vec(len);
        std::vector<char> chars; in.vec(chars);
        if (len.size() != n) return miss("image node-id table inconsistent");
        nodeIds.resize((size_t)n);
        size_t at = 0;
        for (size_t i = 0; i < nodeIds.size(); ++i) {
            if (len[i] > chars.size() - at) return miss("image node-id table inconsistent");
            nodeIds[i].assign(chars.data() + at, len[i]); at += len[i];
        }
    } catch (const std::exception& e) { if (why) *why = e.what(); return false; }
    if (F.parent.size() != F.N || F.nodeEnd > F.N || F.nodeBegin > F.nodeEnd) return miss("image arrays inconsistent");
    return true;
}

}  // namespace pm
