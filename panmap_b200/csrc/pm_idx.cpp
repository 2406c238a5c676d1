// pm_idx.cpp -- reader for panmap's `.idx` container (host side, no CUDA).
//
// Format (reference: /root/reference/src/index_single_mode.cpp:1561-1636, index_single_mode.hpp:24-38,
// main.cpp:193-236, schema src/index_lite.capnp):
//   32-byte header: u32 magic 0x31494D50 "PMI1", u32 version 1, i32 k,s,t,l, u8 hpc, u8 open, u8 uncompressed
//   payload: raw Cap'n Proto flat-array message (when `uncompressed`) whose root is LiteIndex.
// The message is walked with a small schema-less pointer decoder (struct / list / far pointers); the struct
// shapes used are LiteIndex (2 data words, 11 pointers), LiteTree (0,2) and LiteNode (1,1).
#include "pm_host.h"
#include "pm_capnp.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace pm {
namespace {

using namespace capnp_walk;
static const size_t kElemBytes[8] = {0, 0, 1, 2, 4, 8, 8, 0};

// zstd-framed payloads (the reference's default: independent 64 MB frames, index_single_mode.cpp:1615-1633,
// zstd_compression.cpp:31-60).  libzstd is a runtime dependency only: resolved with dlopen, no headers needed.
void inflateZstdFrames(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
    void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) throw Unsupported("zstd-compressed .idx needs libzstd.so.1 at run time (or write the index uncompressed)");
    auto decompress = reinterpret_cast<size_t (*)(void*, size_t, const void*, size_t)>(dlsym(h, "ZSTD_decompress"));
    auto contentSize = reinterpret_cast<unsigned long long (*)(const void*, size_t)>(dlsym(h, "ZSTD_getFrameContentSize"));
    auto frameSize = reinterpret_cast<size_t (*)(const void*, size_t)>(dlsym(h, "ZSTD_findFrameCompressedSize"));
    auto isError = reinterpret_cast<unsigned (*)(size_t)>(dlsym(h, "ZSTD_isError"));
    if (!decompress || !contentSize || !frameSize || !isError) throw Unsupported("libzstd lacks the frame API");
    size_t pos = 0;
    while (pos < n) {
        const size_t csz = frameSize(src + pos, n - pos);
        if (isError(csz)) throw std::runtime_error("Failed to decompress index: bad zstd frame");
        const unsigned long long dsz = contentSize(src + pos, csz);
        if (dsz >= 0xFFFFFFFFFFFFFFFEULL) throw std::runtime_error("Failed to decompress index: frame without content size");
        const size_t o = out.size();
        out.resize(o + static_cast<size_t>(dsz));
        const size_t got = decompress(out.data() + o, static_cast<size_t>(dsz), src + pos, csz);
        if (isError(got) || got != dsz) throw std::runtime_error("Failed to decompress index: zstd error");
        pos += csz;
    }
}

}  // namespace

void readIdxFile(const std::string& path, HostIndex& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw IoError("cannot open index file: " + path);
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.raw.resize(static_cast<size_t>(sz > 0 ? sz : 0));
    if (sz > 0 && std::fread(out.raw.data(), 1, static_cast<size_t>(sz), f) != static_cast<size_t>(sz)) { std::fclose(f); throw IoError("short read: " + path); }
    std::fclose(f);
    if (out.raw.size() < 40) throw std::runtime_error("index file too small: " + path);
    uint32_t magic, ver; std::memcpy(&magic, out.raw.data(), 4); std::memcpy(&ver, out.raw.data() + 4, 4);
    size_t payload = 0;
    std::vector<uint8_t> inflated;
    bool compressed = false;
    if (magic == 0x31494D50u && ver != 1) throw std::runtime_error("unsupported .idx header version " + std::to_string(ver) + ": " + path);
    if (magic == 0x31494D50u && ver == 1) { payload = 32; compressed = out.raw[26] == 0; }
    else if (magic == 0xFD2FB528u) compressed = true;   // header-less legacy file that starts with a zstd frame
    if (compressed) inflateZstdFrames(out.raw.data() + payload, out.raw.size() - payload, inflated);
    Msg m; m.base = compressed ? inflated.data() : out.raw.data() + payload;
    const size_t avail = compressed ? inflated.size() : out.raw.size() - payload;
    if (avail < 8) throw std::runtime_error("index message truncated: " + path);
    uint32_t nseg; std::memcpy(&nseg, m.base, 4); nseg += 1;
    if (nseg == 0 || nseg > 4096 || 4 + 4 * static_cast<size_t>(nseg) > avail) throw std::runtime_error("not a Cap'n Proto index message: " + path);
    size_t hdr = (4 + 4 * static_cast<size_t>(nseg) + 7) & ~size_t(7);
    size_t pos = hdr;
    for (uint32_t i = 0; i < nseg; ++i) {
        uint32_t w; std::memcpy(&w, m.base + 4 + 4 * i, 4);
        m.segStart.push_back(pos); m.segWords.push_back(w); pos += 8 * static_cast<size_t>(w);
    }
    if (pos > avail) throw std::runtime_error("index message truncated: " + path);

    const Ref root = resolve(m, 0, 0);
    if (root.kind != 1 || root.dataWords < 2) throw std::runtime_error("index root is not a LiteIndex struct");
    const uint64_t d0 = m.word(root.seg, root.off), d1 = m.word(root.seg, root.off + 1);
    out.sp.k = static_cast<int>(d0 & 0xffff); out.sp.s = static_cast<int>((d0 >> 16) & 0xffff);
    out.sp.t = static_cast<int>((d0 >> 32) & 0xffff); out.sp.l = static_cast<int>((d0 >> 48) & 0xffff);
    out.sp.open = static_cast<int>(d1 & 1); out.sp.hpc = static_cast<int>((d1 >> 1) & 1);
    const unsigned formatVersion = static_cast<unsigned>((d1 >> 16) & 0xffff);
    // same check and wording as placement.cpp:1013-1019
    if (formatVersion != 4)
        throw std::runtime_error("Index format version " + std::to_string(formatVersion) +
                                 " is incompatible with this panmap (expects 4). Rebuild the index (delete the .idx and rerun).");
    const Ref tree = ptrOf(m, root, 0), hashes = ptrOf(m, root, 1), parents = ptrOf(m, root, 2), childs = ptrOf(m, root, 3),
              offs = ptrOf(m, root, 4);
    if (hashes.kind != 2 || parents.kind != 2 || childs.kind != 2 || offs.kind != 2)
        throw std::runtime_error("Index missing required V3 fields (seedChangeHashes, etc). V2 is no longer supported.");
    const Ref nodes = ptrOf(m, tree, 0);
    const uint64_t N = nodes.kind == 2 ? nodes.count : 0;
    if (offs.count < N + 1)
        throw std::runtime_error("Struct-of-arrays format offsets size mismatch: " + std::to_string(offs.count) + " vs " + std::to_string(N + 1));
    if (offs.elemSize != 5) throw std::runtime_error("index: nodeChangeOffsets is not a list of 64-bit words");
    out.nodeOffsets.resize(N + 1);
    std::memcpy(out.nodeOffsets.data(), m.span(offs.seg, offs.off, 8 * (N + 1)), 8 * (N + 1));
    const uint64_t D = out.nodeOffsets[N];
    {   // D sizes the allocations below: it cannot exceed what the seed-change lists hold
        uint64_t have = 0;
        for (uint64_t sgi = 0; sgi < hashes.count; ++sgi) { const Ref inner = resolve(m, hashes.seg, hashes.off + sgi); if (inner.kind == 2) have += inner.count; }
        if (D > have) throw std::runtime_error("index: nodeChangeOffsets names more seed changes than the file holds");
    }
    out.hash.resize(D); out.parentCount.resize(D); out.childCount.resize(D);
    auto gatherSegs = [&](const Ref& outer, void* dst, size_t elemBytes, unsigned wantCode) {
        uint64_t done = 0;
        for (uint64_t sgi = 0; sgi < outer.count; ++sgi) {  // 5e8-element segments (placement.cpp:1052-1071)
            const Ref inner = resolve(m, outer.seg, outer.off + sgi);
            if (inner.kind != 2) continue;
            if (inner.elemSize != wantCode) throw std::runtime_error("index: unexpected list element size");
            const uint64_t n = inner.count < D - done ? inner.count : D - done;
            std::memcpy(static_cast<uint8_t*>(dst) + done * elemBytes, m.span(inner.seg, inner.off, n * elemBytes), n * elemBytes);
            done += n;
        }
        if (done != D) throw std::runtime_error("index: seed-change arrays shorter than nodeChangeOffsets says");
    };
    (void)kElemBytes;
    gatherSegs(hashes, out.hash.data(), 8, 5);
    gatherSegs(parents, out.parentCount.data(), 2, 3);
    gatherSegs(childs, out.childCount.data(), 2, 3);
    out.parentIndex.resize(N); out.nodeIds.resize(N); out.identicalToParent.assign(N, 0);
    for (uint64_t i = 0; i < N; ++i) {
        const size_t eo = nodes.off + i * (nodes.dataWords + nodes.ptrWords);
        out.parentIndex[i] = nodes.dataWords ? static_cast<uint32_t>(m.word(nodes.seg, eo) & 0xffffffffu) : 0;
        out.identicalToParent[i] = nodes.dataWords ? static_cast<uint8_t>((m.word(nodes.seg, eo) >> 32) & 1u) : 0;
        Ref e; e.kind = 1; e.seg = nodes.seg; e.off = eo; e.dataWords = nodes.dataWords; e.ptrWords = nodes.ptrWords;
        const Ref id = ptrOf(m, e, 0);
        if (id.kind == 2 && id.elemSize == 2 && id.count > 0) out.nodeIds[i].assign(reinterpret_cast<const char*>(m.span(id.seg, id.off, id.count)), id.count - 1);
        if (i > 0 && out.parentIndex[i] >= i) throw std::runtime_error("index: nodes are not in DFS pre-order (parentIndex >= index)");
    }
    out.blockRanges.clear(); out.substitutionMatrix.clear();
    const Ref ranges = ptrOf(m, tree, 1);
    if (ranges.kind == 2 && ranges.elemSize == 7 && ranges.dataWords >= 1) {      // List(BlockRange): two u32 in one data word
        for (uint64_t i = 0; i < ranges.count; ++i) {
            const uint64_t w = m.word(ranges.seg, ranges.off + i * (ranges.dataWords + ranges.ptrWords));
            out.blockRanges.push_back(static_cast<uint32_t>(w & 0xffffffffu)); out.blockRanges.push_back(static_cast<uint32_t>(w >> 32));
        }
    } else if (ranges.kind == 2 && ranges.elemSize == 5) {                       // the same list in its 8-byte-element encoding
        for (uint64_t i = 0; i < ranges.count; ++i) {
            const uint64_t w = m.word(ranges.seg, ranges.off + i);
            out.blockRanges.push_back(static_cast<uint32_t>(w & 0xffffffffu)); out.blockRanges.push_back(static_cast<uint32_t>(w >> 32));
        }
    }
    const Ref sub = ptrOf(m, root, 10);
    if (sub.kind == 2 && sub.elemSize == 5 && sub.count > 0) {
        out.substitutionMatrix.resize(sub.count);
        std::memcpy(out.substitutionMatrix.data(), m.span(sub.seg, sub.off, 8 * sub.count), 8 * sub.count);
    }
    out.raw.clear(); out.raw.shrink_to_fit();
}

// ---- writer ----------------------------------------------------------------------------------------------------------------------
// One Cap'n Proto segment laid out front to back: root pointer, LiteIndex (2 data words, 11 pointers), LiteTree (2 pointers), the
// LiteNode list (1 data word + 1 pointer each, behind its tag word), the id texts, block ranges, nodeChangeOffsets, the three outer
// seed-change lists and their inner lists (at most 5e8 elements each: the reference's SEED_CHANGE_SEGMENT, placement.cpp:1052-1071),
// the substitution matrix.  Pointers are intra-segment (30-bit signed word offsets), so a message is limited to 4 GB here -- ~2.8e8
// seed changes; the reference splits larger messages over several segments.
namespace {
struct Out {
    std::vector<uint64_t> w;
    size_t alloc(size_t words) { const size_t at = w.size(); w.resize(at + words, 0); return at; }
    void structPtr(size_t at, size_t target, unsigned dataWords, unsigned ptrWords) {
        const int64_t off = static_cast<int64_t>(target) - static_cast<int64_t>(at) - 1;
        w[at] = (static_cast<uint64_t>(static_cast<uint32_t>(off << 2))) | (static_cast<uint64_t>(dataWords) << 32) | (static_cast<uint64_t>(ptrWords) << 48);
    }
    void listPtr(size_t at, size_t target, unsigned elemCode, uint64_t count) {
        const int64_t off = static_cast<int64_t>(target) - static_cast<int64_t>(at) - 1;
        if (off >= (1ll << 29) || off < -(1ll << 29) || count >= (1ull << 29)) throw Unsupported("index too large for a single-segment .idx message (4 GB)");
        w[at] = (static_cast<uint64_t>(static_cast<uint32_t>(off << 2)) | 1u) | (static_cast<uint64_t>(elemCode) << 32) | (count << 35);
    }
    template <class T> size_t dataList(size_t ptrAt, const T* src, uint64_t n, unsigned elemCode) {
        const size_t at = alloc((n * sizeof(T) + 7) / 8);
        if (n) std::memcpy(reinterpret_cast<uint8_t*>(w.data() + at), src, n * sizeof(T));
        listPtr(ptrAt, at, elemCode, n);
        return at;
    }
};
constexpr uint64_t kSeedChangeSegment = 500000000ull;

void zstdFrames(const uint8_t* src, size_t n, int level, FILE* f, const std::string& path) {
    void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) throw Unsupported("a zstd-compressed .idx needs libzstd.so.1 at run time (or write it uncompressed)");
    auto compress = reinterpret_cast<size_t (*)(void*, size_t, const void*, size_t, int)>(dlsym(h, "ZSTD_compress"));
    auto bound = reinterpret_cast<size_t (*)(size_t)>(dlsym(h, "ZSTD_compressBound"));
    auto isError = reinterpret_cast<unsigned (*)(size_t)>(dlsym(h, "ZSTD_isError"));
    if (!compress || !bound || !isError) throw Unsupported("libzstd lacks the simple API");
    constexpr size_t kFrame = 64ull << 20;   // independent frames, so that a reader may inflate them in parallel
    std::vector<uint8_t> buf(bound(kFrame < n ? kFrame : n));
    for (size_t pos = 0; pos < n || pos == 0; pos += kFrame) {
        const size_t len = n - pos < kFrame ? n - pos : kFrame;
        const size_t got = compress(buf.data(), buf.size(), src + pos, len, level);
        if (isError(got)) throw std::runtime_error("zstd compression failed");
        if (std::fwrite(buf.data(), 1, got, f) != got) throw IoError("short write: " + path);
        if (n == 0) break;
    }
}
}  // namespace

uint64_t writeIdxFile(const std::string& path, const pm_index_desc& d, const IdxExtras& x, int zstdLevel) {
    const uint64_t N = d.n_nodes, D = d.n_deltas;
    if (!d.node_offsets || !d.parent_index || (D && (!d.delta_hash || !d.delta_parent || !d.delta_child))) throw std::runtime_error("null index arrays");
    if (N == 0 || d.node_offsets[N] != D) throw std::runtime_error("node_offsets[n_nodes] != n_deltas");
    Out o;
    o.w.reserve(16 + 4 * N + D + D / 2 + N + 64);
    o.alloc(1);                                   // root pointer
    const size_t root = o.alloc(2 + 11);
    o.structPtr(0, root, 2, 11);
    const pm_seed_params& sp = d.seed;
    o.w[root] = (static_cast<uint64_t>(sp.k) & 0xffff) | ((static_cast<uint64_t>(sp.s) & 0xffff) << 16) | ((static_cast<uint64_t>(sp.t) & 0xffff) << 32) |
                ((static_cast<uint64_t>(sp.l) & 0xffff) << 48);
    o.w[root + 1] = (sp.open ? 1u : 0u) | (sp.hpc ? 2u : 0u) | (4ull << 16);   // formatVersion 4 (panmap_utils.hpp:27)
    const size_t rp = root + 2;
    // LiteTree
    const size_t tree = o.alloc(2);
    o.structPtr(rp + 0, tree, 0, 2);
    {   // liteNodes: inline-composite list
        const size_t tag = o.alloc(1 + 2 * N);
        if (2 * N >= (1ull << 29)) throw Unsupported("too many nodes for one .idx node list");
        const int64_t off = static_cast<int64_t>(tag) - static_cast<int64_t>(tree) - 1;
        o.w[tree] = (static_cast<uint64_t>(static_cast<uint32_t>(off << 2)) | 1u) | (7ull << 32) | ((2 * N) << 35);
        o.w[tag] = (static_cast<uint64_t>(static_cast<uint32_t>(N << 2))) | (1ull << 32) | (1ull << 48);
        std::string tmp;
        for (uint64_t i = 0; i < N; ++i) {
            const size_t e = tag + 1 + 2 * i;
            o.w[e] = static_cast<uint64_t>(i ? d.parent_index[i] : 0) | ((x.identicalToParent && x.identicalToParent[i]) ? (1ull << 32) : 0);
            const char* id = x.nodeIds ? x.nodeIds[i] : nullptr;
            if (!id) { tmp = "node_" + std::to_string(i); id = tmp.c_str(); }
            const size_t len = std::strlen(id) + 1;           // Text: NUL-terminated byte list
            const size_t at = o.alloc((len + 7) / 8);
            std::memcpy(reinterpret_cast<uint8_t*>(o.w.data() + at), id, len - 1);
            o.listPtr(e + 1, at, 2, len);
        }
    }
    if (x.blockRanges && x.nBlocks) {   // List(BlockRange): structs of one data word, written as an inline-composite list
        const size_t tag = o.alloc(1 + x.nBlocks);
        const int64_t off = static_cast<int64_t>(tag) - static_cast<int64_t>(tree + 1) - 1;
        o.w[tree + 1] = (static_cast<uint64_t>(static_cast<uint32_t>(off << 2)) | 1u) | (7ull << 32) | (x.nBlocks << 35);
        o.w[tag] = (static_cast<uint64_t>(static_cast<uint32_t>(x.nBlocks << 2))) | (1ull << 32);
        for (uint64_t i = 0; i < x.nBlocks; ++i) o.w[tag + 1 + i] = static_cast<uint64_t>(x.blockRanges[2 * i]) | (static_cast<uint64_t>(x.blockRanges[2 * i + 1]) << 32);
    }
    o.dataList(rp + 4, d.node_offsets, N + 1, 5);
    const uint64_t nSeg = D ? (D + kSeedChangeSegment - 1) / kSeedChangeSegment : 1;
    auto outer = [&](size_t ptrAt, const void* src, size_t elemBytes, unsigned code) {
        const size_t at = o.alloc(nSeg);
        o.listPtr(ptrAt, at, 6, nSeg);    // list of pointers
        for (uint64_t sgi = 0; sgi < nSeg; ++sgi) {
            const uint64_t b = sgi * kSeedChangeSegment, n = D - b < kSeedChangeSegment ? D - b : kSeedChangeSegment;
            const size_t body = o.alloc((n * elemBytes + 7) / 8);
            if (n) std::memcpy(reinterpret_cast<uint8_t*>(o.w.data() + body), static_cast<const uint8_t*>(src) + b * elemBytes, n * elemBytes);
            o.listPtr(at + sgi, body, code, n);
        }
    };
    outer(rp + 1, d.delta_hash, 8, 5);
    outer(rp + 2, d.delta_parent, 2, 3);
    outer(rp + 3, d.delta_child, 2, 3);
    if (x.substitutionMatrix) o.dataList(rp + 10, x.substitutionMatrix, 16, 5);
    if (o.w.size() >= (1ull << 32)) throw Unsupported("index too large for a single-segment .idx message");

    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw IoError("cannot create index file: " + path);
    uint8_t hdr[32] = {0};
    auto put32 = [&](size_t at, uint32_t v) { std::memcpy(hdr + at, &v, 4); };
    put32(0, 0x31494D50u); put32(4, 1u); put32(8, static_cast<uint32_t>(sp.k)); put32(12, static_cast<uint32_t>(sp.s));
    put32(16, static_cast<uint32_t>(sp.t)); put32(20, static_cast<uint32_t>(sp.l));
    hdr[24] = sp.hpc ? 1 : 0; hdr[25] = sp.open ? 1 : 0; hdr[26] = zstdLevel < 0 ? 1 : 0;
    uint32_t segTable[2] = {0u, static_cast<uint32_t>(o.w.size())};   // one segment: (count - 1, size in words)
    try {
        if (std::fwrite(hdr, 1, 32, f) != 32) throw IoError("short write: " + path);
        if (zstdLevel < 0) {
            if (std::fwrite(segTable, 1, 8, f) != 8 || std::fwrite(o.w.data(), 8, o.w.size(), f) != o.w.size()) throw IoError("short write: " + path);
        } else {
            std::vector<uint8_t> flat(8 + 8 * o.w.size());
            std::memcpy(flat.data(), segTable, 8); std::memcpy(flat.data() + 8, o.w.data(), 8 * o.w.size());
            std::vector<uint64_t>().swap(o.w);
            zstdFrames(flat.data(), flat.size(), zstdLevel, f, path);
        }
    } catch (...) { std::fclose(f); std::remove(path.c_str()); throw; }
    const long total = std::ftell(f);
    if (std::fclose(f) != 0) throw IoError("cannot finish index file: " + path);
    return static_cast<uint64_t>(total);
}

}  // namespace pm
