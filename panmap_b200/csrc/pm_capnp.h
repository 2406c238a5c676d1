// pm_capnp.h -- schema-less walk of a Cap'n Proto flat-array message (struct / list / far pointers, composite lists): shared by the
// `.idx` reader (pm_idx.cpp) and the `.panman` reader (pm_panman.cpp).  Host side only.  Every access is bounds-checked against the
// segment table; a corrupt or truncated file ends in std::runtime_error, never in an out-of-range read.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace pm {
namespace capnp_walk {

struct Msg {
    const uint8_t* base = nullptr;
    std::vector<size_t> segStart, segWords;
    uint64_t word(uint32_t seg, size_t w) const {
        if (seg >= segStart.size() || w >= segWords[seg]) throw std::runtime_error("index: pointer out of range");
        uint64_t v; std::memcpy(&v, base + segStart[seg] + 8 * w, 8); return v;
    }
    // checked view of `bytes` bytes that start at word w of segment seg (every list body goes through here)
    const uint8_t* span(uint32_t seg, size_t w, uint64_t bytes) const {
        if (seg >= segStart.size() || w > segWords[seg] || bytes > 8ull * (segWords[seg] - w)) throw std::runtime_error("index: list extends past its segment (truncated or corrupt file)");
        return base + segStart[seg] + 8 * w;
    }
};
struct Ref { int kind = 0; uint32_t seg = 0; size_t off = 0; uint32_t dataWords = 0, ptrWords = 0, elemSize = 0; uint64_t count = 0; };

inline Ref decode(const Msg& m, uint64_t p, uint32_t seg, size_t base) {
    Ref r; r.seg = seg;
    const int64_t off = static_cast<int32_t>(p & 0xffffffffu) >> 2;
    r.off = static_cast<size_t>(static_cast<int64_t>(base) + off);
    if ((p & 3) == 0) { r.kind = 1; r.dataWords = (p >> 32) & 0xffff; r.ptrWords = (p >> 48) & 0xffff; }
    else {
        r.kind = 2; r.elemSize = (p >> 32) & 7; r.count = p >> 35;
        if (r.elemSize == 7) {
            const uint64_t tag = m.word(seg, r.off);
            r.count = static_cast<uint32_t>(tag & 0xffffffffu) >> 2;
            r.dataWords = (tag >> 32) & 0xffff; r.ptrWords = (tag >> 48) & 0xffff;
            r.off += 1;
        }
    }
    return r;
}
inline Ref resolve(const Msg& m, uint32_t seg, size_t w) {
    const uint64_t p = m.word(seg, w);
    if (p == 0) return Ref{};
    if ((p & 3) == 2) {
        const bool dbl = (p >> 2) & 1;
        const size_t padOff = (p & 0xffffffffu) >> 3;
        const uint32_t padSeg = static_cast<uint32_t>(p >> 32);
        if (!dbl) { const uint64_t q = m.word(padSeg, padOff); return q ? decode(m, q, padSeg, padOff + 1) : Ref{}; }
        const uint64_t far2 = m.word(padSeg, padOff), tag = m.word(padSeg, padOff + 1);
        return decode(m, tag & 0xFFFFFFFF00000003ULL, static_cast<uint32_t>(far2 >> 32), (far2 & 0xffffffffu) >> 3);
    }
    if ((p & 3) == 3) throw std::runtime_error("index: unexpected capability pointer");
    return decode(m, p, seg, w + 1);
}
inline Ref ptrOf(const Msg& m, const Ref& s, uint32_t i) { return (s.kind == 1 && i < s.ptrWords) ? resolve(m, s.seg, s.off + s.dataWords + i) : Ref{}; }
// element i of a composite (struct) list
inline Ref elemOf(const Ref& l, uint64_t i) {
    Ref e; e.kind = 1; e.seg = l.seg; e.off = l.off + (size_t)i * (l.dataWords + l.ptrWords); e.dataWords = l.dataWords; e.ptrWords = l.ptrWords; return e;
}
inline uint64_t dataOf(const Msg& m, const Ref& s, uint32_t i) { return (s.kind == 1 && i < s.dataWords) ? m.word(s.seg, s.off + i) : 0; }
// segment table of a flat-array message that starts at `base` and is `avail` bytes long
inline void openMessage(Msg& m, const uint8_t* base, size_t avail, const char* what) {
    m.base = base; m.segStart.clear(); m.segWords.clear();
    if (avail < 8) throw std::runtime_error(std::string(what) + ": message truncated");
    uint32_t nseg; std::memcpy(&nseg, base, 4); nseg += 1;
    if (nseg == 0 || nseg > 4096 || 4 + 4 * static_cast<size_t>(nseg) > avail) throw std::runtime_error(std::string(what) + ": not a Cap'n Proto message");
    size_t pos = (4 + 4 * static_cast<size_t>(nseg) + 7) & ~size_t(7);
    for (uint32_t i = 0; i < nseg; ++i) {
        uint32_t w; std::memcpy(&w, base + 4 + 4 * i, 4);
        m.segStart.push_back(pos); m.segWords.push_back(w); pos += 8 * static_cast<size_t>(w);
    }
    if (pos > avail) throw std::runtime_error(std::string(what) + ": message truncated");
}

}  // namespace capnp_walk
}  // namespace pm
