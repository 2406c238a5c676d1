// ORACLE SCAFFOLDING (test infrastructure, not product code).
// Schema-less reader for `.panman` files: an xz stream holding an unpacked Cap'n Proto stream
// message (TreeGroup -> trees[0] -> newick / nodes / consensusSeqMap / gaps).  panman's own schema
// (panman.capnp, TurakhiaLab/panman v0.1.4) is not in this image; the struct layout below is the one
// documented in SURVEY.md Appendix E and is validated in tests by reconstructing the reference's
// fixture genomes (src/test/data/MZ515733.1.fa etc.).  It fills the stub panmanUtils::Tree that the
// reference's unmodified IndexBuilder then consumes.
#pragma once
#include "panmanUtils.hpp"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
namespace panman_loader {

struct Msg {
    std::vector<uint8_t> bytes;
    std::vector<size_t> segStart;  // byte offset of each segment
    std::vector<size_t> segWords;
    uint64_t word(uint32_t seg, size_t w) const {
        if (seg >= segStart.size() || w >= segWords[seg]) throw std::runtime_error("panman: pointer out of range");
        uint64_t v; std::memcpy(&v, bytes.data() + segStart[seg] + 8 * w, 8); return v;
    }
};
struct Ref {  // a resolved pointer
    int kind = 0;            // 0 null, 1 struct, 2 list
    uint32_t seg = 0; size_t off = 0;       // first content word
    uint32_t dataWords = 0, ptrWords = 0;   // struct (or composite element) shape
    uint32_t elemSize = 0; uint32_t count = 0;  // list
};
inline Ref decodeAt(const Msg& m, uint64_t p, uint32_t seg, size_t contentBase) {
    // p is a struct/list pointer word whose offset is relative to contentBase (word after pointer)
    Ref r; r.seg = seg;
    const int32_t off = static_cast<int32_t>(p & 0xffffffffu) >> 2;
    if ((p & 3) == 0) {
        r.kind = 1; r.off = contentBase + off;
        r.dataWords = (p >> 32) & 0xffff; r.ptrWords = (p >> 48) & 0xffff;
    } else {
        r.kind = 2; r.off = contentBase + off;
        r.elemSize = (p >> 32) & 7; r.count = static_cast<uint32_t>(p >> 35);
        if (r.elemSize == 7) {  // composite: tag word first
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
    if ((p & 3) == 2) {  // far pointer
        const bool dbl = (p >> 2) & 1;
        const size_t padOff = (p & 0xffffffffu) >> 3;
        const uint32_t padSeg = static_cast<uint32_t>(p >> 32);
        if (!dbl) {
            const uint64_t q = m.word(padSeg, padOff);
            if (q == 0) return Ref{};
            return decodeAt(m, q, padSeg, padOff + 1);
        }
        const uint64_t far2 = m.word(padSeg, padOff);
        const uint64_t tag = m.word(padSeg, padOff + 1);
        const size_t cOff = (far2 & 0xffffffffu) >> 3;
        const uint32_t cSeg = static_cast<uint32_t>(far2 >> 32);
        // tag has offset 0; content begins exactly at (cSeg, cOff)
        Ref r = decodeAt(m, tag & ~0xfffffffcULL, cSeg, cOff);
        return r;
    }
    if ((p & 3) == 3) throw std::runtime_error("panman: capability pointer unexpected");
    return decodeAt(m, p, seg, w + 1);
}
inline Ref structPtr(const Msg& m, const Ref& s, uint32_t i) {
    if (s.kind != 1 || i >= s.ptrWords) return Ref{};
    return resolve(m, s.seg, s.off + s.dataWords + i);
}
inline uint64_t structData(const Msg& m, const Ref& s, uint32_t i) {
    if (s.kind != 1 || i >= s.dataWords) return 0;
    return m.word(s.seg, s.off + i);
}
inline Ref compositeElem(const Ref& l, uint32_t i) {
    Ref e; e.kind = 1; e.seg = l.seg; e.off = l.off + static_cast<size_t>(i) * (l.dataWords + l.ptrWords);
    e.dataWords = l.dataWords; e.ptrWords = l.ptrWords; return e;
}
inline Ref ptrListElem(const Msg& m, const Ref& l, uint32_t i) { return resolve(m, l.seg, l.off + i); }
inline std::string text(const Msg& m, const Ref& l) {
    if (l.kind != 2 || l.count == 0) return {};
    const uint8_t* p = m.bytes.data() + m.segStart[l.seg] + 8 * l.off;
    return std::string(reinterpret_cast<const char*>(p), l.count - 1);
}
template <class T>
inline std::vector<T> prims(const Msg& m, const Ref& l) {
    std::vector<T> out;
    if (l.kind != 2) return out;
    out.resize(l.count);
    if (l.count) std::memcpy(out.data(), m.bytes.data() + m.segStart[l.seg] + 8 * l.off, sizeof(T) * l.count);
    return out;
}

inline Msg readMessage(const std::string& path) {
    Msg m;
    const std::string cmd = "xz -dc '" + path + "'";
    FILE* fp = popen(cmd.c_str(), "r");
    if (!fp) throw std::runtime_error("panman: cannot run xz on " + path);
    uint8_t buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) m.bytes.insert(m.bytes.end(), buf, buf + n);
    pclose(fp);
    if (m.bytes.size() < 8) throw std::runtime_error("panman: empty stream " + path);
    uint32_t nseg; std::memcpy(&nseg, m.bytes.data(), 4); nseg += 1;
    size_t hdr = 4 + 4 * static_cast<size_t>(nseg); hdr = (hdr + 7) & ~size_t(7);
    size_t pos = hdr;
    for (uint32_t i = 0; i < nseg; ++i) {
        uint32_t sz; std::memcpy(&sz, m.bytes.data() + 4 + 4 * i, 4);
        m.segStart.push_back(pos); m.segWords.push_back(sz); pos += 8 * static_cast<size_t>(sz);
    }
    if (pos > m.bytes.size()) throw std::runtime_error("panman: truncated message");
    return m;
}

// newick -> nodes in pre-order (internal node opened at '(', label follows its ')').
inline void parseNewick(const std::string& nw, panmanUtils::Tree& T, std::vector<panmanUtils::Node*>& pre) {
    std::vector<panmanUtils::Node*> stack;
    panmanUtils::Node* lastClosed = nullptr;
    size_t i = 0;
    auto newNode = [&]() {
        auto* n = new panmanUtils::Node();
        if (!stack.empty()) { n->parent = stack.back(); stack.back()->children.push_back(n); n->level = stack.back()->level + 1; }
        else { T.root = n; n->level = 1; }
        pre.push_back(n); return n;
    };
    auto readLabel = [&](panmanUtils::Node* n) {
        size_t b = i;
        while (i < nw.size() && nw[i] != ':' && nw[i] != ',' && nw[i] != ')' && nw[i] != '(' && nw[i] != ';') ++i;
        n->identifier = nw.substr(b, i - b);
        if (i < nw.size() && nw[i] == ':') {
            ++i; size_t c = i;
            while (i < nw.size() && nw[i] != ',' && nw[i] != ')' && nw[i] != ';') ++i;
            try { n->branchLength = std::stof(nw.substr(c, i - c)); } catch (...) {}
        }
    };
    while (i < nw.size()) {
        const char c = nw[i];
        if (c == '(') { stack.push_back(newNode()); ++i; lastClosed = nullptr; }
        else if (c == ',') { ++i; lastClosed = nullptr; }
        else if (c == ')') { lastClosed = stack.back(); stack.pop_back(); ++i; readLabel(lastClosed); }
        else if (c == ';' || c == ' ' || c == '\n') { ++i; }
        else { auto* n = newNode(); readLabel(n); }
    }
    for (auto* n : pre) T.allNodes[n->identifier] = n;
}

inline void load(const std::string& path, panmanUtils::Tree& T) {
    const Msg m = readMessage(path);
    const Ref root = resolve(m, 0, 0);
    const Ref trees = structPtr(m, root, 0);
    if (trees.kind != 2 || trees.count == 0) throw std::runtime_error("panman: no trees");
    const Ref tree = (trees.elemSize == 7) ? compositeElem(trees, 0) : ptrListElem(m, trees, 0);
    const std::string newick = text(m, structPtr(m, tree, 0));
    std::vector<panmanUtils::Node*> pre;
    parseNewick(newick, T, pre);

    const Ref nodes = structPtr(m, tree, 1);
    if (nodes.count < pre.size()) throw std::runtime_error("panman: node list shorter than newick");
    for (size_t ni = 0; ni < pre.size(); ++ni) {
        const Ref node = compositeElem(nodes, static_cast<uint32_t>(ni));
        const Ref muts = structPtr(m, node, 0);
        for (uint32_t mi = 0; muts.kind == 2 && mi < muts.count; ++mi) {
            const Ref mu = compositeElem(muts, mi);
            const int64_t blockId = static_cast<int64_t>(structData(m, mu, 0));
            const uint64_t flags = structData(m, mu, 1);
            const int32_t primary = static_cast<int32_t>(blockId >> 32);
            const Ref nms = structPtr(m, mu, 0);
            for (uint32_t k = 0; nms.kind == 2 && k < nms.count; ++k) {
                const Ref nm = compositeElem(nms, k);
                const uint64_t w0 = structData(m, nm, 0), w1 = structData(m, nm, 1);
                panmanUtils::NucMut x;
                x.nucPosition = static_cast<int32_t>(w0 & 0xffffffffu);
                x.nucGapPosition = (w1 & 1) ? static_cast<int32_t>(w0 >> 32) : -1;
                x.primaryBlockId = primary;
                const uint32_t raw = static_cast<uint32_t>(w1 >> 32);
                x.mutInfo = static_cast<uint8_t>(raw & 0xff);
                const int len = x.mutInfo >> 4;
                x.nucs = (len <= 6) ? ((raw >> 8) << (24 - 4 * len)) : (raw >> 8);
                pre[ni]->nucMutation.push_back(x);
            }
            if (flags & 2) {
                panmanUtils::BlockMut b;
                b.primaryBlockId = primary;
                b.blockMutInfo = (flags >> 2) & 1;
                b.inversion = (flags >> 3) & 1;
                pre[ni]->blockMutation.push_back(b);
            }
        }
    }
    const Ref cmap = structPtr(m, tree, 2);
    for (uint32_t i = 0; cmap.kind == 2 && i < cmap.count; ++i) {
        const Ref e = compositeElem(cmap, i);
        const auto ids = prims<int64_t>(m, structPtr(m, e, 0));
        const auto seq = prims<uint32_t>(m, structPtr(m, e, 1));
        const std::string chrom = text(m, structPtr(m, e, 3));
        for (int64_t id : ids) {
            panmanUtils::Block b; b.primaryBlockId = id >> 32; b.consensusSeq = seq; b.chromosomeName = chrom;
            T.blocks.push_back(std::move(b));
        }
    }
    std::sort(T.blocks.begin(), T.blocks.end(),
              [](const panmanUtils::Block& a, const panmanUtils::Block& b) { return a.primaryBlockId < b.primaryBlockId; });
    const Ref gaps = structPtr(m, tree, 3);
    for (uint32_t i = 0; gaps.kind == 2 && i < gaps.count; ++i) {
        const Ref e = compositeElem(gaps, i);
        panmanUtils::GapList g;
        g.primaryBlockId = static_cast<int32_t>(static_cast<int64_t>(structData(m, e, 0)) >> 32);
        g.nucGapLength = prims<int32_t>(m, structPtr(m, e, 0));
        g.nucPosition = prims<int32_t>(m, structPtr(m, e, 1));
        T.gaps.push_back(std::move(g));
    }
}
}  // namespace panman_loader
