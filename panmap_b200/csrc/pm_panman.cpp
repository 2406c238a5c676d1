// pm_panman.cpp -- `.panman` reader and node-genome walker (host side, no CUDA): the input half of the index builder
// (pm_build.cu).
//
// File: an xz stream of an unpacked Cap'n Proto message, root TreeGroup -> trees[0] -> newick / nodes / consensusSeqMap / gaps
// (TurakhiaLab/panman v0.1.4's panman.capnp; the struct shapes are listed in SURVEY.md Appendix E).  liblzma is a runtime dependency
// only (dlopen, one function); the message is walked with the schema-less decoder of pm_capnp.h.
//
// Sequence model (reference: /root/reference/src/panmap_utils.cpp:7-131 getSequenceFromReference, :133-180 getStringFromSequence,
// panmap_utils.hpp:121-163 Coordinate, :204-213 forEachConsensusNuc): every block is a vector of (main base, gap bases before it) with a
// trailing sentinel; a node's genome is the root's state plus the block and nucleotide mutations on the path to it; absent blocks are
// skipped, inverted blocks are emitted as their reverse complement, gap characters are dropped.  The reference rebuilds that state from
// scratch for one node at a time; here ONE depth-first walk applies a node's mutations on the way down and takes them back on the way up
// (the nucleotide mutations of a block do not depend on whether the block is present, so applying all of them in path order gives every
// present block the content the from-scratch rule gives it), handing each node's ungapped genome to a callback.
#include "pm_capnp.h"
#include "pm_host.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>

namespace pm {
namespace {

using namespace capnp_walk;

// xz -> bytes.  lzma_stream_buffer_decode(memlimit*, flags, allocator, in, in_pos*, in_size, out, out_pos*, out_size); the output size is
// not known up front: grow and retry (LZMA_BUF_ERROR = 10, LZMA_OK = 0)
std::vector<uint8_t> inflateXz(const std::vector<uint8_t>& in, const std::string& path) {
    void* h = dlopen("liblzma.so.5", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("liblzma.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) throw Unsupported("reading .panman files needs liblzma.so.5 at run time");
    typedef int (*DecodeFn)(uint64_t*, uint32_t, const void*, const uint8_t*, size_t*, size_t, uint8_t*, size_t*, size_t);
    DecodeFn decode = reinterpret_cast<DecodeFn>(dlsym(h, "lzma_stream_buffer_decode"));
    if (!decode) throw Unsupported("liblzma lacks lzma_stream_buffer_decode");
    std::vector<uint8_t> out;
    for (size_t cap = std::max<size_t>(in.size() * 16, 1 << 20);; cap *= 4) {
        if (cap > ((size_t)1 << 36)) throw std::runtime_error("panman: implausible decompressed size: " + path);
        out.resize(cap);
        uint64_t memlimit = ~0ull; size_t ip = 0, op = 0;
        const int rc = decode(&memlimit, 0, nullptr, in.data(), &ip, in.size(), out.data(), &op, out.size());
        if (rc == 0) { out.resize(op); return out; }
        if (rc == 10 && op < out.size()) throw std::runtime_error("panman: xz stream truncated: " + path);   // LZMA_BUF_ERROR with room left: the input ran out
        if (rc != 10) throw std::runtime_error("panman: not an xz stream (liblzma code " + std::to_string(rc) + "): " + path);
    }
}

std::string textOf(const Msg& m, const Ref& l) {
    if (l.kind != 2 || l.elemSize != 2 || l.count == 0) return {};
    return std::string(reinterpret_cast<const char*>(m.span(l.seg, l.off, l.count)), (size_t)l.count - 1);
}
template <class T> std::vector<T> primsOf(const Msg& m, const Ref& l, unsigned wantCode) {
    std::vector<T> v;
    if (l.kind != 2 || l.count == 0) return v;
    if (l.elemSize != wantCode) throw std::runtime_error("panman: unexpected list element size");
    v.resize((size_t)l.count);
    std::memcpy(v.data(), m.span(l.seg, l.off, sizeof(T) * l.count), sizeof(T) * (size_t)l.count);
    return v;
}

// newick -> nodes in pre-order: an internal node is opened at '(' and named by the label after its ')'
void parseNewick(const std::string& nw, PanmanTree& T) {
    std::vector<uint32_t> stack;
    size_t i = 0;
    auto newNode = [&]() {
        PanmanNode n;
        n.parent = stack.empty() ? kNoNode : stack.back();
        T.nodes.push_back(std::move(n));
        const uint32_t v = (uint32_t)T.nodes.size() - 1;
        if (!stack.empty()) T.nodes[stack.back()].children.push_back(v);
        return v;
    };
    auto readLabel = [&](uint32_t v) {
        const size_t b = i;
        while (i < nw.size() && nw[i] != ':' && nw[i] != ',' && nw[i] != ')' && nw[i] != '(' && nw[i] != ';') ++i;
        T.nodes[v].id = nw.substr(b, i - b);
        if (i < nw.size() && nw[i] == ':') while (i < nw.size() && nw[i] != ',' && nw[i] != ')' && nw[i] != ';') ++i;   // branch length: unused
    };
    while (i < nw.size()) {
        const char c = nw[i];
        if (c == '(') { stack.push_back(newNode()); ++i; }
        else if (c == ',') ++i;
        else if (c == ')') {
            if (stack.empty()) throw std::runtime_error("panman: unbalanced newick");
            const uint32_t v = stack.back(); stack.pop_back(); ++i; readLabel(v);
        } else if (c == ';' || c == ' ' || c == '\n' || c == '\r' || c == '\t') ++i;
        else readLabel(newNode());
    }
    if (!stack.empty() || T.nodes.empty()) throw std::runtime_error("panman: unbalanced or empty newick");
}

char nucOfCode(int code) {   // 4-bit IUPAC masks: A 1, C 2, G 4, T 8
    static const char t[16] = {'-', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};
    return t[code & 15];
}
char complementOf(char c) {
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'R': return 'Y'; case 'Y': return 'R'; case 'K': return 'M'; case 'M': return 'K';
        case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
        default: return c;
    }
}

}  // namespace

void readPanman(const std::string& path, PanmanTree& T) {
    T = PanmanTree{};
    std::vector<uint8_t> packed;
    {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw IoError("cannot open panman file: " + path);
        std::fseek(f, 0, SEEK_END);
        const long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        packed.resize((size_t)(sz > 0 ? sz : 0));
        const bool ok = sz <= 0 || std::fread(packed.data(), 1, packed.size(), f) == packed.size();
        std::fclose(f);
        if (!ok) throw IoError("short read: " + path);
    }
    if (packed.size() < 6 || std::memcmp(packed.data(), "\xFD" "7zXZ", 6) != 0) throw std::runtime_error("panman: not an xz stream: " + path);
    const std::vector<uint8_t> bytes = inflateXz(packed, path);
    Msg m;
    openMessage(m, bytes.data(), bytes.size(), "panman");
    const Ref root = resolve(m, 0, 0);
    const Ref trees = ptrOf(m, root, 0);
    if (trees.kind != 2 || trees.count == 0) throw std::runtime_error("panman: no trees in " + path);
    const Ref tree = trees.elemSize == 7 ? elemOf(trees, 0) : resolve(m, trees.seg, trees.off);
    parseNewick(textOf(m, ptrOf(m, tree, 0)), T);

    const Ref nodes = ptrOf(m, tree, 1);
    if (nodes.kind != 2 || nodes.elemSize != 7 || nodes.count < T.nodes.size()) throw std::runtime_error("panman: node list shorter than the newick tree");
    for (size_t ni = 0; ni < T.nodes.size(); ++ni) {
        PanmanNode& N = T.nodes[ni];
        N.nucBegin = (uint32_t)T.nucMuts.size(); N.blockBegin = (uint32_t)T.blockMuts.size();
        const Ref muts = ptrOf(m, elemOf(nodes, ni), 0);
        for (uint64_t mi = 0; muts.kind == 2 && muts.elemSize == 7 && mi < muts.count; ++mi) {
            const Ref mu = elemOf(muts, mi);
            const int32_t primary = (int32_t)((int64_t)dataOf(m, mu, 0) >> 32);
            const uint64_t flags = dataOf(m, mu, 1);
            const Ref nms = ptrOf(m, mu, 0);
            for (uint64_t q = 0; nms.kind == 2 && nms.elemSize == 7 && q < nms.count; ++q) {
                const Ref nm = elemOf(nms, q);
                const uint64_t w0 = dataOf(m, nm, 0), w1 = dataOf(m, nm, 1);
                PanmanNucMut x;
                x.block = primary;
                x.pos = (int32_t)(w0 & 0xffffffffu);
                x.gap = (w1 & 1) ? (int32_t)(w0 >> 32) : -1;
                const uint32_t raw = (uint32_t)(w1 >> 32);
                x.len = (uint8_t)std::min<uint32_t>((raw & 0xff) >> 4, 6u); x.type = (uint8_t)(raw & 0xf);   // 24 bits hold at most six 4-bit codes
                x.nucs = (raw >> 8) << (24 - 4 * x.len);   // base i = (nucs >> 4 (5 - i)) & 15
                T.nucMuts.push_back(x);
            }
            if (flags & 2) T.blockMuts.push_back(PanmanBlockMut{primary, (flags >> 2 & 1) != 0, (flags >> 3 & 1) != 0});
        }
        N.nucEnd = (uint32_t)T.nucMuts.size(); N.blockEnd = (uint32_t)T.blockMuts.size();
    }
    // consensus sequences: several block ids may share one entry; blocks are addressed by their primary id
    struct Cons { int64_t id; std::vector<uint32_t> seq; };
    std::vector<Cons> cons;
    const Ref cmap = ptrOf(m, tree, 2);
    for (uint64_t i = 0; cmap.kind == 2 && cmap.elemSize == 7 && i < cmap.count; ++i) {
        const Ref e = elemOf(cmap, i);
        const std::vector<int64_t> ids = primsOf<int64_t>(m, ptrOf(m, e, 0), 5);
        const std::vector<uint32_t> seq = primsOf<uint32_t>(m, ptrOf(m, e, 1), 4);
        for (int64_t id : ids) cons.push_back(Cons{id >> 32, seq});
    }
    std::sort(cons.begin(), cons.end(), [](const Cons& a, const Cons& b) { return a.id < b.id; });
    T.blocks.resize(cons.size());
    for (size_t b = 0; b < cons.size(); ++b) {
        if (cons[b].id != (int64_t)b) throw std::runtime_error("panman: block ids are not 0 .. n-1");
        std::string& s = T.blocks[b];
        bool done = false;
        for (size_t i = 0; i < cons[b].seq.size() && !done; ++i)
            for (int j = 0; j < 8; ++j) {
                const int code = (int)((cons[b].seq[i] >> (4 * (7 - j))) & 15u);
                if (code == 0) { done = true; break; }   // the first zero nibble ends the block
                s.push_back(nucOfCode(code));
            }
    }
    const Ref gaps = ptrOf(m, tree, 3);
    for (uint64_t i = 0; gaps.kind == 2 && gaps.elemSize == 7 && i < gaps.count; ++i) {
        const Ref e = elemOf(gaps, i);
        PanmanGapList g;
        g.block = (int32_t)((int64_t)dataOf(m, e, 0) >> 32);
        g.length = primsOf<int32_t>(m, ptrOf(m, e, 0), 4);
        g.position = primsOf<int32_t>(m, ptrOf(m, e, 1), 4);
        if (g.length.size() != g.position.size()) throw std::runtime_error("panman: gap list lengths differ");
        T.gaps.push_back(std::move(g));
    }
}

namespace {

// the mutable sequence state of the walk
struct SeqState {
    // block b: main[b] = main bases + sentinel 'x'; gapOff[b][p] .. gapOff[b][p + 1] = the gap slots before main position p in gapChars[b]
    std::vector<std::string> main, gapChars;
    std::vector<std::vector<uint32_t>> gapOff;
    std::vector<std::vector<uint32_t>> gapLive;   // [b][p]: bases (non '-') in the gap run before main position p -- most runs are empty
    std::vector<char> exists, forward;
};
struct Undo { int32_t block; uint32_t idx; char old; bool inGap; uint32_t pos; };
struct BlockUndo { int32_t block; char exists, forward; };

}  // namespace

void walkPanmanGenomes(const PanmanTree& T, const std::function<void(uint32_t, const std::string&)>& visit) {
    walkPanmanGenomesCoords(T, false, [&](uint32_t v, const std::string& g, const std::vector<uint32_t>&) { visit(v, g); });
}
// the same walk, optionally with the global (aligned) coordinate of every emitted base: slots are numbered block after block, inside a
// block the gap slots before main position p come right before p (the reference's scalar coordinates, panmap_utils.hpp:281-420)
void walkPanmanGenomesCoords(const PanmanTree& T, bool wantCoords, const std::function<void(uint32_t, const std::string&, const std::vector<uint32_t>&)>& visit) {
    const size_t B = T.blocks.size();
    SeqState S;
    S.main.resize(B); S.gapChars.resize(B); S.gapOff.resize(B); S.gapLive.resize(B); S.exists.assign(B, 0); S.forward.assign(B, 1);
    std::vector<std::vector<uint32_t>> gapLen(B);
    for (size_t b = 0; b < B; ++b) { S.main[b] = T.blocks[b]; S.main[b].push_back('x'); gapLen[b].assign(S.main[b].size(), 0); }
    for (const PanmanGapList& g : T.gaps) {
        if (g.block < 0 || (size_t)g.block >= B) throw std::runtime_error("panman: gap list names an unknown block");
        for (size_t j = 0; j < g.position.size(); ++j) {
            if (g.position[j] < 0 || (size_t)g.position[j] >= gapLen[g.block].size() || g.length[j] < 0) throw std::runtime_error("panman: gap list position out of range");
            gapLen[g.block][g.position[j]] = (uint32_t)g.length[j];   // resize(len, '-') semantics: the last entry for a position wins
        }
    }
    for (size_t b = 0; b < B; ++b) {
        S.gapOff[b].resize(S.main[b].size() + 1);
        uint32_t acc = 0;
        for (size_t p = 0; p < S.main[b].size(); ++p) { S.gapOff[b][p] = acc; acc += gapLen[b][p]; }
        S.gapOff[b][S.main[b].size()] = acc;
        S.gapChars[b].assign(acc, '-');
        S.gapLive[b].assign(S.main[b].size(), 0);
    }
    std::vector<Undo> undo;
    std::vector<BlockUndo> blockUndo;
    std::string genome;
    std::vector<uint32_t> coords;
    std::vector<uint64_t> blockStart(B + 1, 0);
    for (size_t b = 0; b < B; ++b) blockStart[b + 1] = blockStart[b] + (S.main[b].size() - 1) + S.gapChars[b].size();
    if (blockStart[B] > 0xFFFFFFFFull) throw std::runtime_error("panman: more than 2^32 aligned positions");
    // iterative depth-first walk in newick order; frame = (node, next child, undo marks)
    struct Frame { uint32_t node; size_t child; size_t undoMark, blockMark; };
    std::vector<Frame> stack;
    auto enter = [&](uint32_t v) {
        const PanmanNode& N = T.nodes[v];
        Frame f{v, 0, undo.size(), blockUndo.size()};
        for (uint32_t i = N.blockBegin; i < N.blockEnd; ++i) {   // panmap_utils.cpp:93-111
            const PanmanBlockMut& bm = T.blockMuts[i];
            if (bm.block < 0 || (size_t)bm.block >= B) throw std::runtime_error("panman: block mutation names an unknown block");
            blockUndo.push_back(BlockUndo{bm.block, S.exists[bm.block], S.forward[bm.block]});
            if (bm.insertion) { S.exists[bm.block] = 1; S.forward[bm.block] = bm.inversion ? 0 : 1; }
            else if (bm.inversion) S.forward[bm.block] = S.forward[bm.block] ? 0 : 1;
            else { S.exists[bm.block] = 0; S.forward[bm.block] = 1; }
        }
        for (uint32_t i = N.nucBegin; i < N.nucEnd; ++i) {        // panmap_utils.cpp:113-129
            const PanmanNucMut& nm = T.nucMuts[i];
            if (nm.block < 0 || (size_t)nm.block >= B) throw std::runtime_error("panman: nucleotide mutation names an unknown block");
            std::string& mainB = S.main[nm.block];
            for (int q = 0; q < (int)nm.len; ++q) {
                const int32_t pos = nm.gap == -1 ? nm.pos + q : nm.pos, gp = nm.gap == -1 ? -1 : nm.gap + q;
                if (pos < 0) continue;
                if ((size_t)pos == mainB.size() - 1 && gp == -1) continue;   // the sentinel
                if ((size_t)pos >= mainB.size()) continue;
                const char nuc = nucOfCode((int)((nm.nucs >> (4 * (5 - q))) & 15u));
                if (gp == -1) { undo.push_back(Undo{nm.block, (uint32_t)pos, mainB[pos], false, (uint32_t)pos}); mainB[pos] = nuc; }
                else {
                    const uint32_t a = S.gapOff[nm.block][pos], e = S.gapOff[nm.block][pos + 1];
                    if (gp < 0 || a + (uint32_t)gp >= e) continue;   // outside the gap run (the reference would write out of bounds)
                    char& slot = S.gapChars[nm.block][a + gp];
                    undo.push_back(Undo{nm.block, a + (uint32_t)gp, slot, true, (uint32_t)pos});
                    S.gapLive[nm.block][pos] += (uint32_t)(nuc != '-') - (uint32_t)(slot != '-');
                    slot = nuc;
                }
            }
        }
        // the node's ungapped genome (panmap_utils.cpp:133-180 with aligned = false)
        size_t cap = 0;
        for (size_t b = 0; b < B; ++b) if (S.exists[b]) cap += S.main[b].size() + S.gapChars[b].size();   // upper bound (+1 scratch byte per write)
        genome.resize(cap);
        char* w = &genome[0];
        if (wantCoords) coords.resize(cap);
        uint32_t* cw = wantCoords ? coords.data() : nullptr;
        for (size_t b = 0; b < B; ++b) {
            if (!S.exists[b]) continue;
            const std::string& mb = S.main[b]; const std::string& gc = S.gapChars[b]; const std::vector<uint32_t>& go = S.gapOff[b];
            const std::vector<uint32_t>& live = S.gapLive[b];
            const size_t nMain = mb.size() - 1;   // without the sentinel; its gap run (go[nMain] .. go[nMain + 1]) still counts
            if (S.forward[b] && wantCoords) {
                const uint32_t base = (uint32_t)blockStart[b];
                for (size_t p = 0; p <= nMain; ++p) {
                    if (live[p]) for (uint32_t g = go[p]; g < go[p + 1]; ++g) { const char c = gc[g]; *w = c; *cw = base + (uint32_t)p + g; w += c != '-'; cw += c != '-'; }
                    if (p < nMain) { const char c = mb[p]; *w = c; *cw = base + (uint32_t)p + go[p + 1]; w += c != '-'; cw += c != '-'; }
                }
            } else if (S.forward[b]) {
                for (size_t p = 0; p <= nMain; ++p) {
                    if (live[p]) for (uint32_t g = go[p]; g < go[p + 1]; ++g) { const char c = gc[g]; *w = c; w += c != '-'; }
                    if (p < nMain) { const char c = mb[p]; *w = c; w += c != '-'; }
                }
            } else {
                if (wantCoords) throw Unsupported("aligned coordinates of a genome with an inverted block");
                for (size_t p = nMain + 1; p-- > 0;) {
                    if (p < nMain) { const char c = mb[p]; *w = complementOf(c); w += c != '-'; }
                    if (live[p]) for (uint32_t g = go[p + 1]; g-- > go[p];) { const char c = gc[g]; *w = complementOf(c); w += c != '-'; }
                }
            }
        }
        genome.resize((size_t)(w - genome.data()));
        if (wantCoords) coords.resize(genome.size());
        visit(v, genome, coords);
        stack.push_back(f);
    };
    enter(0);
    while (!stack.empty()) {
        Frame& f = stack.back();
        const PanmanNode& N = T.nodes[f.node];
        if (f.child < N.children.size()) { const uint32_t c = N.children[f.child++]; enter(c); continue; }
        while (undo.size() > f.undoMark) {
            const Undo& u = undo.back();
            if (u.inGap) {
                char& slot = S.gapChars[u.block][u.idx];
                S.gapLive[u.block][u.pos] += (uint32_t)(u.old != '-') - (uint32_t)(slot != '-');
                slot = u.old;
            } else S.main[u.block][u.idx] = u.old;
            undo.pop_back();
        }
        while (blockUndo.size() > f.blockMark) {
            const BlockUndo& u = blockUndo.back();
            S.exists[u.block] = u.exists; S.forward[u.block] = u.forward;
            blockUndo.pop_back();
        }
        stack.pop_back();
    }
}

// The tree flattened for the device pipeline of the builder (pm_build_kernels.cu genome_materialize): aligned template, one point edit per
// mutated slot (the skip rules of panmap_utils.cpp:113-129 applied here), block mutations, parents, depth.
void flattenPanman(const PanmanTree& T, PanmanFlat& F) {
    F = PanmanFlat{};
    const size_t B = T.blocks.size(), N = T.nodes.size();
    std::vector<std::vector<uint32_t>> gapLen(B), go(B);
    for (size_t b = 0; b < B; ++b) gapLen[b].assign(T.blocks[b].size() + 1, 0);
    for (const PanmanGapList& g : T.gaps) {
        if (g.block < 0 || (size_t)g.block >= B) throw std::runtime_error("panman: gap list names an unknown block");
        for (size_t j = 0; j < g.position.size(); ++j) {
            if (g.position[j] < 0 || (size_t)g.position[j] >= gapLen[g.block].size() || g.length[j] < 0) throw std::runtime_error("panman: gap list position out of range");
            gapLen[g.block][g.position[j]] = (uint32_t)g.length[j];
        }
    }
    F.blockStart.assign(B + 1, 0);
    uint64_t slots = 0;
    for (size_t b = 0; b < B; ++b) {
        const size_t nMain = T.blocks[b].size();
        go[b].assign(nMain + 2, 0);
        for (size_t p = 0; p <= nMain; ++p) go[b][p + 1] = go[b][p] + gapLen[b][p];
        F.blockStart[b] = (uint32_t)slots;
        slots += nMain + go[b][nMain + 1];
        if (slots > 0x7FFFFFFFull) throw Unsupported("panman: more than 2^31 aligned positions");
    }
    F.blockStart[B] = (uint32_t)slots;
    F.tmpl.assign((size_t)slots, '-'); F.slotBlock.assign((size_t)slots, 0);
    for (size_t b = 0; b < B; ++b) {
        for (uint32_t q = F.blockStart[b]; q < F.blockStart[b + 1]; ++q) F.slotBlock[q] = (uint32_t)b;
        for (size_t p = 0; p < T.blocks[b].size(); ++p) F.tmpl[F.blockStart[b] + p + go[b][p + 1]] = T.blocks[b][p];
    }
    F.parent.resize(N); F.editBegin.assign(N + 1, 0); F.blockMutBegin.assign(N + 1, 0); F.editSerial.assign(N, 0);
    std::vector<uint32_t> depth(N, 1), seen;
    F.maxDepth = 1;
    for (size_t v = 0; v < N; ++v) {
        const PanmanNode& nd = T.nodes[v];
        F.parent[v] = nd.parent;
        if (nd.parent != kNoNode) { depth[v] = depth[nd.parent] + 1; F.maxDepth = std::max(F.maxDepth, depth[v]); }
        for (uint32_t i = nd.blockBegin; i < nd.blockEnd; ++i) {
            const PanmanBlockMut& bm = T.blockMuts[i];
            if (bm.block < 0 || (size_t)bm.block >= B) throw std::runtime_error("panman: block mutation names an unknown block");
            F.blockMut.push_back(((uint32_t)bm.block << 2) | (bm.inversion ? 2u : 0u) | (bm.insertion ? 1u : 0u));
        }
        F.blockMutBegin[v + 1] = (uint32_t)F.blockMut.size();
        const size_t e0 = F.editSlot.size();
        for (uint32_t i = nd.nucBegin; i < nd.nucEnd; ++i) {
            const PanmanNucMut& nm = T.nucMuts[i];
            if (nm.block < 0 || (size_t)nm.block >= B) throw std::runtime_error("panman: nucleotide mutation names an unknown block");
            const size_t nMain = T.blocks[nm.block].size();
            for (int q = 0; q < (int)nm.len; ++q) {
                const int32_t pos = nm.gap == -1 ? nm.pos + q : nm.pos, gp = nm.gap == -1 ? -1 : nm.gap + q;
                if (pos < 0 || (size_t)pos > nMain) continue;            // beyond the sentinel
                if ((size_t)pos == nMain && gp == -1) continue;          // the sentinel itself
                uint32_t slot;
                if (gp == -1) slot = F.blockStart[nm.block] + (uint32_t)pos + go[nm.block][pos + 1];
                else {
                    if (gp < 0 || (uint32_t)gp >= gapLen[nm.block][pos]) continue;
                    slot = F.blockStart[nm.block] + (uint32_t)pos + go[nm.block][pos] + (uint32_t)gp;
                }
                F.editSlot.push_back(slot); F.editChar.push_back(nucOfCode((int)((nm.nucs >> (4 * (5 - q))) & 15u)));
            }
        }
        F.editBegin[v + 1] = (uint32_t)F.editSlot.size();
        seen.assign(F.editSlot.begin() + (long)e0, F.editSlot.end());
        std::sort(seen.begin(), seen.end());
        F.editSerial[v] = std::adjacent_find(seen.begin(), seen.end()) != seen.end() ? 1 : 0;
    }
    if (F.editSlot.size() > 0xFFFFFFF0ull) throw Unsupported("panman: more than 2^32 point edits");
}

}  // namespace pm
