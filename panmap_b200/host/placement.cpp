// placement.cpp -- host shim: file ingest + result assembly around the C ABI (see placement.hpp).
#include "placement.hpp"
#include "seeding.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <cstdlib>
#include <stdexcept>
#include <thread>

namespace {

// reads of one file in the layout the C ABI takes: read i = data[off[i], off[i+1])
struct FlatReads {
    std::string data; std::vector<uint64_t> off;
    FlatReads() : off(1, 0) {}
    size_t size() const { return off.size() - 1; }
    void push(const std::string& s) { data += s; off.push_back(data.size()); }
};

// kseq semantics (FASTA and FASTQ, multi-line sequences, gz or plain through zlib's transparent gzread)
// quals (optional): the quality string of every record, 'I' * length when the record has none (extractFullFastqData, placement.cpp:199-238)
void readFastxSerial(const std::string& path, FlatReads& out, FlatReads* quals) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Failed to open FASTQ file: " + path);  // mgsr.hpp:175
    gzbuffer(f, 1 << 20);
    std::string data;
    std::vector<char> buf(1 << 22);
    int n;
    while ((n = gzread(f, buf.data(), static_cast<unsigned>(buf.size()))) > 0) data.append(buf.data(), static_cast<size_t>(n));
    gzclose(f);
    // kseq.h semantics: a record starts at '>' or '@'; sequence lines run until a line that starts with '>', '@' or '+';
    // after '+' the quality is read until it is as long as the sequence
    size_t p = 0;
    const size_t N = data.size();
    auto lineEnd = [&](size_t q) { while (q < N && data[q] != '\n') ++q; return q; };
    auto appendLine = [&](std::string& dst, size_t b, size_t e) {   // ks_getuntil(KS_SEP_LINE): the line without its end, a trailing CR dropped
        if (e > b && data[e - 1] == '\r') --e;
        dst.append(data, b, e - b);
    };
    while (p < N && data[p] != '>' && data[p] != '@') p = lineEnd(p) + 1;
    std::string seq, qual;
    while (p < N) {
        p = lineEnd(p) + 1;  // header line
        seq.clear(); qual.clear();
        while (p < N && data[p] != '>' && data[p] != '@' && data[p] != '+') { const size_t e = lineEnd(p); appendLine(seq, p, e); p = e + 1; }
        if (p < N && data[p] == '+') {
            p = lineEnd(p) + 1;
            while (p < N && qual.size() < seq.size()) { const size_t e = lineEnd(p); appendLine(qual, p, e); p = e + 1; }
            while (p < N && data[p] != '>' && data[p] != '@') p = lineEnd(p) + 1;
            if (qual.size() != seq.size()) break;   // kseq_read returns -2 (truncated quality): `while (kseq_read(seq) >= 0)` ends here, with or without quals
        }
        if (quals) quals->push(qual.empty() ? std::string(seq.size(), 'I') : qual);
        out.push(seq);
    }
}

// Uncompressed strict four-line FASTQ (what parallelFastqSeqs, placement.cpp:96-162, takes): the file is mapped, cut at record starts
// into one range per thread, and parsed twice -- first for the record count and base total of every range, then, the prefix sums known,
// every thread copies its sequences (and qualities) to their final place in the flat buffers.  No per-read strings, no merge.
// Returns false for gzip / FASTA / anything that is not four lines per record: the serial parser above takes those.
struct MappedFile {
    const char* d = nullptr; size_t size = 0; int fd = -1;
    explicit MappedFile(const std::string& path) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        if (::fstat(fd, &st) != 0 || st.st_size <= 0) return;
        void* m = ::mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return;
        d = static_cast<const char*>(m); size = static_cast<size_t>(st.st_size);
    }
    ~MappedFile() { if (d) ::munmap(const_cast<char*>(d), size); if (fd >= 0) ::close(fd); }
};
inline size_t fqEol(const char* d, size_t size, size_t p) {
    const void* nl = p < size ? std::memchr(d + p, '\n', size - p) : nullptr;
    return nl ? static_cast<size_t>(static_cast<const char*>(nl) - d) : size;
}
// Where does the first whole record begin at or after byte `from`?  One pass over the newlines keeps the starts of the last four
// lines in a small ring; as soon as four consecutive lines look like (header '@...', bases, '+...', qualities of the same length as the
// bases) the first of them is the answer.  The equal-length test is what tells a header from a quality line that happens to begin with '@'.
size_t fqRecordStart(const char* d, size_t size, size_t from) {
    size_t pos = from;
    if (pos > 0 && pos < size && d[pos - 1] != '\n') pos = fqEol(d, size, pos) + 1;   // move to the next line boundary
    size_t start[4], len[4];   // ring over the last four lines
    int have = 0;
    while (pos < size) {
        const size_t e = fqEol(d, size, pos);
        if (have == 4) { for (int i = 0; i < 3; ++i) { start[i] = start[i + 1]; len[i] = len[i + 1]; } have = 3; }
        start[have] = pos; len[have] = e - pos; ++have;
        if (have == 4 && d[start[0]] == '@' && len[2] > 0 && d[start[2]] == '+' && len[1] == len[3]) return start[0];
        pos = e + 1;
    }
    // fewer than four lines left after the candidate: a last record whose quality line ends the file without a newline was handled above
    return size;
}
bool readFastqParallel(const std::string& path, FlatReads& out, FlatReads* quals) {
    MappedFile mf(path);
    if (!mf.d || mf.size < 4) return false;
    const char* d = mf.d; const size_t size = mf.size;
    if (d[0] != '@') return false;   // gzip streams (1f 8b), FASTA ('>') and anything else: the serial parser takes them
    size_t nT = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (size < (1u << 20)) nT = 1;
    std::vector<size_t> bounds(nT + 1, size);
    bounds[0] = 0;
    for (size_t i = 1; i < nT; ++i) bounds[i] = std::max(bounds[i - 1], fqRecordStart(d, size, (size / nT) * i));
    struct Range { std::vector<uint64_t> seqAt, qualAt; std::vector<uint32_t> len; uint64_t bases = 0; bool bad = false; };
    std::vector<Range> R(nT);
    auto scan = [&](size_t t) {
        Range& r = R[t];
        size_t o = bounds[t];
        const size_t end = bounds[t + 1];
        while (o < end) {
            if (d[o] != '@') {   // only blank lines may follow the last record
                for (size_t i = o; i < end; ++i) if (d[i] != '\n' && d[i] != '\r') { r.bad = true; break; }
                break;
            }
            const size_t s0 = fqEol(d, size, o) + 1;
            if (s0 > size) { r.bad = true; break; }
            size_t s1 = fqEol(d, size, s0);
            const size_t p0 = s1 + 1;
            if (p0 >= size || d[p0] != '+') { r.bad = true; break; }
            const size_t q0 = fqEol(d, size, p0) + 1;
            size_t q1 = fqEol(d, size, q0);
            o = q1 + 1;
            if (s1 > s0 && d[s1 - 1] == '\r') --s1;   // \r\n files
            if (q1 > q0 && q0 <= size && d[q1 - 1] == '\r') --q1;
            const size_t L = s1 - s0;
            if (q0 > size || q1 - q0 != L || L > 0xFFFFFFFFull) { r.bad = true; break; }
            for (size_t i = s0; i < s1; ++i) if (d[i] == ' ' || d[i] == '\t') { r.bad = true; break; }   // kseq would drop these
            if (r.bad) break;
            r.seqAt.push_back(s0); r.len.push_back(static_cast<uint32_t>(L)); r.bases += L;
            if (quals) r.qualAt.push_back(q0);
        }
    };
    auto runAll = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (size_t t = 1; t < nT; ++t) th.emplace_back(fn, t);
        fn(0);
        for (auto& x : th) x.join();
    };
    runAll(scan);
    uint64_t nReads = 0, nBases = 0;
    std::vector<uint64_t> firstRead(nT), firstBase(nT);
    for (size_t t = 0; t < nT; ++t) { if (R[t].bad) return false; firstRead[t] = nReads; firstBase[t] = nBases; nReads += R[t].len.size(); nBases += R[t].bases; }
    out.data.resize(nBases); out.off.assign(nReads + 1, 0);
    if (quals) { quals->data.resize(nBases); quals->off.assign(nReads + 1, 0); }
    auto fill = [&](size_t t) {
        const Range& r = R[t];
        uint64_t b = firstBase[t];
        for (size_t i = 0; i < r.len.size(); ++i) {
            out.off[firstRead[t] + i] = b;
            std::memcpy(&out.data[b], d + r.seqAt[i], r.len[i]);
            if (quals) std::memcpy(&quals->data[b], d + r.qualAt[i], r.len[i]);
            b += r.len[i];
        }
    };
    runAll(fill);
    out.off[nReads] = nBases;
    if (quals) quals->off = out.off;
    return true;
}

void readFastx(const std::string& path, FlatReads& out, FlatReads* quals) {
    if (!readFastqParallel(path, out, quals)) { out = FlatReads(); if (quals) *quals = FlatReads(); readFastxSerial(path, out, quals); }
}

}  // namespace

namespace placement {

void extractReadSequences(const std::string& readPath1, const std::string& readPath2, std::string& bases, std::vector<uint64_t>& offsets,
                          std::string* quals) {
    FlatReads a, qa;
    readFastx(readPath1, a, quals ? &qa : nullptr);
    if (readPath2.empty()) {
        bases = std::move(a.data); offsets = std::move(a.off);
        if (quals) *quals = std::move(qa.data);
        return;
    }
    FlatReads b, qb;
    readFastx(readPath2, b, quals ? &qb : nullptr);
    if (b.size() != a.size()) throw std::runtime_error("File " + readPath2 + " does not contain the same number of reads as " + readPath1);
    // seeding::perfect_shuffle (seeding.hpp:33-43): pair i becomes reads 2i, 2i+1; their places follow from the two offset arrays
    const size_t n = a.size();
    offsets.assign(2 * n + 1, 0);
    bases.resize(a.data.size() + b.data.size());
    if (quals) quals->resize(bases.size());
    for (size_t i = 0; i < n; ++i) {
        const uint64_t o0 = a.off[i] + b.off[i], o1 = a.off[i + 1] + b.off[i];
        offsets[2 * i] = o0; offsets[2 * i + 1] = o1;
        std::memcpy(&bases[o0], &a.data[a.off[i]], a.off[i + 1] - a.off[i]);
        std::memcpy(&bases[o1], &b.data[b.off[i]], b.off[i + 1] - b.off[i]);
        if (quals) {
            std::memcpy(&(*quals)[o0], &qa.data[a.off[i]], a.off[i + 1] - a.off[i]);
            std::memcpy(&(*quals)[o1], &qb.data[b.off[i]], b.off[i + 1] - b.off[i]);
        }
    }
    offsets[2 * n] = bases.size();
}

void placeLite(PlacementResult& result, DeviceIndex& index, const std::string& reads1, const std::string& reads2, std::string& outputPath,
               const TraversalParams& params) {
    if (!index.index || !index.workspace) throw std::runtime_error("placeLite: device index not initialised");
    std::string bases, quals; std::vector<uint64_t> off(1, 0);
    const bool quality = params.minSeedQuality > 0;   // placement.cpp:1131-1137: the quality strings are loaded only then
    if (!reads1.empty()) extractReadSequences(reads1, reads2, bases, off, quality ? &quals : nullptr);
    pm_place_params p{};
    p.trim_start = params.trimStart; p.trim_end = params.trimEnd; p.min_read_support = params.minReadSupport;
    p.dedup_reads = params.dedupReads ? 1 : 0; p.force_leaf = params.forceLeaf ? 1 : 0; p.skip_node_index = PM_NONE;
    p.seed_mask_fraction = params.seedMaskFraction; p.want_node_scores = params.store_diagnostics ? 1 : 0;
    p.min_seed_quality = quality && off.size() > 1 ? params.minSeedQuality : 0;   // no reads: `!allReadQualities.empty()` fails, default path
    pm_place_result r{};
    const int rc = p.min_seed_quality > 0 ? pm_place_quality(index.workspace, bases.data(), quals.data(), off.data(), off.size() - 1, &p, &r)
                                          : pm_place(index.workspace, bases.data(), off.data(), off.size() - 1, &p, &r);
    if (rc != PM_OK) throw std::runtime_error(pm_last_error());
    double* sc[5] = {&result.bestLogRawScore, &result.bestLogCosineScore, &result.bestContainmentScore, &result.bestWeightedContainmentScore,
                     &result.bestLogContainmentScore};
    uint32_t* ix[5] = {&result.bestLogRawNodeIndex, &result.bestLogCosineNodeIndex, &result.bestContainmentNodeIndex,
                       &result.bestWeightedContainmentNodeIndex, &result.bestLogContainmentNodeIndex};
    std::vector<uint32_t>* td[5] = {&result.tiedLogRawNodeIndices, &result.tiedLogCosineNodeIndices, &result.tiedContainmentNodeIndices,
                                    &result.tiedWeightedContainmentNodeIndices, &result.tiedLogContainmentNodeIndices};
    std::string* id[5] = {&result.bestLogRawNodeId, &result.bestLogCosineNodeId, &result.bestContainmentNodeId,
                          &result.bestWeightedContainmentNodeId, &result.bestLogContainmentNodeId};
    auto name = [&](uint32_t v) -> std::string { return (index.nodeIds && v < index.nodeIds->size()) ? (*index.nodeIds)[v] : std::string(); };
    for (int m = 0; m < 5; ++m) {
        *sc[m] = r.best_score[m]; *ix[m] = r.best_index[m];
        td[m]->assign(r.tied_count[m], 0);
        if (r.tied_count[m]) pm_get_tied(index.workspace, m, td[m]->data(), r.tied_count[m]);
        *id[m] = r.best_index[m] != PM_NONE ? name(r.best_index[m]) : std::string();
    }
    if (params.store_diagnostics) {
        const uint64_t N = pm_index_num_nodes(index.index);
        std::vector<double> flat(N * 5);
        if (pm_get_node_scores(index.workspace, flat.data()) != PM_OK) throw std::runtime_error(pm_last_error());
        result.nodeScores.assign(N, std::vector<double>(5));
        for (uint64_t v = 0; v < N; ++v) for (int m = 0; m < 5; ++m) result.nodeScores[v][m] = flat[v * 5 + m];
    }
    result.totalReadsProcessed = static_cast<int64_t>(r.total_reads);
    result.reads1Path = reads1; result.reads2Path = reads2;
    result.readUniqueSeedCount = r.read_unique_seed_count; result.totalReadSeedFrequency = r.total_read_seed_frequency;
    result.readMagnitude = r.read_magnitude;
    result.raw = r;
    // <prefix>.placement.tsv (placement.cpp:1952-1985)
    std::ofstream out(outputPath);
    if (out.is_open()) {
        out << "metric\tscore\tnodes\n";
        const char* names[5] = {"log_raw", "log_cosine", "containment", "weighted_containment", "log_containment"};
        for (int m = 0; m < 5; ++m) {
            out << names[m] << "\t" << std::fixed << std::setprecision(6) << *sc[m] << "\t";
            if (!td[m]->empty()) { for (size_t i = 0; i < td[m]->size(); ++i) { if (i) out << ","; out << name((*td[m])[i]); } }
            else out << *id[m];
            out << "\n";
        }
    }
}

}  // namespace placement

namespace seeding {
std::vector<std::tuple<size_t, bool, bool, int64_t>> rollingSyncmers(std::string_view seq, int k, int s, bool open, int t, bool returnAll, int device) {
    std::vector<std::tuple<size_t, bool, bool, int64_t>> out;
    if (static_cast<int64_t>(seq.size()) < k) return out;
    const uint64_t off[2] = {0, seq.size()};
    const size_t win = seq.size() - k + 1;
    std::vector<uint64_t> h(win); std::vector<uint8_t> rev(win); std::vector<int64_t> pos(win); uint64_t cnt = 0;
    if (pm_rolling_syncmers(device, seq.data(), off, 1, k, s, open ? 1 : 0, t, h.data(), rev.data(), pos.data(), &cnt) != PM_OK)
        throw std::runtime_error(pm_last_error());
    size_t j = 0;
    for (size_t p = 0; p < win; ++p) {
        if (j < cnt && static_cast<size_t>(pos[j]) == p) { out.emplace_back(h[j], rev[j] != 0, true, static_cast<int64_t>(p)); ++j; }
        else if (returnAll) out.emplace_back(SIZE_MAX, false, false, static_cast<int64_t>(p));
    }
    return out;
}
}  // namespace seeding

// ---- ASCII -> 4-bit codes on the host (the layout pm_place_packed takes, see include/panmap_b200.h) ----
namespace {
struct PackLut {
    uint8_t lo[256];
    PackLut() { for (int c = 0; c < 256; ++c) { uint8_t v = 4; switch (c) { case 'A': case 'a': v = 0; break; case 'C': case 'c': v = 1; break; case 'G': case 'g': v = 2; break; case 'T': case 't': v = 3; break; default: break; } lo[c] = v; } }
};
const PackLut kPackLut;
// one read: len bases -> ceil(len / 32) chunks at dst (slots past the end hold 4)
inline void packOneRead(const unsigned char* src, uint64_t len, unsigned char* dst) {
    const uint64_t pairs = len / 2;
    for (uint64_t b = 0; b < pairs; ++b) dst[b] = static_cast<unsigned char>(kPackLut.lo[src[2 * b]] | (kPackLut.lo[src[2 * b + 1]] << 4));
    uint64_t b = pairs;
    const uint64_t bytes = ((len + 31) / 32) * 16;
    if (len & 1) { dst[b] = static_cast<unsigned char>(kPackLut.lo[src[len - 1]] | 0x40); ++b; }
    for (; b < bytes; ++b) dst[b] = 0x44;
}
}  // namespace

extern "C" uint64_t pm_packed_chunks(const uint64_t* read_offsets, uint64_t n_reads) {
    uint64_t c = 0;
    if (!read_offsets) return 0;
    for (uint64_t i = 0; i < n_reads; ++i) c += (read_offsets[i + 1] - read_offsets[i] + 31) >> 5;
    return c;
}
extern "C" int pm_pack_reads(const char* reads, const uint64_t* read_offsets, uint64_t n_reads, void* packed_out, int threads) {
    if (!read_offsets || (!reads && n_reads) || (!packed_out && n_reads)) return PM_ERR_INVALID;
    size_t nT = threads > 0 ? static_cast<size_t>(threads) : std::max(1u, std::thread::hardware_concurrency());
    if (n_reads < 4096) nT = 1;
    nT = std::min<size_t>(nT, 64);
    // thread t takes reads [n t / nT, n (t+1) / nT); its first chunk is the chunk count of everything before
    std::vector<uint64_t> firstChunk(nT + 1, 0);
    for (size_t t = 0; t < nT; ++t) {
        const uint64_t r0 = n_reads * t / nT, r1 = n_reads * (t + 1) / nT;
        uint64_t c = 0;
        for (uint64_t i = r0; i < r1; ++i) c += (read_offsets[i + 1] - read_offsets[i] + 31) >> 5;
        firstChunk[t + 1] = firstChunk[t] + c;
    }
    auto work = [&](size_t t) {
        const uint64_t r0 = n_reads * t / nT, r1 = n_reads * (t + 1) / nT;
        unsigned char* dst = static_cast<unsigned char*>(packed_out) + firstChunk[t] * 16;
        for (uint64_t i = r0; i < r1; ++i) {
            const uint64_t len = read_offsets[i + 1] - read_offsets[i];
            packOneRead(reinterpret_cast<const unsigned char*>(reads) + read_offsets[i], len, dst);
            dst += ((len + 31) >> 5) * 16;
        }
    };
    std::vector<std::thread> th;
    for (size_t t = 1; t < nT; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    return PM_OK;
}

// file ingest alone (no GPU): bases + offsets of reads1 (+ reads2 interleaved); buffers are malloc'd, free with pm_free
extern "C" int pm_read_fastx(const char* reads1, const char* reads2, char** bases, uint64_t** offsets, uint64_t* n_reads, char* err, uint64_t err_cap) {
    try {
        std::string b; std::vector<uint64_t> off;
        placement::extractReadSequences(reads1 ? reads1 : "", reads2 ? reads2 : "", b, off);
        *bases = static_cast<char*>(std::malloc(b.size() + 1)); std::memcpy(*bases, b.data(), b.size());
        *offsets = static_cast<uint64_t*>(std::malloc(off.size() * sizeof(uint64_t))); std::memcpy(*offsets, off.data(), off.size() * sizeof(uint64_t));
        *n_reads = off.size() - 1;
        return PM_OK;
    } catch (const std::exception& e) {
        if (err && err_cap) { std::strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return PM_ERR_IO;
    }
}
extern "C" void pm_free(void* p) { std::free(p); }

// C entry point used by the tests / bindings to drive the C++ shim end to end
extern "C" int pm_place_files(pm_index* idx, pm_workspace* ws, const char* const* node_ids, uint64_t n_ids, const char* reads1, const char* reads2,
                              const char* out_tsv, const pm_place_params* prm, pm_place_result* res_out, char* err, uint64_t err_cap) {
    try {
        std::vector<std::string> ids(n_ids);
        for (uint64_t i = 0; i < n_ids; ++i) ids[i] = node_ids[i];
        if (n_ids == 0 && idx) {   // indexes opened from a file or a cached image carry their LiteNode ids
            const uint64_t n = pm_index_num_nodes(idx);
            if (n && pm_index_node_id(idx, 0)[0]) { ids.resize(n); for (uint64_t i = 0; i < n; ++i) ids[i] = pm_index_node_id(idx, i); }
        }
        placement::DeviceIndex D; D.index = idx; D.workspace = ws; D.nodeIds = &ids;
        placement::TraversalParams tp;
        tp.seedMaskFraction = prm ? prm->seed_mask_fraction : 0.0; tp.trimStart = prm ? prm->trim_start : 0; tp.trimEnd = prm ? prm->trim_end : 0;
        tp.minReadSupport = prm ? prm->min_read_support : -1; tp.forceLeaf = prm && prm->force_leaf; tp.dedupReads = prm && prm->dedup_reads;
        tp.minSeedQuality = prm ? prm->min_seed_quality : 0;
        placement::PlacementResult R;
        std::string out = out_tsv ? out_tsv : "";
        placement::placeLite(R, D, reads1 ? reads1 : "", reads2 ? reads2 : "", out, tp);
        if (res_out) *res_out = R.raw;   // the whole C-ABI result of the sample: scores, best nodes, tie counts (pm_get_tied has the lists), statistics
        return PM_OK;
    } catch (const std::exception& e) {
        if (err && err_cap) { std::strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return PM_ERR_INVALID;
    }
}
