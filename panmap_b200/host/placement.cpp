// placement.cpp -- host shim: file ingest + result assembly around the C ABI (see placement.hpp).
#include "placement.hpp"
#include "seeding.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
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
// pass 1 over a mapped four-line FASTQ: where every record's bases (and qualities) lie, per thread range; then the read offsets inside
// this file by prefix sums.  The bytes themselves are moved by fillFrom() to wherever the caller wants them.
struct FastqScan {
    MappedFile mf;
    struct Range { std::vector<uint64_t> seqAt, qualAt; std::vector<uint32_t> len; uint64_t bases = 0; bool bad = false; };
    std::vector<Range> R;
    std::vector<uint64_t> firstRead;   // [ranges]
    std::vector<uint64_t> off;         // [nReads + 1] offsets of the reads when this file's sequences are laid out back to back
    uint64_t nReads = 0, nBases = 0;
    explicit FastqScan(const std::string& path) : mf(path) {}
};
template <class Fn> void onThreads(size_t nT, Fn&& fn) {
    std::vector<std::thread> th;
    for (size_t t = 1; t < nT; ++t) th.emplace_back(fn, t);
    fn(0);
    for (auto& x : th) x.join();
}
size_t ingestThreads() { return std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16); }

bool scanFastq(FastqScan& S, bool wantQuals) {
    if (!S.mf.d || S.mf.size < 4) return false;
    const char* d = S.mf.d; const size_t size = S.mf.size;
    if (d[0] != '@') return false;   // gzip streams (1f 8b), FASTA ('>') and anything else: the serial parser takes them
    size_t nT = ingestThreads();
    if (size < (1u << 20)) nT = 1;
    std::vector<size_t> bounds(nT + 1, size);
    bounds[0] = 0;
    for (size_t i = 1; i < nT; ++i) bounds[i] = std::max(bounds[i - 1], fqRecordStart(d, size, (size / nT) * i));
    S.R.assign(nT, FastqScan::Range());
    onThreads(nT, [&](size_t t) {
        FastqScan::Range& r = S.R[t];
        size_t o = bounds[t];
        const size_t end = bounds[t + 1];
        const size_t guess = (end - o) / 256 + 16;
        r.seqAt.reserve(guess); r.len.reserve(guess); if (wantQuals) r.qualAt.reserve(guess);
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
            if (std::memchr(d + s0, ' ', L) || std::memchr(d + s0, '\t', L)) { r.bad = true; break; }   // kseq would drop these
            r.seqAt.push_back(s0); r.len.push_back(static_cast<uint32_t>(L)); r.bases += L;
            if (wantQuals) r.qualAt.push_back(q0);
        }
    });
    S.firstRead.assign(nT, 0);
    std::vector<uint64_t> firstBase(nT, 0);
    S.nReads = S.nBases = 0;
    for (size_t t = 0; t < nT; ++t) { if (S.R[t].bad) return false; S.firstRead[t] = S.nReads; firstBase[t] = S.nBases; S.nReads += S.R[t].len.size(); S.nBases += S.R[t].bases; }
    S.off.resize(S.nReads + 1);
    onThreads(nT, [&](size_t t) {
        uint64_t b = firstBase[t];
        const auto& len = S.R[t].len;
        uint64_t* o = S.off.data() + S.firstRead[t];
        for (size_t i = 0; i < len.size(); ++i) { o[i] = b; b += len[i]; }
    });
    S.off[S.nReads] = S.nBases;
    return true;
}
// pass 2: read i of the file goes to bases[place(i)] (and its qualities to quals[place(i)])
template <class Place> void fillFrom(const FastqScan& S, char* bases, char* quals, Place&& place) {
    const char* d = S.mf.d;
    onThreads(S.R.size(), [&](size_t t) {
        const FastqScan::Range& r = S.R[t];
        for (size_t i = 0; i < r.len.size(); ++i) {
            const uint64_t at = place(S.firstRead[t] + i);
            std::memcpy(bases + at, d + r.seqAt[i], r.len[i]);
            if (quals) std::memcpy(quals + at, d + r.qualAt[i], r.len[i]);
        }
    });
}

bool readFastqParallel(const std::string& path, FlatReads& out, FlatReads* quals) {
    FastqScan S(path);
    if (!scanFastq(S, quals != nullptr)) return false;
    out.data.resize(S.nBases);
    if (quals) quals->data.resize(S.nBases);
    fillFrom(S, out.data.data(), quals ? quals->data.data() : nullptr, [&](uint64_t i) { return S.off[i]; });
    out.off = std::move(S.off);
    if (quals) quals->off = out.off;
    return true;
}

void readFastx(const std::string& path, FlatReads& out, FlatReads* quals) {
    if (!readFastqParallel(path, out, quals)) { out = FlatReads(); if (quals) *quals = FlatReads(); readFastxSerial(path, out, quals); }
}

// A sample's files -> (bases, offsets[, qualities]) in buffers obtained from `alloc(bytes, nReads)` -- the workspace's pinned staging
// buffers in placeLite -- with R1/R2 pairs interleaved (seeding::perfect_shuffle, seeding.hpp:33-43).  Plain four-line FASTQ is copied from
// the mapped file(s) straight to its final place by all threads (one copy per base, no intermediate strings); gz / FASTA / multi-line
// files go through the kseq-style parser, the two files of a pair on two threads (inflate is serial inside one gzip stream).
struct Landing { char* bases = nullptr; uint64_t* off = nullptr; char* quals = nullptr; };
struct IngestClock {   // PM_INGEST_TIMING=1: phase times of the parser on stderr (tools/ingest_probe.py)
    const bool on = std::getenv("PM_INGEST_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[pm ingest] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
template <class Alloc> uint64_t ingestSample(const std::string& path1, const std::string& path2, bool wantQuals, Alloc&& alloc, Landing& L) {
    const bool paired = !path2.empty();
    IngestClock clk;
    {
        FastqScan A(path1);
        if (scanFastq(A, wantQuals)) {
            clk.lap("scan (record boundaries)");
            if (!paired) {
                L = alloc(A.nBases, A.nReads);
                clk.lap("landing buffers");
                std::memcpy(L.off, A.off.data(), (A.nReads + 1) * sizeof(uint64_t));
                fillFrom(A, L.bases, L.quals, [&](uint64_t i) { return A.off[i]; });
                clk.lap("fill (bases to their place)");
                return A.nReads;
            }
            FastqScan B(path2);
            if (scanFastq(B, wantQuals)) {
                if (B.nReads != A.nReads) throw std::runtime_error("File " + path2 + " does not contain the same number of reads as " + path1);
                const uint64_t n = A.nReads;
                L = alloc(A.nBases + B.nBases, 2 * n);
                onThreads(ingestThreads(), [&](size_t t) {
                    const size_t nT = ingestThreads();
                    for (uint64_t i = n * t / nT; i < n * (t + 1) / nT; ++i) { L.off[2 * i] = A.off[i] + B.off[i]; L.off[2 * i + 1] = A.off[i + 1] + B.off[i]; }
                });
                L.off[2 * n] = A.nBases + B.nBases;
                fillFrom(A, L.bases, L.quals, [&](uint64_t i) { return A.off[i] + B.off[i]; });
                fillFrom(B, L.bases, L.quals, [&](uint64_t i) { return A.off[i + 1] + B.off[i]; });
                return 2 * n;
            }
        }
    }
    // general route
    FlatReads a, qa, b, qb;
    if (paired) {
        std::exception_ptr err;
        std::thread other([&] { try { readFastx(path2, b, wantQuals ? &qb : nullptr); } catch (...) { err = std::current_exception(); } });
        try { readFastx(path1, a, wantQuals ? &qa : nullptr); } catch (...) { other.join(); throw; }
        other.join();
        if (err) std::rethrow_exception(err);
        if (b.size() != a.size()) throw std::runtime_error("File " + path2 + " does not contain the same number of reads as " + path1);
    } else readFastx(path1, a, wantQuals ? &qa : nullptr);
    const uint64_t n = a.size();
    if (!paired) {
        L = alloc(a.data.size(), n);
        std::memcpy(L.off, a.off.data(), (n + 1) * sizeof(uint64_t));
        std::memcpy(L.bases, a.data.data(), a.data.size());
        if (wantQuals) std::memcpy(L.quals, qa.data.data(), qa.data.size());
        return n;
    }
    L = alloc(a.data.size() + b.data.size(), 2 * n);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t o0 = a.off[i] + b.off[i], o1 = a.off[i + 1] + b.off[i];
        L.off[2 * i] = o0; L.off[2 * i + 1] = o1;
        std::memcpy(L.bases + o0, &a.data[a.off[i]], a.off[i + 1] - a.off[i]);
        std::memcpy(L.bases + o1, &b.data[b.off[i]], b.off[i + 1] - b.off[i]);
        if (wantQuals) {
            std::memcpy(L.quals + o0, &qa.data[a.off[i]], a.off[i + 1] - a.off[i]);
            std::memcpy(L.quals + o1, &qb.data[b.off[i]], b.off[i + 1] - b.off[i]);
        }
    }
    L.off[2 * n] = a.data.size() + b.data.size();
    return 2 * n;
}

}  // namespace

namespace placement {

void extractReadSequences(const std::string& readPath1, const std::string& readPath2, std::string& bases, std::vector<uint64_t>& offsets,
                          std::string* quals) {
    Landing L;
    ingestSample(readPath1, readPath2, quals != nullptr, [&](uint64_t bytes, uint64_t nReads) {
        bases.resize(bytes); offsets.assign(nReads + 1, 0);
        if (quals) quals->resize(bytes);
        Landing x; x.bases = bases.data(); x.off = offsets.data(); x.quals = quals ? quals->data() : nullptr;
        return x;
    }, L);
}

void placeLite(PlacementResult& result, DeviceIndex& index, const std::string& reads1, const std::string& reads2, std::string& outputPath,
               const TraversalParams& params) {
    if (!index.index || !index.workspace) throw std::runtime_error("placeLite: device index not initialised");
    const bool quality = params.minSeedQuality > 0;   // placement.cpp:1131-1137: the quality strings are loaded only then
    // the sample lands in the workspace's pinned staging buffers: the copy engine reads them directly (pageable std::strings would be staged
    // through a bounce buffer at a fraction of the PCIe rate)
    Landing in;
    uint64_t nReads = 0, zero = 0;
    if (!reads1.empty())
        nReads = ingestSample(reads1, reads2, quality, [&](uint64_t bytes, uint64_t n) {
            Landing x;
            if (pm_workspace_staging(index.workspace, bytes, n, quality ? 1 : 0, &x.bases, &x.off, &x.quals) != PM_OK) throw std::runtime_error(pm_last_error());
            return x;
        }, in);
    else in.off = &zero;
    pm_place_params p{};
    p.trim_start = params.trimStart; p.trim_end = params.trimEnd; p.min_read_support = params.minReadSupport;
    p.dedup_reads = params.dedupReads ? 1 : 0; p.force_leaf = params.forceLeaf ? 1 : 0; p.skip_node_index = PM_NONE;
    p.seed_mask_fraction = params.seedMaskFraction; p.want_node_scores = params.store_diagnostics ? 1 : 0;
    p.min_seed_quality = quality && nReads > 0 ? params.minSeedQuality : 0;   // no reads: `!allReadQualities.empty()` fails, default path
    pm_place_result r{};
    IngestClock clk;
    const int rc = p.min_seed_quality > 0 ? pm_place_quality(index.workspace, in.bases, in.quals, in.off, nReads, &p, &r)
                                          : pm_place(index.workspace, in.bases, in.off, nReads, &p, &r);
    clk.lap("pm_place (staging -> result)");
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
