// placement.cpp -- host shim: file ingest + result assembly around the C ABI (see placement.hpp).
#include "placement.hpp"
#include "seeding.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <cstdlib>
#include <stdexcept>

namespace {

// kseq semantics (FASTA and FASTQ, multi-line sequences, gz or plain through zlib's transparent gzread)
// quals (optional): the quality string of every record, 'I' * length when the record has none (extractFullFastqData, placement.cpp:199-238)
void readFastx(const std::string& path, std::vector<std::string>& out, std::vector<std::string>* quals = nullptr) {
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
    auto appendLine = [&](std::string& dst, size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) { const char ch = data[i]; if (ch != '\r' && ch != ' ' && ch != '\t') dst.push_back(ch); }
    };
    while (p < N && data[p] != '>' && data[p] != '@') p = lineEnd(p) + 1;
    while (p < N) {
        p = lineEnd(p) + 1;  // header line
        std::string seq;
        while (p < N && data[p] != '>' && data[p] != '@' && data[p] != '+') { const size_t e = lineEnd(p); appendLine(seq, p, e); p = e + 1; }
        std::string qual;
        if (p < N && data[p] == '+') {
            p = lineEnd(p) + 1;
            while (p < N && qual.size() < seq.size()) { const size_t e = lineEnd(p); appendLine(qual, p, e); p = e + 1; }
            while (p < N && data[p] != '>' && data[p] != '@') p = lineEnd(p) + 1;
            if (quals && qual.size() != seq.size()) break;   // kseq_read returns -2 (truncated quality): the reference's loop ends here
        }
        if (quals) quals->push_back(qual.empty() ? std::string(seq.size(), 'I') : std::move(qual));
        out.push_back(std::move(seq));
    }
}

}  // namespace

namespace placement {

void extractReadSequences(const std::string& readPath1, const std::string& readPath2, std::string& bases, std::vector<uint64_t>& offsets,
                          std::string* quals) {
    std::vector<std::string> r, q;
    readFastx(readPath1, r, quals ? &q : nullptr);
    if (!readPath2.empty()) {
        const size_t fwd = r.size();
        readFastx(readPath2, r, quals ? &q : nullptr);
        if (r.size() != fwd * 2) throw std::runtime_error("File " + readPath2 + " does not contain the same number of reads as " + readPath1);
        auto shuffle = [&](std::vector<std::string>& v) {   // seeding::perfect_shuffle (seeding.hpp:33-43)
            std::vector<std::string> canvas(v.size());
            for (size_t i = 0; i < fwd; ++i) { canvas[2 * i] = std::move(v[i]); canvas[2 * i + 1] = std::move(v[i + fwd]); }
            v.swap(canvas);
        };
        shuffle(r);
        if (quals) shuffle(q);
    }
    offsets.assign(r.size() + 1, 0);
    size_t tot = 0;
    for (size_t i = 0; i < r.size(); ++i) { tot += r[i].size(); offsets[i + 1] = tot; }
    bases.clear(); bases.reserve(tot);
    for (auto& s : r) bases += s;
    if (quals) { quals->clear(); quals->reserve(tot); for (auto& s : q) *quals += s; }
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
        placement::DeviceIndex D; D.index = idx; D.workspace = ws; D.nodeIds = &ids;
        placement::TraversalParams tp;
        tp.seedMaskFraction = prm ? prm->seed_mask_fraction : 0.0; tp.trimStart = prm ? prm->trim_start : 0; tp.trimEnd = prm ? prm->trim_end : 0;
        tp.minReadSupport = prm ? prm->min_read_support : -1; tp.forceLeaf = prm && prm->force_leaf; tp.dedupReads = prm && prm->dedup_reads;
        tp.minSeedQuality = prm ? prm->min_seed_quality : 0;
        placement::PlacementResult R;
        std::string out = out_tsv ? out_tsv : "";
        placement::placeLite(R, D, reads1 ? reads1 : "", reads2 ? reads2 : "", out, tp);
        if (res_out) {
            const double sc[5] = {R.bestLogRawScore, R.bestLogCosineScore, R.bestContainmentScore, R.bestWeightedContainmentScore, R.bestLogContainmentScore};
            const uint32_t ix[5] = {R.bestLogRawNodeIndex, R.bestLogCosineNodeIndex, R.bestContainmentNodeIndex, R.bestWeightedContainmentNodeIndex, R.bestLogContainmentNodeIndex};
            for (int m = 0; m < 5; ++m) { res_out->best_score[m] = sc[m]; res_out->best_index[m] = ix[m]; }
            res_out->total_reads = static_cast<uint64_t>(R.totalReadsProcessed); res_out->read_unique_seed_count = R.readUniqueSeedCount;
            res_out->total_read_seed_frequency = R.totalReadSeedFrequency; res_out->read_magnitude = R.readMagnitude;
        }
        return PM_OK;
    } catch (const std::exception& e) {
        if (err && err_cap) { std::strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return PM_ERR_INVALID;
    }
}
