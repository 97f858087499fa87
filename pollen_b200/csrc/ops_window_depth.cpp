#include "ops_window_depth.hpp"

#include <cstring>

#include "../../include/fgfa_depth.h"
#include "ops_depth.hpp"

namespace flatgfa {
namespace ops {
namespace window_depth {

namespace {
[[noreturn]] void raise(int rc) {
    std::string m = fgfa_strerror(rc);
    const char* detail = fgfa_last_error();
    if (detail && *detail) m += std::string(": ") + detail;
    throw Error(m, rc);
}

struct Arrays {
    std::vector<uint32_t> s, e, len, aligned;
    const uint32_t* steps = nullptr;
    uint32_t n_paths = 0, n_segs = 0;
};
Arrays arrays_of(const FlatGFA& gfa) {
    if (gfa.segs.len() > 0x7FFFFFFFull || gfa.paths.len() > 0xFFFFFFFFull || gfa.steps.len() > 0xFFFFFFFFull)
        raise(FGFA_ERR_TOO_LARGE);
    Arrays a;
    a.n_paths = (uint32_t)gfa.paths.len();
    a.n_segs = (uint32_t)gfa.segs.len();
    a.s.resize(a.n_paths);
    a.e.resize(a.n_paths);
    a.len.resize(a.n_segs);
    for (uint32_t p = 0; p < a.n_paths; ++p) {
        a.s[p] = gfa.paths.data[p].steps.start;
        a.e[p] = gfa.paths.data[p].steps.end;
    }
    for (uint32_t i = 0; i < a.n_segs; ++i) a.len[i] = (uint32_t)gfa.segs.data[i].len();   // flatgfa.rs:84-89
    a.steps = reinterpret_cast<const uint32_t*>(gfa.steps.data);
    if (reinterpret_cast<uintptr_t>(a.steps) & 3u) {
        a.aligned.resize(gfa.steps.len());
        std::memcpy(a.aligned.data(), gfa.steps.data, gfa.steps.len() * 4);
        a.steps = a.aligned.data();
    }
    return a;
}
}  // namespace

void Windows::emit(std::string& out) const {
    if (size == 0) throw Error("window size must be positive");
    uint64_t pos = start;
    while (pos < end) {
        uint64_t stop = pos + size;
        if (stop < pos || stop > end) stop = end;
        out.append(reinterpret_cast<const char*>(name.data), name.len());
        out += '\t';
        out += std::to_string(pos);
        out += '\t';
        out += std::to_string(stop);
        out += '\n';
        pos = stop;
    }
}

void Windows::emit_bed(HeapBEDStore& store) const {
    if (size == 0) throw Error("window size must be positive");
    const Span nm = HeapGFAStore::add_slice(store.name_data, name.data, name.len());   // one shared name, :42
    store.entries.reserve(store.entries.size() + len());
    uint64_t pos = start;
    while (pos < end) {
        uint64_t stop = pos + size;
        if (stop < pos || stop > end) stop = end;
        store.entries.push_back(BEDEntry{nm, pos, stop});
        pos = stop;
    }
}

HeapBEDStore Windows::as_bed() const {
    HeapBEDStore s;
    emit_bed(s);
    return s;
}

size_t Windows::len() const {
    if (size == 0) throw Error("attempt to divide by zero");
    const uint64_t span = end - start;
    return (size_t)(span / size + (span % size ? 1 : 0));
}

void IntervalDepth::emit(std::string& out) const {
    for (size_t i = 0; i < intervals.entries.len(); ++i) {                  // :163-172
        const BEDEntry& e = intervals.entries.data[i];
        const Pool<uint8_t> nm = intervals.get_name_of_entry(e);
        out.append(reinterpret_cast<const char*>(nm.data), nm.len());
        out += '\t';
        out += std::to_string((uint64_t)e.start);
        out += '\t';
        out += std::to_string((uint64_t)e.end);
        out += '\t';
        out += depth::format_float(depths[i], 4);
        out += '\n';
    }
}

void IntervalDepth::emit(FILE* f) const {
    std::string s;
    emit(s);
    std::fwrite(s.data(), 1, s.size(), f);
}

std::pair<HeapBEDStore, std::vector<double>> window_depth(const FlatGFA& gfa, uint32_t path, uint64_t window_size) {
    if (window_size == 0) throw Error("window size must be positive");
    const Arrays a = arrays_of(gfa);
    if (path >= a.n_paths) throw Error("pool index out of bounds");
    double* d = nullptr;
    uint64_t m = 0, total = 0;
    int rc = fgfa_window_depth_steps(a.steps, gfa.steps.len(), a.s.data(), a.e.data(), a.n_paths, a.len.data(),
                                     a.n_segs, path, window_size, &d, &m, &total);
    if (rc) raise(rc);
    std::vector<double> depths(d, d + m);
    fgfa_free(d);
    HeapBEDStore windows = Windows{gfa.get_path_name(gfa.paths[path]), 0, total, window_size}.as_bed();   // :188-194
    return {std::move(windows), std::move(depths)};
}

std::vector<double> bed_depth(const FlatGFA& gfa, const FlatBED& intervals) {
    if (intervals.entries.len() == 0) throw Error("index out of bounds: the len is 0 but the index is 0");   // :207
    const Pool<uint8_t> nm = intervals.get_name_of_entry(intervals.entries.data[0]);
    const int64_t path = gfa.find_path(nm.data, nm.len());
    if (path < 0) throw Error("path not found in graph");                                                     // :208
    const Arrays a = arrays_of(gfa);
    const size_t m = intervals.entries.len();
    std::vector<uint64_t> ws(m), we(m);
    for (size_t i = 0; i < m; ++i) {
        ws[i] = intervals.entries.data[i].start;
        we[i] = intervals.entries.data[i].end;
    }
    std::vector<double> depths(m);
    int rc = fgfa_interval_depth_steps(a.steps, gfa.steps.len(), a.s.data(), a.e.data(), a.n_paths, a.len.data(),
                                       a.n_segs, (uint32_t)path, ws.data(), we.data(), m, depths.data());
    if (rc) raise(rc);
    return depths;
}

}  // namespace window_depth
}  // namespace ops
}  // namespace flatgfa
