// C ABI of libflatgfa: the reference's flatgfa-c surface (flatgfa-c/src/lib.rs:62-172)
// plus the node-depth additions declared in include/flatgfa.h.
#include <climits>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#include "../../include/fgfa_depth.h"
#include "../../include/flatgfa.h"
#include "file.hpp"
#include "ops_depth.hpp"
#include "ops_window_depth.hpp"
#include "parse.hpp"
#include "print.hpp"

// lib.rs:16: the opaque store behind flatgfa_t.  Either a parsed heap store or a
// zero-copy view of a mapped .flatgfa file.
struct CStore {
    flatgfa::HeapGFAStore heap;
    std::unique_ptr<flatgfa::MappedFile> map;
    flatgfa::FlatGFA gfa;
};

struct flatbed {
    flatgfa::HeapBEDStore store;
};

namespace {
thread_local std::string g_err;
flatgfa_string_t null_string() { return flatgfa_string_t{nullptr, 0}; }
flatgfa_string_t to_string(flatgfa::Pool<uint8_t> p) {
    return flatgfa_string_t{p.data, (int)p.len()};
}
}  // namespace

namespace {
int text_out(const std::string& s, char** out, size_t* out_len) {
    char* buf = static_cast<char*>(std::malloc(s.size() + 1));
    if (!buf) return FGFA_ERR_NOMEM;
    std::memcpy(buf, s.data(), s.size());
    buf[s.size()] = 0;
    *out = buf;
    *out_len = s.size();
    return FGFA_OK;
}
// The FGFA_ERR_* a C entry point returns for an exception: the code the device ABI reported when the
// exception carries one (so flatgfa_path_depth / _interval_depth / _window_depth / _bed_depth agree with
// flatgfa_seg_depth), FGFA_ERR_INVALID_ARG for host-side complaints (unknown path, zero window ...).
int code_of(const std::exception& e) {
    g_err = e.what();
    if (const auto* fe = dynamic_cast<const flatgfa::Error*>(&e))
        if (fe->code) return fe->code;
    if (std::strstr(e.what(), "no CUDA device") || std::strstr(e.what(), "no usable CUDA device")) return FGFA_ERR_NO_DEVICE;
    return FGFA_ERR_INVALID_ARG;
}
}  // namespace

extern "C" {

const char* flatgfa_last_error(void) { return g_err.c_str(); }

flatgfa_t flatgfa_parse(const char* filename) {
    if (!filename) { g_err = "null filename"; return nullptr; }
    try {
        flatgfa::MappedFile f(filename);                                   // lib.rs:65
        std::unique_ptr<CStore> s(new CStore());
        s->heap = flatgfa::Parser::parse_mem(f.data(), f.size());          // lib.rs:66
        s->gfa = s->heap.view();
        return s.release();                                                // lib.rs:25-27
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

flatgfa_t flatgfa_load(const char* filename) {
    if (!filename) { g_err = "null filename"; return nullptr; }
    try {
        std::unique_ptr<CStore> s(new CStore());
        s->map.reset(new flatgfa::MappedFile(filename));
        s->gfa = flatgfa::file::view_or_throw(s->map->data(), s->map->size());
        return s.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void flatgfa_free(flatgfa_t gfa) { delete gfa; }   // lib.rs:71-76 (null-safe)

uint32_t flatgfa_get_segment_count(flatgfa_t gfa) { return (uint32_t)gfa->gfa.segs.len(); }

flatgfa_string_t flatgfa_get_seq(flatgfa_t gfa, uint32_t segment_id) {
    const auto& g = gfa->gfa;
    if ((size_t)segment_id >= g.segs.len()) return null_string();          // lib.rs:95-97
    try {
        return to_string(g.get_seq(g.segs.data[segment_id]));
    } catch (const std::exception&) {
        return null_string();
    }
}

uint32_t flatgfa_path_count(flatgfa_t gfa) { return (uint32_t)gfa->gfa.paths.len(); }

flatgfa_string_t flatgfa_get_path_name(flatgfa_t gfa, uint32_t path_index) {
    const auto& g = gfa->gfa;
    if ((size_t)path_index >= g.paths.len()) return null_string();         // lib.rs:121-122
    try {
        return to_string(g.get_path_name(g.paths.data[path_index]));
    } catch (const std::exception&) {
        return null_string();
    }
}

uint32_t flatgfa_get_path_step_count(flatgfa_t gfa, uint32_t path_index) {
    const auto& g = gfa->gfa;
    if ((size_t)path_index >= g.paths.len()) return UINT32_MAX;            // lib.rs:133-134
    return (uint32_t)g.paths.data[path_index].step_count();
}

bool flatgfa_get_step(flatgfa_t gfa, uintptr_t path_index, uintptr_t step_index,
                      flatgfa_handle_t* out) {
    const auto& g = gfa->gfa;
    if (path_index >= g.paths.len()) return false;                         // lib.rs:157-160
    try {
        const auto steps = g.get_path_steps(g.paths.data[path_index]);
        if (step_index >= steps.len()) return false;                       // lib.rs:161-164
        const flatgfa::Handle h = steps.data[step_index];
        out->segment_id = h.segment();                                     // lib.rs:167
        out->is_forward = h.is_forward();                                  // lib.rs:168
        return true;
    } catch (const std::exception&) {
        return false;
    }
}

int flatgfa_seg_depth(flatgfa_t gfa, uint64_t* depth, uint64_t* uniq) {
    if (!gfa || (!depth && gfa->gfa.segs.len())) return FGFA_ERR_INVALID_ARG;
    const auto& g = gfa->gfa;
    flatgfa::ops::depth::PoolArrays a;
    int rc = flatgfa::ops::depth::pool_arrays_of(g, &a);
    if (!rc)
        rc = fgfa_seg_depth_with_uniq_steps(a.steps, a.n_steps, a.start.data(), a.end.data(), a.n_paths, a.n_segs, depth, uniq);
    if (rc) {
        g_err = fgfa_strerror(rc);
        const char* detail = fgfa_last_error();
        if (detail && *detail) g_err += std::string(": ") + detail;
    }
    return rc;
}

int flatgfa_format_seg_depth(flatgfa_t gfa, const uint64_t* depth, const uint64_t* uniq, char** out,
                             size_t* out_len) {
    if (!gfa || !out || !out_len) return FGFA_ERR_INVALID_ARG;
    const size_t n = gfa->gfa.segs.len();
    if (n && (!depth || !uniq)) return FGFA_ERR_INVALID_ARG;
    char* buf = flatgfa::ops::depth::seg_depth_table(gfa->gfa, depth, uniq, out_len);
    if (!buf) return FGFA_ERR_NOMEM;
    *out = buf;
    return FGFA_OK;
}

int flatgfa_path_depth(flatgfa_t gfa, const uint32_t* path_ids, uint32_t n, uint64_t* lengths,
                       double* mean_depths) {
    if (!gfa || (n && (!lengths || !mean_depths))) return FGFA_ERR_INVALID_ARG;
    try {
        std::vector<uint32_t> ids(n);
        for (uint32_t i = 0; i < n; ++i) {
            ids[i] = path_ids ? path_ids[i] : i;
            if (ids[i] >= gfa->gfa.paths.len()) return FGFA_ERR_INVALID_ARG;
        }
        auto ld = flatgfa::ops::depth::path_depth(gfa->gfa, ids);
        for (uint32_t i = 0; i < n; ++i) { lengths[i] = ld.first[i]; mean_depths[i] = ld.second[i]; }
        return FGFA_OK;
    } catch (const std::exception& e) {
        return code_of(e);
    }
}

int flatgfa_format_path_depth(flatgfa_t gfa, const uint32_t* path_ids, uint32_t n, const uint64_t* lengths,
                              const double* mean_depths, char** out, size_t* out_len) {
    if (!gfa || !out || !out_len || (n && (!lengths || !mean_depths))) return FGFA_ERR_INVALID_ARG;
    std::vector<uint32_t> ids(n);
    for (uint32_t i = 0; i < n; ++i) {
        ids[i] = path_ids ? path_ids[i] : i;
        if (ids[i] >= gfa->gfa.paths.len()) return FGFA_ERR_INVALID_ARG;
    }
    flatgfa::ops::depth::PathDepth t{gfa->gfa, std::vector<uint64_t>(lengths, lengths + n),
                                     std::vector<double>(mean_depths, mean_depths + n), ids};
    std::string s;
    try {
        t.emit(s);
    } catch (const std::exception& e) {
        g_err = e.what();
        return FGFA_ERR_INVALID_ARG;
    }
    char* buf = static_cast<char*>(std::malloc(s.size() + 1));
    if (!buf) return FGFA_ERR_NOMEM;
    std::memcpy(buf, s.data(), s.size());
    buf[s.size()] = 0;
    *out = buf;
    *out_len = s.size();
    return FGFA_OK;
}

flatbed_t flatbed_parse_mem(const uint8_t* bed_text, size_t bed_len) {
    if (bed_len && !bed_text) { g_err = "null text"; return nullptr; }
    try {
        std::unique_ptr<flatbed> b(new flatbed);
        b->store = flatgfa::BEDParser::parse_mem(bed_text, bed_len);
        return b.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

flatbed_t flatbed_make_windows(const uint8_t* name, size_t name_len, uint64_t start, uint64_t end, uint64_t size) {
    if (size == 0 || (name_len && !name)) { g_err = "window size must be positive"; return nullptr; }
    try {
        std::unique_ptr<flatbed> b(new flatbed);
        b->store = flatgfa::ops::window_depth::Windows{{name, name_len}, start, end, size}.as_bed();
        return b.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void flatbed_free(flatbed_t bed) { delete bed; }

uint64_t flatbed_entry_count(flatbed_t bed) { return bed ? bed->store.entries.size() : 0; }

bool flatbed_get_entry(flatbed_t bed, uint64_t i, flatgfa_string_t* name, uint64_t* start, uint64_t* end) {
    if (!bed || i >= bed->store.entries.size()) return false;
    const flatgfa::BEDEntry& e = bed->store.entries[i];
    if (name) {
        name->data = bed->store.name_data.data() + e.name.start;
        name->len = (int)(e.name.end - e.name.start);
    }
    if (start) *start = e.start;
    if (end) *end = e.end;
    return true;
}

int flatgfa_interval_depth(flatgfa_t gfa, flatbed_t bed, double* depths) {
    if (!gfa || !bed || (!depths && !bed->store.entries.empty())) return FGFA_ERR_INVALID_ARG;
    try {
        auto d = flatgfa::ops::window_depth::bed_depth(gfa->gfa, bed->store.view());
        std::memcpy(depths, d.data(), d.size() * sizeof(double));
        return FGFA_OK;
    } catch (const std::exception& e) {
        return code_of(e);
    }
}

int flatgfa_window_depth(flatgfa_t gfa, const char* path_name, uint64_t window_size, char** out, size_t* out_len) {
    if (!gfa || !path_name || !out || !out_len) return FGFA_ERR_INVALID_ARG;
    try {
        const int64_t path = gfa->gfa.find_path(reinterpret_cast<const uint8_t*>(path_name), std::strlen(path_name));
        if (path < 0) { g_err = "path not found"; return FGFA_ERR_INVALID_ARG; }              // cmds.rs:489
        auto wd = flatgfa::ops::window_depth::window_depth(gfa->gfa, (uint32_t)path, window_size);
        flatgfa::ops::window_depth::IntervalDepth t{wd.first.view(), std::move(wd.second)};
        std::string s;
        t.emit(s);
        return text_out(s, out, out_len);
    } catch (const std::exception& e) {
        return code_of(e);
    }
}

int flatgfa_bed_depth(flatgfa_t gfa, const uint8_t* bed_text, size_t bed_len, char** out, size_t* out_len) {
    if (!gfa || (bed_len && !bed_text) || !out || !out_len) return FGFA_ERR_INVALID_ARG;
    try {
        const flatgfa::HeapBEDStore store = flatgfa::BEDParser::parse_mem(bed_text, bed_len);   // cmds.rs:248-249
        auto depths = flatgfa::ops::window_depth::bed_depth(gfa->gfa, store.view());            // cmds.rs:250
        flatgfa::ops::window_depth::IntervalDepth t{store.view(), std::move(depths)};
        std::string s;
        t.emit(s);
        return text_out(s, out, out_len);
    } catch (const std::exception& e) {
        return code_of(e);
    }
}

flatgfa_t flatgfa_parse_mem(const uint8_t* gfa_text, size_t len) {
    if (len && !gfa_text) { g_err = "null text"; return nullptr; }
    try {
        std::unique_ptr<CStore> s(new CStore());
        s->heap = flatgfa::Parser::parse_mem(gfa_text, len);
        s->gfa = s->heap.view();
        return s.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

size_t flatgfa_image_size(flatgfa_t gfa) { return gfa ? flatgfa::file::size(gfa->gfa) : 0; }

int flatgfa_dump_mem(flatgfa_t gfa, uint8_t* buf, size_t cap) {
    if (!gfa || !buf || cap < flatgfa::file::size(gfa->gfa)) return FGFA_ERR_INVALID_ARG;
    try {
        flatgfa::file::dump(gfa->gfa, buf);
        return FGFA_OK;
    } catch (const std::exception& e) {
        g_err = e.what();
        return FGFA_ERR_INVALID_ARG;
    }
}

int flatgfa_format_gfa(flatgfa_t gfa, char** out, size_t* out_len) {
    if (!gfa || !out || !out_len) return FGFA_ERR_INVALID_ARG;
    try {
        std::string s;
        flatgfa::print::gfa(gfa->gfa, s);
        return text_out(s, out, out_len);
    } catch (const std::exception& e) {
        g_err = e.what();
        return FGFA_ERR_INVALID_ARG;
    }
}

int flatgfa_dump(flatgfa_t gfa, const char* filename) {
    if (!gfa || !filename) return FGFA_ERR_INVALID_ARG;
    try {
        std::vector<uint8_t> buf(flatgfa::file::size(gfa->gfa));
        flatgfa::file::dump(gfa->gfa, buf.data());
        flatgfa::write_file(filename, buf.data(), buf.size());
        return FGFA_OK;
    } catch (const std::exception& e) {
        g_err = e.what();
        return FGFA_ERR_INVALID_ARG;
    }
}

}  // extern "C"
