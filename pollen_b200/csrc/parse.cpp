#include "parse.hpp"

#include "../../include/fgfa_depth.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace flatgfa {

void NameMap::insert(uint64_t name, uint32_t id) {
    // namemap.rs:17-25 (release-build arithmetic: `name - 1` wraps for name == 0).
    if ((name - 1) == sequential_max_ && (name - 1) == (uint64_t)id) {
        sequential_max_ += 1;
    } else {
        others_[name] = id;   // HashMap::insert overwrites
    }
}

uint32_t NameMap::get(uint64_t name) const {
    if (name <= sequential_max_) return (uint32_t)(name - 1);   // namemap.rs:28-29 (`as u32` truncates)
    auto it = others_.find(name);
    if (it == others_.end()) throw Error("unknown segment name " + std::to_string(name));
    return it->second;
}

namespace {

struct Cursor {
    const uint8_t* p;
    size_t n;
    bool empty() const { return n == 0; }
    void advance(size_t k) { p += k; n -= k; }
};

// atoi::FromRadix10 (unchecked): leading decimal digits, wrapping arithmetic.
template <typename T>
T parse_num(Cursor& c) {   // gfaline.rs:158-164
    T v = 0;
    size_t used = 0;
    while (used < c.n && c.p[used] >= '0' && c.p[used] <= '9') {
        v = (T)(v * 10 + (T)(c.p[used] - '0'));
        ++used;
    }
    if (used == 0) throw Error("expected number");
    c.advance(used);
    return v;
}

void parse_byte(Cursor& c, uint8_t b) {   // gfaline.rs:150-155
    if (c.empty() || c.p[0] != b) throw Error("expected byte");
    c.advance(1);
}

// gfaline.rs:133-147: up to the next tab (consumed) or the end of the line.
Cursor parse_field(Cursor& c) {
    const void* t = c.n ? std::memchr(c.p, '\t', c.n) : nullptr;
    size_t end = t ? (size_t)((const uint8_t*)t - c.p) : c.n;
    Cursor f{c.p, end};
    if (end == c.n) c.advance(c.n); else c.advance(end + 1);
    return f;
}

bool parse_orient(Cursor& c) {   // gfaline.rs:167-177; true = forward
    if (c.empty()) throw Error("expected orientation");
    const uint8_t b = c.p[0];
    if (b != '+' && b != '-') throw Error("expected orient");
    c.advance(1);
    return b == '+';
}

AlignOp parse_align_op(Cursor& c) {   // gfaline.rs:180-190; flatgfa.rs:229-234
    const uint32_t len = parse_num<uint32_t>(c);
    if (c.empty()) throw Error("expected align op");   // reference: index panic
    uint32_t op;
    switch (c.p[0]) {   // enum order Match, Gap, Insertion, Deletion (flatgfa.rs:213-218)
        case 'M': op = 0; break;
        case 'N': op = 1; break;
        case 'D': op = 3; break;   // gfaline.rs:185: 'D' => Deletion
        case 'I': op = 2; break;   // gfaline.rs:186: 'I' => Insertion
        default: throw Error("expected align op");
    }
    if (len & ~0xFFu) throw Error("length too large");   // flatgfa.rs:231
    c.advance(1);
    return AlignOp{(len << 8) | op};
}

std::vector<AlignOp> parse_align(Cursor& c) {   // gfaline.rs:195-204
    std::vector<AlignOp> a;
    while (!c.empty() && c.p[0] >= '0' && c.p[0] <= '9') a.push_back(parse_align_op(c));
    return a;
}

std::vector<std::vector<AlignOp>> parse_maybe_overlap_list(Cursor& c) {   // gfaline.rs:103-126
    std::vector<std::vector<AlignOp>> out;
    if (c.n == 1 && c.p[0] == '*') { c.advance(1); return out; }
    while (!c.empty()) {
        out.push_back(parse_align(c));
        if (!c.empty()) parse_byte(c, ',');
    }
    return out;
}

struct Builder {
    HeapGFAStore flat;
    NameMap seg_ids;

    static Cursor body(const uint8_t* line, size_t n) {   // gfaline.rs:37-41
        if (n < 2 || line[1] != '\t') throw Error("expected marker and tab");
        return Cursor{line + 2, n - 2};
    }

    void header(const uint8_t* line, size_t n) {   // gfaline.rs:52-54; parse.rs:44-46
        Cursor c = body(line, n);
        flat.add_header(c.p, c.n);
    }
    void segment(const uint8_t* line, size_t n) {   // gfaline.rs:57-62; parse.rs:138-141
        Cursor c = body(line, n);
        const uint64_t name = parse_num<uint64_t>(c);
        parse_byte(c, '\t');
        Cursor seq = parse_field(c);
        const uint32_t id = flat.add_seg(name, seq.p, seq.n, c.p, c.n);
        seg_ids.insert(name, id);
    }
    void link(const uint8_t* line, size_t n) {   // gfaline.rs:65-85; parse.rs:143-147
        Cursor c = body(line, n);
        const uint64_t from_seg = parse_num<uint64_t>(c);
        parse_byte(c, '\t');
        const bool from_fwd = parse_orient(c);
        parse_byte(c, '\t');
        const uint64_t to_seg = parse_num<uint64_t>(c);
        parse_byte(c, '\t');
        const bool to_fwd = parse_orient(c);
        parse_byte(c, '\t');
        std::vector<AlignOp> overlap = parse_align(c);
        if (!c.empty()) throw Error("expected end of line");
        const Handle from = Handle::make(seg_ids.get(from_seg), from_fwd);
        const Handle to = Handle::make(seg_ids.get(to_seg), to_fwd);
        flat.add_link(from, to, overlap);
    }
    // A P line parsed but not yet added to the store: the step list tokenisation (the bulk
    // of a pangenome GFA) only reads the finished segment-name map, so many lines can be
    // tokenised concurrently and then added in file order.
    struct ParsedPath {
        Cursor name{nullptr, 0};
        std::vector<Handle> steps;
        std::vector<std::vector<AlignOp>> overlaps;
        std::string error;                 // non-empty: what the reference would have failed with
    };
    ParsedPath parse_path(const uint8_t* line, size_t n) const {   // gfaline.rs:88-100; parse.rs:149-156
        ParsedPath out;
        try {
            Cursor c = body(line, n);
            out.name = parse_field(c);
            Cursor steps = parse_field(c);
            out.overlaps = parse_maybe_overlap_list(c);
            if (!c.empty()) throw Error("expected end of line");
            out.steps.reserve(steps.n / 3 + 1);
            const size_t used = parse_steps(steps.p, steps.n, [&](uint64_t seg, bool fwd) {
                out.steps.push_back(Handle::make(seg_ids.get(seg), fwd));
            });
            if (used != steps.n) throw Error("malformed step list");   // parse.rs:155 assert
        } catch (const std::exception& e) {
            out.error = e.what();
        }
        return out;
    }
    void add_parsed_path(const ParsedPath& pp) {                   // parse.rs:151-158
        if (!pp.error.empty()) throw Error(pp.error);
        const Span span = HeapGFAStore::add_slice(flat.steps, pp.steps.data(), pp.steps.size());
        flat.add_path(pp.name.p, pp.name.n, span, pp.overlaps);
    }
    void path(const uint8_t* line, size_t n) { add_parsed_path(parse_path(line, n)); }
    void other(const uint8_t* line, size_t n) {   // gfaline.rs:37-49 for H / S / anything else
        if (n < 2 || line[1] != '\t') throw Error("expected marker and tab");
        switch (line[0]) {
            case 'H': flat.record_line(kLineHeader); header(line, n); break;
            case 'S': flat.record_line(kLineSegment); segment(line, n); break;
            default: throw Error("unhandled line kind");
        }
    }
};

}  // namespace

static bool gpu_paths(Builder& b, const uint8_t* buf, size_t len,
                      const std::vector<std::pair<const uint8_t*, size_t>>& deferred);
static bool gpu_paths_hook(Builder& b, const uint8_t* buf, size_t len,
                           const std::vector<std::pair<const uint8_t*, size_t>>& deferred) {
    // links must still be added in file order relative to the paths, and a link that fails to
    // parse must fail the same way on both routes: only take the GPU route when it succeeds whole
    return gpu_paths(b, buf, len, deferred);
}

HeapGFAStore Parser::parse_mem(const uint8_t* buf, size_t len) {
    Builder b;
    std::vector<std::pair<const uint8_t*, size_t>> deferred;
    size_t pos = 0;
    while (pos < len) {   // memfile.rs:50-61
        const void* nl = std::memchr(buf + pos, '\n', len - pos);
        if (!nl) break;   // final line without a newline is dropped
        const uint8_t* line = buf + pos;
        const size_t n = (size_t)((const uint8_t*)nl - line);
        pos += n + 1;
        if (n == 0) throw Error("empty line");   // reference: line[0] index panic
        if (line[0] == 'P' || line[0] == 'L') {   // parse.rs:83-91
            b.flat.record_line(line[0] == 'P' ? kLinePath : kLineLink);
            deferred.emplace_back(line, n);
            continue;
        }
        b.other(line, n);
    }
    if (gpu_paths_hook(b, buf, len, deferred)) return std::move(b.flat);
    // parse.rs:110-123: links and paths are added in file order.  The step lists are
    // tokenised ahead of that by a few worker threads when there is enough text to pay for them.
    size_t path_bytes = 0, n_path_lines = 0;
    for (auto& d : deferred)
        if (d.first[0] == 'P') { path_bytes += d.second; ++n_path_lines; }
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (!(hw > 1 && path_bytes > (4u << 20))) {
        for (auto& d : deferred) {
            if (d.first[0] == 'L') b.link(d.first, d.second);
            else b.path(d.first, d.second);
        }
        return std::move(b.flat);
    }
    // Threaded route.  Every P line is split into its fields up front (cheap, serial); the step
    // lists are cut into pieces of about kPieceBytes at commas that follow an orientation sign --
    // where StepsParser is back in its initial state -- so that one very long path spreads over
    // the workers as well as many short ones do.  A piece that is not the last of its field must
    // end cleanly (no early stop): wherever the sequential parser would have stopped, `rest` is
    // non-empty and parse.rs:155 fails the same way.
    constexpr size_t kPieceBytes = 1u << 20;
    struct Line {
        Cursor name{nullptr, 0};
        std::vector<std::vector<AlignOp>> overlaps;
        std::string error;
        size_t first_piece = 0, n_pieces = 0;
    };
    struct Piece {
        const uint8_t* p;
        size_t n;
        bool last;
        std::vector<Handle> steps;
        std::string error;
    };
    std::vector<Line> lines;
    std::vector<Piece> pieces;
    for (auto& d : deferred) {
        if (d.first[0] != 'P') continue;
        Line ln;
        try {                                                   // gfaline.rs:88-100, as in parse_path
            Cursor c = Builder::body(d.first, d.second);
            ln.name = parse_field(c);
            Cursor steps = parse_field(c);
            ln.overlaps = parse_maybe_overlap_list(c);
            if (!c.empty()) throw Error("expected end of line");
            ln.first_piece = pieces.size();
            size_t a = 0;
            while (true) {
                size_t cut = steps.n;                           // default: the rest of the field
                if (steps.n - a > kPieceBytes + (kPieceBytes >> 2)) {
                    for (size_t q = a + kPieceBytes; q + 1 < steps.n; ++q)
                        if (steps.p[q] == ',' && (steps.p[q - 1] == '+' || steps.p[q - 1] == '-')) { cut = q; break; }
                }
                const bool last = cut == steps.n;
                pieces.push_back(Piece{steps.p + a, cut - a, last, {}, {}});
                if (last) break;
                a = cut + 1;                                    // the comma itself is consumed here
            }
            ln.n_pieces = pieces.size() - ln.first_piece;
        } catch (const std::exception& e) {
            ln.error = e.what();
        }
        lines.push_back(std::move(ln));
    }
    {
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t k; (k = next.fetch_add(1)) < pieces.size();) {
                Piece& pc = pieces[k];
                try {
                    pc.steps.reserve(pc.n / 3 + 1);
                    bool clean = false;
                    const size_t used = parse_steps(pc.p, pc.n, [&](uint64_t seg, bool fwd) {
                        pc.steps.push_back(Handle::make(b.seg_ids.get(seg), fwd));
                    }, &clean);
                    if (pc.last ? used != pc.n : !clean) throw Error("malformed step list");   // parse.rs:155 assert
                } catch (const std::exception& e) {
                    pc.error = e.what();
                }
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < std::min<size_t>(hw, pieces.size()); ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    }
    // parse.rs:110-123: links and paths are added in file order; the first failure in that order wins.
    const size_t base = b.flat.steps.size();
    size_t running = base, k = 0;
    std::vector<size_t> piece_off(pieces.size(), 0);
    for (auto& d : deferred) {
        if (d.first[0] == 'L') { b.link(d.first, d.second); continue; }
        const Line& ln = lines[k++];
        if (!ln.error.empty()) throw Error(ln.error);
        const size_t start = running;
        for (size_t i = ln.first_piece; i < ln.first_piece + ln.n_pieces; ++i) {
            if (!pieces[i].error.empty()) throw Error(pieces[i].error);
            piece_off[i] = running;
            running += pieces[i].steps.size();
        }
        b.flat.add_path(ln.name.p, ln.name.n, Span{HeapGFAStore::id(start), HeapGFAStore::id(running)}, ln.overlaps);
    }
    b.flat.steps.resize(running);
    {
        std::atomic<size_t> next{0};
        auto copy = [&]() {
            for (size_t i; (i = next.fetch_add(1)) < pieces.size();)
                if (!pieces[i].steps.empty())
                    std::memcpy(b.flat.steps.data() + piece_off[i], pieces[i].steps.data(), pieces[i].steps.size() * sizeof(Handle));
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < std::min<size_t>(hw, pieces.size()); ++t) pool.emplace_back(copy);
        copy();
        for (auto& t : pool) t.join();
    }
    return std::move(b.flat);
}

// GPU route for the step lists: tokenise every P line's steps field with
// fgfa_tokenizer_* and add the paths with the spans it reports.  Returns false -- leaving the
// store untouched -- when there is no device, the text is small, or the input is outside the
// tokenizer's strict grammar; the caller then takes the host route above, which reproduces the
// reference's quirks and error messages.
static bool gpu_paths(Builder& b, const uint8_t* buf, size_t len,
                      const std::vector<std::pair<const uint8_t*, size_t>>& deferred) {
    const char* env = std::getenv("FGFA_GPU_PARSE");
    if (env && env[0] == '0') return false;
    size_t path_bytes = 0;
    for (auto& d : deferred)
        if (d.first[0] == 'P') path_bytes += d.second;
    // a process that has no CUDA context yet pays ~1 s to create one: below ~1 GiB of step-list
    // text the threaded host route (~0.8 GB/s on 16 cores) wins for a one-shot command
    const size_t threshold = (env && env[0] == '1') ? 0 : ((size_t)1 << 30);
    if (path_bytes <= threshold || fgfa_device_count() <= 0) return false;
    struct Fields { Cursor name; Cursor rest; };
    std::vector<Fields> lines;
    std::vector<uint64_t> off, flen;
    for (auto& d : deferred) {
        if (d.first[0] != 'P') continue;
        if (d.second < 2 || d.first[1] != '\t') return false;
        Cursor c{d.first + 2, d.second - 2};
        Fields f;
        f.name = parse_field(c);
        Cursor steps = parse_field(c);
        f.rest = c;
        off.push_back((uint64_t)(steps.p - buf));
        flen.push_back(steps.n);
        lines.push_back(f);
    }
    fgfa_tokenizer_t* tok = nullptr;
    if (fgfa_tokenizer_create(&tok, buf, len, off.data(), flen.data(), (uint32_t)off.size()) != FGFA_OK) return false;
    struct Guard { fgfa_tokenizer_t* t; ~Guard() { fgfa_tokenizer_destroy(t); } } guard{tok};
    std::vector<uint32_t> ss(off.size()), se(off.size());
    uint64_t total = 0;
    if (fgfa_tokenizer_spans(tok, ss.data(), se.data(), &total) != FGFA_OK) return false;
    const size_t prev = b.flat.steps.size();
    if (prev + total > 0xFFFFFFFFull) return false;
    std::vector<uint64_t> names;
    std::vector<uint32_t> ids;
    for (auto& kv : b.seg_ids.others()) { names.push_back(kv.first); ids.push_back(kv.second); }
    std::vector<Handle> steps(total);
    if (fgfa_tokenizer_parse(tok, b.seg_ids.sequential_max(), names.data(), ids.data(), (uint32_t)names.size(),
                             reinterpret_cast<uint32_t*>(steps.data())) != FGFA_OK)
        return false;
    // overlaps are parsed before anything is committed so that a malformed line leaves the store clean
    std::vector<std::vector<std::vector<AlignOp>>> overlaps(lines.size());
    try {
        for (size_t i = 0; i < lines.size(); ++i) {
            Cursor c = lines[i].rest;
            overlaps[i] = parse_maybe_overlap_list(c);
            if (!c.empty()) return false;
        }
    } catch (const std::exception&) {
        return false;
    }
    b.flat.steps.insert(b.flat.steps.end(), steps.begin(), steps.end());
    size_t k = 0;
    for (auto& d : deferred) {   // parse.rs:110-123: file order
        if (d.first[0] == 'L') { b.link(d.first, d.second); continue; }
        const Span span{HeapGFAStore::id(prev + ss[k]), HeapGFAStore::id(prev + se[k])};
        b.flat.add_path(lines[k].name.p, lines[k].name.n, span, overlaps[k]);
        ++k;
    }
    return true;
}

file::Toc estimate_toc(const uint8_t* buf, size_t len) {
    size_t segs = 0, links = 0, paths = 0, header_bytes = 0, seg_bytes = 0, path_bytes = 0;
    size_t pos = 0;
    while (pos < len) {
        const uint8_t marker = buf[pos];
        const void* nl = std::memchr(buf + pos, '\n', len - pos);
        const size_t rest = len - pos;
        const size_t next = nl ? (size_t)((const uint8_t*)nl - (buf + pos)) : rest + 1;   // parse.rs:187
        switch (marker) {
            case 'H': header_bytes += next; break;
            case 'S': ++segs; seg_bytes += next; break;
            case 'L': ++links; break;
            case 'P': ++paths; path_bytes += next; break;
            default: throw Error("unknown line type");
        }
        if (next >= rest) break;
        pos += next + 1;
    }
    return file::Toc::estimate(segs, links, paths, header_bytes, seg_bytes, path_bytes);
}

HeapGFAStore Parser::parse_stream(FILE* in) {
    Builder b;
    std::vector<std::string> links, paths;
    std::string line;
    auto handle = [&](const std::string& ln) {
        if (ln.empty()) throw Error("empty line");
        const uint8_t* p = reinterpret_cast<const uint8_t*>(ln.data());
        if (ln[0] == 'P') {   // parse.rs:35-39: paths are kept as raw lines
            b.flat.record_line(kLinePath);
            paths.push_back(ln);
        } else if (ln[0] == 'L') {   // parse.rs:42-43,52-54: parsed now in the reference, added later
            if (ln.size() < 2 || ln[1] != '\t') throw Error("expected marker and tab");
            b.flat.record_line(kLineLink);
            links.push_back(ln);
        } else {
            b.other(p, ln.size());
        }
    };
    char chunk[1 << 16];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof chunk, in)) > 0) {
        size_t start = 0;
        for (size_t i = 0; i < got; ++i) {
            if (chunk[i] == '\n') {
                line.append(chunk + start, i - start);
                handle(line);
                line.clear();
                start = i + 1;
            }
        }
        line.append(chunk + start, got - start);
    }
    if (!line.empty()) handle(line);   // BufRead::split yields an unterminated last line
    for (auto& l : links) b.link(reinterpret_cast<const uint8_t*>(l.data()), l.size());   // parse.rs:61-63
    for (auto& p : paths) b.path(reinterpret_cast<const uint8_t*>(p.data()), p.size());   // parse.rs:64-70
    return std::move(b.flat);
}

}  // namespace flatgfa
