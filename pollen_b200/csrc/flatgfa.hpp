// Host-side data model of a FlatGFA graph: eleven flat, pointer-free pools.
//
// C++ restatement of the interface of the reference's `flatgfa` crate for the types
// the node-depth path touches (reference: flatgfa/src/flatgfa.rs:19-67 FlatGFA,
// :71-82 Segment, :99-118 Path, :121-133 Link, :186-209 Handle, :225-251 AlignOp;
// flatgfa/src/pool.rs:9-11 Id, :80-124 Span, :279-347 Pool).  All records are
// byte-packed (Rust `repr(packed)`), little-endian, and may sit at any byte offset
// of an mmapped file, so every struct here has alignment 1.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace flatgfa {

// The reference panics on malformed input; this library throws instead and the
// C ABI / CLI turn the exception into an error code / message.
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
    // FGFA_ERR_* of a failed device call (0 = a host-side error without a code); the C entry points
    // hand it back unchanged so that all of them report the same code for the same failure
    Error(const std::string& what, int code_) : std::runtime_error(what), code(code_) {}
    int code = 0;
};

#pragma pack(push, 1)

// pool.rs:80-86: half-open [start, end) range of u32 pool indices.
struct Span {
    uint32_t start;
    uint32_t end;
    bool is_empty() const { return start == end; }
    size_t len() const { return (size_t)(end - start); }
};

// flatgfa.rs:186-209: segment index << 1 | orientation bit (0 = forward, 1 = backward).
struct Handle {
    uint32_t bits;
    static Handle make(uint32_t segment, bool forward) {
        if (segment & 0x80000000u) throw Error("index too large");  // flatgfa.rs:194
        return Handle{(segment << 1) | (forward ? 0u : 1u)};
    }
    uint32_t segment() const { return bits >> 1; }
    bool is_forward() const { return (bits & 1u) == 0; }
};

// flatgfa.rs:71-82
struct Segment {
    uint64_t name;   // `usize` on the reference's 64-bit targets
    Span seq;        // range in seq_data
    Span optional;   // range in optional_data
    size_t len() const { return seq.len(); }
};

// flatgfa.rs:99-118
struct Path {
    Span name;       // range in name_data
    Span steps;      // range in steps
    Span overlaps;   // range in overlaps
    size_t step_count() const { return (size_t)steps.end - (size_t)steps.start; }
};

// flatgfa.rs:121-133
struct Link {
    Handle from;
    Handle to;
    Span overlap;    // range in alignment
};

// flatgfa.rs:211-251: (len << 8) | opcode; opcodes M=0, N=1, D(Insertion enum slot)=2, I=3
// follow the enum order Match, Gap, Insertion, Deletion (flatgfa.rs:213-218).
struct AlignOp {
    uint32_t bits;
};

#pragma pack(pop)

static_assert(sizeof(Span) == 8 && sizeof(Handle) == 4 && sizeof(Segment) == 24 &&
                  sizeof(Path) == 24 && sizeof(Link) == 16 && sizeof(AlignOp) == 4,
              "record sizes must match the .flatgfa format");

// flatgfa.rs:255-262 LineKind
enum LineKind : uint8_t { kLineHeader = 0, kLineSegment = 1, kLinePath = 2, kLineLink = 3 };

// pool.rs:279-347: a borrowed, fixed-size view of one pool.
template <typename T>
struct Pool {
    const T* data = nullptr;
    size_t count = 0;
    size_t len() const { return count; }
    bool is_empty() const { return count == 0; }
    const T& operator[](size_t i) const {
        if (i >= count) throw Error("pool index out of bounds");
        return data[i];
    }
    // pool.rs:341-347: Index<Span> -> sub-slice, bounds-checked like a Rust slice.
    Pool<T> slice(Span s) const {
        if (s.start > s.end || (size_t)s.end > count) throw Error("span out of bounds");
        return Pool<T>{data + s.start, (size_t)(s.end - s.start)};
    }
    const T* begin() const { return data; }
    const T* end() const { return data + count; }
};

// flatgfa.rs:19-67
struct FlatGFA {
    Pool<uint8_t> header;
    Pool<Segment> segs;
    Pool<Path> paths;
    Pool<Link> links;
    Pool<Handle> steps;
    Pool<uint8_t> seq_data;
    Pool<Span> overlaps;
    Pool<AlignOp> alignment;
    Pool<uint8_t> name_data;
    Pool<uint8_t> optional_data;
    Pool<uint8_t> line_order;

    Pool<uint8_t> get_seq(const Segment& seg) const { return seq_data.slice(seg.seq); }       // flatgfa.rs:354-356
    Pool<uint8_t> get_path_name(const Path& p) const { return name_data.slice(p.name); }       // flatgfa.rs:381-383
    Pool<Handle> get_path_steps(const Path& p) const { return steps.slice(p.steps); }          // flatgfa.rs:385-387
    // flatgfa.rs:387-389 + pool.rs:305-307: index of the first path with this name, or -1.
    int64_t find_path(const uint8_t* name, size_t n) const {
        for (size_t p = 0; p < paths.len(); ++p) {
            const Pool<uint8_t> nm = get_path_name(paths.data[p]);
            if (nm.len() == n && (n == 0 || std::memcmp(nm.data, name, n) == 0)) return (int64_t)p;
        }
        return -1;
    }
};

// flatgfa.rs:426-552: the growable in-memory store (HeapGFAStore) the parser fills.
struct HeapGFAStore {
    std::vector<uint8_t> header;
    std::vector<Segment> segs;
    std::vector<Path> paths;
    std::vector<Link> links;
    std::vector<Handle> steps;
    std::vector<uint8_t> seq_data;
    std::vector<Span> overlaps;
    std::vector<AlignOp> alignment;
    std::vector<uint8_t> name_data;
    std::vector<uint8_t> optional_data;
    std::vector<uint8_t> line_order;

    static uint32_t id(size_t index) {  // pool.rs:51-53 Id::new
        if (index > 0xFFFFFFFFull) throw Error("id too large");
        return (uint32_t)index;
    }
    template <typename T>
    static Span add_slice(std::vector<T>& pool, const T* items, size_t n) {  // pool.rs:205-210
        uint32_t start = id(pool.size());
        pool.insert(pool.end(), items, items + n);
        return Span{start, id(pool.size())};
    }

    void add_header(const uint8_t* v, size_t n) {  // flatgfa.rs:441-444
        if (!header.empty()) throw Error("duplicate header line");
        header.insert(header.end(), v, v + n);
    }
    uint32_t add_seg(uint64_t name, const uint8_t* seq, size_t seq_len, const uint8_t* opt,
                     size_t opt_len) {  // flatgfa.rs:447-453
        Segment s;
        s.name = name;
        s.seq = add_slice(seq_data, seq, seq_len);
        s.optional = add_slice(optional_data, opt, opt_len);
        uint32_t i = id(segs.size());
        segs.push_back(s);
        return i;
    }
    uint32_t add_link(Handle from, Handle to, const std::vector<AlignOp>& overlap) {  // flatgfa.rs:498-504
        Link l{from, to, add_slice(alignment, overlap.data(), overlap.size())};
        uint32_t i = id(links.size());
        links.push_back(l);
        return i;
    }
    uint32_t add_path(const uint8_t* name, size_t name_len, Span steps_span,
                      const std::vector<std::vector<AlignOp>>& ovl) {  // flatgfa.rs:456-476
        uint32_t ostart = id(overlaps.size());
        for (const auto& a : ovl) overlaps.push_back(add_slice(alignment, a.data(), a.size()));
        Path p;
        p.overlaps = Span{ostart, id(overlaps.size())};
        p.name = add_slice(name_data, name, name_len);
        p.steps = steps_span;
        uint32_t i = id(paths.size());
        paths.push_back(p);
        return i;
    }
    void record_line(LineKind k) { line_order.push_back((uint8_t)k); }  // flatgfa.rs:507-509

    FlatGFA view() const {  // flatgfa.rs:512-526 as_ref
        FlatGFA g;
        g.header = {header.data(), header.size()};
        g.segs = {segs.data(), segs.size()};
        g.paths = {paths.data(), paths.size()};
        g.links = {links.data(), links.size()};
        g.steps = {steps.data(), steps.size()};
        g.seq_data = {seq_data.data(), seq_data.size()};
        g.overlaps = {overlaps.data(), overlaps.size()};
        g.alignment = {alignment.data(), alignment.size()};
        g.name_data = {name_data.data(), name_data.size()};
        g.optional_data = {optional_data.data(), optional_data.size()};
        g.line_order = {line_order.data(), line_order.size()};
        return g;
    }
};

}  // namespace flatgfa
