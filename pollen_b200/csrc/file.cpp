#include "file.hpp"

#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace flatgfa {
namespace file {

namespace {
// Element sizes in Toc order (file.rs:64-79).
constexpr size_t kElem[11] = {1, sizeof(Segment), sizeof(Path), sizeof(Link), sizeof(Handle), 1,
                              sizeof(Span), sizeof(AlignOp), 1, 1, 1};
const Size* sizes_of(const Toc& t) { return &t.header; }
}  // namespace

size_t Toc::size() const {
    size_t n = sizeof(Toc);
    const Size* s = sizes_of(*this);
    for (int i = 0; i < 11; ++i) n += (size_t)s[i].capacity * kElem[i];
    return n;
}

Toc Toc::full(const FlatGFA& g) {
    Toc t;
    t.magic = MAGIC_NUMBER;
    auto of = [](size_t n) { return Size{(uint64_t)n, (uint64_t)n}; };
    t.header = of(g.header.len());
    t.segs = of(g.segs.len());
    t.paths = of(g.paths.len());
    t.links = of(g.links.len());
    t.steps = of(g.steps.len());
    t.seq_data = of(g.seq_data.len());
    t.overlaps = of(g.overlaps.len());
    t.alignment = of(g.alignment.len());
    t.name_data = of(g.name_data.len());
    t.optional_data = of(g.optional_data.len());
    t.line_order = of(g.line_order.len());
    return t;
}

ViewError view(const uint8_t* data, size_t len, FlatGFA* out) {
    if (len < sizeof(Toc)) return kViewTooShort;           // file.rs:171 ref_from_prefix().unwrap()
    Toc toc;
    std::memcpy(&toc, data, sizeof(Toc));
    if (toc.magic != MAGIC_NUMBER) return kViewBadMagic;   // file.rs:172-173
    const Size* s = sizes_of(toc);
    size_t off = sizeof(Toc);
    const uint8_t* base[11];
    for (int i = 0; i < 11; ++i) {                         // file.rs:163-167 slice_prefix
        if (s[i].capacity < s[i].len) return kViewTruncated;
        const unsigned __int128 bytes = (unsigned __int128)s[i].capacity * kElem[i];
        if (bytes > (unsigned __int128)(len - off)) return kViewTruncated;
        base[i] = data + off;
        off += (size_t)bytes;
    }
    FlatGFA g;
    g.header = {base[0], (size_t)s[0].len};
    g.segs = {reinterpret_cast<const Segment*>(base[1]), (size_t)s[1].len};
    g.paths = {reinterpret_cast<const Path*>(base[2]), (size_t)s[2].len};
    g.links = {reinterpret_cast<const Link*>(base[3]), (size_t)s[3].len};
    g.steps = {reinterpret_cast<const Handle*>(base[4]), (size_t)s[4].len};
    g.seq_data = {base[5], (size_t)s[5].len};
    g.overlaps = {reinterpret_cast<const Span*>(base[6]), (size_t)s[6].len};
    g.alignment = {reinterpret_cast<const AlignOp*>(base[7]), (size_t)s[7].len};
    g.name_data = {base[8], (size_t)s[8].len};
    g.optional_data = {base[9], (size_t)s[9].len};
    g.line_order = {base[10], (size_t)s[10].len};
    *out = g;
    return kViewOk;
}

FlatGFA view_or_throw(const uint8_t* data, size_t len) {
    FlatGFA g;
    switch (view(data, len, &g)) {
        case kViewOk: return g;
        case kViewTooShort: throw Error("file too short for a FlatGFA table of contents");
        case kViewBadMagic: throw Error("bad magic number: not a FlatGFA file");
        default: throw Error("FlatGFA file is truncated or its table of contents is inconsistent");
    }
}

size_t size(const FlatGFA& g) { return Toc::full(g).size(); }

namespace {
template <typename T>
uint8_t* put(uint8_t* p, const Pool<T>& pool, size_t slack) {
    const size_t n = pool.len() * sizeof(T);
    if (n) std::memcpy(p, pool.data, n);
    if (slack) std::memset(p + n, 0, slack * sizeof(T));
    return p + n + slack * sizeof(T);
}
void dump_impl(const FlatGFA& g, uint8_t* buf, size_t slack) {
    Toc toc = Toc::full(g);
    Size* s = const_cast<Size*>(sizes_of(toc));
    for (int i = 0; i < 11; ++i) s[i].capacity += slack;
    std::memcpy(buf, &toc, sizeof(Toc));
    uint8_t* p = buf + sizeof(Toc);
    p = put(p, g.header, slack);
    p = put(p, g.segs, slack);
    p = put(p, g.paths, slack);
    p = put(p, g.links, slack);
    p = put(p, g.steps, slack);
    p = put(p, g.seq_data, slack);
    p = put(p, g.overlaps, slack);
    p = put(p, g.alignment, slack);
    p = put(p, g.name_data, slack);
    p = put(p, g.optional_data, slack);
    put(p, g.line_order, slack);
}
}  // namespace

void dump(const FlatGFA& g, uint8_t* buf) { dump_impl(g, buf, 0); }

Toc Toc::guess(size_t f) {
    Toc t;
    t.magic = MAGIC_NUMBER;
    auto empty = [](size_t cap) { return Size{0, (uint64_t)cap}; };
    t.header = empty(128);
    t.segs = empty(32 * f * f);
    t.paths = empty(f);
    t.links = empty(32 * f * f);
    t.steps = empty(1024 * f * f);
    t.seq_data = empty(512 * f * f);
    t.overlaps = empty(256 * f);
    t.alignment = empty(64 * f * f);
    t.name_data = empty(64 * f);
    t.optional_data = empty(512 * f * f);
    t.line_order = empty(64 * f * f);
    return t;
}

Toc Toc::estimate(size_t segs, size_t links, size_t paths, size_t header_bytes, size_t seg_bytes, size_t path_bytes) {
    Toc t;
    t.magic = MAGIC_NUMBER;
    auto empty = [](size_t cap) { return Size{0, (uint64_t)cap}; };
    t.header = empty(header_bytes);
    t.segs = empty(segs);
    t.paths = empty(paths);
    t.links = empty(links);
    t.steps = empty(path_bytes / 3);
    t.seq_data = empty(seg_bytes);
    t.overlaps = empty((links + paths) * 2);
    t.alignment = empty(links * 2 + paths * 4);
    t.name_data = empty(paths * 512);
    t.optional_data = empty(links * 16);
    t.line_order = empty(segs + links + paths + 8);
    return t;
}

namespace {
template <typename T>
uint8_t* put_cap(uint8_t* p, const Pool<T>& pool, Size& sz) {
    if (pool.len() > sz.capacity) throw Error("capacity overflow");   // tinyvec SliceVec::push in the reference
    sz.len = pool.len();
    const size_t n = pool.len() * sizeof(T);
    if (n) std::memcpy(p, pool.data, n);
    return p + (size_t)sz.capacity * sizeof(T);                        // the rest of the slots stay zero
}
}  // namespace

std::vector<uint8_t> dump_preallocated(const FlatGFA& g, const Toc& capacities) {
    Toc toc = capacities;
    toc.magic = MAGIC_NUMBER;
    std::vector<uint8_t> out(toc.size(), 0);
    uint8_t* p = out.data() + sizeof(Toc);
    p = put_cap(p, g.header, toc.header);
    p = put_cap(p, g.segs, toc.segs);
    p = put_cap(p, g.paths, toc.paths);
    p = put_cap(p, g.links, toc.links);
    p = put_cap(p, g.steps, toc.steps);
    p = put_cap(p, g.seq_data, toc.seq_data);
    p = put_cap(p, g.overlaps, toc.overlaps);
    p = put_cap(p, g.alignment, toc.alignment);
    p = put_cap(p, g.name_data, toc.name_data);
    p = put_cap(p, g.optional_data, toc.optional_data);
    put_cap(p, g.line_order, toc.line_order);
    std::memcpy(out.data(), &toc, sizeof(Toc));                        // file.rs: Toc::for_fixed_store
    return out;
}

std::vector<uint8_t> dump_with_slack(const FlatGFA& g, size_t extra) {
    Toc toc = Toc::full(g);
    Size* s = const_cast<Size*>(sizes_of(toc));
    for (int i = 0; i < 11; ++i) s[i].capacity += extra;
    std::vector<uint8_t> out(toc.size());
    dump_impl(g, out.data(), extra);
    return out;
}

}  // namespace file

MappedFile::MappedFile(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw Error("cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); throw Error("cannot stat " + path); }
    size_ = (size_t)st.st_size;
    if (size_ == 0) { ::close(fd); data_ = nullptr; return; }
    void* p = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) throw Error("cannot mmap " + path);
    data_ = static_cast<const uint8_t*>(p);
}

MappedFile::~MappedFile() {
    if (data_) ::munmap(const_cast<uint8_t*>(data_), size_);
}

void write_file(const std::string& path, const uint8_t* data, size_t len) {
    int fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) throw Error("cannot create " + path);
    size_t off = 0;
    while (off < len) {
        ssize_t w = ::write(fd, data + off, len - off);
        if (w <= 0) { ::close(fd); throw Error("write failed: " + path); }
        off += (size_t)w;
    }
    ::close(fd);
}

}  // namespace flatgfa
