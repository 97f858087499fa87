// The `.flatgfa` binary format: a table of contents followed by the eleven pools.
// C++ restatement of the read side (`view`) and the compact write side (`dump`) of the
// reference's flatgfa/src/file.rs, plus the read-only mmap of flatgfa/src/memfile.rs:7-10.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "flatgfa.hpp"

namespace flatgfa {
namespace file {

constexpr uint64_t MAGIC_NUMBER = 0xB1011054ull;  // file.rs:9

#pragma pack(push, 1)
struct Size {            // file.rs:29-38
    uint64_t len;        // valid elements
    uint64_t capacity;   // allocated elements; capacity - len slots are empty
};
struct Toc {             // file.rs:11-27
    uint64_t magic;
    Size header, segs, paths, links, steps, seq_data, overlaps, alignment, name_data,
        optional_data, line_order;
    size_t size() const;                 // file.rs:64-79: total file size in bytes
    static Toc full(const FlatGFA& g);   // file.rs:82-97: capacity == len everywhere
    static Toc guess(size_t factor);     // file.rs:117-132: capacities of a fresh file, all pools empty
    // file.rs:136-158: capacities estimated from measurements of the GFA text
    static Toc estimate(size_t segs, size_t links, size_t paths, size_t header_bytes, size_t seg_bytes,
                        size_t path_bytes);
};
#pragma pack(pop)
static_assert(sizeof(Toc) == 184, "Toc is 8 + 11*16 bytes");

enum ViewError { kViewOk = 0, kViewTooShort = 1, kViewBadMagic = 2, kViewTruncated = 3 };

// file.rs:185-213.  Where the reference would panic (short buffer, bad magic, pool
// running past the end) this returns an error code and leaves `out` untouched.
ViewError view(const uint8_t* data, size_t len, FlatGFA* out);
// Throwing form for the C++ API / CLI.
FlatGFA view_or_throw(const uint8_t* data, size_t len);

size_t size(const FlatGFA& g);                 // file.rs:311-313
void dump(const FlatGFA& g, uint8_t* buf);     // file.rs:290-307; buf must hold size(g) bytes
// A compact image with `capacity` slack on chosen pools, as the reference's
// preallocated in-place files have (file.rs:117-158): used to test capacity > len.
std::vector<uint8_t> dump_with_slack(const FlatGFA& g, size_t extra_per_pool);
// The image a graph has after being parsed INTO a preallocated file (file.rs:261-272 `init` +
// `Parser::for_slice`, cli/main.rs:216-248 `prealloc_translate`): pool i occupies capacities.<i>.capacity
// slots, of which the first len hold the graph and the rest are zero.  Throws "capacity overflow" where
// the reference's fixed-size store would (a pool larger than its capacity).
std::vector<uint8_t> dump_preallocated(const FlatGFA& g, const Toc& capacities);

}  // namespace file

// memfile.rs:7-10: read-only mapping of a whole file.
class MappedFile {
public:
    explicit MappedFile(const std::string& path);   // throws Error
    ~MappedFile();
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    const uint8_t* data() const { return data_; }
    size_t size() const { return size_; }
private:
    const uint8_t* data_ = nullptr;
    size_t size_ = 0;
};

// memfile.rs:12-23 map_new_file: create/truncate and write a buffer.
void write_file(const std::string& path, const uint8_t* data, size_t len);

}  // namespace flatgfa
