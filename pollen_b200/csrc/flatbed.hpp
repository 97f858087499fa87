// FlatBED: a flat list of named intervals, and its text parser.
//
// C++ restatement of the reference's flatgfa/src/flatbed.rs for what interval depth needs:
// :10-16 BEDEntry, :19-33 FlatBED, :62-81 BEDStore (heap family only), :118-158 BEDParser.
#pragma once
#include <cstdint>
#include <vector>

#include "flatgfa.hpp"

namespace flatgfa {

#pragma pack(push, 1)
// flatbed.rs:10-16 (repr(C, packed)): 24 bytes.
struct BEDEntry {
    Span name;        // range in name_data
    uint64_t start;
    uint64_t end;
};
#pragma pack(pop)
static_assert(sizeof(BEDEntry) == 24, "BEDEntry is 24 packed bytes");

// flatbed.rs:19-33
struct FlatBED {
    Pool<uint8_t> name_data;
    Pool<BEDEntry> entries;
    size_t get_num_entries() const { return entries.len(); }
    Pool<uint8_t> get_name_of_entry(const BEDEntry& e) const { return name_data.slice(e.name); }
};

// flatbed.rs:62-81, heap family (`HeapBEDStore`, :108).
struct HeapBEDStore {
    std::vector<uint8_t> name_data;
    std::vector<BEDEntry> entries;
    uint32_t add_entry(const uint8_t* name, size_t n, uint64_t start, uint64_t end) {   // :69-72
        const Span nm = HeapGFAStore::add_slice(name_data, name, n);
        const uint32_t id = HeapGFAStore::id(entries.size());
        entries.push_back(BEDEntry{nm, start, end});
        return id;
    }
    FlatBED view() const {                                                              // :74-79 as_ref
        return FlatBED{{name_data.data(), name_data.size()}, {entries.data(), entries.size()}};
    }
};

// flatbed.rs:118-158.  Lines are split like `MemchrSplit` (memfile.rs:51-63): a last line
// without '\n' is dropped.  `#` lines are skipped; a line is `name \t start SEP end ...`
// where SEP is any single byte.  Where the reference panics (missing number, line ending
// right after `start`) this throws flatgfa::Error.
struct BEDParser {
    static HeapBEDStore parse_mem(const uint8_t* buf, size_t len);
};

}  // namespace flatgfa
