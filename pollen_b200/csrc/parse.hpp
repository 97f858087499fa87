// GFA text -> FlatGFA pools.  C++ restatement of the reference's parser
// (flatgfa/src/parse.rs:8-159, gfaline.rs:1-263, namemap.rs:8-43, memfile.rs:33-62),
// kept to what the node-depth path and the C ABI need: H, S, L and P lines.
// Where the reference panics or returns Err on malformed input, this throws
// flatgfa::Error with the reference's message.
#pragma once
#include <cstdint>
#include <cstdio>
#include <unordered_map>

#include "file.hpp"
#include "flatgfa.hpp"

namespace flatgfa {

// namemap.rs:8-43: segment name -> pool index; names 1..k seen in order need no table.
class NameMap {
public:
    void insert(uint64_t name, uint32_t id);
    uint32_t get(uint64_t name) const;   // throws if unknown (reference: HashMap index panic)
    uint64_t sequential_max() const { return sequential_max_; }
    const std::unordered_map<uint64_t, uint32_t>& others() const { return others_; }
private:
    uint64_t sequential_max_ = 0;
    std::unordered_map<uint64_t, uint32_t> others_;
};

class Parser {
public:
    // parse.rs:77-126: whole buffer in memory; L and P lines are deferred, in file order.
    // Mirrors MemchrSplit (memfile.rs:50-61): a final line without '\n' is dropped.
    static HeapGFAStore parse_mem(const uint8_t* buf, size_t len);
    // parse.rs:24-74: streaming; links are unwound before paths; the last line is kept
    // even without a trailing newline (BufRead::split).
    static HeapGFAStore parse_stream(FILE* in);
};

// parse.rs:176-216: one scan over GFA text counting the lines of each kind and the bytes of the H, S and
// P lines, turned into the capacities of a preallocated file by Toc::estimate.  Throws on a line that
// does not start with H, S, L or P (the reference panics with "unknown line type").
file::Toc estimate_toc(const uint8_t* buf, size_t len);

// gfaline.rs:201-263: the `1+,23-,4+` step-list state machine.  Returns the number of
// bytes consumed; emits (name, forward) pairs through `emit`.
template <typename F>
size_t parse_steps(const uint8_t* s, size_t n, F&& emit, bool* clean_end = nullptr) {
    size_t index = 0;
    bool want_seg = true;
    uint64_t seg = 0;
    if (clean_end) *clean_end = false;
    while (index < n) {
        const uint8_t byte = s[index++];
        if (want_seg) {
            if (byte == '+' || byte == '-') {
                want_seg = false;
                emit(seg, byte == '+');
            } else if (byte >= '0' && byte <= '9') {
                seg = seg * 10 + (uint64_t)(byte - '0');
            } else {
                return index;   // gfaline.rs:243-245: stop, the bad byte is consumed
            }
        } else {
            if (byte == ',') {
                want_seg = true;
                seg = 0;
            } else {
                return index;   // gfaline.rs:251-253
            }
        }
    }
    // clean_end: every byte was consumed without an early stop and the last one completed a step,
    // i.e. the state is "expecting a comma" -- what a piece cut just before a comma must end in.
    if (clean_end) *clean_end = !want_seg;
    return index;
}

}  // namespace flatgfa
