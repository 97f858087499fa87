// Interval ("window") depth along one path, host side.  Same names and semantics as the
// reference's flatgfa/src/ops/window_depth.rs (`Windows`, `window_depth`, `bed_depth`,
// `IntervalDepth` + `Emit`), with `weighted_depths` / `assign_depths` executed by the sm_100a
// kernels behind fgfa_depth.h (interval_kernels.cuh).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "flatbed.hpp"
#include "flatgfa.hpp"

namespace flatgfa {
namespace ops {
namespace window_depth {

// window_depth.rs:20-66: equally sized windows from `start` to `end` in steps of `size`.
struct Windows {
    Pool<uint8_t> name;
    uint64_t start, end, size;
    void emit(std::string& out) const;            // :27-38  "name\tpos\tend\n"
    void emit_bed(HeapBEDStore& store) const;     // :41-52
    HeapBEDStore as_bed() const;                  // :54-58
    size_t len() const;                           // :61-63  ceil((end - start) / size)
    bool is_empty() const { return len() == 0; }
};

// window_depth.rs:156-174
struct IntervalDepth {
    FlatBED intervals;
    std::vector<double> depths;
    void emit(std::string& out) const;            // "name\tstart\tend\t{format_float(depth, 4)}\n"
    void emit(FILE* f) const;
    void print() const { emit(stdout); }
};

// window_depth.rs:183-197.  `path` is a path pool index.  Throws flatgfa::Error for
// window_size == 0 (the reference never terminates) and on device failure.
std::pair<HeapBEDStore, std::vector<double>> window_depth(const FlatGFA& gfa, uint32_t path, uint64_t window_size);

// window_depth.rs:203-211: all intervals are taken to lie on the path named by the FIRST
// entry, sorted along it.  Throws where the reference panics (no entries; unknown path).
std::vector<double> bed_depth(const FlatGFA& gfa, const FlatBED& intervals);

}  // namespace window_depth
}  // namespace ops
}  // namespace flatgfa
