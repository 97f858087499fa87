// The node-depth op, host side.  Same names and semantics as the reference's
// flatgfa/src/ops/depth.rs:15-82 (`seg_depth_with_uniq`, `seg_depth`, `SegDepth` +
// `Emit`), with the loop nest executed by the sm_100a kernels behind fgfa_depth.h.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "flatgfa.hpp"

namespace flatgfa {
namespace ops {
namespace depth {

// depth.rs:15-39.  Both vectors have gfa.segs.len() entries, indexed by segment pool
// index; values are `usize` (u64).  Throws flatgfa::Error on device failure or where the
// reference would panic (span / segment index out of range).
std::pair<std::vector<uint64_t>, std::vector<uint64_t>> seg_depth_with_uniq(const FlatGFA& gfa);

// depth.rs:45-56.
std::vector<uint64_t> seg_depth(const FlatGFA& gfa);

// The same op on `n_gpus` devices of this box (whole paths partitioned by step count, partial
// arrays combined with one NCCL all-reduce; fgfa_depth_multi_* in fgfa_depth.h).  n_gpus <= 1 is
// seg_depth_with_uniq(gfa); more than the box has is clamped.
std::pair<std::vector<uint64_t>, std::vector<uint64_t>> seg_depth_with_uniq(const FlatGFA& gfa, int n_gpus);

// What every host entry point hands to the device ABI: the per-path `steps` spans
// (flatgfa.rs:99-112), a 4-byte-aligned view of the steps pool (the pool sits at an arbitrary byte
// offset of a .flatgfa image; it is copied only when misaligned) and the counts.
struct PoolArrays {
    std::vector<uint32_t> start, end;
    const uint32_t* steps = nullptr;
    uint32_t n_paths = 0, n_segs = 0;
    uint64_t n_steps = 0;
    std::vector<uint32_t> realigned;    // backing store of `steps` when the pool was misaligned
};
// FGFA_OK, or FGFA_ERR_TOO_LARGE when a count does not fit the format's u32 ids (pool.rs:9-11).
int pool_arrays_of(const FlatGFA& gfa, PoolArrays* out);

// depth.rs:61-82: the odgi-style TSV table.
struct SegDepth {
    const FlatGFA& gfa;
    std::vector<uint64_t> depths;
    std::vector<uint64_t> uniq_depths;

    // emit.rs:8-19 `Emit`: header `#node.id\tdepth\tdepth.uniq`, then one row per segment
    // in pool order with `seg.name as u32` (truncating cast, depth.rs:71).
    void emit(std::string& out) const;
    void emit(FILE* f) const;
    void print() const { emit(stdout); }   // emit.rs:13-18
};

// The same table for borrowed arrays of gfa.segs.len() entries, in a malloc'ed, NUL-terminated
// buffer (what the C ABI hands out); nullptr if the allocation fails.
char* seg_depth_table(const FlatGFA& gfa, const uint64_t* depths, const uint64_t* uniq, size_t* len);

// depth.rs:88-113: node depth over ALL paths, then (length in bp, mean depth) of each
// queried path.  `paths` are path pool indices (the reference takes an iterator of ids).
std::pair<std::vector<uint64_t>, std::vector<double>> path_depth(const FlatGFA& gfa,
                                                                 const std::vector<uint32_t>& paths);

// depth.rs:192-197.
std::string format_float(double x, int digits);

// depth.rs:136-160: the odgi-style path-depth TSV.
struct PathDepth {
    const FlatGFA& gfa;
    std::vector<uint64_t> lengths;
    std::vector<double> depths;
    std::vector<uint32_t> paths;
    void emit(std::string& out) const;
    void emit(FILE* f) const;
    void print() const { emit(stdout); }
};

}  // namespace depth
}  // namespace ops
}  // namespace flatgfa
