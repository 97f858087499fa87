// sm_100a kernels for interval ("window") depth along one path.
//
// GPU form of the reference's flatgfa/src/ops/window_depth.rs:
//   path_length      (:69-77)    total base pairs of a path
//   weighted_depths  (:84-103)   per step: depth[seg] * len(seg) as f64 and the step's [pos, pos+len) range
//   overlap          (:110-112)
//   assign_depths    (:118-153)  one cursor walks the intervals while the steps stream by; every
//                                interval accumulates, in step order, seg.depth * overlap_fraction / interval_length
// The reference's loop is a sequential merge of two sorted lists.  Here it is restated as
//   W1  k_tile_reduce / k_scan_tile_totals / k_tile_scan : inclusive prefix sum of len(seg) over the
//       path's steps -> seg_end[j] (u64), the end offset of step j; seg_end[n-1] is path_length
//   W2  k_interval_lower_bound : lb[w] = first step j with seg_end[j] >= interval_w.end (binary search)
//       followed by a prefix MAX over w (the same three scan kernels): fin[w] = the step at which the
//       reference's cursor leaves interval w.  For intervals sorted along the path lb is already
//       monotone; the prefix max reproduces the cursor for unsorted or overlapping input as well.
//   W3  k_interval_accumulate : interval w receives the steps fin[w-1] .. min(fin[w], n-1) (both
//       inclusive: the step that finishes w-1 is offered to w too, window_depth.rs:128-150), summed in
//       step order with the reference's exact f64 operation sequence (u64 -> f64 conversions,
//       one divide, one multiply, one divide, one add; no contraction), so the result is bit-exact.
// Short intervals are summed one per thread (W3a); long ones go to a worklist and are summed one per
// warp (W3b: lanes form the terms in parallel, the additions stay in order).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fgfa {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;   // 4096 elements per CTA

struct OpSum { __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a + b; } };
struct OpMax { __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; } };

// element j of the scanned sequence = len(segment of step j)   (window_depth.rs:73-74, :93-94)
struct LoadStepLen {
    const uint32_t* __restrict__ steps;      // first step of the path
    const uint32_t* __restrict__ seg_len;
    uint32_t n_segs;
    uint32_t* __restrict__ err;
    __device__ __forceinline__ uint64_t operator()(uint64_t j) const {
        const uint32_t seg = steps[j] >> 1;
        if (seg >= n_segs) { *err = 1u; return 0; }
        return __ldg(seg_len + seg);
    }
};
struct LoadU32 {                              // no __restrict__: the prefix max runs in place
    const uint32_t* v;
    __device__ __forceinline__ uint64_t operator()(uint64_t j) const { return v[j]; }
};
struct StoreU64 {
    uint64_t* __restrict__ out;
    __device__ __forceinline__ void operator()(uint64_t j, uint64_t x) const { out[j] = x; }
};
struct StoreU32 {
    uint32_t* out;
    __device__ __forceinline__ void operator()(uint64_t j, uint64_t x) const { out[j] = (uint32_t)x; }
};

// Block-wide inclusive scan of one value per thread; returns the inclusive value and the block total.
template <typename Op>
__device__ __forceinline__ uint64_t block_scan_inclusive(uint64_t v, Op op, uint64_t* s_warp, uint64_t& total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t u = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= (uint32_t)o) v = op(u, v);
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    uint64_t before = 0, tot = 0;                    // 0 is the identity of both operators (unsigned sum, unsigned max)
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        const uint64_t x = s_warp[w];
        if ((uint32_t)w < warp) before = op(before, x);
        tot = op(tot, x);
    }
    total = tot;
    __syncthreads();
    return op(before, v);
}

// phase 1: tile_total[t] = op over the tile's elements (identity 0 for both operators used here)
template <typename Load, typename Op>
__global__ void __launch_bounds__(kScanThreads) k_tile_reduce(Load load, Op op, uint64_t n, uint64_t* __restrict__ tile_total) {
    __shared__ uint64_t s_warp[kScanThreads / 32];
    const uint64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t base = t * kScanTile;
        uint64_t acc = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const uint64_t j = base + (uint64_t)i * kScanThreads + threadIdx.x;   // coalesced
            if (j < n) acc = op(acc, load(j));
        }
        uint64_t total;
        block_scan_inclusive(acc, op, s_warp, total);
        if (threadIdx.x == 0) tile_total[t] = total;
    }
}

// phase 2 (one CTA): tile_total[t] <- op over the totals of tiles 0..t-1 (exclusive)
template <typename Op>
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_totals(Op op, uint64_t n_tiles, uint64_t* __restrict__ tile_total) {
    __shared__ uint64_t s_warp[kScanThreads / 32];
    __shared__ uint64_t s_last[kScanThreads / 32];
    uint64_t carry = 0;
    for (uint64_t base = 0; base < n_tiles; base += kScanThreads) {
        const uint64_t t = base + threadIdx.x;
        const uint64_t mine = t < n_tiles ? tile_total[t] : 0;
        uint64_t total;
        const uint64_t incl = block_scan_inclusive(mine, op, s_warp, total);
        // exclusive value = carry op (inclusive of the previous thread)
        const uint64_t prev = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        if ((threadIdx.x & 31) == 31) s_last[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t excl;
        if (threadIdx.x == 0) excl = carry;
        else if ((threadIdx.x & 31) == 0) excl = op(carry, s_last[(threadIdx.x >> 5) - 1]);
        else excl = op(carry, prev);
        if (t < n_tiles) tile_total[t] = excl;
        carry = op(carry, total);
        __syncthreads();
    }
}

// phase 3: out[j] = tile_offset op (inclusive scan inside the tile).  A thread owns kScanItems
// consecutive elements; the tile is staged through shared memory so global loads stay coalesced.
template <typename Load, typename Op, typename Store>
__global__ void __launch_bounds__(kScanThreads) k_tile_scan(Load load, Op op, Store store, uint64_t n,
                                                            const uint64_t* __restrict__ tile_offset) {
    __shared__ uint64_t s_val[kScanTile + kScanTile / kScanItems];   // one pad word per thread: conflict-free rows
    __shared__ uint64_t s_warp[kScanThreads / 32];
    __shared__ uint64_t s_last[kScanThreads / 32];
    const uint64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t base = t * kScanTile;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const uint32_t k = (uint32_t)i * kScanThreads + threadIdx.x;
            const uint64_t j = base + k;
            s_val[k + k / kScanItems] = j < n ? load(j) : 0;
        }
        __syncthreads();
        uint64_t v[kScanItems];
        const uint32_t r0 = threadIdx.x * kScanItems;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const uint64_t x = s_val[r0 + i + threadIdx.x];
            v[i] = i == 0 ? x : op(v[i - 1], x);
        }
        uint64_t total;
        const uint64_t incl = block_scan_inclusive(v[kScanItems - 1], op, s_warp, total);
        // what precedes this thread: tile offset, then the inclusive value of the previous thread
        const uint64_t prev = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        if ((threadIdx.x & 31) == 31) s_last[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t before = tile_offset[t];
        if (threadIdx.x != 0) before = op(before, (threadIdx.x & 31) == 0 ? s_last[(threadIdx.x >> 5) - 1] : prev);
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) s_val[r0 + i + threadIdx.x] = op(before, v[i]);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const uint32_t k = (uint32_t)i * kScanThreads + threadIdx.x;
            const uint64_t j = base + k;
            if (j < n) store(j, s_val[k + k / kScanItems]);
        }
        __syncthreads();
    }
}

// The equally sized windows of `Windows` (window_depth.rs:27-38, :41-52): [w*size, min((w+1)*size, end)).
__global__ void __launch_bounds__(256) k_make_windows(uint64_t start, uint64_t end, uint64_t size, uint64_t n_win,
                                                      uint64_t* __restrict__ win_start, uint64_t* __restrict__ win_end) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const uint64_t a = start + w * size;
    const uint64_t b = a + size;
    win_start[w] = a;
    win_end[w] = (b < a || b > end) ? end : b;
}

// W2: lb[w] = first step j with seg_end[j] >= win_end[w]  (n if the interval ends past the path)
__global__ void __launch_bounds__(256) k_interval_lower_bound(const uint64_t* __restrict__ seg_end, uint32_t n,
                                                              const uint64_t* __restrict__ win_end, uint64_t n_win,
                                                              uint32_t* __restrict__ lb) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const uint64_t key = win_end[w];
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (seg_end[mid] >= key) hi = mid; else lo = mid + 1;
    }
    lb[w] = lo;
}

struct IntervalParams {
    const uint32_t* __restrict__ steps;      // first step of the path
    uint32_t n;                              // steps in the path
    const uint32_t* __restrict__ depth;      // [n_segs] node depth (seg_depth, window_depth.rs:177)
    const uint32_t* __restrict__ seg_len;    // [n_segs]
    uint32_t n_segs;
    const uint64_t* __restrict__ seg_end;    // [n] W1
    const uint64_t* __restrict__ win_start;  // [n_win]
    const uint64_t* __restrict__ win_end;
    uint64_t n_win;
    const uint32_t* __restrict__ fin;        // [n_win] W2
    double* __restrict__ out;                // [n_win]
    uint32_t* __restrict__ long_list;        // [n_win] intervals left to W3b
    uint32_t* __restrict__ long_count;       // zero on entry
};

constexpr uint32_t kLongInterval = 48;       // steps; longer intervals are summed by the whole warp

// One term of assign_depths (window_depth.rs:131-139).  Returns false when the step and the interval
// do not overlap (the reference adds nothing then).
__device__ __forceinline__ bool interval_term(const IntervalParams& P, uint32_t j, uint64_t w0, uint64_t w1, double& term) {
    const uint64_t s1 = P.seg_end[j];
    const uint64_t s0 = j ? P.seg_end[j - 1] : 0;
    const uint64_t a = w0 > s0 ? w0 : s0;                        // overlap(), :110-112
    const uint64_t b = w1 < s1 ? w1 : s1;
    if (!(b > a)) return false;
    const uint32_t seg = P.steps[j] >> 1;
    if (seg >= P.n_segs) return false;                           // flagged by W1 already
    const uint64_t total = (uint64_t)__ldg(P.depth + seg) * (uint64_t)__ldg(P.seg_len + seg);   // :97, wrapping usize
    const double seg_depth = __ull2double_rn(total);                                            // :99
    const double amt = __ddiv_rn(__ull2double_rn(b - a), __ull2double_rn(s1 - s0));             // :134
    term = __ddiv_rn(__dmul_rn(seg_depth, amt), __ull2double_rn(w1 - w0));                      // :136
    return true;
}

// The range of steps interval w is offered (see W3 above); empty ranges come back as j0 > j1.
__device__ __forceinline__ void interval_range(const IntervalParams& P, uint64_t w, uint32_t& j0, uint32_t& j1) {
    j0 = 1;
    j1 = 0;
    if (P.n) {
        const uint32_t a = w ? P.fin[w - 1] : 0u;
        const uint32_t f = P.fin[w];
        const uint32_t b = f < P.n ? f : P.n - 1;
        if (a <= b) { j0 = a; j1 = b; }
    }
}

// W3a: one thread per interval.  Short intervals are summed here; long ones (>= kLongInterval
// steps) are appended to a worklist for W3b, so that a BED file with a few chromosome-sized
// intervals still spreads over the whole GPU instead of serialising inside one warp.
__global__ void __launch_bounds__(256) k_interval_accumulate(IntervalParams P) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.n_win) return;
    const uint64_t w0 = P.win_start[w], w1 = P.win_end[w];
    uint32_t j0, j1;
    interval_range(P, w, j0, j1);
    if (j1 >= j0 && (j1 - j0) >= kLongInterval) {
        P.long_list[atomicAdd(P.long_count, 1u)] = (uint32_t)w;        // n_win <= 2^32 is enforced by the host
        return;
    }
    double acc = 0.0;                                            // :119
    for (uint32_t j = j0; j <= j1; ++j) {                        // j1 <= n-1 < 2^32-1: no wrap
        double t;
        if (interval_term(P, j, w0, w1, t)) acc = __dadd_rn(acc, t);   // :135
    }
    P.out[w] = acc;
}

// W3b: one warp per long interval.  The lanes form 32 terms in parallel (and already fetch the
// next 32 while the current ones are being added); the additions stay in step order.
__global__ void __launch_bounds__(256) k_interval_accumulate_long(IntervalParams P) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_long = *P.long_count;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < n_long; idx += warps) {   // warp-uniform
        const uint64_t w = P.long_list[idx];
        const uint64_t w0 = P.win_start[w], w1 = P.win_end[w];
        uint32_t j0, j1;
        interval_range(P, w, j0, j1);
        double sum = 0.0;
        uint64_t base = j0;
        double t_next = 0.0;
        bool ok_next = base + lane <= j1 && interval_term(P, (uint32_t)(base + lane), w0, w1, t_next);
        while (true) {                                           // warp-uniform trip count
            const double t = t_next;
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, ok_next);
            const uint64_t next = base + 32;
            const bool more = next <= j1;
            if (more) {
                t_next = 0.0;
                ok_next = next + lane <= j1 && interval_term(P, (uint32_t)(next + lane), w0, w1, t_next);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const double tk = __shfl_sync(0xFFFFFFFFu, t, k);
                if ((mask >> k) & 1u) sum = __dadd_rn(sum, tk);
            }
            if (!more) break;
            base = next;
        }
        if (lane == 0) P.out[w] = sum;
    }
}

}  // namespace fgfa
