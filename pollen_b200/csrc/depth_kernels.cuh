// sm_100a kernels for FlatGFA's node-depth query.
//
// Replaces the serial loop of the reference's `seg_depth_with_uniq`
// (flatgfa/src/ops/depth.rs:15-39) and `seg_depth` (depth.rs:45-56):
//
//   kernel A  (k_step_stream)   streams the `steps` pool once.  For every step it
//             decodes Handle -> segment id (flatgfa/src/flatgfa.rs:201-203, `h >> 1`),
//             adds 1 to depth[seg] (depth.rs:29) and sets bit `seg` in the
//             *per-path* `seen` bitmap row (the GPU form of depth.rs:23,26,30-34:
//             one row per path instead of one bitmap cleared between paths).
//   kernel B  (k_uniq_popcount) sums the bitmap rows column-wise with bit-sliced
//             counters: uniq[seg] = #rows with bit seg set (depth.rs:32), and clears
//             the rows it consumed so the next run starts from a zero bitmap.
//
// Work decomposition: a *chunk* is up to kChunk consecutive steps of ONE path
// (a chunk never straddles two paths because the bitmap row depends on the path);
// chunk c of path p starts at the 16-byte-aligned element (start_p & ~3) + c*kChunk
// so that every 128-bit load is aligned even though a path's span start is only
// 4-byte aligned.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fgfa {

constexpr int kThreads = 256;          // threads per CTA in kernel A
constexpr int kItems = 16;             // steps per thread per chunk
constexpr int kChunk = kThreads * kItems;  // 4096 steps = 16 KiB of the pool

struct StreamParams {
    const uint32_t* __restrict__ steps;   // device: the steps pool (Handle words)
    uint64_t n_steps;                     // pool length (for the tail guard)
    const uint32_t* __restrict__ span_start;  // device [n_paths]
    const uint32_t* __restrict__ span_end;    // device [n_paths]
    const uint32_t* __restrict__ chunk_prefix;  // device [n_paths+1]: chunks before path p
    uint32_t path_lo, path_hi;            // this launch covers paths [path_lo, path_hi)
    uint32_t n_segs;
    uint32_t words_per_row;               // bitmap row pitch in 32-bit words
    uint32_t* __restrict__ depth;         // device [n_segs], pre-zeroed
    uint32_t* __restrict__ bitmap;        // device [(path_hi-path_lo) rows][words_per_row], zero
    uint32_t* __restrict__ err;           // sticky error flag (bit 0: segment id out of range)
};

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// Streaming 128-bit load of the steps pool: read-only path, no L1 allocation,
// L2 evict-first so the 1.6 GB stream does not push the depth table and the
// seen-bitmap (both live in L2 between their atomics) out of the cache.
__device__ __forceinline__ uint4 ld_stream_v4(const uint32_t* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p, uint64_t pol) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;"
                 : "=r"(r)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or_b32(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Locate the path that owns global chunk index `c`: the last p with chunk_prefix[p] <= c.
__device__ __forceinline__ uint32_t find_path(const uint32_t* __restrict__ prefix, uint32_t lo,
                                              uint32_t hi, uint32_t c) {
    // invariant: prefix[lo] <= c < prefix[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= c) lo = mid; else hi = mid;
    }
    return lo;
}

enum StreamMode : int {
    kModeDepthAndSeen = 0,   // the product configuration
    kModeDepthOnly = 1,      // seg_depth (depth.rs:45-56) and the roofline split
    kModeSeenOnly = 2,       // measurement only
    kModeReadOnly = 3,       // measurement only: pure stream, no atomics
};

// ---------------------------------------------------------------------------
// kernel A, direct form: every step issues its own L2 reductions, in step order.
// LANE_ORDER=1: lane l of a warp handles step base+l (32 consecutive steps per warp
// instruction, so a near-monotone walk touches few L2 sectors per RED);
// LANE_ORDER=0: each thread handles 4 consecutive steps from one 128-bit load.
// ---------------------------------------------------------------------------
template <int MODE, int LANE_ORDER>
__global__ void __launch_bounds__(kThreads) k_step_stream_direct(StreamParams P) {
    const uint64_t pol = make_evict_first_policy();
    const uint32_t c_lo = __ldg(P.chunk_prefix + P.path_lo);
    const uint32_t c_hi = __ldg(P.chunk_prefix + P.path_hi);
    uint32_t sink = 0;
    for (uint32_t c = c_lo + blockIdx.x; c < c_hi; c += gridDim.x) {
        const uint32_t p = find_path(P.chunk_prefix, P.path_lo, P.path_hi, c);
        const uint32_t s = __ldg(P.span_start + p), e = __ldg(P.span_end + p);
        const uint64_t a = (uint64_t)(s & ~3u) + (uint64_t)(c - __ldg(P.chunk_prefix + p)) * kChunk;
        uint32_t* __restrict__ row = P.bitmap + (size_t)(p - P.path_lo) * P.words_per_row;

        auto visit = [&](uint32_t h, uint64_t idx) {
            if (idx < s || idx >= e) return;
            if (MODE == kModeReadOnly) { sink += h; return; }
            const uint32_t seg = h >> 1;
            if (seg >= P.n_segs) { *P.err = 1u; return; }
            if (MODE == kModeDepthAndSeen || MODE == kModeDepthOnly) red_add_u32(P.depth + seg, 1u);
            if (MODE == kModeDepthAndSeen || MODE == kModeSeenOnly)
                red_or_b32(row + (seg >> 5), 1u << (seg & 31));
        };

        if (LANE_ORDER) {
            uint32_t h[kItems];
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const uint64_t idx = a + (uint64_t)i * kThreads + threadIdx.x;
                h[i] = (idx >= s && idx < e) ? ld_stream_u32(P.steps + idx, pol) : 0u;
            }
#pragma unroll
            for (int i = 0; i < kItems; ++i) visit(h[i], a + (uint64_t)i * kThreads + threadIdx.x);
        } else {
            uint4 v[kItems / 4];
#pragma unroll
            for (int i = 0; i < kItems / 4; ++i) {
                const uint64_t idx = a + ((uint64_t)i * kThreads + threadIdx.x) * 4;
                if (idx + 4 <= P.n_steps && idx < e) {
                    v[i] = ld_stream_v4(P.steps + idx, pol);
                } else {
                    v[i].x = idx + 0 < e ? P.steps[idx + 0] : 0u;
                    v[i].y = idx + 1 < e ? P.steps[idx + 1] : 0u;
                    v[i].z = idx + 2 < e ? P.steps[idx + 2] : 0u;
                    v[i].w = idx + 3 < e ? P.steps[idx + 3] : 0u;
                }
            }
#pragma unroll
            for (int i = 0; i < kItems / 4; ++i) {
                const uint64_t idx = a + ((uint64_t)i * kThreads + threadIdx.x) * 4;
                visit(v[i].x, idx + 0);
                visit(v[i].y, idx + 1);
                visit(v[i].z, idx + 2);
                visit(v[i].w, idx + 3);
            }
        }
    }
    if (MODE == kModeReadOnly && sink == 0xDEADBEEFu) *P.err = 2u;
}

// ---------------------------------------------------------------------------
// kernel B: uniq[seg] (+)= number of bitmap rows whose bit `seg` is set.
// One thread owns one 32-segment column word; rows are added into bit-sliced
// (vertical) counters, 7 planes = up to 127 rows per group, then the planes are
// expanded into 32 per-segment counts.  Consumed bitmap words are zeroed.
// ---------------------------------------------------------------------------
struct PopcountParams {
    uint32_t* __restrict__ bitmap;   // [n_rows][words_per_row]
    uint32_t n_rows;
    uint32_t words_per_row;
    uint32_t n_words;                // ceil(n_segs / 32)
    uint32_t n_segs;
    uint32_t* __restrict__ uniq;     // [n_segs]
    int accumulate;                  // 0: uniq = count, 1: uniq += count
};

__global__ void __launch_bounds__(256) k_uniq_popcount(PopcountParams P) {
    __shared__ uint32_t tile[8][32][33];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) cnt[j] = 0;
    if (w < P.n_words) {
        uint32_t* col = P.bitmap + w;
        for (uint32_t r0 = 0; r0 < P.n_rows; r0 += 127) {
            const uint32_t r1 = min(P.n_rows, r0 + 127);
            uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0;
            for (uint32_t r = r0; r < r1; ++r) {
                uint32_t* q = col + (size_t)r * P.words_per_row;
                uint32_t x = *q;
                if (x) {
                    *q = 0u;
                    uint32_t t;
                    t = c0 & x; c0 ^= x; x = t;
                    t = c1 & x; c1 ^= x; x = t;
                    t = c2 & x; c2 ^= x; x = t;
                    t = c3 & x; c3 ^= x; x = t;
                    t = c4 & x; c4 ^= x; x = t;
                    t = c5 & x; c5 ^= x; x = t;
                    c6 ^= x;
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                cnt[j] += ((c0 >> j) & 1u) | (((c1 >> j) & 1u) << 1) | (((c2 >> j) & 1u) << 2) |
                          (((c3 >> j) & 1u) << 3) | (((c4 >> j) & 1u) << 4) |
                          (((c5 >> j) & 1u) << 5) | (((c6 >> j) & 1u) << 6);
            }
        }
    }
    // Transpose through shared memory so that each warp store covers 128 contiguous bytes.
#pragma unroll
    for (int j = 0; j < 32; ++j) tile[warp][lane][j] = cnt[j];
    __syncwarp();
    const uint32_t w_base = blockIdx.x * blockDim.x + warp * 32;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const uint64_t seg = ((uint64_t)(w_base + k) << 5) + lane;
        if (seg < P.n_segs) {
            const uint32_t v = tile[warp][k][lane];
            if (P.accumulate) P.uniq[seg] += v; else P.uniq[seg] = v;
        }
    }
}

}  // namespace fgfa
