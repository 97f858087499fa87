// sm_100a kernels for FlatGFA's node-depth query.
//
// Replaces the serial loop of the reference's `seg_depth_with_uniq`
// (flatgfa/src/ops/depth.rs:15-39) and `seg_depth` (depth.rs:45-56):
//
//   kernel A  (k_step_stream_merged) streams the `steps` pool once.  For every step it
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

// 16-byte-vector index swizzle for the staged chunk: conflict-free for the coalesced
// staging stores, the per-thread 64-byte reads of pass 2 and the lane-order reads of pass 3.
__device__ __forceinline__ uint32_t swz(uint32_t v) {
    return (v & ~7u) | ((v & 7u) ^ ((v >> 3) & 7u));
}
// 1 << (c & 31) as a single funnel shift in wrap mode.
__device__ __forceinline__ uint32_t bit_of(uint32_t c) {
    uint32_t r;
    asm("shf.l.wrap.b32 %0, 0, 1, %1;" : "=r"(r) : "r"(c));
    return r;
}
// Opaque copies: stop ptxas from rematerialising cheap values (S2R, constant-bank
// loads) inside the hot loops.
__device__ __forceinline__ uint32_t keep(uint32_t v) { asm volatile("mov.b32 %0, %0;" : "+r"(v)); return v; }
template <typename T>
__device__ __forceinline__ T* keep_ptr(T* p) { asm volatile("mov.b64 %0, %0;" : "+l"(p)); return p; }

__device__ __forceinline__ void load_thread_steps(uint32_t (&h)[kItems], const uint4* ld_base,
                                                  uint32_t ld_lo, uint32_t swc) {
#pragma unroll
    for (int j = 0; j < kItems / 4; ++j) {
        const uint4 x = ld_base[(ld_lo + j) ^ swc];
        h[4 * j + 0] = x.x; h[4 * j + 1] = x.y; h[4 * j + 2] = x.z; h[4 * j + 3] = x.w;
    }
}


// ---------------------------------------------------------------------------
// kernel A, merged-direct form.
// Same staging as the first-touch form, but no returning atomics: pass 2 merges the
// seen-bitmap updates of each thread's consecutive steps into one RED.OR per run,
// pass 3 issues one depth RED.ADD per step in lane order.  uniq comes from kernel B;
// depth is complete after this kernel (kernel B must NOT add the first visits).
// ---------------------------------------------------------------------------
template <int BLOCKS_PER_SM, bool WITH_SEEN>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM) k_step_stream_merged(StreamParams P) {
    __shared__ uint4 s_steps[kChunk / 4];
    const uint64_t pol = make_evict_first_policy();
    const uint32_t c_lo = __ldg(P.chunk_prefix + P.path_lo);
    const uint32_t c_hi = __ldg(P.chunk_prefix + P.path_hi);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg_limit = P.n_segs * 2u;
    uint4* const st_ptr = s_steps + swz(tid);
    const uint32_t swc = (tid >> 1) & 7u;
    const uint4* const ld_base = s_steps + ((tid * 4u) & ~7u);
    const uint32_t ld_lo = 4u * (tid & 1u);
    const uint32_t* const s_words = reinterpret_cast<const uint32_t*>(s_steps);
    const uint32_t* const p3_ptr = keep_ptr(s_words + ((warp * 8u + ((lane >> 2) ^ (warp & 7u))) * 4u + (lane & 3u)));
    uint32_t* const depth_ptr = keep_ptr(P.depth);
    const uint32_t one = keep(1u);

    for (uint32_t c = c_lo + blockIdx.x; c < c_hi; c += gridDim.x) {
        const uint32_t p = find_path(P.chunk_prefix, P.path_lo, P.path_hi, c);
        const uint32_t s = __ldg(P.span_start + p), e = __ldg(P.span_end + p);
        const uint64_t a = (uint64_t)(s & ~3u) + (uint64_t)(c - __ldg(P.chunk_prefix + p)) * kChunk;
        uint32_t* __restrict__ row = P.bitmap + (size_t)(p - P.path_lo) * P.words_per_row;
        const bool full = a >= s && a + kChunk <= e && a + kChunk <= P.n_steps;

        // ---- pass 1: stage; out-of-span slots are filled with an invalid handle ----
        if (full) {
            const uint32_t* src = P.steps + a + tid * 4u;
#pragma unroll
            for (int j = 0; j < kItems / 4; ++j)
                st_ptr[j * kThreads] = ld_stream_v4(src + j * kThreads * 4, pol);
        } else {
#pragma unroll
            for (int j = 0; j < kItems / 4; ++j) {
                const uint64_t idx = a + (uint64_t)(tid + j * kThreads) * 4;
                uint32_t x[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    x[k] = (idx + k >= s && idx + k < e) ? P.steps[idx + k] : 0xFFFFFFFFu;
                st_ptr[j * kThreads] = make_uint4(x[0], x[1], x[2], x[3]);
            }
        }
        __syncthreads();

        // ---- pass 2: thread order -- one RED.OR per run of steps in the same bitmap word ----
        if (WITH_SEEN) {
            uint32_t h[kItems];
            load_thread_steps(h, ld_base, ld_lo, swc);
            uint32_t hmax = 0;
#pragma unroll
            for (int i = 0; i < kItems; ++i) hmax = max(hmax, h[i]);
            if (hmax < seg_limit) {
                uint32_t acc = 0;
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    const uint32_t bit = bit_of(h[i] >> 1);
                    acc = ((i > 0 && ((h[i] ^ h[i - 1]) < 64u)) ? acc : 0u) | bit;
                    if (i == kItems - 1 || ((h[i] ^ h[i + (i < kItems - 1)]) >= 64u))
                        red_or_b32(row + (h[i] >> 6), acc);
                }
            } else {
                // edge chunk (invalid filler) or out-of-range segment id: per-step, unmerged
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    if (h[i] < seg_limit) red_or_b32(row + (h[i] >> 6), bit_of(h[i] >> 1));
                    else if (h[i] != 0xFFFFFFFFu) *P.err = 1u;
                }
            }
        }

        // ---- pass 3: lane order -- one depth RED.ADD per step ----
        if (full) {
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const uint32_t hh = p3_ptr[i * kThreads];
                if (hh < seg_limit) red_add_u32(depth_ptr + (hh >> 1), one);
                else if (!WITH_SEEN) *P.err = 1u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const uint32_t hh = p3_ptr[i * kThreads];
                if (hh < seg_limit) red_add_u32(depth_ptr + (hh >> 1), one);
                else if (!WITH_SEEN && hh != 0xFFFFFFFFu) *P.err = 1u;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// kernel B: column-wise population count of the seen-bitmap.
//   cnt[seg] = number of bitmap rows (paths of this batch) whose bit `seg` is set
//   uniq[seg]  = cnt  (first batch)  or  += cnt (later batches)      -- depth.rs:32
//   depth[seg] += cnt                 (first visits; kernel A added the repeats)
// One thread owns one 32-segment column word; rows are added into bit-sliced
// (vertical) counters -- 7 planes, flushed every 120 rows -- with 8 independent row
// loads in flight, then the planes are expanded into 32 per-segment counts and
// transposed through shared memory so that every warp store covers 128 contiguous
// bytes.  Consumed bitmap words are zeroed so the next run starts from a clean bitmap.
// ---------------------------------------------------------------------------
struct PopcountParams {
    uint32_t* __restrict__ bitmap;   // [n_rows][words_per_row]
    uint32_t n_rows;
    uint32_t words_per_row;
    uint32_t n_words;                // ceil(n_segs / 32)
    uint32_t n_segs;
    uint32_t* __restrict__ uniq;     // [n_segs] or nullptr (seg_depth: no unique depth)
    uint32_t* __restrict__ depth;    // [n_segs] or nullptr
    int accumulate;                  // 0: uniq = cnt, 1: uniq += cnt
};

constexpr int kPopThreads = 128;
constexpr int kPopRowsInFlight = 8;

__global__ void __launch_bounds__(kPopThreads) k_uniq_popcount(PopcountParams P) {
    __shared__ uint32_t tile[kPopThreads / 32][32][33];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kPopThreads + threadIdx.x;
    uint32_t cnt[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) cnt[j] = 0;
    if (w < P.n_words) {
        uint32_t* col = P.bitmap + w;
        for (uint32_t g0 = 0; g0 < P.n_rows; g0 += 120) {
            const uint32_t g1 = min(P.n_rows, g0 + 120);
            uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0;
            for (uint32_t r0 = g0; r0 < g1; r0 += kPopRowsInFlight) {
                uint32_t x[kPopRowsInFlight];
#pragma unroll
                for (int k = 0; k < kPopRowsInFlight; ++k)
                    x[k] = (r0 + k < g1) ? col[(size_t)(r0 + k) * P.words_per_row] : 0u;
#pragma unroll
                for (int k = 0; k < kPopRowsInFlight; ++k) {
                    uint32_t v = x[k];
                    if (v) {
                        col[(size_t)(r0 + k) * P.words_per_row] = 0u;
                        uint32_t t;
                        t = c0 & v; c0 ^= v; v = t;
                        t = c1 & v; c1 ^= v; v = t;
                        t = c2 & v; c2 ^= v; v = t;
                        t = c3 & v; c3 ^= v; v = t;
                        t = c4 & v; c4 ^= v; v = t;
                        t = c5 & v; c5 ^= v; v = t;
                        c6 ^= v;
                    }
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
#pragma unroll
    for (int j = 0; j < 32; ++j) tile[warp][lane][j] = cnt[j];
    __syncwarp();
    const uint32_t w_base = blockIdx.x * kPopThreads + warp * 32;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const uint64_t seg = ((uint64_t)(w_base + k) << 5) + lane;
        if (seg < P.n_segs) {
            const uint32_t v = tile[warp][k][lane];
            if (P.uniq) { if (P.accumulate) P.uniq[seg] += v; else P.uniq[seg] = v; }
            if (P.depth && v) P.depth[seg] += v;
        }
    }
}

}  // namespace fgfa
