// Experimental kernel-A variants kept for the microbenchmark only (tools/ubench.cu).
// None of these is used by the product library; see DESIGN.md "What was tried".
#pragma once
#include "../pollen_b200/csrc/depth_kernels.cuh"

namespace fgfa {

// The experimental kernels keep the earlier work decomposition (chunk prefix + binary search).
struct ExpParams {
    const uint32_t* __restrict__ steps;
    uint64_t n_steps;
    const uint32_t* __restrict__ span_start;
    const uint32_t* __restrict__ span_end;
    const uint32_t* __restrict__ chunk_prefix;
    uint32_t path_lo, path_hi;
    uint32_t n_segs;
    uint32_t words_per_row;
    uint32_t* __restrict__ depth;
    uint32_t* __restrict__ bitmap;
    uint32_t* __restrict__ err;
};
__device__ __forceinline__ uint32_t find_path(const uint32_t* __restrict__ prefix, uint32_t lo,
                                              uint32_t hi, uint32_t c) {
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= c) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ void load_thread_steps(uint32_t (&h)[kItems], const uint4* ld_base,
                                                  uint32_t ld_lo, uint32_t swc) {
#pragma unroll
    for (int j = 0; j < kItems / 4; ++j) {
        const uint4 x = ld_base[(ld_lo + j) ^ swc];
        h[4 * j + 0] = x.x; h[4 * j + 1] = x.y; h[4 * j + 2] = x.z; h[4 * j + 3] = x.w;
    }
}

enum StreamMode : int {
    kModeDepthAndSeen = 0,   // the product configuration
    kModeDepthOnly = 1,      // seg_depth (depth.rs:45-56) and the roofline split
    kModeSeenOnly = 2,       // measurement only
    kModeReadOnly = 3,       // measurement only: pure stream, no atomics
    kModeDepthHalfLanes = 4, // measurement only: RED on even lanes only (lane- vs request-bound?)
    kModeDepthPairs = 5,     // measurement only: even lanes add 2 (same sectors, half the lanes)
};

// ---------------------------------------------------------------------------
// kernel A, direct form: every step issues its own L2 reductions, in step order.
// LANE_ORDER=1: lane l of a warp handles step base+l (32 consecutive steps per warp
// instruction, so a near-monotone walk touches few L2 sectors per RED);
// LANE_ORDER=0: each thread handles 4 consecutive steps from one 128-bit load.
// ---------------------------------------------------------------------------
template <int MODE, int LANE_ORDER>
__global__ void __launch_bounds__(kThreads) k_step_stream_direct(ExpParams P) {
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
            if (MODE == kModeDepthHalfLanes) { if ((seg & 1u) == 0) red_add_u32(P.depth + seg, 1u); }
            if (MODE == kModeDepthPairs) { if ((threadIdx.x & 1u) == 0) red_add_u32(P.depth + seg, 2u); }
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

__device__ __forceinline__ uint32_t atom_or_b32(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.relaxed.gpu.global.or.b32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void red_add_u32_if_nz(uint32_t x, uint32_t* p, uint32_t v) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t"
        "@q red.relaxed.gpu.global.add.u32 [%0], %1;\n\t}"
        :: "l"(p), "r"(v), "r"(x) : "memory");
}
__device__ __forceinline__ uint32_t shr_wrap(uint32_t x, uint32_t c) {
    uint32_t r;
    asm("shf.r.wrap.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(x), "r"(c));
    return r;
}
constexpr uint32_t kInvalidWord = 0xFFFFFFFFu;

// Generic pass 2 for one thread: per-step validity, bounds check, run merge, fetch-or.
// Returns the 16-bit repeat mask.  Only edge chunks come here, so it is kept out of
// line to leave the fast path's register allocation alone.
__device__ __noinline__ uint32_t pass2_generic(const uint4* ld_base, uint32_t ld_lo, uint32_t swc,
                                               uint64_t idx0, uint32_t s, uint32_t e,
                                               uint32_t n_segs, uint32_t* row, uint32_t* err) {
    uint32_t h[kItems];
    load_thread_steps(h, ld_base, ld_lo, swc);
    uint32_t word[kItems];
    uint32_t validmask = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const uint64_t idx = idx0 + i;
        const uint32_t seg = h[i] >> 1;
        bool valid = idx >= s && idx < e;
        if (valid && seg >= n_segs) { *err = 1u; valid = false; }
        word[i] = valid ? (seg >> 5) : kInvalidWord;
        validmask |= (uint32_t)valid << i;
    }
    if (!validmask) return 0u;
    uint32_t old[kItems];
    uint32_t acc = 0, dup = 0, flags = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const uint32_t bit = 1u << ((h[i] >> 1) & 31);
        if (i > 0 && word[i] != word[i - 1]) acc = 0;
        dup |= ((acc & bit) ? 1u : 0u) << i;
        acc |= bit;
        const bool fl = (i == kItems - 1) || (word[i + (i < kItems - 1)] != word[i]);
        old[i] = 0;
        if (fl && word[i] != kInvalidWord) old[i] = atom_or_b32(row + word[i], acc);
    }
    uint32_t cur = 0;
#pragma unroll
    for (int i = kItems - 1; i >= 0; --i) {
        if (i == kItems - 1 || word[i + (i < kItems - 1)] != word[i]) cur = old[i];
        flags |= (((dup >> i) | (cur >> ((h[i] >> 1) & 31))) & 1u) << i;
    }
    return flags & validmask;
}

// In-run duplicate mask (a path stepping on the same segment twice inside one merged
// run); rare, so recomputed out of line only when the fast path saw any duplicate.
__device__ __noinline__ uint32_t dup_mask(const uint4* ld_base, uint32_t ld_lo, uint32_t swc) {
    uint32_t h[kItems];
    load_thread_steps(h, ld_base, ld_lo, swc);
    uint32_t acc = 0, dup = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const uint32_t bit = 1u << ((h[i] >> 1) & 31);
        if (i > 0 && (h[i] >> 6) != (h[i - 1] >> 6)) acc = 0;
        dup |= ((acc & bit) ? 1u : 0u) << i;
        acc |= bit;
    }
    return dup;
}

template <int BLOCKS_PER_SM, int EXPERIMENT = 0>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM) k_step_stream_first_touch(ExpParams P) {
    __shared__ uint4 s_steps[kChunk / 4];
    __shared__ uint32_t s_flags[kChunk / 32];   // bit l of word m: step 32m+l is a repeat visit
    const uint64_t pol = make_evict_first_policy();
    const uint32_t c_lo = __ldg(P.chunk_prefix + P.path_lo);
    const uint32_t c_hi = __ldg(P.chunk_prefix + P.path_hi);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg_limit = P.n_segs * 2u;   // h < seg_limit  <=>  (h >> 1) < n_segs  (n_segs < 2^31)

    // chunk-invariant shared-memory addresses
    uint4* const st_ptr = s_steps + swz(tid);                 // pass 1: + j*kThreads vectors
    const uint32_t swc = (tid >> 1) & 7u;
    const uint4* const ld_base = s_steps + ((tid * 4u) & ~7u);  // pass 2: + ((4*(tid&1)+j) ^ swc)
    const uint32_t ld_lo = 4u * (tid & 1u);
    const uint32_t* const s_words = reinterpret_cast<const uint32_t*>(s_steps);
    // pass 3: element el = i*kThreads + tid lives in vector i*64 + warp*8 + lane/4
    const uint32_t* const p3_ptr = keep_ptr(s_words + ((warp * 8u + ((lane >> 2) ^ (warp & 7u))) * 4u + (lane & 3u)));
    uint16_t* const flag16 = reinterpret_cast<uint16_t*>(s_flags) + tid;
    const uint32_t lanebit = keep(1u << lane), one = keep(1u);
    uint32_t* const depth_ptr = keep_ptr(P.depth);
    const uint32_t* const flag_row = keep_ptr(s_flags + warp);

    for (uint32_t c = c_lo + blockIdx.x; c < c_hi; c += gridDim.x) {
        const uint32_t p = find_path(P.chunk_prefix, P.path_lo, P.path_hi, c);
        const uint32_t s = __ldg(P.span_start + p), e = __ldg(P.span_end + p);
        const uint64_t a = (uint64_t)(s & ~3u) + (uint64_t)(c - __ldg(P.chunk_prefix + p)) * kChunk;
        uint32_t* __restrict__ row = P.bitmap + (size_t)(p - P.path_lo) * P.words_per_row;
        const bool full = a >= s && a + kChunk <= e && a + kChunk <= P.n_steps;

        // ---- pass 1: stage the chunk ----
        if (full) {
            const uint32_t* src = P.steps + a + tid * 4u;
#pragma unroll
            for (int j = 0; j < kItems / 4; ++j)
                st_ptr[j * kThreads] = ld_stream_v4(src + j * kThreads * 4, pol);
        } else {
#pragma unroll
            for (int j = 0; j < kItems / 4; ++j) {
                const uint64_t idx = a + (uint64_t)(tid + j * kThreads) * 4;
                uint4 x = make_uint4(0u, 0u, 0u, 0u);
                if (idx < e) {
                    if (idx + 4 <= P.n_steps) {
                        x = ld_stream_v4(P.steps + idx, pol);
                    } else {
                        x.x = P.steps[idx];
                        if (idx + 1 < P.n_steps) x.y = P.steps[idx + 1];
                        if (idx + 2 < P.n_steps) x.z = P.steps[idx + 2];
                    }
                }
                st_ptr[j * kThreads] = x;
            }
        }
        __syncthreads();

        // ---- pass 2: thread order -- run merge + fetch-or + repeat flags ----
        {
            uint32_t h[kItems];
            load_thread_steps(h, ld_base, ld_lo, swc);
            uint32_t hmax = 0;
#pragma unroll
            for (int i = 0; i < kItems; ++i) hmax = max(hmax, h[i]);
            uint32_t flags;
            if (full && hmax < seg_limit) {
                uint32_t old[kItems];
                uint32_t acc = 0, anydup = 0;
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    const uint32_t bit = bit_of(h[i] >> 1);
                    const uint32_t prev = (i > 0 && ((h[i] ^ h[i - 1]) < 64u)) ? acc : 0u;
                    anydup |= prev & bit;
                    acc = prev | bit;
                    const uint32_t xn = (i == kItems - 1) ? 64u : (h[i] ^ h[i + (i < kItems - 1)]);
                    if (EXPERIMENT == 2) { if (xn >= 64u) red_or_b32(row + (h[i] >> 6), acc); }
                    else if (EXPERIMENT == 3) old[i] = acc * 3u;
                    else if (xn >= 64u) old[i] = atom_or_b32(row + (h[i] >> 6), acc);
                }
                uint32_t cur = 0;
                flags = 0;
#pragma unroll
                for (int i = kItems - 1; i >= 0; --i) {
                    if (i == kItems - 1 || ((h[i] ^ h[i + (i < kItems - 1)]) >= 64u)) cur = old[i];
                    flags |= (shr_wrap(cur, h[i] >> 1) & 1u) << i;
                }
                if (anydup) flags |= dup_mask(ld_base, ld_lo, swc);
            } else {
                flags = pass2_generic(ld_base, ld_lo, swc, a + (uint64_t)tid * kItems, s, e, P.n_segs, row, P.err);
            }
            *flag16 = (uint16_t)flags;
        }
        __syncthreads();

        // ---- pass 3: lane order -- depth REDs for repeat visits only ----
        if (EXPERIMENT == 0) {
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const uint32_t f = flag_row[i * (kThreads / 32)];
                if (f == 0u) continue;   // warp-uniform: none of these 32 steps is a repeat
                red_add_u32_if_nz(f & lanebit, depth_ptr + (p3_ptr[i * kThreads] >> 1), one);
            }
        } else if (EXPERIMENT == 3) {
            if (s_flags[tid & 127] == 0xABCD1234u) *P.err = 4u;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// kernel A, warp-aggregated first-touch form (no shared memory, no block barriers).
// Lane l of a warp owns step base+l of a 32-step row.  Lanes whose steps fall into the
// same bitmap word are grouped with match.any, their bits are OR-reduced with
// redux.sync, the group's lowest lane issues one returning atom.or and broadcasts the
// old word; each lane then knows whether its step is a first or a repeat visit and
// repeat visits issue their depth RED in lane order straight away.
// ---------------------------------------------------------------------------
template <int BLOCKS_PER_SM>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM) k_step_stream_warp_agg(ExpParams P) {
    const uint64_t pol = make_evict_first_policy();
    const uint32_t c_lo = __ldg(P.chunk_prefix + P.path_lo);
    const uint32_t c_hi = __ldg(P.chunk_prefix + P.path_hi);
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    for (uint32_t c = c_lo + blockIdx.x; c < c_hi; c += gridDim.x) {
        const uint32_t p = find_path(P.chunk_prefix, P.path_lo, P.path_hi, c);
        const uint32_t s = __ldg(P.span_start + p), e = __ldg(P.span_end + p);
        const uint64_t a = (uint64_t)(s & ~3u) + (uint64_t)(c - __ldg(P.chunk_prefix + p)) * kChunk;
        uint32_t* __restrict__ row = P.bitmap + (size_t)(p - P.path_lo) * P.words_per_row;
        uint32_t h[kItems];
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const uint64_t idx = a + (uint64_t)i * kThreads + tid;
            h[i] = (idx >= s && idx < e) ? ld_stream_u32(P.steps + idx, pol) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const uint64_t idx = a + (uint64_t)i * kThreads + tid;
            const uint32_t seg = h[i] >> 1;
            bool valid = idx >= s && idx < e;
            if (valid && seg >= P.n_segs) { *P.err = 1u; valid = false; }
            // invalid lanes get a unique key so they form singleton groups
            const uint32_t word = valid ? (seg >> 5) : (0x80000000u | lane);
            const uint32_t bit = valid ? bit_of(seg) : 0u;
            const uint32_t gm = __match_any_sync(0xFFFFFFFFu, word);
            const uint32_t acc = __reduce_or_sync(gm, bit);
            const int leader = __ffs(gm) - 1;
            uint32_t old = 0;
            if ((int)lane == leader && valid) old = atom_or_b32(row + word, acc);
            old = __shfl_sync(gm, old, leader);
            uint32_t rep = shr_wrap(old, seg) & 1u;
            if (__popc(acc) != __popc(gm)) {   // some segment appears twice inside this group
                const uint32_t dm = __match_any_sync(__activemask(), valid ? seg : (0x80000000u | lane));
                rep |= (dm & lanemask_lt) ? 1u : 0u;
            }
            if (valid && rep) red_add_u32(P.depth + seg, 1u);
        }
    }
}

// ---------------------------------------------------------------------------
// Decode of the packed depth counters (-DFGFA_DEPTH_PACK=16|8, see depth_kernels.cuh): adds every
// field to the u32 depth table, clears the packed word and accumulates the decoded total, which
// equals the number of steps added iff no field overflowed.
// ---------------------------------------------------------------------------
template <int PACK>
__global__ void __launch_bounds__(256) k_depth_unpack(uint32_t* __restrict__ packed, uint32_t* __restrict__ depth,
                                                      uint32_t n_segs, unsigned long long* __restrict__ total) {
    constexpr uint32_t kPer = 32 / PACK;                   // fields per word
    constexpr uint32_t kMask = PACK == 16 ? 0xFFFFu : 0xFFu;
    const uint32_t n_words = (n_segs + kPer - 1) / kPer;
    unsigned long long sum = 0;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
        const uint32_t v = packed[w];
        if (v == 0) continue;
        packed[w] = 0;
#pragma unroll
        for (uint32_t f = 0; f < kPer; ++f) {
            const uint32_t c = (v >> (PACK * f)) & kMask;
            const uint32_t seg = w * kPer + f;
            if (c && seg < n_segs) { depth[seg] += c; sum += c; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(total, sum);
}

}  // namespace fgfa
