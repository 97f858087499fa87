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
//   kernel C  (k_path_measure) the per-path weighted sums of path-depth mode
//             (`measure_path`, depth.rs:116-131).
//   kernel X  (k_uniq_exchange) kernel B fused with the multi-GPU reduce-scatter /
//             all-gather of depth and uniq over NVLink peer memory.
// (kernels T1/T2, the step-list tokenizer, live in tokenize_kernels.cuh.)
//
// Work decomposition: a *chunk* is up to kChunk consecutive steps of ONE path
// (a chunk never straddles two paths because the bitmap row depends on the path);
// chunk k of path p starts at the 16-byte-aligned element (start_p & ~3) + k*kChunk
// so that every 128-bit copy is aligned even though a path's span start is only
// 4-byte aligned.  The host builds the chunk table once per plan.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fgfa {

constexpr int kThreads = 256;          // threads per CTA in kernel A
#ifndef FGFA_ITEMS
#define FGFA_ITEMS 16
#endif
constexpr int kItems = FGFA_ITEMS;     // steps per thread per chunk
constexpr int kChunk = kThreads * kItems;  // 4096 steps = 16 KiB of the pool

// One unit of work of kernel A: up to kChunk consecutive steps of one path.
struct __align__(16) ChunkDesc {
    uint32_t a;      // first element of the chunk (16-byte aligned relative to `steps`)
    uint32_t s, e;   // the owning path's span [s, e) in the steps pool
    uint32_t path;   // path index (selects the seen-bitmap row)
};

struct StreamParams {
    const uint32_t* __restrict__ steps;       // device: the steps pool (Handle words), 16-byte aligned
    const ChunkDesc* __restrict__ chunks;     // device: chunk table of the whole plan
    uint32_t chunk_lo, chunk_hi;              // this launch covers chunks [chunk_lo, chunk_hi)
    uint32_t path_lo;                         // path whose bitmap row is row 0 of `bitmap`
    uint32_t n_segs;
    uint32_t words_per_row;                   // bitmap row pitch in 32-bit words
    uint32_t* __restrict__ depth;             // device [n_segs], pre-zeroed
    uint32_t* __restrict__ bitmap;            // device [rows][words_per_row], zero on entry
    uint32_t* __restrict__ err;               // sticky error flag (bit 0: segment id out of range)
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
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v));
}
// Experiment for the next round (tools/ubench.cu built with -DFGFA_DEPTH_PACK=16 or 8; the product
// builds with 32 and is unchanged): several depth counters per 32-bit word make the adds denser in
// segment space -- a 32-byte sector then covers 16 or 32 segments instead of 8, and sectors, not
// lanes, are what the L2 reduction path charges for.  A field that overflows carries into its
// neighbour, which LOWERS the sum of the decoded fields, so `sum(decoded) == steps added` proves
// that no field overflowed (tools/experimental_kernels.cuh: k_depth_unpack).
#ifndef FGFA_DEPTH_PACK
#define FGFA_DEPTH_PACK 32
#endif
__device__ __forceinline__ void depth_add(uint32_t* depth, uint32_t h) {   // h = Handle word, (h >> 1) < n_segs
#if FGFA_DEPTH_PACK == 32
    red_add_u32(depth + (h >> 1), 1u);
#elif FGFA_DEPTH_PACK == 16
    red_add_u32(depth + (h >> 2), 1u << ((h & 2u) << 3));
#elif FGFA_DEPTH_PACK == 8
    red_add_u32(depth + (h >> 3), 1u << ((h & 6u) << 2));
#else
#error "FGFA_DEPTH_PACK must be 32, 16 or 8"
#endif
}
#if FGFA_DEPTH_PACK != 32
// Same experiment, second form (-DFGFA_PACK_MERGE=1): measured without it, the packed adds were
// SLOWER than one counter per word (0.76 ms with 16-bit fields, 0.94 ms with 8-bit fields against
// 0.66 ms at config C) although they touch half / a quarter of the sectors -- the lanes of one
// instruction that hit the same word serialise in the L2 atomic unit.  Here neighbouring lanes that
// hit the same word are first merged with shuffles (groups of 32/FGFA_DEPTH_PACK lanes inside a run
// of equal words), so every word is added to by one lane per instruction.  All 32 lanes must call.
__device__ __forceinline__ void depth_add_merged(uint32_t* depth, uint32_t h, bool valid, uint32_t lane) {
    constexpr uint32_t kShift = FGFA_DEPTH_PACK == 16 ? 2u : 3u;      // handle -> word index
    constexpr uint32_t kGroup = 32u / FGFA_DEPTH_PACK;                 // 2 or 4 lanes merge
    const uint32_t w = valid ? (h >> kShift) : (0xFFFFFFFFu - lane);  // invalid lanes never merge
    uint32_t v = !valid ? 0u : FGFA_DEPTH_PACK == 16 ? 1u << ((h & 2u) << 3) : 1u << ((h & 6u) << 2);
    const uint32_t wp = __shfl_up_sync(0xFFFFFFFFu, w, 1);
    const bool head = lane == 0 || wp != w;
    const uint32_t heads = __ballot_sync(0xFFFFFFFFu, head);
    const uint32_t start = 31u - __clz(heads & (0xFFFFFFFFu >> (31u - lane)));   // head lane of my run
    const uint32_t off = lane - start;                                 // my position inside the run
    // level 1: even positions absorb the next lane of the same run
    uint32_t vn = __shfl_down_sync(0xFFFFFFFFu, v, 1);
    const bool next_same = lane < 31u && !((heads >> (lane + 1u)) & 1u);
    if ((off & 1u) == 0u && next_same) v += vn;
    if (kGroup == 4u) {                                                // level 2: positions 0 mod 4 absorb position +2
        vn = __shfl_down_sync(0xFFFFFFFFu, v, 2);
        const bool next2_same = lane < 30u && !((heads >> (lane + 1u)) & 3u);
        if ((off & 3u) == 0u && next2_same) v += vn;
    }
    if (valid && (off & (kGroup - 1u)) == 0u) red_add_u32(depth + w, v);
}
#endif
__device__ __forceinline__ void red_or_b32(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v));
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

// ---------------------------------------------------------------------------
// kernel A (k_step_stream_merged): one pass over the steps pool.
//
// Per chunk (kChunk consecutive steps of one path):
//   stage   the chunk is copied global -> shared with 16-byte cp.async (L2 evict-first),
//           double buffered: the copy of the block's NEXT chunk is in flight while the
//           current one is processed, so HBM latency is off the critical path;
//   pass 2  thread order: each thread walks its kItems consecutive steps, merges the
//           seen-bitmap updates of steps that fall into the same 32-segment word in a
//           register and issues ONE RED.OR per run (depth.rs:30-34 without the serial test);
//   pass 3  lane order: lane l of each warp instruction handles consecutive step l, so
//           the depth RED.ADDs (depth.rs:29) of a near-monotone walk coalesce into few
//           L2 sector requests.
// uniq comes from kernel B; depth is complete after this kernel.
// Interior ("full") chunks take straight-line code; chunks that touch a path boundary
// or the pool end are staged element-wise with an invalid-handle filler.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, uint64_t pol) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr uint32_t kFiller = 0xFFFFFFFFu;   // never a valid handle: segment ids are < 2^31

// How kernel A records the seen-bits:
//   kSeenNone    seg_depth only (depth.rs:45-56), no bitmap;
//   kSeenDirect  one RED.OR per thread-level run, straight to L2;
//   kSeenDeferred like kSeenDirect, but a thread parks its runs in a private shared-memory
//                stack and the warp issues everybody's k-th run in the SAME instruction:
//                neighbouring threads' k-th runs fall into the same 32-byte bitmap sector
//                (a sector covers 256 segments, a thread ~19), so the L2 sees one request
//                per sector instead of one per run.  No barrier, no atomics on shared memory.
//   kSeenWindow  for looping paths (tandem repeats, config E).  Thread-level runs are first
//                OR-ed into a block-wide shared-memory window (direct-mapped by bitmap-word
//                index, one tag per 32-byte bitmap sector), then every touched sector is
//                flushed with ONE coalesced RED.OR request.  In addition the depth counters
//                of a chunk whose segments span < kCacheSegs are accumulated in a
//                shared-memory cache that PERSISTS across the block's chunks as long as they
//                stay inside the same segment window (a loop does), and is flushed with
//                coalesced RED.ADDs when the window moves or the kernel ends: the hot
//                segments' adds never serialise in L2.
enum SeenMode : int { kSeenNone = 0, kSeenDirect = 1, kSeenWindow = 2, kSeenDeferred = 3 };

constexpr uint32_t kCacheSegs = 8192;             // kSeenWindow: depth counters cached in shared memory
constexpr uint32_t kWinShift = 11;
constexpr uint32_t kWinWords = 1u << kWinShift;   // window: 2048 bitmap words = 65536 segments
constexpr uint32_t kWinGroups = kWinWords / 8;    // one tag per bitmap sector (8 words)
constexpr uint32_t kTagEmpty = 0xFFFFFFFFu;

constexpr int kRunSlots = 4;                      // kSeenDeferred: parked runs per thread per chunk (3, 4, 6 measure alike)

__host__ __device__ constexpr size_t stream_smem_bytes(int seen_mode) {
    return 2 * (size_t)kChunk * 4 +
           (seen_mode == kSeenWindow ? (size_t)kWinWords * 4 + kWinGroups * 4 + kWinGroups * 4 + 16 + (size_t)kCacheSegs * 4 : 0) +
           (seen_mode == kSeenDeferred ? (size_t)kRunSlots * kThreads * 8 : 0);
}

__device__ __forceinline__ void red_shared_or(uint32_t* p, uint32_t v) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// P2 = consecutive steps one thread walks per pass-2 round (kItems/P2 rounds per chunk).
// Smaller P2 packs the lanes of one RED.OR instruction closer together in segment space
// (fewer bitmap sectors per instruction) at the price of shorter runs.
template <int BLOCKS_PER_SM, int SEEN_MODE, int P2 = kItems>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM) k_step_stream_merged(StreamParams P) {
    constexpr bool WITH_SEEN = SEEN_MODE != kSeenNone;
    static_assert(P2 == 32 || P2 == 16 || P2 == 8 || P2 == 4, "P2 must divide kItems and be a multiple of 4");
    extern __shared__ uint4 smem_dyn[];
    uint4 (*s_steps)[kChunk / 4] = reinterpret_cast<uint4 (*)[kChunk / 4]>(smem_dyn);
    uint32_t* const s_bits = reinterpret_cast<uint32_t*>(smem_dyn + 2 * (kChunk / 4));   // [kWinWords]
    uint32_t* const s_tag = s_bits + kWinWords;                                          // [kWinGroups]
    uint32_t* const s_list = s_tag + kWinGroups;                                         // [kWinGroups]
    uint32_t* const s_count = s_list + kWinGroups;                                       // [0] list length, [1] min, [2] max
    uint32_t* const s_cnt = s_count + 4;                                                 // [kCacheSegs]
    uint32_t cache_base = kFiller;               // first segment of the cached depth window (block-uniform)
    uint2* const s_runs = reinterpret_cast<uint2*>(smem_dyn + 2 * (kChunk / 4)) + threadIdx.x;   // [kRunSlots][kThreads]
    if (SEEN_MODE == kSeenWindow) {
        for (uint32_t i = threadIdx.x; i < kWinWords; i += kThreads) s_bits[i] = 0u;
        for (uint32_t i = threadIdx.x; i < kWinGroups; i += kThreads) s_tag[i] = kTagEmpty;
        for (uint32_t i = threadIdx.x; i < kCacheSegs; i += kThreads) s_cnt[i] = 0u;
        if (threadIdx.x == 0) { s_count[0] = 0u; s_count[1] = 0xFFFFFFFFu; s_count[2] = 0u; }
    }
    const uint64_t pol = make_evict_first_policy();
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg_limit = P.n_segs * 2u;   // h < seg_limit  <=>  (h >> 1) < n_segs  (n_segs < 2^31)
    const uint32_t st_off = swz(tid);                               // staging: vector tid + j*kThreads
    constexpr uint32_t kVecs = P2 / 4;                              // pass 2: vectors kVecs*tid + j (+ round)
    const uint32_t swc = ((tid * kVecs) >> 3) & 7u;
    const uint32_t ld_base = (tid * kVecs) & ~7u, ld_lo = (tid * kVecs) & 7u;
    // pass 3: element i*kThreads + tid lives in vector i*64 + warp*8 + lane/4
    const uint32_t p3_off = (warp * 8u + ((lane >> 2) ^ (warp & 7u))) * 4u + (lane & 3u);
    uint32_t* const depth_ptr = keep_ptr(P.depth);

    auto stage = [&](uint4* buf, const ChunkDesc& d) {
        const bool full = d.a >= d.s && (uint64_t)d.a + kChunk <= d.e;
        if (full) {
            const uint32_t* src = P.steps + d.a + tid * 4u;
#pragma unroll
            for (int j = 0; j < kItems / 4; ++j) cp_async_16(buf + st_off + j * kThreads, src + j * kThreads * 4, pol);
        } else {
#pragma unroll
            for (int j = 0; j < kItems / 4; ++j) {
                const uint64_t idx = (uint64_t)d.a + (uint64_t)(tid + j * kThreads) * 4;
                uint32_t x[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] = (idx + k >= d.s && idx + k < d.e) ? P.steps[idx + k] : kFiller;
                buf[st_off + j * kThreads] = make_uint4(x[0], x[1], x[2], x[3]);
            }
        }
    };

    uint32_t c = P.chunk_lo + blockIdx.x;
    if (c >= P.chunk_hi) return;
    ChunkDesc cur = P.chunks[c];
    stage(s_steps[0], cur);
    cp_async_commit();
    // descriptors are fetched two chunks ahead so their latency never gates a copy
    ChunkDesc nxt = cur;
    if (c + gridDim.x < P.chunk_hi) nxt = P.chunks[c + gridDim.x];
    uint32_t b = 0;
    for (; c < P.chunk_hi; c += gridDim.x) {
        const uint32_t cn = c + gridDim.x;
        ChunkDesc nxt2 = nxt;
        if (cn < P.chunk_hi) {
            if (cn + gridDim.x < P.chunk_hi) nxt2 = P.chunks[cn + gridDim.x];
            stage(s_steps[b ^ 1], nxt);          // prefetch the next chunk while this one is processed
        }
        cp_async_commit();
        cp_async_wait<1>();                      // this chunk's copies have landed
        __syncthreads();

        const uint4* buf = s_steps[b];
        bool cached = false;                     // block-uniform: this chunk's depth adds go to the cache
        if (SEEN_MODE == kSeenWindow) {
            // chunk-wide min / max handle (fillers and out-of-range ids make the chunk uncacheable)
            const uint32_t* w = reinterpret_cast<const uint32_t*>(buf) + p3_off;
            uint32_t lo = 0xFFFFFFFFu, hi = 0u;
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const uint32_t v = w[i * kThreads];
                lo = min(lo, v);
                hi = max(hi, v);
            }
            lo = __reduce_min_sync(0xFFFFFFFFu, lo);
            hi = __reduce_max_sync(0xFFFFFFFFu, hi);
            if (lane == 0) { atomicMin(&s_count[1], lo); atomicMax(&s_count[2], hi); }
            __syncthreads();
            const uint32_t seg_lo = s_count[1] >> 1, seg_hi = s_count[2] >> 1;
            if (s_count[2] < seg_limit && seg_hi - (seg_lo & ~7u) < kCacheSegs) {
                cached = true;
                if (cache_base == kFiller || seg_lo < cache_base || seg_hi >= cache_base + kCacheSegs) {
                    if (cache_base != kFiller) {     // the window moves: flush the old one
                        for (uint32_t i = tid; i < kCacheSegs; i += kThreads) {
                            const uint32_t v = s_cnt[i];
                            if (v) { red_add_u32(depth_ptr + cache_base + i, v); s_cnt[i] = 0u; }
                        }
                        __syncthreads();
                    }
                    cache_base = seg_lo & ~7u;
                }
            }
        }
        // ---- pass 2: thread order -- one RED.OR per run of steps in the same bitmap word ----
        if (WITH_SEEN) {
            uint32_t* __restrict__ row = P.bitmap + (size_t)(cur.path - P.path_lo) * P.words_per_row;
#pragma unroll
            for (int round = 0; round < kItems / P2; ++round) {
            uint32_t h[P2];
#pragma unroll
            for (int j = 0; j < P2 / 4; ++j) {
                const uint4 x = buf[round * (kThreads * kVecs) + ld_base + ((ld_lo + j) ^ swc)];
                h[4 * j + 0] = x.x; h[4 * j + 1] = x.y; h[4 * j + 2] = x.z; h[4 * j + 3] = x.w;
            }
            uint32_t hmax = 0;
#pragma unroll
            for (int i = 0; i < P2; ++i) hmax = max(hmax, h[i]);
            uint32_t cnt_out = 0;                 // runs parked by this thread (kSeenDeferred)
            if (hmax < seg_limit) {
                uint32_t acc = 0, cnt = 0;
#pragma unroll
                for (int i = 0; i < P2; ++i) {
                    const uint32_t bit = bit_of(h[i] >> 1);
                    acc = ((i > 0 && ((h[i] ^ h[i - 1]) < 64u)) ? acc : 0u) | bit;
                    if (i == P2 - 1 || ((h[i] ^ h[i + (i < P2 - 1)]) >= 64u)) {
                        const uint32_t w = h[i] >> 6;
                        if (SEEN_MODE == kSeenDeferred) {
                            if (cnt < kRunSlots) s_runs[cnt * kThreads] = make_uint2(w, acc);
                            else red_or_b32(row + w, acc);
                            ++cnt;
                        } else if (SEEN_MODE == kSeenWindow) {
                            const uint32_t g = (w >> 3) & (kWinGroups - 1), tag = w >> kWinShift;
                            uint32_t t = s_tag[g];
                            if (t == kTagEmpty) {
                                t = atomicCAS(&s_tag[g], kTagEmpty, tag);
                                if (t == kTagEmpty) {            // this thread claimed the sector
                                    t = tag;
                                    s_list[atomicAdd(s_count, 1u)] = g;
                                }
                            }
                            if (t == tag) red_shared_or(&s_bits[w & (kWinWords - 1)], acc);
                            else red_or_b32(row + w, acc);       // window slot taken by another sector
                        } else {
                            red_or_b32(row + w, acc);
                        }
                    }
                }
                cnt_out = cnt;
            } else {
                // filler (edge chunk) or out-of-range segment id somewhere: per step, unmerged
#pragma unroll
                for (int i = 0; i < P2; ++i) {
                    if (h[i] < seg_limit) red_or_b32(row + (h[i] >> 6), bit_of(h[i] >> 1));
                    else if (h[i] != kFiller) *P.err = 1u;
                }
            }
            // issue the parked runs by ordinal: everybody's k-th run in the same instruction
            if (SEEN_MODE == kSeenDeferred) {
                const uint32_t cnt = cnt_out;
#pragma unroll
                for (int k = 0; k < kRunSlots; ++k) {
                    if ((uint32_t)k < cnt) {
                        const uint2 r = s_runs[k * kThreads];
                        red_or_b32(row + r.x, r.y);
                    }
                }
            }
            }   // round
        }
        // ---- pass 3: lane order -- one depth RED.ADD per step ----
        {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(buf) + p3_off;
            uint32_t hh[kItems];
#pragma unroll
            for (int i = 0; i < kItems; ++i) hh[i] = w[i * kThreads];
            if (cached) {                        // every handle is valid and inside the cached window
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    const uint32_t a = (uint32_t)__cvta_generic_to_shared(s_cnt + ((hh[i] >> 1) - cache_base));
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
                }
            } else {
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
#if FGFA_DEPTH_PACK != 32 && defined(FGFA_PACK_MERGE)
                    depth_add_merged(depth_ptr, hh[i], hh[i] < seg_limit, lane);
                    if (!WITH_SEEN && hh[i] >= seg_limit && hh[i] != kFiller) *P.err = 1u;
#else
                    if (hh[i] < seg_limit) depth_add(depth_ptr, hh[i]);
                    else if (!WITH_SEEN && hh[i] != kFiller) *P.err = 1u;
#endif
                }
            }
        }
        if (SEEN_MODE == kSeenWindow) {
            // ---- flush: one coalesced RED.OR request per touched bitmap sector ----
            uint32_t* __restrict__ row = P.bitmap + (size_t)(cur.path - P.path_lo) * P.words_per_row;
            __syncthreads();
            const uint32_t n = *s_count * 8u;            // 8 lanes per touched sector, block-uniform
            for (uint32_t base = warp * 32u; base < n; base += kThreads) {   // warp-uniform trip count
                const uint32_t i = base + lane;
                const bool on = i < n;
                const uint32_t g = on ? s_list[i >> 3] : 0u, slot = (g << 3) | (i & 7u);
                const uint32_t v = on ? s_bits[slot] : 0u, tag = s_tag[g];
                if (v) {
                    red_or_b32(row + ((tag << kWinShift) | slot), v);
                    s_bits[slot] = 0u;
                }
                __syncwarp();                            // every lane has read the tag
                if (on && (i & 7u) == 0u) s_tag[g] = kTagEmpty;
            }
            __syncthreads();
            if (tid == 0) { s_count[0] = 0u; s_count[1] = 0xFFFFFFFFu; s_count[2] = 0u; }
        } else {
            __syncthreads();                     // buffer b is free for the prefetch after next
        }
        cur = nxt;
        nxt = nxt2;
        b ^= 1;
    }
    if (SEEN_MODE == kSeenWindow && cache_base != kFiller) {   // flush what is still cached
        for (uint32_t i = tid; i < kCacheSegs; i += kThreads) {
            const uint32_t v = s_cnt[i];
            if (v) red_add_u32(depth_ptr + cache_base + i, v);
        }
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
    void* __restrict__ uniq;         // [n_segs] u32 (or u8, see uniq_bytes) or nullptr
    uint32_t* __restrict__ depth;    // [n_segs] or nullptr
    int accumulate;                  // 0: uniq = cnt, 1: uniq += cnt
    int uniq_bytes;                  // 4: u32 counters; 1: u8 counters (caller guarantees <= 255 paths)
};

constexpr int kPopThreads = 128;
constexpr int kPopRowsInFlight = 8;

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kPopThreads, MIN_BLOCKS) k_uniq_popcount(PopcountParams P) {
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
            if (P.uniq) {
                if (P.uniq_bytes == 1) {
                    uint8_t* u = static_cast<uint8_t*>(P.uniq);
                    u[seg] = (uint8_t)(P.accumulate ? u[seg] + v : v);
                } else {
                    uint32_t* u = static_cast<uint32_t*>(P.uniq);
                    if (P.accumulate) u[seg] += v; else u[seg] = v;
                }
            }
            if (P.depth && v) P.depth[seg] += v;
        }
    }
}

// ---------------------------------------------------------------------------
// kernel C (k_path_measure): the per-path weighted sums of the reference's path-depth mode,
//   sums[2p]   = sum over the steps of path p of depth[seg] * len(seg)     (depth.rs:122-123)
//   sums[2p+1] = sum over the steps of path p of len(seg)                  (depth.rs:124)
// (`measure_path`, flatgfa/src/ops/depth.rs:116-131; the one f64 divide is done on the host).
// A gather + segmented reduction: steps are streamed in lane order (coalesced), the
// interleaved {depth, len} table (8 bytes per segment, L2-resident) is gathered once per
// step, every CTA keeps two u64 partial sums per thread and adds them to its path's
// accumulators when its contiguous share of the chunk table crosses into the next path.
// Arithmetic is wrapping u64, like `usize` in a release build of the reference.
// ---------------------------------------------------------------------------
struct MeasureParams {
    const uint32_t* __restrict__ steps;
    const ChunkDesc* __restrict__ chunks;
    uint32_t chunk_lo, chunk_hi;
    uint32_t n_segs;
    const uint2* __restrict__ depth_len;          // [n_segs] {depth, len}
    unsigned long long* __restrict__ sums;        // [2 * n_paths], zero on entry
    uint32_t* __restrict__ err;
};

__global__ void __launch_bounds__(256) k_interleave_depth_len(const uint32_t* __restrict__ depth,
                                                              const uint32_t* __restrict__ len,
                                                              uint2* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_uint2(depth[i], len[i]);
}

// 8 resident CTAs per SM (32 registers): the kernel is latency-bound and lives on occupancy —
// caps of 6 / 5 / 4 CTAs measured 0.50 / 0.59 / 0.72 ms against 0.44 ms; ld.global.cg gathers measured
// the same as ld.global.nc, L1::no_allocate 0.54 ms.
__global__ void __launch_bounds__(kThreads, 8) k_path_measure(MeasureParams P) {
    __shared__ unsigned long long s_part[2][kThreads / 32];
    const uint64_t pol = make_evict_first_policy();
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Every CTA owns a CONTIGUOUS share of the chunk table, so its consecutive chunks belong
    // to the same path and the partial sums stay in registers until the path changes: two
    // u64 atomics per (CTA, path) instead of two per chunk, and no barrier inside the loop.
    const uint32_t total = P.chunk_hi - P.chunk_lo;
    const uint32_t per = (total + gridDim.x - 1) / gridDim.x;
    const uint32_t c0 = P.chunk_lo + min(total, blockIdx.x * per);
    const uint32_t c1 = P.chunk_lo + min(total, (blockIdx.x + 1) * per);
    unsigned long long wsum = 0, lsum = 0;
    uint32_t cur_path = kFiller;
    auto flush = [&](uint32_t path) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, o);
            lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
        }
        if (lane == 0) { s_part[0][warp] = wsum; s_part[1][warp] = lsum; }
        __syncthreads();
        if (tid < 2) {
            unsigned long long t = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) t += s_part[tid][w];
            atomicAdd(P.sums + 2 * (size_t)path + tid, t);
        }
        __syncthreads();
        wsum = 0;
        lsum = 0;
    };
    for (uint32_t c = c0; c < c1; ++c) {
        const ChunkDesc d = P.chunks[c];
        if (d.path != cur_path) {                              // block-uniform
            if (cur_path != kFiller) flush(cur_path);
            cur_path = d.path;
        }
        // valid offsets inside the chunk: [lo, lo + span) relative to d.a (32-bit index math)
        const uint32_t* __restrict__ base = P.steps + d.a;
        const uint32_t lo = d.s > d.a ? d.s - d.a : 0u;
        const uint32_t hi = min(d.e - d.a, (uint32_t)kChunk);
        const uint32_t span = hi > lo ? hi - lo : 0u;
        uint32_t h[kItems];
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const uint32_t off = (uint32_t)i * kThreads + tid;
            h[i] = (off - lo < span) ? ld_stream_u32(base + off, pol) : kFiller;
        }
        // (issuing the next chunk's loads before these gathers was measured: 64 registers, half the
        // resident warps, 0.45 -> 0.76 ms)
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const uint32_t seg = h[i] >> 1;
            if (seg < P.n_segs) {
                const uint2 dl = __ldg(P.depth_len + seg);
                wsum += (unsigned long long)dl.x * dl.y;
                lsum += dl.y;
            } else if (h[i] != kFiller) {
                *P.err = 1u;
            }
        }
    }
    if (cur_path != kFiller) flush(cur_path);
}

// ---------------------------------------------------------------------------
// kernel X (k_uniq_exchange): kernel B fused with the multi-GPU exchange, over NVLink peer
// memory.  Every rank owns one slice of the segment axis.  For its slice it
//   * sums the ranks' partial depths                       (the reduce-scatter of depth),
//   * popcounts the seen-bitmap rows of ALL ranks          (uniq never needs reducing),
//   * stores the final depth (u32) and uniq (u8) slice into EVERY rank's result buffer
//     (the all-gather), with plain peer loads/stores on symmetric-memory pointers.
// A rank moves (N-1)/N of [4 B depth + rows/8 B bitmap] per slice segment in and
// (N-1) x 5 B per slice segment out, instead of NCCL's allreduce of the whole 5 B/segment
// buffer; inter-rank ordering is two stream barriers around the launch (host side).
// ---------------------------------------------------------------------------
constexpr int kMaxRanks = 16;
constexpr int kMaxRows = 255;                   // u8 uniq counters: at most 255 paths in the graph

struct ExchangeParams {
    const uint32_t* row_ptr[kMaxRows];          // every rank's seen-bitmap rows, flattened (peer pointers)
    const uint32_t* partial_depth[kMaxRanks];   // rank q's partial depth [n_segs]
    uint32_t* final_depth[kMaxRanks];           // rank q's result depth [n_segs]
    uint8_t* final_uniq[kMaxRanks];             // rank q's result uniq  [n_segs]
    int n_ranks;
    uint32_t n_rows;                            // total rows over all ranks
    uint32_t n_segs;
    uint32_t w_lo, w_hi;                        // this rank's slice of bitmap words (32 segments each)
    uint32_t uniq_blocks;                       // blocks [0, uniq_blocks) popcount, the rest sum depths
    // NVLS: multicast mapping of the symmetric buffer (nullptr = plain peer loads/stores) and
    // the byte offsets of the three regions inside it (identical on every rank)
    uint8_t* mc_base;
    uint64_t off_partial, off_final_depth, off_final_uniq;
    int mc_reduce;                              // 1: depth sum by multimem.ld_reduce, 0: by peer loads
};

// multimem.*: one instruction on the multicast address reaches every rank's copy through
// the NVSwitch -- ld_reduce returns the in-switch SUM of all ranks' words, st stores to all.
__device__ __forceinline__ uint64_t mc_ld_reduce_add_u64(const void* mc) {
    uint64_t r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(r) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void mc_st_v4(void* mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

constexpr int kXThreads = 128;
constexpr int kXRowsInFlight = 32;

__device__ __forceinline__ uint4 ld_peer_v4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_peer_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(kXThreads) k_uniq_exchange(ExchangeParams P) {
    if (blockIdx.x < P.uniq_blocks) {
        // ---- role 1: uniq = bit-sliced count over every rank's rows of one column word ----
        const uint32_t w = P.w_lo + blockIdx.x * kXThreads + threadIdx.x;
        if (w >= P.w_hi) return;
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
        for (uint32_t r0 = 0; r0 < P.n_rows; r0 += kXRowsInFlight) {
            uint32_t x[kXRowsInFlight];
#pragma unroll
            for (int k = 0; k < kXRowsInFlight; ++k)
                x[k] = (r0 + k < P.n_rows) ? ld_peer_u32(P.row_ptr[r0 + k] + w) : 0u;
#pragma unroll
            for (int k = 0; k < kXRowsInFlight; ++k) {
                uint32_t v = x[k], t;
                t = c0 & v; c0 ^= v; v = t;
                t = c1 & v; c1 ^= v; v = t;
                t = c2 & v; c2 ^= v; v = t;
                t = c3 & v; c3 ^= v; v = t;
                t = c4 & v; c4 ^= v; v = t;
                t = c5 & v; c5 ^= v; v = t;
                t = c6 & v; c6 ^= v; v = t;
                c7 ^= v;                        // <= 255 rows in total: eight planes never overflow
            }
        }
        uint32_t packed[8];                     // 32 u8 counts, segment 32w + j in byte j
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t cnt = ((c0 >> j) & 1u) | (((c1 >> j) & 1u) << 1) | (((c2 >> j) & 1u) << 2) |
                                 (((c3 >> j) & 1u) << 3) | (((c4 >> j) & 1u) << 4) | (((c5 >> j) & 1u) << 5) |
                                 (((c6 >> j) & 1u) << 6) | (((c7 >> j) & 1u) << 7);
            if ((j & 3) == 0) packed[j >> 2] = cnt; else packed[j >> 2] |= cnt << (8 * (j & 3));
        }
        const uint32_t seg0 = w << 5;
        if (seg0 + 32u <= P.n_segs && P.mc_base) {     // all-gather through the switch
            uint8_t* du = P.mc_base + P.off_final_uniq + seg0;
            mc_st_v4(du, make_uint4(packed[0], packed[1], packed[2], packed[3]));
            mc_st_v4(du + 16, make_uint4(packed[4], packed[5], packed[6], packed[7]));
        } else if (seg0 + 32u <= P.n_segs) {
            for (int q = 0; q < P.n_ranks; ++q) {      // all-gather by peer stores
                uint4* du = reinterpret_cast<uint4*>(P.final_uniq[q] + seg0);
                du[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                du[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            }
        } else {
            for (uint32_t j = 0; seg0 + j < P.n_segs; ++j)
                for (int q = 0; q < P.n_ranks; ++q)
                    P.final_uniq[q][seg0 + j] = (uint8_t)(packed[j >> 2] >> (8 * (j & 3)));
        }
    } else {
        // ---- role 2: depth = sum of the ranks' partials, four segments per thread ----
        const uint64_t s_lo = (uint64_t)P.w_lo << 5, s_hi = min((uint64_t)P.w_hi << 5, (uint64_t)P.n_segs);
        const uint64_t seg = s_lo + ((uint64_t)(blockIdx.x - P.uniq_blocks) * kXThreads + threadIdx.x) * 4;
        if (seg >= s_hi) return;
        if (seg + 4 <= s_hi && P.mc_base && P.mc_reduce) {
            // two packed-u32 sums per 64-bit in-switch reduction: a true depth never carries
            // out of its 32 bits (depth <= n_steps < 2^32)
            const uint8_t* src = P.mc_base + P.off_partial + seg * 4;
            const uint64_t lo = mc_ld_reduce_add_u64(src), hi = mc_ld_reduce_add_u64(src + 8);
            mc_st_v4(P.mc_base + P.off_final_depth + seg * 4,
                     make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)));
        } else if (seg + 4 <= s_hi) {
            uint4 t[kMaxRanks];
#pragma unroll
            for (int q = 0; q < kMaxRanks; ++q)
                if (q < P.n_ranks) t[q] = ld_peer_v4(reinterpret_cast<const uint4*>(P.partial_depth[q] + seg));
            uint4 d = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int q = 0; q < kMaxRanks; ++q)
                if (q < P.n_ranks) { d.x += t[q].x; d.y += t[q].y; d.z += t[q].z; d.w += t[q].w; }
            if (P.mc_base) {
                mc_st_v4(P.mc_base + P.off_final_depth + seg * 4, d);
            } else {
#pragma unroll
                for (int q = 0; q < kMaxRanks; ++q)
                    if (q < P.n_ranks) *reinterpret_cast<uint4*>(P.final_depth[q] + seg) = d;
            }
        } else {
            for (uint64_t s2 = seg; s2 < s_hi; ++s2) {
                uint32_t sum = 0;
                for (int q = 0; q < P.n_ranks; ++q) sum += ld_peer_u32(P.partial_depth[q] + s2);
                for (int q = 0; q < P.n_ranks; ++q) P.final_depth[q][s2] = sum;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// kernels P + R: the multi-GPU exchange as PUSH + local reduce (the shipped fused form for N > 2).
//
// Kernel X above PULLS: a rank reads its slice of every peer's partial depth and bitmap rows over NVLink.  Reads
// are round trips; X sustains ~450-510 GB/s of ingress per GPU (profiles/r2_x_nvlink_n8.csv) and the result
// slices it multicasts afterwards arrive on the same ingress.  Here the first half of the traffic travels as
// posted stores instead, and the bitmaps never leave their GPU:
//   kernel P (k_push_partials)  every rank popcounts ITS OWN rows over the WHOLE segment axis (the bit-sliced
//        counters of kernel B; <= 255 paths, so u8 per segment) and clears them, and copies its partial depth;
//        each 32-segment word goes straight into slot `rank` of its OWNER's receive buffer -- a peer store for
//        (N-1)/N of the words.  Out: (N-1)/N x 5 bytes per segment, nothing in.
//   [inter-rank barrier]
//   kernel R (k_reduce_slices)  the owner adds the N slots of its slice (local loads only) and stores the final
//        depth (u32) and uniq (u8) slice to every rank: one multimem.st through the NVSwitch when the buffers
//        have a multicast mapping, N peer stores otherwise.
//   [inter-rank barrier]
// The segment axis is cut into N slices of `per` bitmap words (per = a multiple of 32, i.e. whole 128-byte
// lines); receive slot s of a rank = { u32 depth[per * 32], u8 uniq[per * 32] } written by rank s.
// ---------------------------------------------------------------------------
struct PushParams {
    uint32_t* __restrict__ bitmap;              // this rank's seen rows [n_rows][words_per_row]; cleared
    const uint32_t* __restrict__ partial_depth; // this rank's partial depth [n_segs]
    const uint8_t* __restrict__ partial_uniq;   // or nullptr: this rank's u8 uniq counts, already popcounted (window
                                                // engine + kernel B2), readable up to a multiple of 32 bytes
    uint8_t* recv[kMaxRanks];                   // every rank's receive buffer (peer pointers)
    int n_ranks, rank;
    uint32_t n_rows, words_per_row;
    uint32_t n_words, n_segs;
    uint32_t per;                               // bitmap words per slice
    uint32_t uniq_blocks;                       // blocks [0, uniq_blocks) popcount, the rest copy depth
};
struct ReduceParams {
    const uint8_t* __restrict__ recv;           // this rank's receive buffer (local)
    uint32_t* final_depth[kMaxRanks];
    uint8_t* final_uniq[kMaxRanks];
    int n_ranks, rank;
    uint32_t n_words, n_segs, per;
    uint8_t* mc_base;                           // multicast mapping holding final depth / uniq at the offsets below, or nullptr
    uint64_t off_final_depth, off_final_uniq;
};
__host__ __device__ inline uint64_t push_slot_bytes(uint32_t per) { return (uint64_t)per * 160u; }   // 32 x (4 + 1) bytes per word

__device__ __forceinline__ void st_peer_v4(void* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(kXThreads) k_push_partials(PushParams P) {
    if (blockIdx.x < P.uniq_blocks) {
        // ---- role 1: u8 counts of this rank's rows for one bitmap word, to the word's owner ----
        const uint32_t w = blockIdx.x * kXThreads + threadIdx.x;
        if (w >= P.n_words) return;
        const uint32_t owner = w / P.per, lw = w - owner * P.per;
        uint8_t* dst = P.recv[owner] + push_slot_bytes(P.per) * P.rank + (size_t)P.per * 128u + (size_t)lw * 32u;
        if (P.partial_uniq) {                   // the counts exist already: forward the word's 32 bytes
            const uint4* src = reinterpret_cast<const uint4*>(P.partial_uniq + (size_t)w * 32u);
            st_peer_v4(dst, src[0]);
            st_peer_v4(dst + 16, src[1]);
            return;
        }
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
        uint32_t* col = P.bitmap + w;
        for (uint32_t r0 = 0; r0 < P.n_rows; r0 += kXRowsInFlight) {
            uint32_t x[kXRowsInFlight];
#pragma unroll
            for (int k = 0; k < kXRowsInFlight; ++k)
                x[k] = (r0 + k < P.n_rows) ? col[(size_t)(r0 + k) * P.words_per_row] : 0u;
#pragma unroll
            for (int k = 0; k < kXRowsInFlight; ++k) {
                uint32_t v = x[k], t;
                if (v) col[(size_t)(r0 + k) * P.words_per_row] = 0u;      // zero bitmap for the next run
                t = c0 & v; c0 ^= v; v = t;
                t = c1 & v; c1 ^= v; v = t;
                t = c2 & v; c2 ^= v; v = t;
                t = c3 & v; c3 ^= v; v = t;
                t = c4 & v; c4 ^= v; v = t;
                t = c5 & v; c5 ^= v; v = t;
                t = c6 & v; c6 ^= v; v = t;
                c7 ^= v;
            }
        }
        uint32_t packed[8];                     // 32 u8 counts, segment 32w + j in byte j
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t cnt = ((c0 >> j) & 1u) | (((c1 >> j) & 1u) << 1) | (((c2 >> j) & 1u) << 2) |
                                 (((c3 >> j) & 1u) << 3) | (((c4 >> j) & 1u) << 4) | (((c5 >> j) & 1u) << 5) |
                                 (((c6 >> j) & 1u) << 6) | (((c7 >> j) & 1u) << 7);
            if ((j & 3) == 0) packed[j >> 2] = cnt; else packed[j >> 2] |= cnt << (8 * (j & 3));
        }
        st_peer_v4(dst, make_uint4(packed[0], packed[1], packed[2], packed[3]));
        st_peer_v4(dst + 16, make_uint4(packed[4], packed[5], packed[6], packed[7]));
    } else {
        // ---- role 2: partial depth, four segments per thread, to the owner of their word ----
        const uint64_t seg = ((uint64_t)(blockIdx.x - P.uniq_blocks) * kXThreads + threadIdx.x) * 4;
        if (seg >= ((uint64_t)P.n_words << 5)) return;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (seg + 4 <= P.n_segs) {
            v = *reinterpret_cast<const uint4*>(P.partial_depth + seg);
        } else if (seg < P.n_segs) {
            v.x = P.partial_depth[seg];
            if (seg + 1 < P.n_segs) v.y = P.partial_depth[seg + 1];
            if (seg + 2 < P.n_segs) v.z = P.partial_depth[seg + 2];
        }
        const uint32_t w = (uint32_t)(seg >> 5), owner = w / P.per;
        const uint64_t lseg = seg - (uint64_t)owner * P.per * 32u;
        st_peer_v4(P.recv[owner] + push_slot_bytes(P.per) * P.rank + lseg * 4u, v);
    }
}

__global__ void __launch_bounds__(kXThreads) k_reduce_slices(ReduceParams P) {
    const uint32_t w_lo = min(P.per * (uint32_t)P.rank, P.n_words), w_hi = min(w_lo + P.per, P.n_words);
    const uint64_t slice_segs = (uint64_t)(w_hi - w_lo) << 5;          // including the padding of the last word
    const uint64_t s_base = (uint64_t)w_lo << 5;
    const uint64_t depth_threads = slice_segs / 4;                     // 4 segments (16 bytes of u32) per thread
    const uint64_t t = (uint64_t)blockIdx.x * kXThreads + threadIdx.x;
    if (t < depth_threads) {
        const uint64_t lseg = t * 4, seg = s_base + lseg;
        uint4 d = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < kMaxRanks; ++q)
            if (q < P.n_ranks) {
                const uint4 x = *reinterpret_cast<const uint4*>(P.recv + push_slot_bytes(P.per) * q + lseg * 4u);
                d.x += x.x; d.y += x.y; d.z += x.z; d.w += x.w;
            }
        if (seg + 4 <= P.n_segs) {
            if (P.mc_base) {
                mc_st_v4(P.mc_base + P.off_final_depth + seg * 4, d);
            } else {
#pragma unroll
                for (int q = 0; q < kMaxRanks; ++q)
                    if (q < P.n_ranks) st_peer_v4(P.final_depth[q] + seg, d);
            }
        } else {
            const uint32_t vals[4] = {d.x, d.y, d.z, d.w};
            for (uint32_t k = 0; k < 4 && seg + k < P.n_segs; ++k)
                for (int q = 0; q < P.n_ranks; ++q) P.final_depth[q][seg + k] = vals[k];
        }
    } else {
        // ---- uniq: 16 segments (16 bytes of u8) per thread; packed bytes never carry (<= 255 paths) ----
        const uint64_t lseg = (t - depth_threads) * 16, seg = s_base + lseg;
        if (lseg >= slice_segs) return;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < kMaxRanks; ++q)
            if (q < P.n_ranks) {
                const uint4 x = *reinterpret_cast<const uint4*>(P.recv + push_slot_bytes(P.per) * q + (size_t)P.per * 128u + lseg);
                u.x += x.x; u.y += x.y; u.z += x.z; u.w += x.w;
            }
        if (seg + 16 <= P.n_segs) {
            if (P.mc_base) {
                mc_st_v4(P.mc_base + P.off_final_uniq + seg, u);
            } else {
#pragma unroll
                for (int q = 0; q < kMaxRanks; ++q)
                    if (q < P.n_ranks) st_peer_v4(P.final_uniq[q] + seg, u);
            }
        } else {
            const uint32_t words[4] = {u.x, u.y, u.z, u.w};
            for (uint32_t k = 0; k < 16 && seg + k < P.n_segs; ++k)
                for (int q = 0; q < P.n_ranks; ++q) P.final_uniq[q][seg + k] = (uint8_t)(words[k >> 2] >> (8 * (k & 3)));
        }
    }
}

}  // namespace fgfa
