// sm_100a kernels for FlatGFA's node-depth query, segment-major form.
//
// Same job as kernel A of depth_kernels.cuh -- the loop nest of the reference's
// `seg_depth_with_uniq` (flatgfa/src/ops/depth.rs:25-35) and `seg_depth` (depth.rs:48-52) --
// organised by SEGMENT WINDOW instead of by position in the steps pool, so that the ~80 adds a
// segment receives from different paths and different wraps of one path meet in shared
// memory and reach L2 once:
//
//   kernel S1 (k_bin_rank)     a *sub-chunk* is 32*ROWS consecutive steps of one path.  One thread
//                              per sub-chunk reads its first handle and the first handle of its
//                              successor, takes the midpoint as the sub-chunk's centre and the
//                              centre's window (kWinBin segments) as its bin; sub-chunks whose
//                              two samples lie further apart than the halo allows go to the
//                              "scattered" key.  key = bin * n_batches + (path / 32); rank inside
//                              the block + per-block histogram (key-major).
//   kernel S2 (k_bin_keyscan)  exclusive scan of the key totals -> key_begin[] (a block's place inside a
//                              key is the return value of S1's atomicAdd on the key's total).
//   kernel S3 (k_bin_scatter)  writes the sub-chunk entries {first element, path} in key order.
//   kernel W  (k_window_count) persistent, one CTA per SM; every CTA takes an EQUAL share of the
//                              sorted entry list (load balance does not depend on how the steps
//                              are distributed over the segments).  For the window of its current
//                              entries it keeps, per segment of the window (bin + halo on both
//                              sides), a u32 depth counter and a u32 *path mask* in shared memory:
//                              a step of path p is `red.shared.add [cnt], 1` and
//                              `red.shared.or [mask], 1 << (p % 32)` -- bit b of the mask is
//                              depth.rs:23's `seen` bit of path 32*batch + b, so 32 paths are
//                              in flight at once and the warps never synchronise per path.  At
//                              the end of a key (window, batch) the masks are OR-ed into the
//                              batch's global mask plane, at the end of a window the counters are
//                              added to depth[]; both flushes are coalesced.  Steps outside the
//                              window (and every step of a scattered sub-chunk) go straight to
//                              L2: one RED.ADD and one RED.OR.
//   kernel B2 (k_uniq_from_masks) uniq[s] = sum over the planes of popc(mask[s]) (depth.rs:32),
//                              and clears the planes for the next run.
//
// The pre-pass depends on the step data, so it runs with every query.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "depth_kernels.cuh"

namespace fgfa {

constexpr int kWinThreads = 1024;                 // kernel W: one CTA per SM
#ifndef FGFA_WIN_HALO
#define FGFA_WIN_HALO 6144
#endif
constexpr uint32_t kWinHalo = FGFA_WIN_HALO;      // segments on each side of the bin
// A sub-chunk whose two samples lie further apart than this is not binned ("scattered": every step goes to L2).
// Measured on config C (profiles/r2_ubench_win_14_span.log): with 2 * halo, 2 % of the sub-chunks were scattered,
// 3.9 % of all steps went to L2 and their tail cost kernel W 0.1 ms; with 4 * halo 0.19 % of the steps leave
// shared memory -- the midpoint of the samples is a good window centre even when they are 24 K segments apart,
// because the window (bin + 2 halos) is 28 K wide.  Larger limits change nothing on C and hurt config E
// (phase changes of its paths land in windows that hold none of their steps).
constexpr uint32_t kWinMaxSpan = 4 * kWinHalo;
// segments per shared-memory window: u32 counter + u32 path mask with uniq, u32 counter alone without
constexpr uint32_t kWinSegsSeen = 28672, kWinSegsDepth = 57344;
constexpr uint32_t kBinThreads = 1024;
constexpr uint32_t kBinRounds = 4;
constexpr uint32_t kBinBlock = kBinThreads * kBinRounds;   // sub-chunks ranked by one CTA of S1
constexpr uint32_t kRankBits = 12;                // rank inside a block < kBinBlock
constexpr uint32_t kMaxKeys = 12000;              // S1 keeps one counter per key in shared memory
constexpr uint32_t kEdgeBit = 0x80000000u;        // entry.y: the sub-chunk touches its path's span boundary

__host__ __device__ constexpr uint32_t win_segs(bool with_seen) { return with_seen ? kWinSegsSeen : kWinSegsDepth; }
__host__ __device__ constexpr uint32_t win_bin(bool with_seen) { return win_segs(with_seen) - 2 * kWinHalo; }
static_assert(kBinBlock <= (1u << kRankBits), "rank field too narrow");
static_assert(kWinSegsSeen % 32 == 0 && kWinSegsDepth % 32 == 0 && kWinHalo % 32 == 0, "window geometry");

struct BinParams {
    const uint32_t* __restrict__ steps;        // 128-byte aligned base of the pool
    const uint32_t* __restrict__ sub_prefix;   // [n_paths + 1] sub-chunks before path p
    const uint32_t* __restrict__ span_s;       // [n_paths] span start (relative to `steps`)
    const uint32_t* __restrict__ span_e;       // [n_paths] span end
    uint32_t path_lo, path_hi;                 // this launch covers the sub-chunks of paths [path_lo, path_hi)
    uint32_t mask_path_lo;                     // path that owns bit 0 of batch 0 (<= path_lo)
    uint32_t sub_shift;                        // log2(steps per sub-chunk)
    uint32_t n_segs;
    uint32_t bin_segs;                         // win_bin(with_seen)
    uint32_t n_bins;
    uint32_t n_batches;                        // ceil((path_hi - path_lo) / 32), or 1 without uniq
    uint32_t n_keys;                           // n_bins * n_batches; key n_keys is the scattered key
    uint32_t n_blocks;                         // CTAs of S1/S3
    uint32_t max_span;                         // |h1 - h0| above this -> scattered
    uint32_t* __restrict__ keyrank;            // [n_sub] key << kRankBits | rank
    uint32_t* __restrict__ hist;               // [(n_keys + 1) * n_blocks], key-major: first entry of block b inside key k (set where b holds k)
    uint32_t* __restrict__ key_total;          // [n_keys + 1], all-zero between pre-passes
    uint32_t* __restrict__ key_begin;          // [n_keys + 2]
    uint32_t* __restrict__ ticket;             // zero between launches
    uint2* __restrict__ entry_tmp;             // [n_sub] the entries in sub-chunk order (S1 -> S3)
    uint2* __restrict__ entries;               // [n_sub] {first element (128-byte aligned), path | edge}
};

__device__ __forceinline__ uint32_t find_sub_path(const uint32_t* __restrict__ prefix, uint32_t lo, uint32_t hi, uint32_t d) {
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= d) lo = mid; else hi = mid;
    }
    return lo;
}

// One handle of the pool for the pre-pass: read-only path, no L1 allocation, 64-byte L2 prefetch hint (a sample
// uses 4 bytes of whatever the L2 fetches from DRAM for it).  S1 0.060 -> 0.056 ms on config C; the hint size
// itself makes no difference, and neither does the CTA count (one wave or two): S1 is bound by one DRAM row
// activation per sample (profiles/r2_ubench_win_11_prepass.log).
__device__ __forceinline__ uint32_t ld_sample(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------
// S1: key + rank inside the block + per-block histogram.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBinThreads) k_bin_rank(BinParams P) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t* const s_cnt = s_dyn;                 // [n_keys + 1]
    __shared__ uint32_t s_h0[kBinBlock + 1];       // first handle (>> 1) of every sub-chunk of the block
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    for (uint32_t i = tid; i <= P.n_keys; i += kBinThreads) s_cnt[i] = 0u;
    const uint32_t d_lo = __ldg(P.sub_prefix + P.path_lo), d_hi = __ldg(P.sub_prefix + P.path_hi);
    const uint32_t sub = 1u << P.sub_shift;
    const uint32_t d0 = d_lo + blockIdx.x * kBinBlock;
    // every sample is fetched ONCE (it costs a DRAM row activation): a sub-chunk's second sample is
    // its successor's first handle, taken from shared memory
    uint32_t path[kBinRounds], last_d[kBinRounds], span_e[kBinRounds];
#pragma unroll
    for (uint32_t round = 0; round < kBinRounds; ++round) {
        const uint32_t d = d0 + round * kBinThreads + tid;
        path[round] = 0; last_d[round] = 0; span_e[round] = 0;
        if (d < d_hi) {
            const uint32_t p = find_sub_path(P.sub_prefix, P.path_lo, P.path_hi, d);
            const uint32_t s = __ldg(P.span_s + p), e = __ldg(P.span_e + p);
            const uint32_t a = (s & ~31u) + ((d - __ldg(P.sub_prefix + p)) << P.sub_shift);
            path[round] = p; span_e[round] = e;
            last_d[round] = __ldg(P.sub_prefix + p + 1) - 1u;       // last sub-chunk of this path
            s_h0[round * kBinThreads + tid] = ld_sample(P.steps + max(a, s)) >> 1;
            P.entry_tmp[d - d_lo] = make_uint2(a, p | ((a < s || (uint64_t)a + sub > e) ? kEdgeBit : 0u));
        }
    }
    if (tid == 0) {                                // the successor of the block's last sub-chunk
        const uint32_t d = d0 + kBinBlock;
        uint32_t v = 0;
        if (d < d_hi) {
            const uint32_t p = find_sub_path(P.sub_prefix, P.path_lo, P.path_hi, d);
            const uint32_t s = __ldg(P.span_s + p);
            v = ld_sample(P.steps + max((s & ~31u) + ((d - __ldg(P.sub_prefix + p)) << P.sub_shift), s)) >> 1;
        }
        s_h0[kBinBlock] = v;
    }
    __syncthreads();
#pragma unroll
    for (uint32_t round = 0; round < kBinRounds; ++round) {
        const uint32_t idx = round * kBinThreads + tid, d = d0 + idx;
        uint32_t key = 0xFFFFFFFFu - lane;         // invalid lanes never match anybody
        if (d < d_hi) {
            const uint32_t h0 = s_h0[idx];
            const uint32_t h1 = d < last_d[round] ? s_h0[idx + 1] : __ldg(P.steps + span_e[round] - 1u) >> 1;
            const uint32_t span = h1 > h0 ? h1 - h0 : h0 - h1;
            if (max(h0, h1) >= P.n_segs || span > P.max_span) key = P.n_keys;
            else key = (uint32_t)(((uint64_t)h0 + h1) >> 1) / P.bin_segs * P.n_batches + (P.n_batches > 1 ? (path[round] - P.mask_path_lo) >> 5 : 0u);
        }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (d < d_hi && before == 0u) base = atomicAdd(&s_cnt[key], __popc(peers));
        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(peers) - 1);
        if (d < d_hi) P.keyrank[d - d_lo] = (key << kRankBits) | (base + before);
    }
    __syncthreads();
    // this block's place inside every key it holds: one returning atomic per (block, key).  Which block comes first
    // inside a key is decided by the atomics -- the counting does not care -- and no histogram row has to be scanned.
    for (uint32_t i = tid; i <= P.n_keys; i += kBinThreads) {
        const uint32_t c = s_cnt[i];
        if (c) P.hist[(size_t)i * P.n_blocks + blockIdx.x] = atomicAdd(P.key_total + i, c);
    }
}

// Engine probe: `samples` evenly spaced sub-chunks; ticket[0] += sampled, ticket[1] += those S1 would
// send to the scattered key.
__global__ void __launch_bounds__(256) k_sample_spans(BinParams P, uint32_t samples, uint32_t stride) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= samples) return;
    const uint32_t d_lo = __ldg(P.sub_prefix + P.path_lo), d_hi = __ldg(P.sub_prefix + P.path_hi);
    const uint32_t d = d_lo + i * stride;
    if (d >= d_hi) return;
    const uint32_t sub = 1u << P.sub_shift;
    const uint32_t p = find_sub_path(P.sub_prefix, P.path_lo, P.path_hi, d);
    const uint32_t s = __ldg(P.span_s + p), e = __ldg(P.span_e + p);
    const uint32_t a = (s & ~31u) + ((d - __ldg(P.sub_prefix + p)) << P.sub_shift);
    const uint64_t nxt = (uint64_t)a + sub;
    const uint32_t h0 = __ldg(P.steps + max(a, s)) >> 1;
    const uint32_t h1 = __ldg(P.steps + (nxt < e ? (uint32_t)nxt : e - 1u)) >> 1;
    const uint32_t span = h1 > h0 ? h1 - h0 : h0 - h1;
    atomicAdd(P.ticket, 1u);
    if (max(h0, h1) >= P.n_segs || span > P.max_span) atomicAdd(P.ticket + 1, 1u);
}

// ---------------------------------------------------------------------------
// S2: one CTA: exclusive scan of the key totals S1 accumulated -> key_begin[]; clears the totals.
// (Until late in round 2 S1 wrote a keys x blocks histogram and S2 scanned every key's row with one CTA per
// key: 7.8 us for 919 keys; a block's place inside a key is now the return value of S1's atomicAdd.)
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;

// exclusive block scan of one value per thread (kScanThreads threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (uint32_t)o) inc += y;
    }
    __syncthreads();                               // s_warp may still be read by the previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t w = lane < kScanThreads / 32 ? s_warp[lane] : 0u, winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, winc, o);
        if (lane >= (uint32_t)o) winc += y;
    }
    *total = __shfl_sync(0xFFFFFFFFu, winc, 31);
    const uint32_t wex = __shfl_sync(0xFFFFFFFFu, winc - w, warp);
    return wex + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) k_bin_keyscan(BinParams P) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const uint32_t tid = threadIdx.x;
    uint32_t carry = 0;
    const uint32_t n = P.n_keys + 1;
    for (uint32_t t0 = 0; t0 < n; t0 += kScanThreads) {
        const uint32_t i = t0 + tid;
        const uint32_t v = i < n ? P.key_total[i] : 0u;
        uint32_t total;
        const uint32_t ex = carry + block_exclusive_scan(v, s_warp, &total);
        if (i < n) { P.key_begin[i] = ex; P.key_total[i] = 0u; }       // key_total is all-zero between pre-passes
        carry += total;
    }
    if (tid == 0) P.key_begin[n] = carry;
}

// ---------------------------------------------------------------------------
// S3: scatter the entries into key order.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBinThreads) k_bin_scatter(BinParams P) {
    const uint32_t d_lo = __ldg(P.sub_prefix + P.path_lo), d_hi = __ldg(P.sub_prefix + P.path_hi);
#pragma unroll
    for (uint32_t round = 0; round < kBinRounds; ++round) {
        const uint32_t d = d_lo + blockIdx.x * kBinBlock + round * kBinThreads + threadIdx.x;
        if (d >= d_hi) continue;
        const uint32_t kr = P.keyrank[d - d_lo];
        const uint32_t key = kr >> kRankBits, rank = kr & ((1u << kRankBits) - 1u);
        const uint32_t pos = P.key_begin[key] + P.hist[(size_t)key * P.n_blocks + blockIdx.x] + rank;
        P.entries[pos] = P.entry_tmp[d - d_lo];
    }
}

// ---------------------------------------------------------------------------
// kernel W: count the entries window by window in shared memory.
// ---------------------------------------------------------------------------
struct WindowParams {
    const uint32_t* __restrict__ steps;
    const uint2* __restrict__ entries;         // key-sorted sub-chunk entries
    const uint32_t* __restrict__ key_begin;    // [n_keys + 2]; key n_keys = scattered
    const uint32_t* __restrict__ span_s;
    const uint32_t* __restrict__ span_e;
    uint32_t n_keys, n_batches;
    uint32_t path_lo;                          // path that owns bit 0 of plane 0
    uint32_t n_segs;
    uint64_t plane_pitch;                      // u32 words between two mask planes
    uint32_t unit;                             // must be 1 (see kernel W)
    uint32_t zero;                             // must be 0: an operand ptxas cannot fold (orders the loads, see kernel W)
    uint32_t* __restrict__ depth;              // [n_segs], zero on entry (or holding earlier batches)
    uint32_t* __restrict__ masks;              // [n_batches][plane_pitch] path-mask planes (WITH_SEEN), zero on entry
    uint32_t* __restrict__ err;
    unsigned long long* __restrict__ stats;    // optional: [0] steps counted in shared memory, [1] steps sent to L2
};

constexpr size_t window_smem_bytes(bool with_seen) {
    return with_seen ? (size_t)(kWinSegsSeen + 32) * 8 : (size_t)(kWinSegsDepth + 32) * 4;
}

// DBG (measurement only, wrong results): 2 = no mask ORs, 3 = neither counters nor masks (loads + address math only),
// 4 = a byte store instead of the mask OR (what a byte-map design would pay per step).
//
// OVL (needs STAGES == 2): the software pipeline ptxas can actually keep.  ptxas tracks every step load of every
// stage on ONE scoreboard (tools/sass_ctrl.py on the STAGES = 2..4 builds), and a scoreboard is a counter: the
// first use of one stage's registers waits for ALL outstanding loads, including the ones issued a moment ago for
// the other stage.  "process, then refill" therefore exposes the full DRAM latency in every iteration, whatever
// STAGES says.  OVL orders an iteration as: wait for the current buffer -> issue the loads of the next entry ->
// count the current buffer, so that a warp's loads fly under its own ATOMS burst.  The order is forced with a
// data dependency ptxas cannot remove: the next entry's address is offset by (OR of the current words) & P.zero.
template <int ROWS, int STAGES, bool WITH_SEEN, bool STATS = false, int DBG = 0, bool OVL = false>
__global__ void __launch_bounds__(kWinThreads, 1) k_window_count(WindowParams P) {
    static_assert(STAGES >= 1 && STAGES <= 4, "stages");
    static_assert(!OVL || STAGES == 2, "the overlapped pipeline is a double buffer");
    constexpr uint32_t kSegs = win_segs(WITH_SEEN), kBin = win_bin(WITH_SEEN);
    constexpr uint32_t kPitch = kSegs + 32;            // slot kSegs = dummy for steps outside the window
    extern __shared__ uint4 smem_w[];
    uint32_t* const s_cnt = reinterpret_cast<uint32_t*>(smem_w);      // [kPitch]
    uint32_t* const s_msk = s_cnt + kPitch;                           // [kPitch] (WITH_SEEN)
    constexpr uint32_t kSub = 32u * ROWS;
    constexpr uint32_t NW = kWinThreads / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t pol = make_evict_first_policy();
    uint32_t* const depth_ptr = keep_ptr(P.depth);
    const uint32_t cnt_addr = (uint32_t)__cvta_generic_to_shared(s_cnt);
    const uint32_t one = P.unit;         // = 1, but opaque to ptxas: a literal 1 becomes ATOMS.POPC.INC, which is slower
                                         // (3.0 vs 2.5 cycles per warp instruction, tools/ubench_smem.cu)
    unsigned long long n_in = 0, n_out = 0;

    const long long dbg_t0 = DBG == 7 ? clock64() : 0;
    for (uint32_t i = tid; i < (WITH_SEEN ? 2 * kPitch : kPitch); i += kWinThreads) s_cnt[i] = 0u;
    __syncthreads();

    const uint32_t n_binned = __ldg(P.key_begin + P.n_keys), n_entries = __ldg(P.key_begin + P.n_keys + 1);

    // One contiguous range of the sorted entry list.  scattered = every step goes to L2.
    auto run_range = [&](const uint32_t begin, const uint32_t end, const bool scattered) {
        if (begin >= end) return;                          // block-uniform
        // --- per-warp pipeline: entry begin + warp + k*NW; STAGES sub-chunks of loads in flight ---
        uint32_t h[STAGES][ROWS];
        uint32_t e_path[STAGES];
        uint32_t j = begin + warp;                         // this warp's next entry to PROCESS
        auto issue_steps = [&](const uint32_t idx, uint32_t (&dst)[ROWS], uint32_t& path_out) {
            if (idx >= end) return;
            const uint2 en = __ldg(P.entries + idx);
            path_out = en.y & ~kEdgeBit;
            const uint32_t* src = P.steps + en.x + lane;
            if (!(en.y & kEdgeBit)) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) dst[r] = ld_stream_u32(src + 32 * r, pol);
            } else {
                const uint32_t s = __ldg(P.span_s + path_out), e = __ldg(P.span_e + path_out);
                const uint32_t lo = s > en.x ? s - en.x : 0u;
                const uint32_t hi = min(e - en.x, kSub);
                const uint32_t span = hi > lo ? hi - lo : 0u;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint32_t off = 32u * r + lane;
                    dst[r] = (off - lo < span) ? ld_stream_u32(src + 32 * r, pol) : kFiller;
                }
            }
        };
        // OVL: lane l of the warp holds the descriptor of the warp's (kb + l)-th entry; one coalesced-by-CTA load
        // per 32 iterations, a shuffle per iteration -- no descriptor load sits between the step loads and their use.
        uint2 en_batch = make_uint2(0u, 0u);
        uint32_t kq = 0, kb = 0;                           // this warp's iteration counter; first iteration of en_batch
        auto load_batch = [&](const uint32_t k0) {
            const uint64_t idx = (uint64_t)begin + warp + (uint64_t)(k0 + lane) * NW;
            en_batch = idx < end ? __ldg(P.entries + idx) : make_uint2(0u, 0u);
            kb = k0;
        };
        auto issue_with = [&](const uint32_t idx, const uint2 en, uint32_t (&dst)[ROWS], uint32_t& path_out, const uint32_t dep) {
            if (idx >= end) return;
            path_out = en.y & ~kEdgeBit;
            const uint32_t* src = P.steps + en.x + lane + dep;
            if (!(en.y & kEdgeBit)) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) dst[r] = ld_stream_u32(src + 32 * r, pol);
            } else {
                const uint32_t s = __ldg(P.span_s + path_out), e = __ldg(P.span_e + path_out);
                const uint32_t lo = s > en.x ? s - en.x : 0u;
                const uint32_t hi = min(e - en.x, kSub);
                const uint32_t span = hi > lo ? hi - lo : 0u;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint32_t off = 32u * r + lane;
                    dst[r] = (off - lo < span) ? ld_stream_u32(src + 32 * r, pol) : kFiller;
                }
            }
        };
        if (OVL) {                                         // one buffer in flight
            load_batch(0);
            issue_with(j, make_uint2(__shfl_sync(0xFFFFFFFFu, en_batch.x, 0), __shfl_sync(0xFFFFFFFFu, en_batch.y, 0)), h[0], e_path[0], 0u);
        } else {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) issue_steps(j + NW * s, h[s], e_path[s]);
        }
        uint32_t phase = 0;

        uint32_t i = begin, key = P.n_keys;
        if (!scattered) {                                  // key of the first entry
            uint32_t lo = 0, hi = P.n_keys;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(P.key_begin + mid) <= i) lo = mid; else hi = mid;
            }
            key = lo;
        }
        uint32_t cur_bin = 0xFFFFFFFFu;                    // window whose counters are in shared memory
        uint32_t w_lo = 0, w_n = 0;
        auto flush_counters = [&]() {                      // counters -> depth[]; leaves them zero
            for (uint32_t k = tid; k < w_n; k += kWinThreads) {
                const uint32_t v = s_cnt[k];
                if (v) { red_add_u32(depth_ptr + w_lo + k, v); s_cnt[k] = 0u; }
            }
        };
        while (i < end) {
            const uint32_t kend = scattered ? end : min(end, __ldg(P.key_begin + key + 1));
            if (i >= kend) { ++key; continue; }
            const uint32_t bin = scattered ? 0xFFFFFFFEu : key / P.n_batches;
            const uint32_t batch = scattered ? 0u : key - bin * P.n_batches;
            if (bin != cur_bin) {                          // block-uniform
                if (cur_bin < 0xFFFFFFFEu) {
                    if (!WITH_SEEN) __syncthreads();       // (with uniq the mask flush has already synchronised)
                    flush_counters();
                    __syncthreads();
                }
                cur_bin = bin;
                w_lo = scattered ? 0u : (bin * kBin > kWinHalo ? bin * kBin - kWinHalo : 0u);
                w_n = scattered ? 0u : min(kSegs, P.n_segs - w_lo);    // segments this window really holds
            }
            // ---- all warps count the entries [i, kend) ----
            auto process = [&](uint32_t (&hh)[ROWS], const uint32_t epath) {
                uint32_t mx = 0;
                const uint32_t bit = bit_of(epath - P.path_lo);        // 1 << ((path - path_lo) % 32)
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint32_t seg = hh[r] >> 1, loc = seg - w_lo;
                    mx = max(mx, loc);
                    const uint32_t lc = min(loc, kSegs);               // outside the window -> the dummy slot
                    if (DBG == 3) { if (loc == 0xFFFFFFF0u) *P.err = 2u; continue; }
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(cnt_addr + 4u * lc), "r"(one) : "memory");
                    if (WITH_SEEN && DBG == 4)       // measurement only: a byte store where the shipped kernel ORs a mask
                        asm volatile("st.shared.u8 [%0+%2], %1;" ::"r"(cnt_addr + lc), "r"(one), "n"(kPitch * 4) : "memory");
                    else if (WITH_SEEN && DBG != 2)
                        asm volatile("red.shared.or.b32 [%0+%2], %1;" ::"r"(cnt_addr + 4u * lc), "r"(bit), "n"(kPitch * 4) : "memory");
                }
                if (STATS) {
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) { const uint32_t seg = hh[r] >> 1; if (seg - w_lo < w_n) ++n_in; else if (seg < P.n_segs) ++n_out; }
                }
                if (__any_sync(0xFFFFFFFFu, mx >= w_n)) {              // rare: steps outside the window
                    const uint32_t rel = epath - P.path_lo;
                    uint32_t* __restrict__ plane = WITH_SEEN ? P.masks + (size_t)(rel >> 5) * P.plane_pitch : nullptr;
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        const uint32_t seg = hh[r] >> 1;
                        if (seg - w_lo < w_n) continue;
                        if (seg < P.n_segs) {
                            red_add_u32(depth_ptr + seg, 1u);
                            if (WITH_SEEN) red_or_b32(plane + seg, bit);
                        } else if (hh[r] != kFiller) {
                            *P.err = 1u;
                        }
                    }
                }
            };
            // OVL: consume the current buffer's words (scoreboard wait), refill the other buffer, then count.
            auto ovl_step = [&](uint32_t (&hc)[ROWS], const uint32_t pc, uint32_t (&hn)[ROWS], uint32_t& pn) {
                uint32_t x = hc[0];
#pragma unroll
                for (int r = 1; r < ROWS; ++r) x |= hc[r];
                const uint32_t dep = x & P.zero;
                if (kq + 1 - kb == 32u) load_batch(kq + 1);                    // warp-uniform, once per 32 iterations
                const uint32_t sl = kq + 1 - kb + dep;
                const uint2 en = make_uint2(__shfl_sync(0xFFFFFFFFu, en_batch.x, sl), __shfl_sync(0xFFFFFFFFu, en_batch.y, sl));
                issue_with(j + NW, en, hn, pn, dep);
                process(hc, pc);
            };
            while (OVL && j < kend) {
                if (phase == 0) ovl_step(h[0], e_path[0], h[STAGES > 1 ? 1 : 0], e_path[STAGES > 1 ? 1 : 0]);
                else ovl_step(h[STAGES > 1 ? 1 : 0], e_path[STAGES > 1 ? 1 : 0], h[0], e_path[0]);
                phase ^= 1u;
                j += NW;
                ++kq;
            }
            while (!OVL && j < kend) {
                switch (phase) {
#define FGFA_STAGE_CASE(S)                                                                                   \
    case S:                                                                                                  \
        if (S < STAGES) {                                                                                    \
            process(h[S < STAGES ? S : 0], e_path[S < STAGES ? S : 0]);                                      \
            issue_steps(j + NW * STAGES, h[S < STAGES ? S : 0], e_path[S < STAGES ? S : 0]);                 \
        }                                                                                                    \
        break;
                    FGFA_STAGE_CASE(0)
                    FGFA_STAGE_CASE(1)
                    FGFA_STAGE_CASE(2)
                    FGFA_STAGE_CASE(3)
#undef FGFA_STAGE_CASE
                }
                phase = phase + 1 == STAGES ? 0 : phase + 1;
                j += NW;
            }
            i = kend;
            if (WITH_SEEN && !scattered) {
                // ---- the key is counted: masks -> this batch's plane; leaves them zero ----
                __syncthreads();
                uint32_t* __restrict__ plane = P.masks + (size_t)batch * P.plane_pitch + w_lo;
                for (uint32_t k = tid; k < w_n; k += kWinThreads) {
                    const uint32_t v = s_msk[k];
                    if (v) { red_or_b32(plane + k, v); s_msk[k] = 0u; }
                }
                __syncthreads();
            }
            ++key;
        }
        if (cur_bin < 0xFFFFFFFEu) {
            if (!WITH_SEEN) __syncthreads();
            flush_counters();
            __syncthreads();
        }
    };

    // equal shares of the binned entries and of the scattered entries for every CTA
    {
        const uint32_t per = (n_binned + gridDim.x - 1) / gridDim.x;
        const uint32_t b0 = min(n_binned, blockIdx.x * per);
        run_range(b0, min(n_binned, b0 + per), false);
    }
    if (DBG == 7 && tid == 0) P.stats[2 + 2 * blockIdx.x] = (unsigned long long)(clock64() - dbg_t0);
    {
        const uint32_t n_sc = n_entries - n_binned;
        const uint32_t per = (n_sc + gridDim.x - 1) / gridDim.x;
        const uint32_t s0 = n_binned + min(n_sc, blockIdx.x * per);
        run_range(s0, min(n_entries, s0 + per), true);
    }
    if (DBG == 7 && tid == 0) P.stats[3 + 2 * blockIdx.x] = (unsigned long long)(clock64() - dbg_t0);
    if (STATS) {
        atomicAdd(P.stats, n_in);
        atomicAdd(P.stats + 1, n_out);
    }
}

// ---------------------------------------------------------------------------
// kernel B2: uniq[s] = number of set path bits of segment s over all planes (depth.rs:32);
// the planes are cleared for the next run.
// ---------------------------------------------------------------------------
struct MaskCountParams {
    uint32_t* __restrict__ masks;     // [n_planes][plane_pitch]
    uint32_t n_planes;
    uint64_t plane_pitch;
    uint32_t n_segs;
    void* __restrict__ uniq;          // [n_segs] u32 or u8
    int accumulate;                   // 0: uniq = cnt, 1: uniq += cnt
    int uniq_bytes;                   // 4 or 1
};

__global__ void __launch_bounds__(256) k_uniq_from_masks(MaskCountParams P) {
    const uint32_t s0 = (blockIdx.x * 256u + threadIdx.x) * 4u;
    if (s0 >= P.n_segs) return;
    uint32_t c[4] = {0u, 0u, 0u, 0u};
    if (s0 + 4u <= P.n_segs) {
        for (uint32_t b = 0; b < P.n_planes; ++b) {
            uint4* p = reinterpret_cast<uint4*>(P.masks + (size_t)b * P.plane_pitch + s0);
            const uint4 v = *p;
            if (v.x | v.y | v.z | v.w) *p = make_uint4(0u, 0u, 0u, 0u);
            c[0] += __popc(v.x); c[1] += __popc(v.y); c[2] += __popc(v.z); c[3] += __popc(v.w);
        }
    } else {
        for (uint32_t b = 0; b < P.n_planes; ++b)
            for (uint32_t k = 0; s0 + k < P.n_segs; ++k) {
                uint32_t* p = P.masks + (size_t)b * P.plane_pitch + s0 + k;
                c[k] += __popc(*p);
                *p = 0u;
            }
    }
    if (P.uniq_bytes == 1) {
        uint8_t* u = static_cast<uint8_t*>(P.uniq) + s0;
        if (s0 + 4u <= P.n_segs) {
            uint32_t packed = c[0] | c[1] << 8 | c[2] << 16 | c[3] << 24;
            if (P.accumulate) packed += *reinterpret_cast<uint32_t*>(u);     // no byte can carry: uniq <= 255
            *reinterpret_cast<uint32_t*>(u) = packed;
        } else {
            for (uint32_t k = 0; s0 + k < P.n_segs; ++k) u[k] = (uint8_t)((P.accumulate ? u[k] : 0u) + c[k]);
        }
    } else {
        uint32_t* u = static_cast<uint32_t*>(P.uniq) + s0;
        if (s0 + 4u <= P.n_segs) {
            uint4 o = make_uint4(c[0], c[1], c[2], c[3]);
            if (P.accumulate) { const uint4 q = *reinterpret_cast<uint4*>(u); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
            *reinterpret_cast<uint4*>(u) = o;
        } else {
            for (uint32_t k = 0; s0 + k < P.n_segs; ++k) u[k] = (P.accumulate ? u[k] : 0u) + c[k];
        }
    }
}

}  // namespace fgfa
