// Rejected variants of kernel W (development evidence, reachable only from tools/ubench_win.cu; not part of the
// library).  See DESIGN.md section 5 and profiles/r2_ubench_win_10_tma_ring.log.
#pragma once
#include "../pollen_b200/csrc/window_kernels.cuh"

namespace fgfa {

// ---------------------------------------------------------------------------
// kernel W, ring form (k_window_ring): the same counting, but the sub-chunks arrive through a per-warp ring of D
// 1 KiB shared-memory slots filled by TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx).
//
// Why: the register pipeline of k_window_count cannot run more than one sub-chunk ahead -- ptxas tracks the step
// loads of every stage on one scoreboard (tools/sass_ctrl.py), so a warp has at most 1 KiB in flight and the SM
// 32 KiB, against the ~64 KiB the HBM latency needs (tools/ubench_ld.cu: 0.33 ms for the register form, 0.27 /
// 0.245 ms for this ring with D = 1 / 2 on config C's 1.6 GB, loads only).  A bulk copy is tracked by its mbarrier,
// not by a scoreboard, costs one instruction of one lane per KiB, and lands in shared memory without passing
// through the LSU pipe the ATOMS are queued in.  Each warp is its own producer: as soon as it has read a slot
// into registers (8 conflict-free LDS) it issues the copy for the entry D iterations ahead into that slot and
// only then counts -- no producer warp, no "empty" barriers.  More than 64 bulk copies in flight per SM collapse
// the TMA throughput (D = 3: 0.455 ms), so D is 1 or 2.  The ring takes 32*D KiB from the window:
//   D = 1: 24576 segments with path masks (bin 12288), 49152 without (bin 36864)
//   D = 2: 20480 / 40960                   (bin  8192 / 28672)
// Sub-chunks on a span boundary are not copied (their 1 KiB may leave the pool): the consumer loads them with
// predicated LDGs as k_window_count does; their slot's barrier is completed by a plain arrive so that the
// parity of every slot keeps advancing once per use.
// ---------------------------------------------------------------------------
constexpr uint32_t kSmemMax = 232448;               // 227 KiB of dynamic shared memory per CTA on sm_100
template <int D> __host__ __device__ constexpr uint32_t ring_bytes() { return 32u * D * 1024u + 32u * D * 16u; }   // slots + {mbarrier, descriptor}
template <int D> __host__ __device__ constexpr uint32_t ring_win_segs(bool with_seen) {
    return ((kSmemMax - ring_bytes<D>()) / (with_seen ? 8u : 4u) - 32u) / 1024u * 1024u;
}
template <int D> __host__ __device__ constexpr uint32_t ring_win_bin(bool with_seen) { return ring_win_segs<D>(with_seen) - 2 * kWinHalo; }
template <int D> constexpr size_t ring_smem_bytes(bool with_seen) {
    return ring_bytes<D>() + (size_t)(ring_win_segs<D>(with_seen) + 32) * (with_seen ? 8 : 4);
}

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra W_%=;\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "l"(pol) : "memory");
}

template <int D, bool WITH_SEEN, bool STATS = false, int DBG = 0>
__global__ void __launch_bounds__(kWinThreads, 1) k_window_ring(WindowParams P) {
    static_assert(D == 1 || D == 2, "more than 64 bulk copies in flight per SM collapse the TMA throughput");
    constexpr int ROWS = 8;
    constexpr uint32_t kSegs = ring_win_segs<D>(WITH_SEEN), kBin = ring_win_bin<D>(WITH_SEEN);
    constexpr uint32_t kPitch = kSegs + 32;            // slot kSegs = dummy for steps outside the window
    constexpr uint32_t kSub = 32u * ROWS;
    constexpr uint32_t NW = kWinThreads / 32;
    extern __shared__ uint4 smem_w[];
    // [ring 32 x D x 1 KiB][meta 32 x D x {mbarrier u64, descriptor uint2}][counters kPitch][masks kPitch]
    uint8_t* const s_base = reinterpret_cast<uint8_t*>(smem_w);
    uint32_t* const s_cnt = reinterpret_cast<uint32_t*>(s_base + ring_bytes<D>());
    uint32_t* const s_msk = s_cnt + kPitch;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s_base) + warp * D * 1024u;
    const uint32_t meta = (uint32_t)__cvta_generic_to_shared(s_base) + 32u * D * 1024u + warp * D * 16u;   // +0 mbarrier, +8 descriptor
    const uint2* const s_desc = reinterpret_cast<const uint2*>(s_base + 32u * D * 1024u + warp * D * 16u + 8u);
    const uint64_t pol = make_evict_first_policy();
    uint32_t* const depth_ptr = keep_ptr(P.depth);
    const uint32_t cnt_addr = (uint32_t)__cvta_generic_to_shared(s_cnt);
    const uint32_t one = P.unit;                       // see k_window_count
    unsigned long long n_in = 0, n_out = 0;

    for (uint32_t i = tid; i < (WITH_SEEN ? 2 * kPitch : kPitch); i += kWinThreads) s_cnt[i] = 0u;
    if (lane < D) mbar_init(meta + 16u * lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t n_binned = __ldg(P.key_begin + P.n_keys), n_entries = __ldg(P.key_begin + P.n_keys + 1);
    uint32_t kt = 0;                                       // slot uses of this warp so far: slot = kt % D, parity = (kt / D) & 1

    auto run_range = [&](const uint32_t begin, const uint32_t end, const bool scattered) {
        if (begin >= end) return;                          // block-uniform
        uint32_t j = begin + warp;                         // this warp's next entry to PROCESS
        // lane l holds the descriptor of the warp's (kb + l)-th entry of this range: one load per 32 issues
        uint2 en_batch = make_uint2(0u, 0u);
        uint32_t ki = 0, kb = 0;                           // issues so far in this range; first issue index of en_batch
        auto load_batch = [&](const uint32_t k0) {
            const uint64_t idx = (uint64_t)begin + warp + (uint64_t)(k0 + lane) * NW;
            en_batch = idx < end ? __ldg(P.entries + idx) : make_uint2(0u, 0u);
            kb = k0;
        };
        // issue the copy of this warp's ki-th entry of the range into `slot` (warp-uniform)
        auto issue = [&](const uint32_t slot) {
            const uint64_t idx = (uint64_t)begin + warp + (uint64_t)ki * NW;
            if (idx < end) {
                if (ki - kb == 32u) load_batch(ki);
                const uint32_t ex = __shfl_sync(0xFFFFFFFFu, en_batch.x, ki - kb), ey = __shfl_sync(0xFFFFFFFFu, en_batch.y, ki - kb);
                if (lane == 0) {
                    const uint32_t bar = meta + 16u * slot;
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(bar + 8u), "r"(ex), "r"(ey) : "memory");
                    if (!(ey & kEdgeBit)) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot was read through the generic proxy
                        mbar_expect_tx(bar, kSub * 4u);
                        bulk_g2s(ring + slot * 1024u, P.steps + ex, kSub * 4u, bar, pol);
                    } else {
                        mbar_arrive(bar);                  // nothing to copy: complete the phase
                    }
                }
            }
            ++ki;
        };
        load_batch(0);
#pragma unroll
        for (int d = 0; d < D; ++d) issue((kt + d) % D);
        __syncwarp();

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
            while (j < kend) {
                // ---- take the sub-chunk out of its slot ----
                const uint32_t slot = kt % D;
                uint32_t hh[ROWS];
                mbar_wait(meta + 16u * slot, (kt / D) & 1u);
                const uint2 en = s_desc[2 * slot];                      // lane 0 stored it before it armed the barrier
                const uint32_t epath = en.y & ~kEdgeBit;
                if (!(en.y & kEdgeBit)) {
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(s_base + (warp * D + slot) * 1024u) + lane;
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) hh[r] = src[32 * r];
                } else {
                    const uint32_t* src = P.steps + en.x + lane;
                    const uint32_t s = __ldg(P.span_s + epath), e = __ldg(P.span_e + epath);
                    const uint32_t lo = s > en.x ? s - en.x : 0u;
                    const uint32_t hi = min(e - en.x, kSub);
                    const uint32_t span = hi > lo ? hi - lo : 0u;
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        const uint32_t off = 32u * r + lane;
                        hh[r] = (off - lo < span) ? ld_stream_u32(src + 32 * r, pol) : kFiller;
                    }
                }
                __syncwarp();                                          // every lane has its words: the slot is free
                issue(slot);
                ++kt;
                // ---- count it (as k_window_count) ----
                uint32_t mx = 0;
                const uint32_t bit = bit_of(epath - P.path_lo);        // 1 << ((path - path_lo) % 32)
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint32_t seg = hh[r] >> 1, loc = seg - w_lo;
                    mx = max(mx, loc);
                    const uint32_t lc = min(loc, kSegs);               // outside the window -> the dummy slot
                    if (DBG == 3) { if (loc == 0xFFFFFFF0u) *P.err = 2u; continue; }
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(cnt_addr + 4u * lc), "r"(one) : "memory");
                    if (WITH_SEEN && DBG != 2)
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

    {
        const uint32_t per = (n_binned + gridDim.x - 1) / gridDim.x;
        const uint32_t b0 = min(n_binned, blockIdx.x * per);
        run_range(b0, min(n_binned, b0 + per), false);
    }
    {
        const uint32_t n_sc = n_entries - n_binned;
        const uint32_t per = (n_sc + gridDim.x - 1) / gridDim.x;
        const uint32_t s0 = n_binned + min(n_sc, blockIdx.x * per);
        run_range(s0, min(n_entries, s0 + per), true);
    }
    if (STATS) {
        atomicAdd(P.stats, n_in);
        atomicAdd(P.stats + 1, n_out);
    }
}

}  // namespace fgfa
