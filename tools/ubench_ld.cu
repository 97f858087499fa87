// Load-path microbenchmark for the window engine (development tool): how fast can the steps
// pool be read in the order kernel W reads it (key-sorted 1 KiB sub-chunks), and by which
// mechanism?  Every kernel only loads and folds the words into a checksum; nothing is counted.
//   ubench_ld <cfg: C|R|E> [reps]
//   K0  pool order, grid-stride LDG.128                    (the streaming reference)
//   K1  sorted entries, many small CTAs, LDG.128 x 2 per lane per entry, U entries in flight per warp
//   K2  sorted entries, persistent 148 x 1024, LDG.32 x 8, STAGES register stages, dependent
//       descriptor load  (kernel W's load structure)
//   K3  sorted entries, persistent 148 x 1024, one producer warp issuing 1 KiB cp.async.bulk
//       copies into an NSLOT-deep shared-memory ring (mbarrier full/empty), 31 consumer warps
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../pollen_b200/csrc/window_kernels.cuh"

extern "C" {
int fgfa_synth_spans(uint32_t, uint64_t, uint32_t, uint64_t, uint32_t*, uint32_t*);
int fgfa_synth_steps(int, uint32_t, uint32_t, const uint32_t*, const uint32_t*, uint64_t, uint32_t*, int);
}

#define CK(x)                                                                                        \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) {                                                                     \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                                 \
        }                                                                                            \
    } while (0)

using namespace fgfa;

__global__ void __launch_bounds__(256) k0_stream(const uint4* __restrict__ p, size_t n16, uint32_t* out) {
    const uint64_t pol = make_evict_first_policy();
    uint32_t acc = 0;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * 256;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = ld_stream_v4(reinterpret_cast<const uint32_t*>(p + i), pol);
        uint4 b = ld_stream_v4(reinterpret_cast<const uint32_t*>(p + i + stride), pol);
        uint4 c = ld_stream_v4(reinterpret_cast<const uint32_t*>(p + i + 2 * stride), pol);
        uint4 d = ld_stream_v4(reinterpret_cast<const uint32_t*>(p + i + 3 * stride), pol);
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n16; i += stride) {
        uint4 a = ld_stream_v4(reinterpret_cast<const uint32_t*>(p + i), pol);
        acc ^= a.x ^ a.y ^ a.z ^ a.w;
    }
    if (acc == 0x9E3779B9u) *out = acc;
}

template <int U>
__global__ void __launch_bounds__(128) k1_sorted_v4(const uint32_t* __restrict__ steps, const uint2* __restrict__ entries,
                                                     uint32_t n_entries, uint32_t* out) {
    const uint64_t pol = make_evict_first_policy();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * 128 + threadIdx.x) >> 5, nw = gridDim.x * 4;
    uint32_t acc = 0;
    for (uint32_t e = gw; e < n_entries; e += nw * U) {
        uint4 v[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t idx = e + u * nw;
            if (idx < n_entries) {
                const uint2 en = __ldg(entries + idx);
                const uint32_t* src = steps + en.x + 4 * lane;
                v[u][0] = ld_stream_v4(src, pol);
                v[u][1] = ld_stream_v4(src + 128, pol);
            } else {
                v[u][0] = v[u][1] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            acc ^= v[u][0].x ^ v[u][0].y ^ v[u][0].z ^ v[u][0].w ^ v[u][1].x ^ v[u][1].y ^ v[u][1].z ^ v[u][1].w;
    }
    if (acc == 0x9E3779B9u) *out = acc;
}

template <int STAGES, bool V4>
__global__ void __launch_bounds__(1024, 1) k2_persist_regs(const uint32_t* __restrict__ steps, const uint2* __restrict__ entries,
                                                           uint32_t n_entries, uint32_t* out) {
    extern __shared__ uint4 smem_dummy[];
    const uint64_t pol = make_evict_first_policy();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = (n_entries + gridDim.x - 1) / gridDim.x;
    const uint32_t begin = min(n_entries, blockIdx.x * per), end = min(n_entries, begin + per);
    uint32_t h[STAGES][8];
    uint32_t acc = 0;
    auto issue = [&](uint32_t idx, uint32_t (&dst)[8]) {
        if (idx >= end) return;
        const uint2 en = __ldg(entries + idx);
        if (V4) {
            const uint32_t* src = steps + en.x + 4 * lane;
            const uint4 a = ld_stream_v4(src, pol), b = ld_stream_v4(src + 128, pol);
            dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
        } else {
            const uint32_t* src = steps + en.x + lane;
#pragma unroll
            for (int r = 0; r < 8; ++r) dst[r] = ld_stream_u32(src + 32 * r, pol);
        }
    };
    uint32_t j = begin + warp;
#pragma unroll
    for (int s = 0; s < STAGES; ++s) issue(j + 32 * s, h[s]);
    while (j < end) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            if (j < end) {
#pragma unroll
                for (int r = 0; r < 8; ++r) acc ^= h[s][r];
                issue(j + 32 * STAGES, h[s]);
                j += 32;
            }
        }
    }
    if (acc == 0x9E3779B9u) *out = acc + smem_dummy[0].x;
}

// ---- TMA ring ----
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

// Consumer warp c owns slots {c, c + NC, ...} (SPW of them) and the stream positions k = c (mod NC), so every slot
// is filled and drained round after round by the same consumer: no parity aliasing.
template <int SPW>
__global__ void __launch_bounds__(1024, 1) k3_ring(const uint32_t* __restrict__ steps, const uint2* __restrict__ entries,
                                                   uint32_t n_entries, uint32_t* out) {
    extern __shared__ __align__(1024) uint8_t smem_r[];
    constexpr uint32_t NC = 31, NSLOT = NC * SPW;
    // [ring NSLOT x 1024][desc NSLOT x 8][full NSLOT x 8][empty NSLOT x 8]
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem_r);
    uint2* const s_desc = reinterpret_cast<uint2*>(smem_r + NSLOT * 1024);
    const uint32_t full = ring + NSLOT * 1024 + NSLOT * 8, empty = full + NSLOT * 8;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = (n_entries + gridDim.x - 1) / gridDim.x;
    const uint32_t begin = min(n_entries, blockIdx.x * per), end = min(n_entries, begin + per);
    const uint32_t total = end - begin;
    if (threadIdx.x < NSLOT) { mbar_init(full + 8 * threadIdx.x, 1); mbar_init(empty + 8 * threadIdx.x, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (warp == NC) {
        // producer warp: lane l issues stream positions base + l; one iteration = 32 consecutive positions
        const uint64_t pol = make_evict_first_policy();
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t k = base + lane;
            if (k < total) {
                const uint32_t c = k % NC, j = k / NC;           // j-th item of consumer c
                const uint32_t s = (j % SPW) * NC + c, round = j / SPW;
                const uint2 en = __ldg(entries + begin + k);
                if (round > 0) mbar_wait(empty + 8 * s, (round - 1) & 1);
                s_desc[s] = en;
                mbar_expect_tx(full + 8 * s, 1024);
                bulk_g2s(ring + s * 1024, steps + en.x, 1024, full + 8 * s, pol);
            }
            __syncwarp();
        }
    } else {
        uint32_t acc = 0, j = 0;
        for (uint32_t k = warp; k < total; k += NC, ++j) {
            const uint32_t s = (j % SPW) * NC + warp, round = j / SPW;
            mbar_wait(full + 8 * s, round & 1);
            const uint2 en = s_desc[s];
            const uint32_t* src = reinterpret_cast<const uint32_t*>(smem_r + s * 1024) + lane;
            uint32_t h[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) h[r] = src[32 * r];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + 8 * s);
#pragma unroll
            for (int r = 0; r < 8; ++r) acc ^= h[r];
            acc += en.y;
        }
        if (acc == 0x9E3779B9u) *out = acc;
    }
}

// K4: every warp runs its own ring of D 1 KiB slots: lane 0 issues a cp.async.bulk for the entry D iterations ahead
// into the slot it has just drained (no producer warp, no empty barriers: the consumer is the producer).
template <int D, bool FENCE>
__global__ void __launch_bounds__(1024, 1) k4_self_tma(const uint32_t* __restrict__ steps, const uint2* __restrict__ entries,
                                                       uint32_t n_entries, uint32_t* out) {
    extern __shared__ __align__(1024) uint8_t smem_r[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem_r) + warp * D * 1024;
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem_r) + 32 * D * 1024 + warp * D * 8;
    const uint32_t per = (n_entries + gridDim.x - 1) / gridDim.x;
    const uint32_t begin = min(n_entries, blockIdx.x * per), end = min(n_entries, begin + per);
    const uint64_t pol = make_evict_first_policy();
    if (lane < D) mbar_init(bars + 8 * lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](uint32_t idx, uint32_t slot) {
        if (idx < end && lane == 0) {
            const uint2 en = __ldg(entries + idx);
            if (FENCE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bars + 8 * slot, 1024);
            bulk_g2s(ring + slot * 1024, steps + en.x, 1024, bars + 8 * slot, pol);
        }
    };
    uint32_t j = begin + warp;
#pragma unroll
    for (int d = 0; d < D; ++d) issue(j + 32 * d, d);
    uint32_t acc = 0, k = 0;
    for (; j < end; j += 32, ++k) {
        const uint32_t slot = k % D;
        mbar_wait(bars + 8 * slot, (k / D) & 1);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(smem_r + (warp * D + slot) * 1024) + lane;
        uint32_t h[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) h[r] = src[32 * r];
        __syncwarp();
        issue(j + 32 * D, slot);
#pragma unroll
        for (int r = 0; r < 8; ++r) acc ^= h[r];
    }
    if (acc == 0x9E3779B9u) *out = acc;
}

// K5: the same ring filled with per-lane 16-byte cp.async (LDGSTS), one commit group per entry, wait_group D-1.
template <int D>
__global__ void __launch_bounds__(1024, 1) k5_self_cpasync(const uint32_t* __restrict__ steps, const uint2* __restrict__ entries,
                                                           uint32_t n_entries, uint32_t* out) {
    extern __shared__ __align__(1024) uint8_t smem_r[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem_r) + warp * D * 1024;
    const uint32_t per = (n_entries + gridDim.x - 1) / gridDim.x;
    const uint32_t begin = min(n_entries, blockIdx.x * per), end = min(n_entries, begin + per);
    auto issue = [&](uint32_t idx, uint32_t slot) {
        if (idx < end) {
            const uint2 en = __ldg(entries + idx);
            const uint32_t* src = steps + en.x + 4 * lane;
            const uint32_t dst = ring + slot * 1024 + 16 * lane;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512), "l"(src + 128) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t j = begin + warp;
#pragma unroll
    for (int d = 0; d < D; ++d) issue(j + 32 * d, d);
    uint32_t acc = 0, k = 0;
    for (; j < end; j += 32, ++k) {
        const uint32_t slot = k % D;
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        __syncwarp();
        const uint32_t* src = reinterpret_cast<const uint32_t*>(smem_r + (warp * D + slot) * 1024) + lane;
        uint32_t h[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) h[r] = src[32 * r];
        __syncwarp();
        issue(j + 32 * D, slot);
#pragma unroll
        for (int r = 0; r < 8; ++r) acc ^= h[r];
    }
    if (acc == 0x9E3779B9u) *out = acc;
}

struct Cfg { const char* name; uint32_t n_segs, n_paths; uint64_t n_steps; int kind; uint32_t jitter; };

int main(int argc, char** argv) {
    std::string which = argc > 1 ? argv[1] : "C";
    int reps = argc > 2 ? atoi(argv[2]) : 10;
    Cfg cfg;
    if (which == "C") cfg = {"C", 5000000, 90, 400000000ull, 0, 20};
    else if (which == "E") cfg = {"E", 5000000, 8, 400000000ull, 1, 0};
    else if (which == "R") cfg = {"R", 5000000, 90, 400000000ull, 3, 20};
    else if (which == "S") cfg = {"S", 200000, 16, 4000000ull, 0, 20};
    else { fprintf(stderr, "unknown cfg\n"); return 2; }
    int n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads > 64) n_threads = 64;
    std::vector<uint32_t> ss(cfg.n_paths), se(cfg.n_paths);
    std::vector<uint32_t> steps(cfg.n_steps);
    fgfa_synth_spans(cfg.n_paths, cfg.n_steps, cfg.jitter, 0xB1011054ull, ss.data(), se.data());
    fgfa_synth_steps(cfg.kind, cfg.n_segs, cfg.n_paths, ss.data(), se.data(), 0xB1011054ull, steps.data(), n_threads);
    printf("cfg %s: n_segs=%u n_paths=%u n_steps=%llu\n", cfg.name, cfg.n_segs, cfg.n_paths, (unsigned long long)cfg.n_steps);

    uint32_t *d_steps, *d_ss, *d_se, *d_out;
    CK(cudaMalloc(&d_steps, cfg.n_steps * 4 + 4096));
    CK(cudaMalloc(&d_ss, cfg.n_paths * 4));
    CK(cudaMalloc(&d_se, cfg.n_paths * 4));
    CK(cudaMalloc(&d_out, 4));
    CK(cudaMemset(d_steps + cfg.n_steps, 0xFF, 4096));
    CK(cudaMemcpy(d_steps, steps.data(), cfg.n_steps * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ss, ss.data(), cfg.n_paths * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_se, se.data(), cfg.n_paths * 4, cudaMemcpyHostToDevice));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;

    // ---- pre-pass (as ubench_win): key-sorted entries ----
    const uint32_t sub_shift = 8, sub = 256;
    std::vector<uint32_t> prefix(cfg.n_paths + 1, 0);
    for (uint32_t p = 0; p < cfg.n_paths; ++p) {
        const uint64_t a = ss[p] & ~31u;
        const uint64_t n = se[p] > ss[p] ? (se[p] - a + sub - 1) / sub : 0;
        prefix[p + 1] = prefix[p] + (uint32_t)n;
    }
    const uint32_t n_sub = prefix[cfg.n_paths];
    BinParams B{};
    B.steps = d_steps; B.span_s = d_ss; B.span_e = d_se; B.path_lo = 0; B.path_hi = cfg.n_paths; B.mask_path_lo = 0;
    B.sub_shift = sub_shift; B.n_segs = cfg.n_segs;
    B.bin_segs = win_bin(true);
    B.n_bins = (cfg.n_segs + B.bin_segs - 1) / B.bin_segs;
    B.n_batches = (cfg.n_paths + 31) / 32;
    B.n_keys = B.n_bins * B.n_batches;
    B.n_blocks = (n_sub + kBinBlock - 1) / kBinBlock;
    B.max_span = kWinMaxSpan;
    uint32_t *d_prefix, *d_keyrank, *d_hist, *d_key_begin, *d_key_total, *d_ticket;
    uint2 *d_entries, *d_entry_tmp;
    CK(cudaMalloc(&d_prefix, (cfg.n_paths + 1) * 4));
    CK(cudaMemcpy(d_prefix, prefix.data(), (cfg.n_paths + 1) * 4, cudaMemcpyHostToDevice));
    B.sub_prefix = d_prefix;
    CK(cudaMalloc(&d_key_total, (size_t)(B.n_keys + 1) * 4));
        CK(cudaMemset(d_key_total, 0, (size_t)(B.n_keys + 1) * 4));
    CK(cudaMalloc(&d_ticket, 4));
    CK(cudaMemset(d_ticket, 0, 4));
    CK(cudaMalloc(&d_keyrank, (size_t)n_sub * 4));
    CK(cudaMalloc(&d_hist, (size_t)(B.n_keys + 1) * B.n_blocks * 4));
    CK(cudaMalloc(&d_key_begin, (size_t)(B.n_keys + 2) * 4));
    CK(cudaMalloc(&d_entries, (size_t)n_sub * 8));
    CK(cudaMalloc(&d_entry_tmp, (size_t)n_sub * 8));
    B.keyrank = d_keyrank; B.hist = d_hist; B.key_begin = d_key_begin; B.entries = d_entries; B.entry_tmp = d_entry_tmp;
    B.key_total = d_key_total; B.ticket = d_ticket;
    k_bin_rank<<<B.n_blocks, kBinThreads, (B.n_keys + 1) * 4>>>(B);
    k_bin_keyscan<<<1, kScanThreads>>>(B);
    k_bin_scatter<<<B.n_blocks, kBinThreads>>>(B);
    CK(cudaDeviceSynchronize());
    // drop the entries that touch a span boundary (their 1 KiB may leave the pool): point them at sub-chunk 0
    {
        std::vector<uint2> h(n_sub);
        CK(cudaMemcpy(h.data(), d_entries, (size_t)n_sub * 8, cudaMemcpyDeviceToHost));
        uint32_t edges = 0;
        for (auto& e : h) if (e.y & kEdgeBit) { e.x = 0; ++edges; }
        CK(cudaMemcpy(d_entries, h.data(), (size_t)n_sub * 8, cudaMemcpyHostToDevice));
        printf("n_sub=%u (edge entries redirected: %u)\n", n_sub, edges);
    }
    const char* only = getenv("UBENCH_ONLY");
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double bytes = (double)n_sub * 1024.0;
    auto timeit = [&](const char* name, auto&& launch) {
        if (only && !strstr(name, only)) return;
        float best = 1e30f, sum = 0;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) { best = std::min(best, ms); sum += ms; }
        }
        printf("%-44s best %.3f avg %.3f ms  %.0f GB/s\n", name, best, sum / reps, bytes / (sum / reps * 1e6));
        fflush(stdout);
    };
    const size_t n16 = cfg.n_steps / 4;
    timeit("K0 pool order LDG.128 grid 148x8", [&] { k0_stream<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(d_steps), n16, d_out); });
    timeit("K1 sorted v4 U=1 grid 148x16", [&] { k1_sorted_v4<1><<<sms * 16, 128>>>(d_steps, d_entries, n_sub, d_out); });
    timeit("K1 sorted v4 U=2 grid 148x16", [&] { k1_sorted_v4<2><<<sms * 16, 128>>>(d_steps, d_entries, n_sub, d_out); });
    timeit("K1 sorted v4 U=4 grid 148x8", [&] { k1_sorted_v4<4><<<sms * 8, 128>>>(d_steps, d_entries, n_sub, d_out); });
    const size_t big = 200 * 1024;
#define K2(S, V4)                                                                                               \
    do {                                                                                                        \
        CK(cudaFuncSetAttribute(k2_persist_regs<S, V4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big)); \
        timeit("K2 persistent regs stages=" #S " v4=" #V4, [&] { k2_persist_regs<S, V4><<<sms, 1024, big>>>(d_steps, d_entries, n_sub, d_out); }); \
    } while (0)
    K2(2, false); K2(3, false); K2(4, false); K2(2, true); K2(3, true);
#define K3(SPW)                                                                                                 \
    do {                                                                                                        \
        CK(cudaFuncSetAttribute(k3_ring<SPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));          \
        timeit("K3 TMA ring slots/warp=" #SPW, [&] { k3_ring<SPW><<<sms, 1024, big>>>(d_steps, d_entries, n_sub, d_out); }); \
    } while (0)
    K3(2); K3(4);
#define K4(D, F)                                                                                                \
    do {                                                                                                        \
        CK(cudaFuncSetAttribute(k4_self_tma<D, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));     \
        timeit("K4 self-service TMA ring D=" #D " fence=" #F, [&] { k4_self_tma<D, F><<<sms, 1024, big>>>(d_steps, d_entries, n_sub, d_out); }); \
    } while (0)
    K4(1, false); K4(2, false); K4(3, false); K4(4, false); K4(2, true); K4(3, true);
#define K5(D)                                                                                                   \
    do {                                                                                                        \
        CK(cudaFuncSetAttribute(k5_self_cpasync<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));    \
        timeit("K5 self-service cp.async ring D=" #D, [&] { k5_self_cpasync<D><<<sms, 1024, big>>>(d_steps, d_entries, n_sub, d_out); }); \
    } while (0)
    K5(1); K5(2); K5(3); K5(4);
    return 0;
}
