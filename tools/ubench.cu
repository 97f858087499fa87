// Kernel microbenchmark for the depth path (development tool; not the product bench).
// Generates one synthetic config on the host, uploads it, times each kernel variant
// with CUDA events and checks the full pipeline against the C oracle.
//
//   ubench <cfg: B|C|E|U> [reps] [verify 0/1]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "experimental_kernels.cuh"

extern "C" {
int fgfa_synth_spans(uint32_t, uint64_t, uint32_t, uint64_t, uint32_t*, uint32_t*);
int fgfa_synth_steps(int, uint32_t, uint32_t, const uint32_t*, const uint32_t*, uint64_t,
                     uint32_t*, int);
int oracle_seg_depth_with_uniq(const uint32_t*, uint64_t, const uint32_t*, uint32_t, uint32_t,
                               uint64_t*, uint64_t*);
}

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

using namespace fgfa;

struct Cfg { const char* name; uint32_t n_segs, n_paths; uint64_t n_steps; int kind; uint32_t jitter; };

int main(int argc, char** argv) {
    std::string which = argc > 1 ? argv[1] : "B";
    int reps = argc > 2 ? atoi(argv[2]) : 10;
    int verify = argc > 3 ? atoi(argv[3]) : 1;
    Cfg cfg;
    if (which == "B") cfg = {"B", 1000000, 16, 20000000ull, 0, 0};
    else if (which == "C") cfg = {"C", 5000000, 90, 400000000ull, 0, 20};
    else if (which == "E") cfg = {"E", 5000000, 8, 400000000ull, 1, 0};
    else if (which == "U") cfg = {"U", 5000000, 90, 400000000ull, 2, 20};
    else if (which == "S") cfg = {"S", 5000, 7, 100003ull, 0, 30};
    else if (which == "R") cfg = {"R", 5000000, 90, 400000000ull, 3, 20};
    else { fprintf(stderr, "unknown cfg\n"); return 2; }

    int n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads > 64) n_threads = 64;
    std::vector<uint32_t> ss(cfg.n_paths), se(cfg.n_paths);
    std::vector<uint32_t> steps(cfg.n_steps);
    auto t0 = std::chrono::steady_clock::now();
    fgfa_synth_spans(cfg.n_paths, cfg.n_steps, cfg.jitter, 0xB1011054ull, ss.data(), se.data());
    fgfa_synth_steps(cfg.kind, cfg.n_segs, cfg.n_paths, ss.data(), se.data(), 0xB1011054ull,
                     steps.data(), n_threads);
    auto t1 = std::chrono::steady_clock::now();
    printf("cfg %s: n_segs=%u n_paths=%u n_steps=%llu gen %.2fs (%d threads)\n", cfg.name,
           cfg.n_segs, cfg.n_paths, (unsigned long long)cfg.n_steps,
           std::chrono::duration<double>(t1 - t0).count(), n_threads);

    // chunk table
    std::vector<uint32_t> prefix(cfg.n_paths + 1, 0);
    for (uint32_t p = 0; p < cfg.n_paths; ++p) {
        uint64_t a = ss[p] & ~3u;
        uint64_t n = se[p] > ss[p] ? (se[p] - a + kChunk - 1) / kChunk : 0;
        prefix[p + 1] = prefix[p] + (uint32_t)n;
    }
    uint32_t n_chunks = prefix[cfg.n_paths];
    uint32_t n_words = (cfg.n_segs + 31) / 32;
    uint32_t wpr = (n_words + 31) & ~31u;

    uint32_t *d_steps, *d_ss, *d_se, *d_prefix, *d_depth, *d_uniq, *d_bitmap, *d_err;
    CK(cudaMalloc(&d_steps, cfg.n_steps * 4 + 16));
    CK(cudaMalloc(&d_ss, cfg.n_paths * 4));
    CK(cudaMalloc(&d_se, cfg.n_paths * 4));
    CK(cudaMalloc(&d_prefix, (cfg.n_paths + 1) * 4));
    CK(cudaMalloc(&d_depth, (size_t)cfg.n_segs * 4));
    CK(cudaMalloc(&d_uniq, (size_t)cfg.n_segs * 4));
    CK(cudaMalloc(&d_bitmap, (size_t)cfg.n_paths * wpr * 4));
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMemcpy(d_steps, steps.data(), cfg.n_steps * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ss, ss.data(), cfg.n_paths * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_se, se.data(), cfg.n_paths * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_prefix, prefix.data(), (cfg.n_paths + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_bitmap, 0, (size_t)cfg.n_paths * wpr * 4));
    CK(cudaMemset(d_err, 0, 4));

    ExpParams P{};
    P.steps = d_steps; P.n_steps = cfg.n_steps; P.span_start = d_ss; P.span_end = d_se;
    P.chunk_prefix = d_prefix; P.path_lo = 0; P.path_hi = cfg.n_paths; P.n_segs = cfg.n_segs;
    P.words_per_row = wpr; P.depth = d_depth; P.bitmap = d_bitmap; P.err = d_err;
    std::vector<ChunkDesc> table(n_chunks);
    for (uint32_t p = 0; p < cfg.n_paths; ++p) {
        uint64_t a = ss[p] & ~3u;
        for (uint32_t c = prefix[p]; c < prefix[p + 1]; ++c, a += kChunk) table[c] = ChunkDesc{(uint32_t)a, ss[p], se[p], p};
    }
    // experiment: the ORDER of the chunk table decides which chunks run concurrently (kernel A walks it
    // with stride gridDim.x).  UBENCH_PERMUTE=contig:<grid> gives every CTA a contiguous share of the
    // original order, UBENCH_PERMUTE=random shuffles it.
    if (const char* pm = getenv("UBENCH_PERMUTE")) {
        std::vector<ChunkDesc> t2;
        t2.reserve(n_chunks);
        if (!strncmp(pm, "contig:", 7)) {
            const uint32_t G = (uint32_t)atoi(pm + 7);
            const uint32_t per = (n_chunks + G - 1) / G;
            for (uint32_t r = 0; r < per; ++r)
                for (uint32_t b = 0; b < G; ++b) {
                    const uint64_t i = (uint64_t)b * per + r;
                    if (i < n_chunks) t2.push_back(table[i]);
                }
        } else {
            t2 = table;
            uint64_t x = 0x9E3779B97F4A7C15ull;
            for (size_t i = t2.size(); i > 1; --i) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                std::swap(t2[i - 1], t2[x % i]);
            }
        }
        table.swap(t2);
        printf("chunk table permuted: %s\n", pm);
    }
    ChunkDesc* d_chunks;
    CK(cudaMalloc(&d_chunks, std::max<size_t>(n_chunks, 1) * sizeof(ChunkDesc)));
    CK(cudaMemcpy(d_chunks, table.data(), (size_t)n_chunks * sizeof(ChunkDesc), cudaMemcpyHostToDevice));
    StreamParams S{};
    S.steps = d_steps; S.chunks = d_chunks; S.chunk_lo = 0; S.chunk_hi = n_chunks; S.path_lo = 0;
    S.n_segs = cfg.n_segs; S.words_per_row = wpr; S.depth = d_depth; S.bitmap = d_bitmap; S.err = d_err;
    PopcountParams Q{};
    Q.bitmap = d_bitmap; Q.n_rows = cfg.n_paths; Q.words_per_row = wpr; Q.n_words = n_words;
    Q.n_segs = cfg.n_segs; Q.uniq = d_uniq; Q.depth = nullptr; Q.accumulate = 0;

    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, chunks=%u, bitmap %.1f MB\n", prop.name, sms, n_chunks,
           (double)cfg.n_paths * wpr * 4 / 1e6);

    if (const char* pv = getenv("UBENCH_PERSIST")) {
        // experiment: pin the depth table (and, with 2, the bitmap too) in L2 with a persisting window
        printf("persistingL2CacheMaxSize = %.1f MB, accessPolicyMaxWindowSize = %.1f MB\n",
               prop.persistingL2CacheMaxSize / 1e6, prop.accessPolicyMaxWindowSize / 1e6);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize));
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = atoi(pv) == 2 ? (void*)d_bitmap : (void*)d_depth;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(atoi(pv) == 2 ? (size_t)cfg.n_paths * wpr * 4 : (size_t)cfg.n_segs * 4,
                                                           prop.accessPolicyMaxWindowSize);
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(0, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double alg_bytes = 4.0 * cfg.n_steps + 8.0 * cfg.n_paths + 8.0 * cfg.n_segs;

    auto time_it = [&](const char* name, auto&& launch, bool clear_bitmap) {
        float best = 1e30f, sum = 0;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaMemsetAsync(d_depth, 0, (size_t)cfg.n_segs * 4));
            if (clear_bitmap) CK(cudaMemsetAsync(d_bitmap, 0, (size_t)cfg.n_paths * wpr * 4));
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) { best = std::min(best, ms); sum += ms; }
        }
        float avg = sum / reps;
        printf("%-34s best %8.3f ms  avg %8.3f ms  %8.1f Gstep/s  alg %7.1f GB/s (%.1f%% of 6540)\n",
               name, best, avg, cfg.n_steps / (avg * 1e6), alg_bytes / (avg * 1e6),
               100.0 * alg_bytes / (avg * 1e6) / 6540.2);
    };

    const bool only_merged = getenv("UBENCH_ONLY_MERGED") != nullptr;   // skip the experimental kernels
    if (!only_merged)
    for (int bps : {2, 4, 8}) {
        int grid = sms * bps;
        char nm[96];
        snprintf(nm, sizeof nm, "read-only lane-order g=%dxSM", bps);
        time_it(nm, [&] { k_step_stream_direct<kModeReadOnly, 1><<<grid, kThreads>>>(P); }, false);
        snprintf(nm, sizeof nm, "read-only v4 g=%dxSM", bps);
        time_it(nm, [&] { k_step_stream_direct<kModeReadOnly, 0><<<grid, kThreads>>>(P); }, false);
    }
    if (!only_merged)
    for (int bps : {4, 8}) {
        int grid = sms * bps;
        char nm[96];
        snprintf(nm, sizeof nm, "depth-only lane-order g=%dxSM", bps);
        time_it(nm, [&] { k_step_stream_direct<kModeDepthOnly, 1><<<grid, kThreads>>>(P); }, false);
        snprintf(nm, sizeof nm, "depth-only v4 g=%dxSM", bps);
        time_it(nm, [&] { k_step_stream_direct<kModeDepthOnly, 0><<<grid, kThreads>>>(P); }, false);
        snprintf(nm, sizeof nm, "seen-only lane-order g=%dxSM", bps);
        time_it(nm, [&] { k_step_stream_direct<kModeSeenOnly, 1><<<grid, kThreads>>>(P); }, true);
        snprintf(nm, sizeof nm, "seen-only v4 g=%dxSM", bps);
        time_it(nm, [&] { k_step_stream_direct<kModeSeenOnly, 0><<<grid, kThreads>>>(P); }, true);
        snprintf(nm, sizeof nm, "depth+seen lane-order g=%dxSM", bps);
        time_it(nm, [&] { k_step_stream_direct<kModeDepthAndSeen, 1><<<grid, kThreads>>>(P); }, true);
        snprintf(nm, sizeof nm, "depth+seen v4 g=%dxSM", bps);
        time_it(nm, [&] { k_step_stream_direct<kModeDepthAndSeen, 0><<<grid, kThreads>>>(P); }, true);
    }
    if (!only_merged) {
        int pgrid = (n_words + kPopThreads - 1) / kPopThreads;
        // populate bitmap, then time popcount (it clears as it goes, so refill each rep)
        auto time_pop = [&](const char* nm, auto&& launch) {
            float best = 1e30f;
            for (int r = 0; r < reps; ++r) {
                k_step_stream_direct<kModeSeenOnly, 1><<<sms * 8, kThreads>>>(P);
                CK(cudaEventRecord(e0));
                launch();
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                best = std::min(best, ms);
            }
            printf("%-34s best %8.3f ms\n", nm, best);
        };
        time_pop("uniq popcount <1>", [&] { k_uniq_popcount<1><<<pgrid, kPopThreads>>>(Q); });
        time_pop("uniq popcount <4>", [&] { k_uniq_popcount<4><<<pgrid, kPopThreads>>>(Q); });
        time_pop("uniq popcount <6>", [&] { k_uniq_popcount<6><<<pgrid, kPopThreads>>>(Q); });
        time_pop("uniq popcount <8>", [&] { k_uniq_popcount<8><<<pgrid, kPopThreads>>>(Q); });
        time_pop("uniq popcount <12>", [&] { k_uniq_popcount<12><<<pgrid, kPopThreads>>>(Q); });
        float best = 1e30f;
        // full pipeline: memset + A + B
        float sum = 0; best = 1e30f;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaEventRecord(e0));
            CK(cudaMemsetAsync(d_depth, 0, (size_t)cfg.n_segs * 4));
            k_step_stream_direct<kModeDepthAndSeen, 1><<<sms * 8, kThreads>>>(P);
            k_uniq_popcount<4><<<pgrid, kPopThreads>>>(Q);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) { best = std::min(best, ms); sum += ms; }
        }
        float avg = sum / reps;
        printf("%-34s best %8.3f ms  avg %8.3f ms  %8.1f Gstep/s  alg %7.1f GB/s (%.1f%% of 6540)\n",
               "PIPELINE memset+A+B", best, avg, cfg.n_steps / (avg * 1e6), alg_bytes / (avg * 1e6),
               100.0 * alg_bytes / (avg * 1e6) / 6540.2);
    }
    {
        int pgrid = (n_words + kPopThreads - 1) / kPopThreads;
        PopcountParams Q2 = Q; Q2.depth = d_depth;
        auto run_ft = [&](int bps) {
            int grid = sms * bps;
            if (bps == 2) k_step_stream_first_touch<2><<<grid, kThreads>>>(P);
            else if (bps == 3) k_step_stream_first_touch<3><<<grid, kThreads>>>(P);
            else if (bps == 4) k_step_stream_first_touch<4><<<grid, kThreads>>>(P);
            else if (bps == 6) k_step_stream_first_touch<6><<<grid, kThreads>>>(P);
            else k_step_stream_first_touch<8><<<grid, kThreads>>>(P);
        };
        for (int bps : {2, 3, 4, 6, 8}) {
            if (only_merged) break;
            char nm[96];
            snprintf(nm, sizeof nm, "first-touch A only g=%dxSM", bps);
            time_it(nm, [&] { run_ft(bps); }, true);
        }
        if (!only_merged) {
        time_it("depth-only half-lanes(seg even) g=8xSM", [&] { k_step_stream_direct<kModeDepthHalfLanes, 1><<<sms * 8, kThreads>>>(P); }, false);
        time_it("depth-only pairs(+2 even tid) g=8xSM", [&] { k_step_stream_direct<kModeDepthPairs, 1><<<sms * 8, kThreads>>>(P); }, false);
        }
        const size_t smD = stream_smem_bytes(kSeenDirect), smW = stream_smem_bytes(kSeenWindow);
        CK(cudaFuncSetAttribute(k_step_stream_merged<8, kSeenWindow>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smW));
        CK(cudaFuncSetAttribute(k_step_stream_merged<2, kSeenDeferred>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_step_stream_merged<3, kSeenDeferred>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_step_stream_merged<4, kSeenDeferred>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_step_stream_merged<5, kSeenDeferred>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_step_stream_merged<4, kSeenWindow>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smW));
        if (!only_merged) {
        time_it("merged direct-OR g=4xSM", [&] { k_step_stream_merged<4, kSeenDirect><<<sms * 4, kThreads, smD>>>(S); }, true);
        time_it("merged direct-OR g=8xSM", [&] { k_step_stream_merged<8, kSeenDirect><<<sms * 8, kThreads, smD>>>(S); }, true);
        }
        const size_t smF = stream_smem_bytes(kSeenDeferred);
        time_it("merged deferred-OR g=4xSM", [&] { k_step_stream_merged<4, kSeenDeferred><<<sms * 4, kThreads, smF>>>(S); }, true);
        time_it("merged deferred-OR <2> grid 4x", [&] { k_step_stream_merged<2, kSeenDeferred><<<sms * 4, kThreads, smF>>>(S); }, true);
        time_it("merged deferred-OR <3> grid 6x", [&] { k_step_stream_merged<3, kSeenDeferred><<<sms * 6, kThreads, smF>>>(S); }, true);
        time_it("merged deferred-OR <4> grid 8x", [&] { k_step_stream_merged<4, kSeenDeferred><<<sms * 8, kThreads, smF>>>(S); }, true);
        time_it("merged deferred-OR <5> grid 5x", [&] { k_step_stream_merged<5, kSeenDeferred><<<sms * 5, kThreads, smF>>>(S); }, true);
        time_it("deferred-OR P2=8 <5> grid 10x", [&] { k_step_stream_merged<5, kSeenDeferred, 8><<<sms * 10, kThreads, smF>>>(S); }, true);
        time_it("deferred-OR P2=8 <4> grid 8x", [&] { k_step_stream_merged<4, kSeenDeferred, 8><<<sms * 8, kThreads, smF>>>(S); }, true);
        time_it("deferred-OR P2=8 <6> grid 10x", [&] { k_step_stream_merged<6, kSeenDeferred, 8><<<sms * 10, kThreads, smF>>>(S); }, true);
        time_it("deferred-OR P2=4 <5> grid 10x", [&] { k_step_stream_merged<5, kSeenDeferred, 4><<<sms * 10, kThreads, smF>>>(S); }, true);
        time_it("merged deferred-OR <5> grid 10x", [&] { k_step_stream_merged<5, kSeenDeferred><<<sms * 10, kThreads, smF>>>(S); }, true);
        time_it("merged deferred-OR <6> grid 5x", [&] { k_step_stream_merged<6, kSeenDeferred><<<sms * 5, kThreads, smF>>>(S); }, true);
        time_it("merged deferred-OR g=8xSM", [&] { k_step_stream_merged<8, kSeenDeferred><<<sms * 8, kThreads, smF>>>(S); }, true);
        if (!only_merged) {
        time_it("merged window-OR g=4xSM", [&] { k_step_stream_merged<4, kSeenWindow><<<sms * 4, kThreads, smW>>>(S); }, true);
        time_it("merged window-OR g=8xSM", [&] { k_step_stream_merged<8, kSeenWindow><<<sms * 8, kThreads, smW>>>(S); }, true);
        time_it("merged depth-only g=8xSM", [&] { k_step_stream_merged<8, kSeenNone><<<sms * 8, kThreads, smD>>>(S); }, true);
        }
        if (!only_merged) {
        time_it("warp-agg A only g=4xSM", [&] { k_step_stream_warp_agg<4><<<sms * 4, kThreads>>>(P); }, true);
        time_it("warp-agg A only g=6xSM", [&] { k_step_stream_warp_agg<6><<<sms * 6, kThreads>>>(P); }, true);
        time_it("warp-agg A only g=8xSM", [&] { k_step_stream_warp_agg<8><<<sms * 8, kThreads>>>(P); }, true);
        time_it("FT exp1 (no pass3) g=4xSM", [&] { k_step_stream_first_touch<4, 1><<<sms * 4, kThreads>>>(P); }, true);
        time_it("FT exp2 (RED.OR, no pass3) g=4xSM", [&] { k_step_stream_first_touch<4, 2><<<sms * 4, kThreads>>>(P); }, true);
        time_it("FT exp3 (no atomics) g=4xSM", [&] { k_step_stream_first_touch<4, 3><<<sms * 4, kThreads>>>(P); }, true);
        }
        const bool use_wagg = getenv("UBENCH_WAGG") != nullptr;
        const bool use_merged = getenv("UBENCH_MERGED") != nullptr;
        if (use_merged) Q2.depth = nullptr;
        for (int bps : {3, 4, 6, 8}) {
            float sum = 0, best = 1e30f;
            for (int r = 0; r < reps + 2; ++r) {
                CK(cudaEventRecord(e0));
                CK(cudaMemsetAsync(d_depth, 0, (size_t)cfg.n_segs * 4));
                if (use_merged) k_step_stream_merged<5, kSeenDeferred><<<sms * 10, kThreads, stream_smem_bytes(kSeenDeferred)>>>(S); else if (use_wagg) k_step_stream_warp_agg<8><<<sms * 8, kThreads>>>(P); else run_ft(bps);
                k_uniq_popcount<4><<<pgrid, kPopThreads>>>(Q2);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r >= 2) { best = std::min(best, ms); sum += ms; }
            }
            float avg = sum / reps;
            printf("FT PIPELINE memset+A+B g=%dxSM      best %8.3f ms  avg %8.3f ms  %8.1f Gstep/s  alg %7.1f GB/s (%.1f%% of 6540)\n",
                   bps, best, avg, cfg.n_steps / (avg * 1e6), alg_bytes / (avg * 1e6),
                   100.0 * alg_bytes / (avg * 1e6) / 6540.2);
        }
    }
#if FGFA_DEPTH_PACK != 32
    // packed-counter experiment: memset(packed) + kernel A (packed adds) + kernel B + unpack into a u32 table
    uint32_t* d_depth32;
    unsigned long long* d_total;
    CK(cudaMalloc(&d_depth32, (size_t)cfg.n_segs * 4));
    CK(cudaMalloc(&d_total, 8));
    {
        PopcountParams Qp = Q; Qp.depth = nullptr;
        const uint32_t n_words_p = (cfg.n_segs + 31) / 32;
        const int pgrid = (int)((n_words_p + kPopThreads - 1) / kPopThreads);
        const size_t packed_bytes = ((size_t)cfg.n_segs * FGFA_DEPTH_PACK / 8 + 3) / 4 * 4;
        float sum = 0, best = 1e30f;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaMemsetAsync(d_depth, 0, packed_bytes));
            CK(cudaMemsetAsync(d_total, 0, 8));
            CK(cudaEventRecord(e0));
            CK(cudaMemsetAsync(d_depth32, 0, (size_t)cfg.n_segs * 4));
            k_step_stream_merged<5, kSeenDeferred><<<sms * 10, kThreads, stream_smem_bytes(kSeenDeferred)>>>(S);
            k_uniq_popcount<4><<<pgrid, kPopThreads>>>(Qp);
            k_depth_unpack<FGFA_DEPTH_PACK><<<sms * 8, 256>>>(d_depth, d_depth32, cfg.n_segs, d_total);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) { best = std::min(best, ms); sum += ms; }
        }
        unsigned long long total = 0;
        CK(cudaMemcpy(&total, d_total, 8, cudaMemcpyDeviceToHost));
        printf("PACKED%d PIPELINE memset+A+B+unpack  best %8.3f ms  avg %8.3f ms   decoded total %llu vs %llu steps: %s\n",
               FGFA_DEPTH_PACK, best, sum / reps, total, (unsigned long long)cfg.n_steps,
               total == cfg.n_steps ? "no field overflowed" : "OVERFLOW (fall back to 32-bit counters)");
        CK(cudaMemcpy(d_depth, d_depth32, (size_t)cfg.n_segs * 4, cudaMemcpyDeviceToDevice));   // verified below
    }
#endif
    CK(cudaDeviceSynchronize());
    uint32_t err;
    CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    printf("err flag = %u\n", err);

    if (verify) {
        std::vector<uint32_t> g_depth(cfg.n_segs), g_uniq(cfg.n_segs);
        CK(cudaMemcpy(g_depth.data(), d_depth, (size_t)cfg.n_segs * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_uniq.data(), d_uniq, (size_t)cfg.n_segs * 4, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> spans(2 * cfg.n_paths);
        for (uint32_t p = 0; p < cfg.n_paths; ++p) { spans[2 * p] = ss[p]; spans[2 * p + 1] = se[p]; }
        std::vector<uint64_t> o_depth(cfg.n_segs), o_uniq(cfg.n_segs);
        auto c0 = std::chrono::steady_clock::now();
        int rc = oracle_seg_depth_with_uniq(steps.data(), cfg.n_steps, spans.data(), cfg.n_paths,
                                            cfg.n_segs, o_depth.data(), o_uniq.data());
        auto c1 = std::chrono::steady_clock::now();
        double cs = std::chrono::duration<double>(c1 - c0).count();
        uint64_t bad = 0, sum_d = 0, sum_u = 0, max_d = 0;
        for (uint32_t i = 0; i < cfg.n_segs; ++i) max_d = std::max<uint64_t>(max_d, o_depth[i]);
        printf("max depth of a segment = %llu\n", (unsigned long long)max_d);
        for (uint32_t i = 0; i < cfg.n_segs; ++i) {
            bad += (g_depth[i] != o_depth[i]) + (g_uniq[i] != o_uniq[i]);
            sum_d += o_depth[i]; sum_u += o_uniq[i];
        }
        printf("oracle rc=%d  %.3f s (%.1f Mstep/s, 1 core)  sum_depth=%llu sum_uniq=%llu  mismatches=%llu %s\n",
               rc, cs, cfg.n_steps / cs / 1e6, (unsigned long long)sum_d, (unsigned long long)sum_u,
               (unsigned long long)bad, bad ? "FAIL" : "PARITY OK");
    }
    return 0;
}
