// Microbenchmark + oracle check of the segment-major window kernels (development tool).
//   ubench_win <cfg: S|B|C|E|U|R> [reps] [verify 0/1]
// Pipeline timed: memset(depth) + S1 + S2 + S3 + W + kernel B; every variant is compared with
// the C oracle (PARITY OK / FAIL).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../pollen_b200/csrc/window_kernels.cuh"
#include "experimental_window.cuh"

extern "C" {
int fgfa_synth_spans(uint32_t, uint64_t, uint32_t, uint64_t, uint32_t*, uint32_t*);
int fgfa_synth_steps(int, uint32_t, uint32_t, const uint32_t*, const uint32_t*, uint64_t, uint32_t*, int);
int oracle_seg_depth_with_uniq(const uint32_t*, uint64_t, const uint32_t*, uint32_t, uint32_t, uint64_t*, uint64_t*);
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
    else if (which == "T") cfg = {"T", 100003, 1000, 3000017ull, 0, 50};   // many short paths
    else if (which == "C8") cfg = {"C8", 5000000, 11, 50000000ull, 0, 20};   // one rank's share of C at 8 GPUs
    else if (which == "E8") cfg = {"E8", 5000000, 1, 50000000ull, 1, 0};     // one rank's share of E at 8 GPUs
    else { fprintf(stderr, "unknown cfg\n"); return 2; }

    int n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads > 64) n_threads = 64;
    std::vector<uint32_t> ss(cfg.n_paths), se(cfg.n_paths);
    std::vector<uint32_t> steps(cfg.n_steps);
    fgfa_synth_spans(cfg.n_paths, cfg.n_steps, cfg.jitter, 0xB1011054ull, ss.data(), se.data());
    fgfa_synth_steps(cfg.kind, cfg.n_segs, cfg.n_paths, ss.data(), se.data(), 0xB1011054ull, steps.data(), n_threads);
    printf("cfg %s: n_segs=%u n_paths=%u n_steps=%llu\n", cfg.name, cfg.n_segs, cfg.n_paths, (unsigned long long)cfg.n_steps);

    const uint32_t n_words = (cfg.n_segs + 31) / 32, wpr = (n_words + 31) & ~31u;
    uint32_t *d_steps, *d_ss, *d_se, *d_depth, *d_uniq, *d_bitmap, *d_err;
    CK(cudaMalloc(&d_steps, cfg.n_steps * 4 + 4096));
    CK(cudaMalloc(&d_ss, cfg.n_paths * 4));
    CK(cudaMalloc(&d_se, cfg.n_paths * 4));
    CK(cudaMalloc(&d_depth, (size_t)cfg.n_segs * 4));
    CK(cudaMalloc(&d_uniq, (size_t)cfg.n_segs * 4));
    CK(cudaMalloc(&d_bitmap, (size_t)cfg.n_paths * wpr * 4));
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMemcpy(d_steps, steps.data(), cfg.n_steps * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ss, ss.data(), cfg.n_paths * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_se, se.data(), cfg.n_paths * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_bitmap, 0, (size_t)cfg.n_paths * wpr * 4));
    CK(cudaMemset(d_err, 0, 4));
    unsigned long long* d_stats;
    CK(cudaMalloc(&d_stats, 16 + 16 * 256));

    if (const char* g = getenv("UBENCH_L2_FETCH")) {
        CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g)));
        size_t v = 0;
        CK(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
        printf("cudaLimitMaxL2FetchGranularity = %zu\n", v);
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    cudaEvent_t ev[8];
    for (auto& e : ev) CK(cudaEventCreate(&e));
    const double alg_bytes = 4.0 * cfg.n_steps + 8.0 * cfg.n_paths + 8.0 * cfg.n_segs;

    std::vector<uint64_t> o_depth, o_uniq;
    if (verify) {
        std::vector<uint32_t> spans(2 * cfg.n_paths);
        for (uint32_t p = 0; p < cfg.n_paths; ++p) { spans[2 * p] = ss[p]; spans[2 * p + 1] = se[p]; }
        o_depth.resize(cfg.n_segs); o_uniq.resize(cfg.n_segs);
        oracle_seg_depth_with_uniq(steps.data(), cfg.n_steps, spans.data(), cfg.n_paths, cfg.n_segs, o_depth.data(), o_uniq.data());
    }

    const char* only = getenv("UBENCH_ONLY");
    // seen_kind: 0 = depth only, 1 = shared-memory path masks + kernel B2
    auto run_variant = [&](const char* name, int rows, int seen_kind, auto&& launch_w, uint32_t max_span, uint32_t bin_override = 0) {
        if (only && !strstr(name, only)) return;
        const bool with_seen = seen_kind == 1;
        const uint32_t sub_shift = rows == 8 ? 8 : rows == 16 ? 9 : rows == 4 ? 7 : 10;
        const uint32_t sub = 1u << sub_shift;
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
        B.bin_segs = bin_override ? bin_override : win_bin(with_seen);
        B.n_bins = (cfg.n_segs + B.bin_segs - 1) / B.bin_segs;
        B.n_batches = with_seen ? (cfg.n_paths + 31) / 32 : 1;
        B.n_keys = B.n_bins * B.n_batches;
        B.n_blocks = (n_sub + kBinBlock - 1) / kBinBlock;
        B.max_span = getenv("UBENCH_SPAN") ? (uint32_t)atoll(getenv("UBENCH_SPAN")) : max_span;
        if (B.n_keys + 1 > kMaxKeys) { printf("%s: too many keys (%u)\n", name, B.n_keys); return; }
        uint32_t* d_prefix;
        CK(cudaMalloc(&d_prefix, (cfg.n_paths + 1) * 4));
        CK(cudaMemcpy(d_prefix, prefix.data(), (cfg.n_paths + 1) * 4, cudaMemcpyHostToDevice));
        B.sub_prefix = d_prefix;
        uint32_t *d_keyrank, *d_hist, *d_key_begin, *d_key_total, *d_ticket, *d_masks;
        uint2 *d_entries, *d_entry_tmp;
        const uint64_t pitch = ((uint64_t)cfg.n_segs + 31) & ~31ull;
        CK(cudaMalloc(&d_key_total, (size_t)(B.n_keys + 1) * 4));
        CK(cudaMemset(d_key_total, 0, (size_t)(B.n_keys + 1) * 4));
        CK(cudaMalloc(&d_ticket, 4));
        CK(cudaMemset(d_ticket, 0, 4));
        CK(cudaMalloc(&d_keyrank, (size_t)std::max(n_sub, 1u) * 4));
        CK(cudaMalloc(&d_hist, (size_t)(B.n_keys + 1) * std::max(B.n_blocks, 1u) * 4));
        CK(cudaMalloc(&d_key_begin, (size_t)(B.n_keys + 2) * 4));
        CK(cudaMalloc(&d_entries, (size_t)std::max(n_sub, 1u) * 8));
        CK(cudaMalloc(&d_entry_tmp, (size_t)std::max(n_sub, 1u) * 8));
        CK(cudaMalloc(&d_masks, (size_t)B.n_batches * pitch * 4));
        CK(cudaMemset(d_masks, 0, (size_t)B.n_batches * pitch * 4));
        B.keyrank = d_keyrank; B.hist = d_hist; B.key_begin = d_key_begin; B.entries = d_entries; B.entry_tmp = d_entry_tmp;
        B.key_total = d_key_total; B.ticket = d_ticket;
        WindowParams W{};
        W.steps = d_steps; W.entries = d_entries; W.key_begin = d_key_begin; W.span_s = d_ss; W.span_e = d_se;
        W.n_keys = B.n_keys; W.n_batches = B.n_batches; W.path_lo = 0; W.n_segs = cfg.n_segs; W.plane_pitch = pitch;
        W.depth = d_depth; W.masks = with_seen ? d_masks : nullptr; W.err = d_err; W.stats = d_stats; W.unit = 1; W.zero = 0;
        MaskCountParams Q{};
        Q.masks = d_masks; Q.n_planes = B.n_batches; Q.plane_pitch = pitch; Q.n_segs = cfg.n_segs; Q.uniq = d_uniq;
        Q.accumulate = 0; Q.uniq_bytes = 4;
        const uint32_t qgrid = (cfg.n_segs + 1023) / 1024;
        const uint32_t grid_w = std::min<uint32_t>((uint32_t)sms, std::max(1u, (n_sub + 31) / 32));
        float best = 1e30f, sum = 0, t_pre = 0, t_w = 0, t_b = 0;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaEventRecord(ev[0]));
            CK(cudaMemsetAsync(d_depth, 0, (size_t)cfg.n_segs * 4));
            CK(cudaEventRecord(ev[1]));
            if (n_sub) {
                k_bin_rank<<<B.n_blocks, kBinThreads, (B.n_keys + 1) * 4>>>(B);
                CK(cudaEventRecord(ev[5]));
                k_bin_keyscan<<<1, kScanThreads>>>(B);
                CK(cudaEventRecord(ev[6]));
                k_bin_scatter<<<B.n_blocks, kBinThreads>>>(B);
            }
            CK(cudaEventRecord(ev[2]));
            if (n_sub) launch_w(grid_w, W);
            CK(cudaEventRecord(ev[3]));
            if (with_seen) k_uniq_from_masks<<<qgrid, 256>>>(Q);
            CK(cudaEventRecord(ev[4]));
            CK(cudaEventSynchronize(ev[4]));
            CK(cudaGetLastError());
            float ms, a, b, c;
            CK(cudaEventElapsedTime(&ms, ev[0], ev[4]));
            CK(cudaEventElapsedTime(&a, ev[1], ev[2]));
            CK(cudaEventElapsedTime(&b, ev[2], ev[3]));
            CK(cudaEventElapsedTime(&c, ev[3], ev[4]));
            if (r >= 2) { best = std::min(best, ms); sum += ms; t_pre += a; t_w += b; t_b += c; }
            if (r == reps + 1 && n_sub) {
                float s1, s2, s3;
                CK(cudaEventElapsedTime(&s1, ev[1], ev[5]));
                CK(cudaEventElapsedTime(&s2, ev[5], ev[6]));
                CK(cudaEventElapsedTime(&s3, ev[6], ev[2]));
                printf("  [S1 %.3f  S2 %.3f  S3 %.3f ms]\n", s1, s2, s3);
            }
        }
        const float avg = sum / reps;
        printf("%-28s n_sub=%u keys=%u  total best %.3f avg %.3f ms (pre %.3f  W %.3f  B %.3f)  %.1f Gstep/s  %.1f%% of 6540\n",
               name, n_sub, B.n_keys, best, avg, t_pre / reps, t_w / reps, t_b / reps, cfg.n_steps / (avg * 1e6),
               100.0 * alg_bytes / (avg * 1e6) / 6540.2);
        uint32_t err;
        CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
        if (err) { printf("  err flag = %u\n", err); CK(cudaMemset(d_err, 0, 4)); }
        if (verify) {
            std::vector<uint32_t> g_depth(cfg.n_segs), g_uniq(cfg.n_segs);
            CK(cudaMemcpy(g_depth.data(), d_depth, (size_t)cfg.n_segs * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(g_uniq.data(), d_uniq, (size_t)cfg.n_segs * 4, cudaMemcpyDeviceToHost));
            uint64_t bad_d = 0, bad_u = 0;
            for (uint32_t i = 0; i < cfg.n_segs; ++i) { bad_d += g_depth[i] != o_depth[i]; bad_u += seen_kind != 0 && g_uniq[i] != o_uniq[i]; }
            printf("  mismatches depth=%llu uniq=%llu %s\n", (unsigned long long)bad_d, (unsigned long long)bad_u,
                   (bad_d | bad_u) ? "FAIL" : "PARITY OK");
        }
        cudaFree(d_key_total); cudaFree(d_ticket); cudaFree(d_prefix); cudaFree(d_keyrank); cudaFree(d_hist);
        cudaFree(d_key_begin); cudaFree(d_entries); cudaFree(d_entry_tmp); cudaFree(d_masks);
    };

#define SETUP(K, SM) CK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM)))
#define VARIANT(NAME, ROWS, STAGES, SEEN, STATS, DBG, SPAN)                                                   \
    do {                                                                                                      \
        SETUP((k_window_count<ROWS, STAGES, SEEN, STATS, DBG>), window_smem_bytes(SEEN));                     \
        run_variant(NAME, ROWS, SEEN ? 1 : 0, [&](uint32_t g, WindowParams& W) {                              \
            k_window_count<ROWS, STAGES, SEEN, STATS, DBG><<<g, kWinThreads, window_smem_bytes(SEEN)>>>(W); }, SPAN); \
    } while (0)
#define VARIANT_OVL(NAME, ROWS, SEEN, DBG, SPAN)                                                              \
    do {                                                                                                      \
        SETUP((k_window_count<ROWS, 2, SEEN, false, DBG, true>), window_smem_bytes(SEEN));             \
        run_variant(NAME, ROWS, SEEN ? 1 : 0, [&](uint32_t g, WindowParams& W) {                              \
            k_window_count<ROWS, 2, SEEN, false, DBG, true><<<g, kWinThreads, window_smem_bytes(SEEN)>>>(W); }, SPAN); \
    } while (0)
#define VARIANT_RING(NAME, D, SEEN, DBG, SPAN)                                                                \
    do {                                                                                                      \
        SETUP((k_window_ring<D, SEEN, false, DBG>), ring_smem_bytes<D>(SEEN));                                \
        run_variant(NAME, 8, SEEN ? 1 : 0, [&](uint32_t g, WindowParams& W) {                                 \
            k_window_ring<D, SEEN, false, DBG><<<g, kWinThreads, ring_smem_bytes<D>(SEEN)>>>(W); }, SPAN, ring_win_bin<D>(SEEN)); \
    } while (0)
    const uint32_t ms_def = kWinMaxSpan;
    {   // where the steps go (statistics build, not timed meaningfully)
        unsigned long long st[2];
        CK(cudaMemset(d_stats, 0, 16));
        VARIANT("stats rows=8", 8, 2, true, true, 0, ms_def);
        CK(cudaMemcpy(st, d_stats, 16, cudaMemcpyDeviceToHost));
        if (st[0] + st[1]) printf("  rows=8: in-window %.3f%%  to-L2 %.3f%% of steps\n", 100.0 * st[0] / (double)(st[0] + st[1]), 100.0 * st[1] / (double)(st[0] + st[1]));
    }
    VARIANT("W r8 s2", 8, 2, true, false, 0, ms_def);
    VARIANT_OVL("OVL W r8", 8, true, 0, ms_def);
    VARIANT_OVL("OVL W r8 span=inf", 8, true, 0, 0xFFFFFFFFu);
    VARIANT_OVL("OVL W r8 span=2halo", 8, true, 0, 2 * kWinHalo);
    VARIANT_OVL("OVL depth-only r8", 8, false, 0, ms_def);
    if (only && strstr("OVL DBG7 per-CTA cycles", only)) {
        VARIANT_OVL("OVL DBG7 per-CTA cycles", 8, true, 7, ms_def);
        std::vector<unsigned long long> st(2 + 2 * 256);
        CK(cudaMemcpy(st.data(), d_stats, st.size() * 8, cudaMemcpyDeviceToHost));
        double mn = 1e30, mx = 0, sum = 0, mnb = 1e30, mxb = 0, sumb = 0;
        for (int b = 0; b < sms; ++b) {
            const double tb = (double)st[2 + 2 * b], t = (double)st[3 + 2 * b];
            mn = std::min(mn, t); mx = std::max(mx, t); sum += t; mnb = std::min(mnb, tb); mxb = std::max(mxb, tb); sumb += tb;
        }
        printf("  per-CTA cycles: binned range min %.0f avg %.0f max %.0f | whole kernel min %.0f avg %.0f max %.0f\n", mnb, sumb / sms, mxb, mn, sum / sms, mx);
    }
    VARIANT_OVL("OVL DBG3 loads only", 8, true, 3, ms_def);
    VARIANT_OVL("OVL DBG2 no mask ORs", 8, true, 2, ms_def);
    VARIANT_RING("RING D=1 W", 1, true, 0, ms_def);
    VARIANT_RING("RING D=2 W", 2, true, 0, ms_def);
    VARIANT_RING("RING D=1 depth-only", 1, false, 0, ms_def);
    VARIANT_RING("RING D=2 depth-only", 2, false, 0, ms_def);
    VARIANT_RING("RING D=1 DBG3 loads only", 1, true, 3, ms_def);
    VARIANT_RING("RING D=2 DBG3 loads only", 2, true, 3, ms_def);
    VARIANT_RING("RING D=1 DBG2 no mask ORs", 1, true, 2, ms_def);
    VARIANT_RING("RING D=2 DBG2 no mask ORs", 2, true, 2, ms_def);
    VARIANT("W r8 s3", 8, 3, true, false, 0, ms_def);
    VARIANT("W r8 s4", 8, 4, true, false, 0, ms_def);
    VARIANT("W r16 s2", 16, 2, true, false, 0, ms_def);
    VARIANT("W r4 s4", 4, 4, true, false, 0, ms_def);
    VARIANT("W r8 s3 span=inf", 8, 3, true, false, 0, 0xFFFFFFFFu);
    VARIANT("depth-only r8 s3", 8, 3, false, false, 0, ms_def);
    VARIANT("depth-only r16 s2", 16, 2, false, false, 0, ms_def);
    // measurement-only variants (results are wrong by construction)
    VARIANT("DBG2 no mask ORs", 8, 3, true, false, 2, ms_def);
    VARIANT("DBG3 loads only", 8, 3, true, false, 3, ms_def);
    VARIANT("DBG4 byte store for OR s2", 8, 2, true, false, 4, ms_def);
    VARIANT("DBG2 no mask ORs s2", 8, 2, true, false, 2, ms_def);
    return 0;
}
