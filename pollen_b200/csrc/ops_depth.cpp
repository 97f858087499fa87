#include "ops_depth.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fgfa_depth.h"
#include "file.hpp"

namespace {

// Process-wide workspace of the host-buffer entry points: device staging for the steps
// pool and the outputs, pinned download staging, two streams, and the last plan.  It is
// kept between calls (a 1.6 GB cudaMalloc + a pinned allocation + plan set-up would
// otherwise cost more than the upload they serve) and released by
// fgfa_release_workspace() or at process exit.  FGFA_WORKSPACE=0 disables the caching.
struct Workspace {
    std::mutex mu;
    int device = -1;
    uint32_t* steps = nullptr;  size_t steps_cap = 0;     // elements
    uint32_t* out = nullptr;    size_t out_cap = 0;       // elements ([depth | uniq])
    uint32_t* h_out = nullptr;  size_t h_out_cap = 0;     // pinned, elements
    void* aux = nullptr;        size_t aux_cap = 0;       // bytes (path-depth scratch / step offsets)
    void* aux2 = nullptr;       size_t aux2_cap = 0;      // bytes (interval depth: intervals, scratch, results)
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // pinned ring for callers whose steps live in pageable memory (an mmapped .flatgfa, the
    // reference's normal case: memfile.rs:7-10): a group is copied into a slot by host threads
    // while the previous group's slot is on its way over PCIe
    uint32_t* ring[2] = {nullptr, nullptr};  size_t ring_cap = 0;   // elements per slot
    cudaEvent_t ring_free[2] = {nullptr, nullptr};
    fgfa_depth_plan_t* plan = nullptr;
    uint64_t plan_key[4] = {0, 0, 0, 0};
    std::vector<uint32_t> plan_start, plan_end;   // the cached plan's span table (exact comparison)
    std::string plan_env;                         // FGFA_ENGINE / FGFA_SEEN_MODE the plan was created under

    void release() {
        if (plan) fgfa_depth_plan_destroy(plan);
        plan = nullptr;
        cudaFree(steps); steps = nullptr; steps_cap = 0;
        cudaFree(out); out = nullptr; out_cap = 0;
        if (h_out) cudaFreeHost(h_out);
        h_out = nullptr; h_out_cap = 0;
        cudaFree(aux); aux = nullptr; aux_cap = 0;
        cudaFree(aux2); aux2 = nullptr; aux2_cap = 0;
        for (auto& e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
        for (auto& e : ring_free) { if (e) cudaEventDestroy(e); e = nullptr; }
        for (auto& r : ring) { if (r) cudaFreeHost(r); r = nullptr; }
        ring_cap = 0;
        if (copy) cudaStreamDestroy(copy);
        if (compute) cudaStreamDestroy(compute);
        copy = compute = nullptr;
        device = -1;
    }
};
Workspace g_ws;

int cuda_rc(cudaError_t e) {
    if (e == cudaSuccess) return FGFA_OK;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return FGFA_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return FGFA_ERR_NOMEM;
    return FGFA_ERR_CUDA;
}
#define CUH(x) do { int rc_ = cuda_rc(x); if (rc_) return rc_; } while (0)

constexpr uint64_t kUploadGroupSteps = 16ull << 20;   // 64 MiB of Handle words per upload

// The workspace's plan for this graph shape: re-used when counts, span tables (compared exactly,
// not by hash) and the engine-selecting environment are unchanged, rebuilt otherwise.
int cached_plan(Workspace& W, const uint32_t* h_span_start, const uint32_t* h_span_end, uint32_t n_paths,
                uint32_t n_segs, uint64_t n_steps) {
    const uint64_t key[4] = {n_paths, n_segs, n_steps, 0};
    std::string env;
    for (const char* name : {"FGFA_ENGINE", "FGFA_SEEN_MODE", "FGFA_BITMAP_BUDGET_MB"}) {
        const char* v = std::getenv(name);
        env += v ? v : "";
        env += '|';
    }
    const bool same = W.plan && std::memcmp(key, W.plan_key, sizeof key) == 0 && env == W.plan_env &&
                      W.plan_start.size() == n_paths &&
                      (n_paths == 0 || (std::memcmp(W.plan_start.data(), h_span_start, (size_t)n_paths * 4) == 0 &&
                                        std::memcmp(W.plan_end.data(), h_span_end, (size_t)n_paths * 4) == 0));
    if (same) return FGFA_OK;
    if (W.plan) fgfa_depth_plan_destroy(W.plan);
    W.plan = nullptr;
    int rc = fgfa_depth_plan_create(&W.plan, h_span_start, h_span_end, n_paths, n_segs, n_steps, 0);
    if (rc) return rc;
    std::memcpy(W.plan_key, key, sizeof key);
    W.plan_start.assign(h_span_start, h_span_start + n_paths);
    W.plan_end.assign(h_span_end, h_span_end + n_paths);
    W.plan_env = env;
    return FGFA_OK;
}

// Engine choice for host steps (the device-resident form is fgfa_depth_plan_autotune): a pool whose
// sub-chunks jump further than a shared-memory window can hold (uniformly random ids) gains nothing
// from the window engine.  4096 evenly spaced 256-step sub-chunks, first against last handle.
bool host_pool_is_scattered(const uint32_t* h_steps, uint64_t n_steps) {
    constexpr uint64_t kSub = 256, kSamples = 4096, kMaxSpan = 2 * 6144;
    if (n_steps < 2 * kSub * kSamples) return false;
    const uint64_t stride = n_steps / kSamples;
    uint64_t far = 0;
    for (uint64_t i = 0; i < kSamples; ++i) {
        const uint64_t a = i * stride;
        const uint32_t h0 = h_steps[a] >> 1, h1 = h_steps[a + kSub - 1] >> 1;
        far += (h0 > h1 ? h0 - h1 : h1 - h0) > kMaxSpan;
    }
    return 2 * far > kSamples;
}

int ensure_workspace(Workspace& w, uint64_t n_steps, uint32_t n_segs, size_t aux_bytes = 0) {
    int dev = 0;
    CUH(cudaGetDevice(&dev));
    if (w.device != dev) {
        w.release();
        CUH(cudaStreamCreateWithFlags(&w.copy, cudaStreamNonBlocking));
        CUH(cudaStreamCreateWithFlags(&w.compute, cudaStreamNonBlocking));
        CUH(cudaEventCreateWithFlags(&w.ev[0], cudaEventDisableTiming));
        CUH(cudaEventCreateWithFlags(&w.ev[1], cudaEventDisableTiming));
        CUH(cudaEventCreateWithFlags(&w.ring_free[0], cudaEventDisableTiming));
        CUH(cudaEventCreateWithFlags(&w.ring_free[1], cudaEventDisableTiming));
        w.device = dev;
    }
    const size_t need_steps = std::max<size_t>((size_t)n_steps, 4);
    if (w.steps_cap < need_steps) {
        cudaFree(w.steps); w.steps = nullptr; w.steps_cap = 0;
        CUH(cudaMalloc(&w.steps, need_steps * 4));
        w.steps_cap = need_steps;
    }
    const size_t need_out = std::max<size_t>((size_t)n_segs * 2, 2);
    if (w.out_cap < need_out) {
        cudaFree(w.out); w.out = nullptr; w.out_cap = 0;
        CUH(cudaMalloc(&w.out, need_out * 4));
        w.out_cap = need_out;
    }
    if (w.h_out_cap < need_out) {
        if (w.h_out) cudaFreeHost(w.h_out);
        w.h_out = nullptr; w.h_out_cap = 0;
        CUH(cudaMallocHost(&w.h_out, need_out * 4));
        w.h_out_cap = need_out;
    }
    if (w.aux_cap < aux_bytes) {
        cudaFree(w.aux); w.aux = nullptr; w.aux_cap = 0;
        CUH(cudaMalloc(&w.aux, aux_bytes));
        w.aux_cap = aux_bytes;
    }
    return FGFA_OK;
}

// Is this host pointer page-locked (cudaMallocHost / cudaHostRegister)?  Pageable memory is staged
// through the workspace's pinned ring instead of being handed to cudaMemcpyAsync directly.
bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void parallel_copy(uint32_t* dst, const uint32_t* src, size_t n) {
    const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (n < (1u << 20) || hw == 1) { std::memcpy(dst, src, n * 4); return; }
    std::vector<std::thread> th;
    const size_t per = (n + hw - 1) / hw;
    for (unsigned t = 0; t < hw; ++t) {
        const size_t a = t * per, b = std::min(n, a + per);
        if (a >= b) break;
        th.emplace_back([=] { std::memcpy(dst + a, src + a, (b - a) * 4); });
    }
    for (auto& t : th) t.join();
}

void widen(const uint32_t* src, uint64_t* dst, size_t n) {
    const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (n < (1u << 20) || hw == 1) {
        for (size_t i = 0; i < n; ++i) dst[i] = src[i];
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (n + hw - 1) / hw;
    for (unsigned t = 0; t < hw; ++t) {
        const size_t a = t * per, b = std::min(n, a + per);
        if (a >= b) break;
        th.emplace_back([=] { for (size_t i = a; i < b; ++i) dst[i] = src[i]; });
    }
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

int fgfa_seg_depth_with_uniq_steps(const uint32_t* h_steps, uint64_t n_steps,
                                   const uint32_t* h_span_start, const uint32_t* h_span_end,
                                   uint32_t n_paths, uint32_t n_segs, uint64_t* depth_out,
                                   uint64_t* uniq_out) {
    if ((n_segs && !depth_out) || (n_steps && !h_steps) || (n_paths && (!h_span_start || !h_span_end)))
        return FGFA_ERR_INVALID_ARG;
    if (fgfa_device_count() <= 0) return FGFA_ERR_NO_DEVICE;
    Workspace& W = g_ws;
    std::lock_guard<std::mutex> lock(W.mu);
    const char* env = std::getenv("FGFA_WORKSPACE");
    const bool keep = !(env && env[0] == '0');
    struct Releaser { Workspace& w; bool on; ~Releaser() { if (on) w.release(); } } releaser{W, !keep};

    const bool want_uniq = uniq_out != nullptr;
    int rc = ensure_workspace(W, n_steps, n_segs);
    if (rc) return rc;
    rc = cached_plan(W, h_span_start, h_span_end, n_paths, n_segs, n_steps);
    if (rc) return rc;
    if (!std::getenv("FGFA_ENGINE") && fgfa_depth_plan_engine(W.plan) == FGFA_ENGINE_WINDOW &&
        host_pool_is_scattered(h_steps, n_steps)) {
        rc = fgfa_depth_plan_set_engine(W.plan, FGFA_ENGINE_STREAM);
        if (rc) return rc;
    }
    uint32_t* d_depth = W.out;
    uint32_t* d_uniq = want_uniq ? W.out + n_segs : nullptr;

    rc = fgfa_depth_plan_begin(W.plan, d_depth, W.compute);
    if (rc) return rc;
    // Are the spans laid out like the parser leaves them (pool order, disjoint)?  Then
    // uploads and kernels are pipelined group by group; otherwise upload everything first.
    bool monotone = true;
    for (uint32_t p = 1; p < n_paths && monotone; ++p) monotone = h_span_start[p] >= h_span_end[p - 1];
    if (monotone && n_paths) {
        const bool pinned = n_steps == 0 || host_pointer_is_pinned(h_steps);
        if (!pinned) {                    // size the ring for the largest upload group
            uint64_t largest = 0;
            for (uint32_t lo = 0; lo < n_paths;) {
                uint32_t hi = lo;
                uint64_t acc = 0;
                while (hi < n_paths && (acc == 0 || acc < kUploadGroupSteps)) { acc += (uint64_t)h_span_end[hi] - h_span_start[hi]; ++hi; }
                largest = std::max<uint64_t>(largest, (uint64_t)h_span_end[hi - 1] - h_span_start[lo]);
                lo = hi;
            }
            if (W.ring_cap < largest) {
                for (auto& r : W.ring) { if (r) cudaFreeHost(r); r = nullptr; }
                W.ring_cap = 0;
                CUH(cudaMallocHost(&W.ring[0], std::max<size_t>((size_t)largest * 4, 16)));
                CUH(cudaMallocHost(&W.ring[1], std::max<size_t>((size_t)largest * 4, 16)));
                W.ring_cap = (size_t)largest;
            }
        }
        uint32_t lo = 0;
        int slot = 0;
        while (lo < n_paths) {
            uint32_t hi = lo;
            uint64_t acc = 0;
            while (hi < n_paths && (acc == 0 || acc < kUploadGroupSteps)) {
                acc += (uint64_t)h_span_end[hi] - h_span_start[hi];
                ++hi;
            }
            const uint64_t a = h_span_start[lo], b = h_span_end[hi - 1];
            if (b > a && pinned) {
                CUH(cudaMemcpyAsync(W.steps + a, h_steps + a, (size_t)(b - a) * 4, cudaMemcpyHostToDevice, W.copy));
            } else if (b > a) {
                CUH(cudaEventSynchronize(W.ring_free[slot]));          // the slot's previous upload has left it
                parallel_copy(W.ring[slot], h_steps + a, (size_t)(b - a));
                CUH(cudaMemcpyAsync(W.steps + a, W.ring[slot], (size_t)(b - a) * 4, cudaMemcpyHostToDevice, W.copy));
                CUH(cudaEventRecord(W.ring_free[slot], W.copy));
            }
            CUH(cudaEventRecord(W.ev[slot], W.copy));
            CUH(cudaStreamWaitEvent(W.compute, W.ev[slot], 0));
            rc = fgfa_depth_plan_feed(W.plan, W.steps, lo, hi, d_depth, d_uniq, W.compute);
            if (rc) return rc;
            slot ^= 1;
            lo = hi;
        }
    } else {
        if (n_steps) CUH(cudaMemcpyAsync(W.steps, h_steps, (size_t)n_steps * 4, cudaMemcpyHostToDevice, W.compute));
        rc = fgfa_depth_plan_feed(W.plan, W.steps, 0, n_paths, d_depth, d_uniq, W.compute);
        if (rc) return rc;
    }
    rc = fgfa_depth_plan_finish(W.plan, d_uniq, W.compute);
    if (rc) return rc;
    if (n_segs)
        CUH(cudaMemcpyAsync(W.h_out, W.out, (size_t)n_segs * 4 * (want_uniq ? 2 : 1), cudaMemcpyDeviceToHost, W.compute));
    rc = fgfa_depth_plan_status(W.plan, W.compute);   // synchronises the compute stream
    if (rc) return rc;
    CUH(cudaStreamSynchronize(W.copy));
    widen(W.h_out, depth_out, n_segs);
    if (want_uniq) widen(W.h_out + n_segs, uniq_out, n_segs);
    return FGFA_OK;
}

int fgfa_path_depth_steps(const uint32_t* h_steps, uint64_t n_steps, const uint32_t* h_span_start,
                          const uint32_t* h_span_end, uint32_t n_paths, const uint32_t* h_seg_len,
                          uint32_t n_segs, const uint32_t* path_ids, uint32_t n_query,
                          uint64_t* length_out, uint64_t* weighted_out, double* mean_out) {
    if ((n_steps && !h_steps) || (n_paths && (!h_span_start || !h_span_end)) || (n_segs && !h_seg_len) ||
        (n_query && (!length_out || !mean_out)))
        return FGFA_ERR_INVALID_ARG;
    if (path_ids)
        for (uint32_t q = 0; q < n_query; ++q)
            if (path_ids[q] >= n_paths) return FGFA_ERR_INVALID_ARG;
    if (!path_ids && n_query != n_paths) return FGFA_ERR_INVALID_ARG;
    if (fgfa_device_count() <= 0) return FGFA_ERR_NO_DEVICE;
    Workspace& W = g_ws;
    std::lock_guard<std::mutex> lock(W.mu);
    const char* env = std::getenv("FGFA_WORKSPACE");
    const bool keep = !(env && env[0] == '0');
    struct Releaser { Workspace& w; bool on; ~Releaser() { if (on) w.release(); } } releaser{W, !keep};

    const size_t sums_bytes = std::max<size_t>((size_t)n_paths * 16, 16);
    const size_t scratch_bytes = std::max<size_t>((size_t)n_segs * 8, 16);
    int rc = ensure_workspace(W, n_steps, n_segs, scratch_bytes + sums_bytes);
    if (rc) return rc;
    rc = cached_plan(W, h_span_start, h_span_end, n_paths, n_segs, n_steps);
    if (rc) return rc;
    uint32_t* d_depth = W.out;
    uint32_t* d_len = W.out + n_segs;
    uint64_t* d_sums = reinterpret_cast<uint64_t*>(static_cast<char*>(W.aux) + scratch_bytes);
    if (n_steps) CUH(cudaMemcpyAsync(W.steps, h_steps, (size_t)n_steps * 4, cudaMemcpyHostToDevice, W.compute));
    if (n_segs) CUH(cudaMemcpyAsync(d_len, h_seg_len, (size_t)n_segs * 4, cudaMemcpyHostToDevice, W.compute));
    rc = fgfa_depth_plan_run(W.plan, W.steps, d_depth, nullptr, W.compute);          // depth.rs:93-99
    if (rc) return rc;
    rc = fgfa_depth_plan_path_sums(W.plan, W.steps, d_depth, d_len, W.aux, d_sums, W.compute);
    if (rc) return rc;
    std::vector<uint64_t> sums((size_t)n_paths * 2);
    if (n_paths) CUH(cudaMemcpyAsync(sums.data(), d_sums, (size_t)n_paths * 16, cudaMemcpyDeviceToHost, W.compute));
    rc = fgfa_depth_plan_status(W.plan, W.compute);
    if (rc) return rc;
    for (uint32_t q = 0; q < n_query; ++q) {
        const uint32_t p = path_ids ? path_ids[q] : q;
        const uint64_t weighted = sums[2 * (size_t)p], length = sums[2 * (size_t)p + 1];
        length_out[q] = length;
        if (weighted_out) weighted_out[q] = weighted;
        mean_out[q] = (double)weighted / (double)length;                              // depth.rs:129
    }
    return FGFA_OK;
}

void fgfa_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_ws.mu);
    g_ws.release();
}

int fgfa_flatgfa_counts(const void* bytes, size_t len, uint64_t* n_segs, uint64_t* n_paths,
                        uint64_t* n_steps) {
    flatgfa::FlatGFA g;
    switch (flatgfa::file::view(static_cast<const uint8_t*>(bytes), len, &g)) {
        case flatgfa::file::kViewOk: break;
        case flatgfa::file::kViewBadMagic: return FGFA_ERR_BAD_MAGIC;
        default: return FGFA_ERR_TRUNCATED;
    }
    if (n_segs) *n_segs = g.segs.len();
    if (n_paths) *n_paths = g.paths.len();
    if (n_steps) *n_steps = g.steps.len();
    return FGFA_OK;
}

static int depth_of_image(const void* bytes, size_t len, uint64_t* depth_out, uint64_t* uniq_out) {
    if (!bytes) return FGFA_ERR_INVALID_ARG;
    flatgfa::FlatGFA g;
    switch (flatgfa::file::view(static_cast<const uint8_t*>(bytes), len, &g)) {
        case flatgfa::file::kViewOk: break;
        case flatgfa::file::kViewBadMagic: return FGFA_ERR_BAD_MAGIC;
        default: return FGFA_ERR_TRUNCATED;
    }
    flatgfa::ops::depth::PoolArrays a;
    const int rc = flatgfa::ops::depth::pool_arrays_of(g, &a);
    if (rc) return rc;
    return fgfa_seg_depth_with_uniq_steps(a.steps, a.n_steps, a.start.data(), a.end.data(), a.n_paths, a.n_segs,
                                          depth_out, uniq_out);
}

int fgfa_seg_depth_with_uniq(const void* bytes, size_t len, uint64_t* depth_out, uint64_t* uniq_out) {
    if (!uniq_out) return FGFA_ERR_INVALID_ARG;
    return depth_of_image(bytes, len, depth_out, uniq_out);
}

int fgfa_seg_depth(const void* bytes, size_t len, uint64_t* depth_out) {
    return depth_of_image(bytes, len, depth_out, nullptr);
}

}  // extern "C"

namespace {
size_t up256(size_t v) { return (v + 255) / 256 * 256; }

// Shared body of fgfa_interval_depth_steps (h_win_* given) and fgfa_window_depth_steps
// (window_size given): window_depth.rs:176-197.
int interval_depth_host(const uint32_t* h_steps, uint64_t n_steps, const uint32_t* h_span_start,
                        const uint32_t* h_span_end, uint32_t n_paths, const uint32_t* h_seg_len,
                        uint32_t n_segs, uint32_t path, const uint64_t* h_win_start, const uint64_t* h_win_end,
                        uint64_t n_intervals, uint64_t window_size, double* depth_out, double** depth_alloc,
                        uint64_t* n_windows_out, uint64_t* path_length_out) {
    const bool uniform = depth_alloc != nullptr;
    if ((n_steps && !h_steps) || !h_span_start || !h_span_end || (n_segs && !h_seg_len) || path >= n_paths)
        return FGFA_ERR_INVALID_ARG;
    if (uniform ? window_size == 0 : (n_intervals && (!h_win_start || !h_win_end || !depth_out)))
        return FGFA_ERR_INVALID_ARG;
    if (fgfa_device_count() <= 0) return FGFA_ERR_NO_DEVICE;
    Workspace& W = g_ws;
    std::lock_guard<std::mutex> lock(W.mu);
    const char* env = std::getenv("FGFA_WORKSPACE");
    const bool keep = !(env && env[0] == '0');
    struct Releaser { Workspace& w; bool on; ~Releaser() { if (on) w.release(); } } releaser{W, !keep};

    if (h_span_start[path] > h_span_end[path] || h_span_end[path] > n_steps) return FGFA_ERR_SPAN_OOB;
    const uint32_t n = h_span_end[path] - h_span_start[path];
    // aux: [seg_end u64 x n][scan scratch for the step offsets]
    const size_t off_bytes = up256(std::max<size_t>((size_t)n * 8, 8));
    const size_t scr1 = fgfa_interval_scratch_bytes(n, 0);
    int rc = ensure_workspace(W, n_steps, n_segs, off_bytes + scr1);
    if (rc) return rc;
    rc = cached_plan(W, h_span_start, h_span_end, n_paths, n_segs, n_steps);
    if (rc) return rc;
    uint32_t* d_depth = W.out;
    uint32_t* d_len = W.out + n_segs;
    uint64_t* d_seg_end = static_cast<uint64_t*>(W.aux);
    void* d_scr1 = static_cast<char*>(W.aux) + off_bytes;
    const uint32_t* d_path = W.steps + h_span_start[path];
    if (n_steps) CUH(cudaMemcpyAsync(W.steps, h_steps, (size_t)n_steps * 4, cudaMemcpyHostToDevice, W.compute));
    if (n_segs) CUH(cudaMemcpyAsync(d_len, h_seg_len, (size_t)n_segs * 4, cudaMemcpyHostToDevice, W.compute));
    rc = fgfa_depth_plan_run(W.plan, W.steps, d_depth, nullptr, W.compute);                 // window_depth.rs:177
    if (rc) return rc;
    rc = fgfa_path_offsets_device(d_path, n, d_len, n_segs, d_seg_end, d_scr1, scr1, W.compute);
    if (rc) return rc;
    uint64_t total = 0;                                                                     // path_length, :69-77
    if (n) CUH(cudaMemcpyAsync(&total, d_seg_end + (n - 1), 8, cudaMemcpyDeviceToHost, W.compute));
    rc = fgfa_depth_plan_status(W.plan, W.compute);    // synchronises; reports OOB segments
    if (rc) return rc;
    if (path_length_out) *path_length_out = total;

    uint64_t m = n_intervals;
    if (uniform) {
        m = total / window_size + (total % window_size ? 1 : 0);                            // Windows::len, :59-61
        if (n_windows_out) *n_windows_out = m;
        *depth_alloc = static_cast<double*>(std::malloc(std::max<size_t>((size_t)m * 8, 8)));
        if (!*depth_alloc) return FGFA_ERR_NOMEM;
        depth_out = *depth_alloc;
    }
    if (m == 0) return FGFA_OK;
    // aux2: [win_start u64 x m][win_end u64 x m][out f64 x m][scratch]
    const size_t col = up256((size_t)m * 8);
    const size_t scr2 = fgfa_interval_scratch_bytes(n, m);
    if (W.aux2_cap < 3 * col + scr2) {
        cudaFree(W.aux2); W.aux2 = nullptr; W.aux2_cap = 0;
        rc = cuda_rc(cudaMalloc(&W.aux2, 3 * col + scr2));
        if (rc) { if (uniform) { std::free(*depth_alloc); *depth_alloc = nullptr; } return rc; }
        W.aux2_cap = 3 * col + scr2;
    }
    uint64_t* d_ws = static_cast<uint64_t*>(W.aux2);
    uint64_t* d_we = reinterpret_cast<uint64_t*>(static_cast<char*>(W.aux2) + col);
    double* d_out = reinterpret_cast<double*>(static_cast<char*>(W.aux2) + 2 * col);
    void* d_scr2 = static_cast<char*>(W.aux2) + 3 * col;
    auto bail = [&](int code) { if (uniform) { std::free(*depth_alloc); *depth_alloc = nullptr; } return code; };
    if (uniform) {
        rc = fgfa_make_windows_device(0, total, window_size, m, d_ws, d_we, W.compute);     // :188-194
        if (rc) return bail(rc);
    } else {
        if ((rc = cuda_rc(cudaMemcpyAsync(d_ws, h_win_start, (size_t)m * 8, cudaMemcpyHostToDevice, W.compute)))) return rc;
        if ((rc = cuda_rc(cudaMemcpyAsync(d_we, h_win_end, (size_t)m * 8, cudaMemcpyHostToDevice, W.compute)))) return rc;
    }
    if ((rc = cuda_rc(cudaMemsetAsync(d_scr2, 0, 4, W.compute)))) return bail(rc);
    rc = fgfa_interval_depth_device(d_path, n, d_depth, d_len, n_segs, d_seg_end, d_ws, d_we, m, d_out, d_scr2,
                                    scr2, W.compute);
    if (rc) return bail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(depth_out, d_out, (size_t)m * 8, cudaMemcpyDeviceToHost, W.compute)))) return bail(rc);
    rc = fgfa_interval_status(d_scr2, W.compute);
    if (rc) return bail(rc);
    return FGFA_OK;
}
}  // namespace

extern "C" {

int fgfa_interval_depth_steps(const uint32_t* h_steps, uint64_t n_steps, const uint32_t* h_span_start,
                              const uint32_t* h_span_end, uint32_t n_paths, const uint32_t* h_seg_len,
                              uint32_t n_segs, uint32_t path, const uint64_t* h_win_start,
                              const uint64_t* h_win_end, uint64_t n_intervals, double* depth_out) {
    return interval_depth_host(h_steps, n_steps, h_span_start, h_span_end, n_paths, h_seg_len, n_segs, path,
                               h_win_start, h_win_end, n_intervals, 0, depth_out, nullptr, nullptr, nullptr);
}

int fgfa_window_depth_steps(const uint32_t* h_steps, uint64_t n_steps, const uint32_t* h_span_start,
                            const uint32_t* h_span_end, uint32_t n_paths, const uint32_t* h_seg_len,
                            uint32_t n_segs, uint32_t path, uint64_t window_size, double** depth_out,
                            uint64_t* n_windows_out, uint64_t* path_length_out) {
    if (!depth_out) return FGFA_ERR_INVALID_ARG;
    *depth_out = nullptr;
    if (n_windows_out) *n_windows_out = 0;
    return interval_depth_host(h_steps, n_steps, h_span_start, h_span_end, n_paths, h_seg_len, n_segs, path,
                               nullptr, nullptr, 0, window_size, nullptr, depth_out, n_windows_out, path_length_out);
}

void fgfa_free(void* p) { std::free(p); }

}  // extern "C"

namespace flatgfa {
namespace ops {
namespace depth {

namespace {
int run_on(const FlatGFA& gfa, std::vector<uint64_t>& d, std::vector<uint64_t>* u) {
    PoolArrays a;
    const int rc = pool_arrays_of(gfa, &a);
    if (rc) return rc;
    d.assign(gfa.segs.len(), 0);
    if (u) u->assign(gfa.segs.len(), 0);
    return fgfa_seg_depth_with_uniq_steps(a.steps, a.n_steps, a.start.data(), a.end.data(), a.n_paths, a.n_segs,
                                          d.data(), u ? u->data() : nullptr);
}
[[noreturn]] void raise(int rc) {
    std::string m = fgfa_strerror(rc);
    const char* detail = fgfa_last_error();
    if (detail && *detail) m += std::string(": ") + detail;
    throw Error(m, rc);
}
}  // namespace

std::pair<std::vector<uint64_t>, std::vector<uint64_t>> seg_depth_with_uniq(const FlatGFA& gfa) {
    std::vector<uint64_t> d, u;
    int rc = run_on(gfa, d, &u);
    if (rc) raise(rc);
    return {std::move(d), std::move(u)};
}

int pool_arrays_of(const FlatGFA& gfa, PoolArrays* a) {
    if (gfa.segs.len() > 0x7FFFFFFFull || gfa.paths.len() > 0xFFFFFFFFull || gfa.steps.len() > 0xFFFFFFFFull)
        return FGFA_ERR_TOO_LARGE;
    a->n_paths = (uint32_t)gfa.paths.len();
    a->n_segs = (uint32_t)gfa.segs.len();
    a->n_steps = gfa.steps.len();
    a->start.resize(a->n_paths);
    a->end.resize(a->n_paths);
    for (uint32_t p = 0; p < a->n_paths; ++p) {   // flatgfa.rs:99-112: Path.steps
        a->start[p] = gfa.paths.data[p].steps.start;
        a->end[p] = gfa.paths.data[p].steps.end;
    }
    a->steps = reinterpret_cast<const uint32_t*>(gfa.steps.data);
    if (reinterpret_cast<uintptr_t>(a->steps) & 3u) {
        a->realigned.resize(gfa.steps.len());
        std::memcpy(a->realigned.data(), gfa.steps.data, gfa.steps.len() * 4);
        a->steps = a->realigned.data();
    }
    return FGFA_OK;
}

std::pair<std::vector<uint64_t>, std::vector<uint64_t>> seg_depth_with_uniq(const FlatGFA& gfa, int n_gpus) {
    n_gpus = std::min(n_gpus, fgfa_device_count());
    if (n_gpus <= 1) return seg_depth_with_uniq(gfa);
    PoolArrays a;
    int rc = pool_arrays_of(gfa, &a);
    if (rc) raise(rc);
    std::vector<int> devices((size_t)n_gpus);
    for (int i = 0; i < n_gpus; ++i) devices[i] = i;
    fgfa_depth_multi_t* m = nullptr;
    rc = fgfa_depth_multi_create(&m, devices.data(), n_gpus, a.start.data(), a.end.data(), a.n_paths, a.n_segs, a.n_steps,
                                 FGFA_EXCHANGE_NCCL);
    if (rc) throw Error(std::string(fgfa_strerror(rc)) + ": " + fgfa_depth_multi_last_error(), rc);
    std::vector<uint64_t> d(a.n_segs), u(a.n_segs);
    rc = fgfa_depth_multi_run_host(m, a.steps, d.data(), u.data());
    const std::string detail = rc ? fgfa_depth_multi_last_error() : "";
    fgfa_depth_multi_destroy(m);
    if (rc) throw Error(std::string(fgfa_strerror(rc)) + ": " + detail, rc);
    return {std::move(d), std::move(u)};
}

std::vector<uint64_t> seg_depth(const FlatGFA& gfa) {
    std::vector<uint64_t> d;
    int rc = run_on(gfa, d, nullptr);
    if (rc) raise(rc);
    return d;
}

std::pair<std::vector<uint64_t>, std::vector<double>> path_depth(const FlatGFA& gfa,
                                                                 const std::vector<uint32_t>& paths) {
    if (gfa.segs.len() > 0x7FFFFFFFull || gfa.paths.len() > 0xFFFFFFFFull || gfa.steps.len() > 0xFFFFFFFFull)
        raise(FGFA_ERR_TOO_LARGE);
    const uint32_t n_paths = (uint32_t)gfa.paths.len(), n_segs = (uint32_t)gfa.segs.len();
    std::vector<uint32_t> s(n_paths), e(n_paths), len(n_segs);
    for (uint32_t p = 0; p < n_paths; ++p) {
        s[p] = gfa.paths.data[p].steps.start;
        e[p] = gfa.paths.data[p].steps.end;
    }
    for (uint32_t i = 0; i < n_segs; ++i) len[i] = (uint32_t)gfa.segs.data[i].len();   // flatgfa.rs:84-89
    const uint32_t* steps = reinterpret_cast<const uint32_t*>(gfa.steps.data);
    std::vector<uint32_t> aligned;
    if (reinterpret_cast<uintptr_t>(steps) & 3u) {
        aligned.resize(gfa.steps.len());
        std::memcpy(aligned.data(), gfa.steps.data, gfa.steps.len() * 4);
        steps = aligned.data();
    }
    std::vector<uint64_t> lengths(paths.size());
    std::vector<double> depths(paths.size());
    int rc = fgfa_path_depth_steps(steps, gfa.steps.len(), s.data(), e.data(), n_paths, len.data(), n_segs,
                                   paths.data(), (uint32_t)paths.size(), lengths.data(), nullptr, depths.data());
    if (rc) raise(rc);
    return {std::move(lengths), std::move(depths)};
}

// depth.rs:192-197: `{:.digits$}` then trim trailing zeroes, then a trailing '.'.
std::string format_float(double x, int digits) {
    if (x != x) return "NaN";                                  // Rust's Display for f64
    if (x == std::numeric_limits<double>::infinity()) return "inf";
    if (x == -std::numeric_limits<double>::infinity()) return "-inf";
    char buf[512];
    std::snprintf(buf, sizeof buf, "%.*f", digits, x);
    std::string s(buf);
    while (!s.empty() && s.back() == '0') s.pop_back();
    while (!s.empty() && s.back() == '.') s.pop_back();
    return s;
}

void PathDepth::emit(std::string& out) const {
    out += "#path\tstart\tend\tmean.depth\n";                              // depth.rs:148
    for (size_t i = 0; i < paths.size(); ++i) {                            // depth.rs:149-157
        const Pool<uint8_t> name = gfa.get_path_name(gfa.paths[paths[i]]);
        out.append(reinterpret_cast<const char*>(name.data), name.len());
        out += "\t0\t";
        out += std::to_string(lengths[i]);
        out += '\t';
        out += format_float(depths[i], 2);
        out += '\n';
    }
}

void PathDepth::emit(FILE* f) const {
    std::string s;
    emit(s);
    std::fwrite(s.data(), 1, s.size(), f);
}

namespace {
inline char* put_u64(char* p, uint64_t v) {
    char tmp[20];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
}  // namespace

namespace {
// Rows [lo, hi) of the table into dst (at least (hi - lo) * kRowMax bytes); returns bytes written.
constexpr size_t kRowMax = 54;   // 10 + 20 + 20 digits + 3 separators, worst case
size_t format_rows(const FlatGFA& gfa, const uint64_t* depths, const uint64_t* uniq, size_t lo, size_t hi, char* dst) {
    char* p = dst;
    for (size_t i = lo; i < hi; ++i) {             // depth.rs:70-78
        p = put_u64(p, (uint32_t)gfa.segs.data[i].name);   // `seg.name as u32`
        *p++ = '\t';
        p = put_u64(p, depths[i]);
        *p++ = '\t';
        p = put_u64(p, uniq[i]);
        *p++ = '\n';
    }
    return (size_t)(p - dst);
}

// The table body as a list of independently formatted blocks, in row order.  Large tables are
// formatted by a few threads (the reference writes row by row through a locked stdout,
// emit.rs:13-18; for a 5 M-segment graph the text is ~70 MB).
struct Block {
    std::unique_ptr<char[]> data;
    size_t len = 0;
};
std::vector<Block> format_table(const FlatGFA& gfa, const uint64_t* depths, const uint64_t* uniq) {
    const size_t n = gfa.segs.len();
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t n_blocks = n < (1u << 18) ? 1 : std::min<size_t>(hw * 4, (n + (1u << 16) - 1) >> 16);
    std::vector<Block> blocks(n_blocks);
    const size_t per = (n + n_blocks - 1) / std::max<size_t>(n_blocks, 1);
    auto work = [&](std::atomic<size_t>& next) {
        for (size_t b; (b = next.fetch_add(1)) < n_blocks;) {
            const size_t lo = std::min(n, b * per), hi = std::min(n, lo + per);
            blocks[b].data.reset(new char[std::max<size_t>((hi - lo) * kRowMax, 1)]);
            blocks[b].len = format_rows(gfa, depths, uniq, lo, hi, blocks[b].data.get());
        }
    };
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < std::min<size_t>(hw, n_blocks); ++t) pool.emplace_back([&] { work(next); });
    work(next);
    for (auto& t : pool) t.join();
    return blocks;
}
const char kSegDepthHeader[] = "#node.id\tdepth\tdepth.uniq\n";   // depth.rs:69
}  // namespace

char* seg_depth_table(const FlatGFA& gfa, const uint64_t* depths, const uint64_t* uniq, size_t* len) {
    const std::vector<Block> blocks = format_table(gfa, depths, uniq);
    size_t total = sizeof(kSegDepthHeader) - 1;
    for (const Block& b : blocks) total += b.len;
    char* buf = static_cast<char*>(std::malloc(total + 1));
    if (!buf) return nullptr;
    char* p = buf;
    std::memcpy(p, kSegDepthHeader, sizeof(kSegDepthHeader) - 1);
    p += sizeof(kSegDepthHeader) - 1;
    for (const Block& b : blocks) { std::memcpy(p, b.data.get(), b.len); p += b.len; }
    *p = 0;
    *len = total;
    return buf;
}

void SegDepth::emit(std::string& out) const {
    if (depths.size() < gfa.segs.len() || uniq_depths.size() < gfa.segs.len()) throw Error("depth table shorter than the segment pool");
    const std::vector<Block> blocks = format_table(gfa, depths.data(), uniq_depths.data());
    size_t total = sizeof(kSegDepthHeader) - 1;
    for (const Block& b : blocks) total += b.len;
    out.reserve(out.size() + total);
    out.append(kSegDepthHeader, sizeof(kSegDepthHeader) - 1);
    for (const Block& b : blocks) out.append(b.data.get(), b.len);
}

void SegDepth::emit(FILE* f) const {
    if (depths.size() < gfa.segs.len() || uniq_depths.size() < gfa.segs.len()) throw Error("depth table shorter than the segment pool");
    const std::vector<Block> blocks = format_table(gfa, depths.data(), uniq_depths.data());
    std::fwrite(kSegDepthHeader, 1, sizeof(kSegDepthHeader) - 1, f);
    for (const Block& b : blocks) std::fwrite(b.data.get(), 1, b.len, f);
}

}  // namespace depth
}  // namespace ops
}  // namespace flatgfa
