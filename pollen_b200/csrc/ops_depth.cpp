#include "ops_depth.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fgfa_depth.h"
#include "file.hpp"

namespace {

struct DeviceBuffers {
    uint32_t* steps = nullptr;
    uint32_t* depth = nullptr;
    uint32_t* uniq = nullptr;
    uint32_t* h_out = nullptr;   // pinned download staging
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    fgfa_depth_plan_t* plan = nullptr;
    ~DeviceBuffers() {
        if (plan) fgfa_depth_plan_destroy(plan);
        cudaFree(steps);
        cudaFree(depth);
        cudaFree(uniq);
        if (h_out) cudaFreeHost(h_out);
        if (ev[0]) cudaEventDestroy(ev[0]);
        if (ev[1]) cudaEventDestroy(ev[1]);
        if (copy) cudaStreamDestroy(copy);
        if (compute) cudaStreamDestroy(compute);
    }
};

int cuda_rc(cudaError_t e) {
    if (e == cudaSuccess) return FGFA_OK;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return FGFA_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return FGFA_ERR_NOMEM;
    return FGFA_ERR_CUDA;
}
#define CUH(x) do { int rc_ = cuda_rc(x); if (rc_) return rc_; } while (0)

constexpr uint64_t kUploadGroupSteps = 16ull << 20;   // 64 MiB of Handle words per upload

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
    if ((n_segs && !depth_out) || (n_steps && !h_steps)) return FGFA_ERR_INVALID_ARG;
    if (fgfa_device_count() <= 0) return FGFA_ERR_NO_DEVICE;
    DeviceBuffers B;
    int rc = fgfa_depth_plan_create(&B.plan, h_span_start, h_span_end, n_paths, n_segs, n_steps, 0);
    if (rc) return rc;
    const bool want_uniq = uniq_out != nullptr;
    CUH(cudaStreamCreateWithFlags(&B.copy, cudaStreamNonBlocking));
    CUH(cudaStreamCreateWithFlags(&B.compute, cudaStreamNonBlocking));
    CUH(cudaEventCreateWithFlags(&B.ev[0], cudaEventDisableTiming));
    CUH(cudaEventCreateWithFlags(&B.ev[1], cudaEventDisableTiming));
    CUH(cudaMalloc(&B.steps, std::max<size_t>((size_t)n_steps * 4, 16)));
    CUH(cudaMalloc(&B.depth, std::max<size_t>((size_t)n_segs * 4, 4)));
    if (want_uniq) CUH(cudaMalloc(&B.uniq, std::max<size_t>((size_t)n_segs * 4, 4)));
    CUH(cudaMallocHost(&B.h_out, std::max<size_t>((size_t)n_segs * 4 * (want_uniq ? 2 : 1), 4)));

    rc = fgfa_depth_plan_begin(B.plan, B.depth, B.compute);
    if (rc) return rc;

    // Are the spans laid out like the parser leaves them (pool order, disjoint)?  Then
    // uploads and kernels can be pipelined group by group; otherwise upload everything first.
    bool monotone = true;
    for (uint32_t p = 1; p < n_paths && monotone; ++p) monotone = h_span_start[p] >= h_span_end[p - 1];
    if (monotone && n_paths) {
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
            if (b > a)
                CUH(cudaMemcpyAsync(B.steps + a, h_steps + a, (size_t)(b - a) * 4, cudaMemcpyHostToDevice, B.copy));
            CUH(cudaEventRecord(B.ev[slot], B.copy));
            CUH(cudaStreamWaitEvent(B.compute, B.ev[slot], 0));
            rc = fgfa_depth_plan_feed(B.plan, B.steps, lo, hi, B.depth, B.uniq, B.compute);
            if (rc) return rc;
            slot ^= 1;
            lo = hi;
        }
    } else {
        if (n_steps) CUH(cudaMemcpyAsync(B.steps, h_steps, (size_t)n_steps * 4, cudaMemcpyHostToDevice, B.compute));
        rc = fgfa_depth_plan_feed(B.plan, B.steps, 0, n_paths, B.depth, B.uniq, B.compute);
        if (rc) return rc;
    }
    rc = fgfa_depth_plan_finish(B.plan, B.uniq, B.compute);
    if (rc) return rc;
    if (n_segs) {
        CUH(cudaMemcpyAsync(B.h_out, B.depth, (size_t)n_segs * 4, cudaMemcpyDeviceToHost, B.compute));
        if (want_uniq)
            CUH(cudaMemcpyAsync(B.h_out + n_segs, B.uniq, (size_t)n_segs * 4, cudaMemcpyDeviceToHost, B.compute));
    }
    rc = fgfa_depth_plan_status(B.plan, B.compute);   // synchronises
    if (rc) return rc;
    CUH(cudaStreamSynchronize(B.copy));
    widen(B.h_out, depth_out, n_segs);
    if (want_uniq) widen(B.h_out + n_segs, uniq_out, n_segs);
    return FGFA_OK;
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
    if (g.segs.len() > 0x7FFFFFFFull || g.paths.len() > 0xFFFFFFFFull || g.steps.len() > 0xFFFFFFFFull)
        return FGFA_ERR_TOO_LARGE;
    const uint32_t n_paths = (uint32_t)g.paths.len();
    std::vector<uint32_t> s(n_paths), e(n_paths);
    for (uint32_t p = 0; p < n_paths; ++p) {   // flatgfa.rs:99-112: Path.steps
        s[p] = g.paths.data[p].steps.start;
        e[p] = g.paths.data[p].steps.end;
    }
    // The steps pool sits at an arbitrary byte offset of the image; a
    // 4-byte-aligned host pointer is all the upload needs, so realign only if necessary.
    const uint32_t* steps = reinterpret_cast<const uint32_t*>(g.steps.data);
    std::vector<uint32_t> aligned;
    if (reinterpret_cast<uintptr_t>(steps) & 3u) {
        aligned.resize(g.steps.len());
        std::memcpy(aligned.data(), g.steps.data, g.steps.len() * 4);
        steps = aligned.data();
    }
    return fgfa_seg_depth_with_uniq_steps(steps, g.steps.len(), s.data(), e.data(), n_paths,
                                          (uint32_t)g.segs.len(), depth_out, uniq_out);
}

int fgfa_seg_depth_with_uniq(const void* bytes, size_t len, uint64_t* depth_out, uint64_t* uniq_out) {
    if (!uniq_out) return FGFA_ERR_INVALID_ARG;
    return depth_of_image(bytes, len, depth_out, uniq_out);
}

int fgfa_seg_depth(const void* bytes, size_t len, uint64_t* depth_out) {
    return depth_of_image(bytes, len, depth_out, nullptr);
}

}  // extern "C"

namespace flatgfa {
namespace ops {
namespace depth {

namespace {
int run_on(const FlatGFA& gfa, std::vector<uint64_t>& d, std::vector<uint64_t>* u) {
    if (gfa.segs.len() > 0x7FFFFFFFull || gfa.paths.len() > 0xFFFFFFFFull || gfa.steps.len() > 0xFFFFFFFFull)
        return FGFA_ERR_TOO_LARGE;
    const uint32_t n_paths = (uint32_t)gfa.paths.len();
    std::vector<uint32_t> s(n_paths), e(n_paths);
    for (uint32_t p = 0; p < n_paths; ++p) {
        s[p] = gfa.paths.data[p].steps.start;
        e[p] = gfa.paths.data[p].steps.end;
    }
    const uint32_t* steps = reinterpret_cast<const uint32_t*>(gfa.steps.data);
    std::vector<uint32_t> aligned;
    if (reinterpret_cast<uintptr_t>(steps) & 3u) {
        aligned.resize(gfa.steps.len());
        std::memcpy(aligned.data(), gfa.steps.data, gfa.steps.len() * 4);
        steps = aligned.data();
    }
    d.assign(gfa.segs.len(), 0);
    if (u) u->assign(gfa.segs.len(), 0);
    return fgfa_seg_depth_with_uniq_steps(steps, gfa.steps.len(), s.data(), e.data(), n_paths,
                                          (uint32_t)gfa.segs.len(), d.data(), u ? u->data() : nullptr);
}
[[noreturn]] void raise(int rc) {
    std::string m = fgfa_strerror(rc);
    const char* detail = fgfa_last_error();
    if (detail && *detail) m += std::string(": ") + detail;
    throw Error(m);
}
}  // namespace

std::pair<std::vector<uint64_t>, std::vector<uint64_t>> seg_depth_with_uniq(const FlatGFA& gfa) {
    std::vector<uint64_t> d, u;
    int rc = run_on(gfa, d, &u);
    if (rc) raise(rc);
    return {std::move(d), std::move(u)};
}

std::vector<uint64_t> seg_depth(const FlatGFA& gfa) {
    std::vector<uint64_t> d;
    int rc = run_on(gfa, d, nullptr);
    if (rc) raise(rc);
    return d;
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

void SegDepth::emit(std::string& out) const {
    static const char hdr[] = "#node.id\tdepth\tdepth.uniq\n";   // depth.rs:69
    const size_t n = gfa.segs.len();
    const size_t base = out.size();
    out.resize(base + sizeof(hdr) - 1 + n * 54);   // 10 + 20 + 20 digits + 3 separators, worst case
    char* p = &out[base];
    std::memcpy(p, hdr, sizeof(hdr) - 1);
    p += sizeof(hdr) - 1;
    for (size_t i = 0; i < n; ++i) {               // depth.rs:70-78
        p = put_u64(p, (uint32_t)gfa.segs.data[i].name);   // `seg.name as u32`
        *p++ = '\t';
        p = put_u64(p, depths[i]);
        *p++ = '\t';
        p = put_u64(p, uniq_depths[i]);
        *p++ = '\n';
    }
    out.resize((size_t)(p - out.data()));
}

void SegDepth::emit(FILE* f) const {
    std::string s;
    emit(s);
    std::fwrite(s.data(), 1, s.size(), f);
}

}  // namespace depth
}  // namespace ops
}  // namespace flatgfa
