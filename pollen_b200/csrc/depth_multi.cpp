// Multi-GPU node depth behind the C ABI (include/fgfa_depth.h, fgfa_depth_multi_*).
//
// The reference's host is one compiled process (flatgfa/src/cli/main.rs:57-188 -> cmds.rs:234-245
// -> ops/depth.rs:15-39), so the multi-GPU form of the path is one process driving N devices:
//   * whole paths are partitioned over the devices by step count (LPT; `Path::step_count`,
//     flatgfa/src/flatgfa.rs:114-118): depth is a sum over steps, uniq a sum over paths of
//     indicator vectors (depth.rs:25-35), so both shard exactly when no path is split;
//   * every device runs a depth plan (fgfa_depth_plan_*) over its packed shard of the pool;
//   * the partial [depth | uniq] arrays are combined either by ONE ncclAllReduce(sum) per device
//     inside a group (FGFA_EXCHANGE_NCCL; libnccl is loaded on first use, the library has no link
//     dependency on it) or by kernel X, popcount + reduce-scatter/all-gather over peer-mapped
//     memory (FGFA_EXCHANGE_PEER), ordered between the devices with CUDA events.
// No CPU compute path: without CUDA every entry point returns FGFA_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fgfa_depth.h"

namespace {

// ---- the five NCCL entry points this file needs, resolved at run time -------------------------
struct Nccl {
    void* so = nullptr;
    int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
    int (*CommDestroy)(void* comm) = nullptr;
    int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string* why) {
        if (so) return true;
        // NCCL writes its NCCL_DEBUG output (even the version banner) to stdout unless told otherwise;
        // stdout belongs to the caller (`fgfa depth -d` prints the table there)
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
        // FGFA_NCCL_LIB names the library explicitly.  A process that also hosts PyTorch must let torch
        // load ITS bundled libnccl.so.2 first (same SONAME, newer symbols): the Python binding imports
        // torch before it creates an NCCL handle for that reason.
        const char* forced = std::getenv("FGFA_NCCL_LIB");
        for (const char* name : {forced ? forced : "libnccl.so.2", "libnccl.so.2", "libnccl.so"}) {
            so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (so) break;
        }
        if (!so) { *why = std::string("libnccl not found: ") + dlerror(); return false; }
        auto sym = [&](const char* n) { return dlsym(so, n); };
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd) {
            *why = "libnccl lacks a required symbol";
            return false;
        }
        return true;
    }
};
Nccl g_nccl;
constexpr int kNcclInt32 = 2, kNcclSum = 0;     // nccl.h: ncclInt32, ncclSum (two's complement == u32 addition)

thread_local std::string g_multi_error;
int fail(int code, const std::string& msg) { g_multi_error = msg; return code; }
int cuda_code(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return FGFA_OK;
    g_multi_error = std::string(what) + ": " + cudaGetErrorString(e);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return FGFA_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return FGFA_ERR_NOMEM;
    return FGFA_ERR_CUDA;
}
#define CU(x) do { int rc_ = cuda_code((x), #x); if (rc_) return rc_; } while (0)
#define RC(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

struct Shard {
    int device = 0;
    std::vector<uint32_t> paths;                  // global path ids, ascending
    std::vector<uint32_t> local_start, local_end; // spans inside the packed shard
    uint64_t n_steps = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t counted = nullptr, exchanged = nullptr;
    fgfa_depth_plan_t* plan = nullptr;
    uint32_t* d_steps = nullptr;
    // NCCL form: out = [depth u32 x n_segs | uniq (u8 packed four to a word, or u32) ...]
    // PEER form: partial depth, bitmap rows, final depth, final uniq (u8)
    uint32_t* d_out = nullptr;
    uint32_t* d_partial = nullptr;
    void* d_bitmap = nullptr;
    uint8_t* d_final_uniq = nullptr;
    void* nccl_comm = nullptr;
};

}  // namespace

struct fgfa_depth_multi {
    int n = 0;
    int exchange = FGFA_EXCHANGE_NCCL;
    uint32_t n_paths = 0, n_segs = 0;
    uint64_t n_steps = 0;
    bool compact = false;                         // uniq carried as u8 (<= 255 paths in the graph)
    size_t out_words = 0;                         // u32 words of the NCCL exchange buffer
    std::vector<uint32_t> h_start, h_end;
    std::vector<uint32_t> path_device;
    std::vector<Shard> shards;
    bool resident = false;
    uint32_t* h_pinned = nullptr;                 // download staging
};

namespace {

// Longest-processing-time-first, ties by lower index: the same partition as
// pollen_b200/sharding.py:lpt_partition, so both front ends shard a graph identically.
void lpt(const std::vector<uint32_t>& start, const std::vector<uint32_t>& end, int parts, std::vector<uint32_t>* owner) {
    const uint32_t n = (uint32_t)start.size();
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return (uint64_t)end[a] - start[a] > (uint64_t)end[b] - start[b];
    });
    std::vector<uint64_t> load((size_t)parts, 0);
    owner->assign(n, 0);
    for (uint32_t p : order) {
        int k = 0;
        for (int j = 1; j < parts; ++j)
            if (load[j] < load[k]) k = j;
        (*owner)[p] = (uint32_t)k;
        load[k] += (uint64_t)end[p] - start[p];
    }
}

int set_device(const Shard& s) { CU(cudaSetDevice(s.device)); return FGFA_OK; }

// u32 / u8 device counters -> the u64 (`usize`, depth.rs:17-18) arrays of the ABI, on a few threads
template <typename T>
void widen(const T* src, uint64_t* dst, size_t n) {
    const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (n < (1u << 20) || hw == 1) { for (size_t i = 0; i < n; ++i) dst[i] = src[i]; return; }
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

const char* fgfa_depth_multi_last_error(void) { return g_multi_error.c_str(); }

int fgfa_lpt_partition(const uint32_t* h_span_start, const uint32_t* h_span_end, uint32_t n_paths, int n_parts,
                       uint32_t* path_part) {
    if (n_parts < 1 || (n_paths && (!h_span_start || !h_span_end || !path_part))) return fail(FGFA_ERR_INVALID_ARG, "bad arguments");
    for (uint32_t p = 0; p < n_paths; ++p)
        if (h_span_start[p] > h_span_end[p]) return fail(FGFA_ERR_SPAN_OOB, "span start beyond its end");
    std::vector<uint32_t> s(h_span_start, h_span_start + n_paths), e(h_span_end, h_span_end + n_paths), owner;
    lpt(s, e, n_parts, &owner);
    std::copy(owner.begin(), owner.end(), path_part);
    return FGFA_OK;
}

void fgfa_depth_multi_destroy(fgfa_depth_multi_t* m) {
    if (!m) return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (Shard& s : m->shards) {
        cudaSetDevice(s.device);
        if (s.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s.nccl_comm);
        if (s.plan) fgfa_depth_plan_destroy(s.plan);
        cudaFree(s.d_steps); cudaFree(s.d_out); cudaFree(s.d_partial); cudaFree(s.d_bitmap); cudaFree(s.d_final_uniq);
        if (s.counted) cudaEventDestroy(s.counted);
        if (s.exchanged) cudaEventDestroy(s.exchanged);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (m->h_pinned) cudaFreeHost(m->h_pinned);
    cudaSetDevice(prev);
    delete m;
}

int fgfa_depth_multi_create(fgfa_depth_multi_t** out, const int* devices, int n_devices,
                            const uint32_t* h_span_start, const uint32_t* h_span_end, uint32_t n_paths,
                            uint32_t n_segs, uint64_t n_steps, int exchange) {
    if (!out || !devices || n_devices < 1 || n_devices > 16 || (n_paths && (!h_span_start || !h_span_end)))
        return fail(FGFA_ERR_INVALID_ARG, "bad arguments");
    *out = nullptr;
    if (exchange != FGFA_EXCHANGE_NCCL && exchange != FGFA_EXCHANGE_PEER) return fail(FGFA_ERR_INVALID_ARG, "unknown exchange");
    if (n_steps > 0xFFFFFFFFull || n_segs > 0x7FFFFFFFu) return fail(FGFA_ERR_TOO_LARGE, "counts exceed u32 ids");
    const int visible = fgfa_device_count();
    if (visible <= 0) return fail(FGFA_ERR_NO_DEVICE, "no CUDA device");
    for (int i = 0; i < n_devices; ++i)
        if (devices[i] < 0 || devices[i] >= visible) return fail(FGFA_ERR_INVALID_ARG, "device ordinal out of range");
    for (uint32_t p = 0; p < n_paths; ++p)
        if (h_span_start[p] > h_span_end[p] || (uint64_t)h_span_end[p] > n_steps)
            return fail(FGFA_ERR_SPAN_OOB, "a path's steps span lies outside the pool");
    if (exchange == FGFA_EXCHANGE_PEER && n_paths > 255)
        return fail(FGFA_ERR_INVALID_ARG, "the peer exchange carries u8 uniq counters (<= 255 paths); use FGFA_EXCHANGE_NCCL");
    if (exchange == FGFA_EXCHANGE_NCCL && n_devices > 1) {
        for (int i = 0; i < n_devices; ++i)
            for (int j = 0; j < i; ++j)
                if (devices[i] == devices[j]) return fail(FGFA_ERR_INVALID_ARG, "NCCL needs distinct devices");
        std::string why;
        if (!g_nccl.load(&why)) return fail(FGFA_ERR_INVALID_ARG, why);
    }
    int prev = 0;
    CU(cudaGetDevice(&prev));
    fgfa_depth_multi* m = new fgfa_depth_multi();
    auto bail = [&](int rc) { fgfa_depth_multi_destroy(m); cudaSetDevice(prev); return rc; };
    m->n = n_devices;
    m->exchange = exchange;
    m->n_paths = n_paths;
    m->n_segs = n_segs;
    m->n_steps = n_steps;
    m->compact = n_paths <= 255;
    m->out_words = (size_t)n_segs + (m->compact ? ((size_t)n_segs + 3) / 4 : (size_t)n_segs);
    m->h_start.assign(h_span_start, h_span_start + n_paths);
    m->h_end.assign(h_span_end, h_span_end + n_paths);
    lpt(m->h_start, m->h_end, n_devices, &m->path_device);
    m->shards.resize((size_t)n_devices);
    for (int i = 0; i < n_devices; ++i) m->shards[i].device = devices[i];
    for (uint32_t p = 0; p < n_paths; ++p) m->shards[m->path_device[p]].paths.push_back(p);
    const uint32_t n_words = (n_segs + 31) / 32, words_per_row = (n_words + 31) & ~31u;
#define CUB(x) do { int rc_ = cuda_code((x), #x); if (rc_) return bail(rc_); } while (0)
    for (Shard& s : m->shards) {
        uint64_t acc = 0;
        for (uint32_t p : s.paths) {
            s.local_start.push_back((uint32_t)acc);
            acc += (uint64_t)m->h_end[p] - m->h_start[p];
            s.local_end.push_back((uint32_t)acc);
        }
        s.n_steps = acc;
        CUB(cudaSetDevice(s.device));
        CUB(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CUB(cudaEventCreateWithFlags(&s.counted, cudaEventDisableTiming));
        CUB(cudaEventCreateWithFlags(&s.exchanged, cudaEventDisableTiming));
        CUB(cudaMalloc(&s.d_steps, std::max<size_t>((size_t)acc * 4, 16)));
        int rc = fgfa_depth_plan_create(&s.plan, s.local_start.data(), s.local_end.data(), (uint32_t)s.paths.size(), n_segs, acc, 0);
        if (rc) { g_multi_error = fgfa_last_error(); return bail(rc); }
        CUB(cudaMalloc(&s.d_out, std::max<size_t>(m->out_words * 4, 16)));
        if (exchange == FGFA_EXCHANGE_PEER) {
            const size_t bm = std::max<size_t>((size_t)words_per_row * 4 * std::max<size_t>(s.paths.size(), 1), 128);
            CUB(cudaMalloc(&s.d_partial, std::max<size_t>((size_t)n_segs * 4, 16)));
            CUB(cudaMalloc(&s.d_bitmap, bm));
            CUB(cudaMemset(s.d_bitmap, 0, bm));
            CUB(cudaMalloc(&s.d_final_uniq, std::max<size_t>(n_segs, 16)));
            rc = fgfa_depth_plan_use_bitmap(s.plan, s.d_bitmap, bm);
            if (rc) { g_multi_error = fgfa_last_error(); return bail(rc); }
        } else if (m->compact) {
            rc = fgfa_depth_plan_set_uniq_width(s.plan, 1);
            if (rc) { g_multi_error = fgfa_last_error(); return bail(rc); }
        }
    }
    if (exchange == FGFA_EXCHANGE_PEER) {
        for (Shard& a : m->shards)
            for (Shard& b : m->shards) {
                if (a.device == b.device) continue;
                int can = 0;
                CUB(cudaDeviceCanAccessPeer(&can, a.device, b.device));
                if (!can) return bail(fail(FGFA_ERR_INVALID_ARG, "devices cannot map each other's memory; use FGFA_EXCHANGE_NCCL"));
                CUB(cudaSetDevice(a.device));
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else CUB(e);
            }
    } else if (n_devices > 1) {
        std::vector<void*> comms((size_t)n_devices, nullptr);
        // NCCL prints its version banner on STDOUT at the first communicator creation (whatever
        // NCCL_DEBUG_FILE says); stdout belongs to the caller -- `fgfa depth -d` prints the table there --
        // so it is pointed at stderr for the duration of the call
        std::fflush(stdout);
        const int saved_stdout = dup(1);
        if (saved_stdout >= 0) dup2(2, 1);
        const int nrc = g_nccl.CommInitAll(comms.data(), n_devices, devices);
        if (saved_stdout >= 0) { std::fflush(stdout); dup2(saved_stdout, 1); close(saved_stdout); }
        if (nrc != 0)
            return bail(fail(FGFA_ERR_CUDA, std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error")));
        for (int i = 0; i < n_devices; ++i) m->shards[i].nccl_comm = comms[i];
    }
    CUB(cudaMallocHost(&m->h_pinned, std::max<size_t>((size_t)n_segs * 8, 16)));
#undef CUB
    CU(cudaSetDevice(prev));
    *out = m;
    return FGFA_OK;
}

int fgfa_depth_multi_partition(const fgfa_depth_multi_t* m, uint32_t* path_device, uint64_t* device_steps) {
    if (!m) return fail(FGFA_ERR_INVALID_ARG, "null handle");
    if (path_device) std::copy(m->path_device.begin(), m->path_device.end(), path_device);
    if (device_steps)
        for (int i = 0; i < m->n; ++i) device_steps[i] = m->shards[i].n_steps;
    return FGFA_OK;
}

int fgfa_depth_multi_upload(fgfa_depth_multi_t* m, const uint32_t* h_steps) {
    if (!m || (m->n_steps && !h_steps)) return fail(FGFA_ERR_INVALID_ARG, "null argument");
    int prev = 0;
    CU(cudaGetDevice(&prev));
    for (Shard& s : m->shards) {
        RC(set_device(s));
        for (size_t k = 0; k < s.paths.size(); ++k) {
            const uint32_t p = s.paths[k];
            const size_t len = (size_t)m->h_end[p] - m->h_start[p];
            if (len)
                CU(cudaMemcpyAsync(s.d_steps + s.local_start[k], h_steps + m->h_start[p], len * 4, cudaMemcpyHostToDevice, s.stream));
        }
    }
    CU(cudaSetDevice(prev));
    m->resident = true;
    return FGFA_OK;
}

int fgfa_depth_multi_device_steps(fgfa_depth_multi_t* m, int index, uint32_t** d_steps, uint64_t* n_steps) {
    if (!m || index < 0 || index >= m->n) return fail(FGFA_ERR_INVALID_ARG, "bad shard index");
    if (d_steps) *d_steps = m->shards[index].d_steps;
    if (n_steps) *n_steps = m->shards[index].n_steps;
    m->resident = true;                            // the caller fills the shard itself
    return FGFA_OK;
}

int fgfa_depth_multi_run(fgfa_depth_multi_t* m, int with_uniq) {
    if (!m) return fail(FGFA_ERR_INVALID_ARG, "null handle");
    if (!m->resident) return fail(FGFA_ERR_INVALID_ARG, "no steps resident: call fgfa_depth_multi_upload first");
    if (m->exchange == FGFA_EXCHANGE_PEER && !with_uniq) with_uniq = 1;   // kernel X always produces both
    int prev = 0;
    CU(cudaGetDevice(&prev));
    const uint32_t n_segs = m->n_segs;
    // ---- every device counts its shard ----
    for (Shard& s : m->shards) {
        RC(set_device(s));
        int rc;
        if (m->exchange == FGFA_EXCHANGE_PEER) {
            rc = fgfa_depth_plan_run_stream_only(s.plan, s.d_steps, s.d_partial, s.stream);
        } else {
            void* d_uniq = with_uniq ? static_cast<void*>(s.d_out + n_segs) : nullptr;
            rc = fgfa_depth_plan_run(s.plan, s.d_steps, s.d_out, static_cast<uint32_t*>(d_uniq), s.stream);
        }
        if (rc) { if (g_multi_error.empty()) g_multi_error = fgfa_last_error(); cudaSetDevice(prev); return rc; }
        CU(cudaEventRecord(s.counted, s.stream));
    }
    // ---- exchange ----
    if (m->exchange == FGFA_EXCHANGE_PEER) {
        std::vector<const void*> bitmaps, partials;
        std::vector<void*> fdepth, funiq;
        std::vector<uint32_t> rows;
        for (Shard& s : m->shards) {
            bitmaps.push_back(s.d_bitmap);
            partials.push_back(s.d_partial);
            fdepth.push_back(s.d_out);
            funiq.push_back(s.d_final_uniq);
            rows.push_back((uint32_t)s.paths.size());
        }
        const uint32_t n_words = (n_segs + 31) / 32, words_per_row = (n_words + 31) & ~31u;
        for (int i = 0; i < m->n; ++i) {
            Shard& s = m->shards[i];
            RC(set_device(s));
            for (Shard& o : m->shards) CU(cudaStreamWaitEvent(s.stream, o.counted, 0));      // every partial is complete
            int rc = fgfa_exchange_uniq_depth(m->n, i, bitmaps.data(), rows.data(), partials.data(), fdepth.data(),
                                              funiq.data(), n_segs, nullptr, 0, 0, 0, s.stream);
            if (rc) { g_multi_error = fgfa_last_error(); cudaSetDevice(prev); return rc; }
            CU(cudaEventRecord(s.exchanged, s.stream));
        }
        for (Shard& s : m->shards) {
            RC(set_device(s));
            for (Shard& o : m->shards) CU(cudaStreamWaitEvent(s.stream, o.exchanged, 0));    // every slice has landed, every bitmap was read
            const size_t bm = (size_t)words_per_row * 4 * std::max<size_t>(s.paths.size(), 1);
            CU(cudaMemsetAsync(s.d_bitmap, 0, bm, s.stream));                                // clean seen-bits for the next run
        }
    } else if (m->n > 1) {
        const size_t count = with_uniq ? m->out_words : (size_t)n_segs;
        int nrc = g_nccl.GroupStart();
        for (Shard& s : m->shards)
            if (!nrc) nrc = g_nccl.AllReduce(s.d_out, s.d_out, count, kNcclInt32, kNcclSum, s.nccl_comm, s.stream);
        const int erc = g_nccl.GroupEnd();
        if (!nrc) nrc = erc;
        if (nrc) {
            cudaSetDevice(prev);
            return fail(FGFA_ERR_CUDA, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error"));
        }
    }
    CU(cudaSetDevice(prev));
    return FGFA_OK;
}

int fgfa_depth_multi_sync(fgfa_depth_multi_t* m) {
    if (!m) return fail(FGFA_ERR_INVALID_ARG, "null handle");
    int prev = 0, first = FGFA_OK;
    CU(cudaGetDevice(&prev));
    for (Shard& s : m->shards) {
        RC(set_device(s));
        const int rc = fgfa_depth_plan_status(s.plan, s.stream);     // synchronises the stream
        if (rc && !first) { first = rc; g_multi_error = fgfa_last_error(); }
    }
    CU(cudaSetDevice(prev));
    return first;
}

int fgfa_depth_multi_download(fgfa_depth_multi_t* m, uint64_t* depth_out, uint64_t* uniq_out) {
    if (!m || (m->n_segs && !depth_out)) return fail(FGFA_ERR_INVALID_ARG, "null argument");
    RC(fgfa_depth_multi_sync(m));
    int prev = 0;
    CU(cudaGetDevice(&prev));
    Shard& s = m->shards[0];                          // the result is replicated: take device 0's copy
    RC(set_device(s));
    const uint32_t n = m->n_segs;
    uint32_t* h = m->h_pinned;
    CU(cudaMemcpyAsync(h, s.d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, s.stream));
    const bool u8 = m->exchange == FGFA_EXCHANGE_PEER || m->compact;
    if (uniq_out) {
        const void* src = m->exchange == FGFA_EXCHANGE_PEER ? static_cast<const void*>(s.d_final_uniq)
                                                            : static_cast<const void*>(s.d_out + n);
        CU(cudaMemcpyAsync(h + n, src, u8 ? (size_t)n : (size_t)n * 4, cudaMemcpyDeviceToHost, s.stream));
    }
    CU(cudaStreamSynchronize(s.stream));
    CU(cudaSetDevice(prev));
    widen(h, depth_out, n);
    if (uniq_out) {
        if (u8) widen(reinterpret_cast<const uint8_t*>(h + n), uniq_out, n);
        else widen(h + n, uniq_out, n);
    }
    return FGFA_OK;
}

int fgfa_depth_multi_run_host(fgfa_depth_multi_t* m, const uint32_t* h_steps, uint64_t* depth_out, uint64_t* uniq_out) {
    RC(fgfa_depth_multi_upload(m, h_steps));
    RC(fgfa_depth_multi_run(m, uniq_out != nullptr));
    return fgfa_depth_multi_download(m, depth_out, uniq_out);
}

int fgfa_depth_multi_result_device(fgfa_depth_multi_t* m, int index, const uint32_t** d_depth, const void** d_uniq, int* uniq_bytes) {
    if (!m || index < 0 || index >= m->n) return fail(FGFA_ERR_INVALID_ARG, "bad shard index");
    const Shard& s = m->shards[index];
    if (d_depth) *d_depth = s.d_out;
    if (d_uniq) *d_uniq = m->exchange == FGFA_EXCHANGE_PEER ? static_cast<const void*>(s.d_final_uniq) : static_cast<const void*>(s.d_out + m->n_segs);
    if (uniq_bytes) *uniq_bytes = (m->exchange == FGFA_EXCHANGE_PEER || m->compact) ? 1 : 4;
    return FGFA_OK;
}

}  // extern "C"
