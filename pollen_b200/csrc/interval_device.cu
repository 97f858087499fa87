// Device entry points of interval ("window") depth (see interval_kernels.cuh and
// include/fgfa_depth.h).  Replaces the merge loop of the reference's `assign_depths`
// (flatgfa/src/ops/window_depth.rs:118-153) and the position bookkeeping of
// `weighted_depths` / `path_length` (:69-103).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "../../include/fgfa_depth.h"
#include "interval_kernels.cuh"

namespace {

int rc_of(cudaError_t e) {
    if (e == cudaSuccess) return FGFA_OK;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return FGFA_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return FGFA_ERR_NOMEM;
    return FGFA_ERR_CUDA;
}
#define CI(x) do { int rc_ = rc_of(x); if (rc_) return rc_; } while (0)

constexpr size_t kAlign = 256;
size_t round_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }
uint64_t tiles_of(uint64_t n) { return (n + fgfa::kScanTile - 1) / fgfa::kScanTile; }

// scratch layout: [err u32, long_count u32, padded][tile totals u64 x max(tiles(n), tiles(m))][lb/fin u32 x m][long list u32 x m]
struct Scratch {
    uint32_t* err;
    uint32_t* long_count;
    uint64_t* tile_totals;
    uint32_t* fin;
    uint32_t* long_list;
};
Scratch carve(void* base, uint64_t n, uint64_t m) {
    char* p = static_cast<char*>(base);
    Scratch s;
    s.err = reinterpret_cast<uint32_t*>(p);
    s.long_count = s.err + 1;
    p += kAlign;
    s.tile_totals = reinterpret_cast<uint64_t*>(p);
    p += round_up(std::max<uint64_t>(1, std::max(tiles_of(n), tiles_of(m))) * 8);
    s.fin = reinterpret_cast<uint32_t*>(p);
    p += round_up(std::max<uint64_t>(1, m) * 4);
    s.long_list = reinterpret_cast<uint32_t*>(p);
    return s;
}

int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

template <typename Load, typename Op, typename Store>
int scan(Load load, Op op, Store store, uint64_t n, uint64_t* tile_totals, cudaStream_t st) {
    if (n == 0) return FGFA_OK;
    const uint64_t n_tiles = tiles_of(n);
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sm_count() * 8);
    fgfa::k_tile_reduce<<<grid, fgfa::kScanThreads, 0, st>>>(load, op, n, tile_totals);
    CI(cudaGetLastError());
    fgfa::k_scan_tile_totals<<<1, fgfa::kScanThreads, 0, st>>>(op, n_tiles, tile_totals);
    CI(cudaGetLastError());
    fgfa::k_tile_scan<<<grid, fgfa::kScanThreads, 0, st>>>(load, op, store, n, tile_totals);
    CI(cudaGetLastError());
    return FGFA_OK;
}

}  // namespace

extern "C" {

size_t fgfa_interval_scratch_bytes(uint64_t n_path_steps, uint64_t n_intervals) {
    return kAlign + round_up(std::max<uint64_t>(1, std::max(tiles_of(n_path_steps), tiles_of(n_intervals))) * 8) +
           2 * round_up(std::max<uint64_t>(1, n_intervals) * 4);
}

int fgfa_path_offsets_device(const uint32_t* d_path_steps, uint32_t n, const uint32_t* d_seg_len,
                             uint32_t n_segs, uint64_t* d_seg_end, void* d_scratch, size_t scratch_bytes,
                             void* cuda_stream) {
    if (!d_scratch || scratch_bytes < fgfa_interval_scratch_bytes(n, 0)) return FGFA_ERR_INVALID_ARG;
    if (n && (!d_path_steps || !d_seg_len || !d_seg_end)) return FGFA_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(d_path_steps) & 3u) return FGFA_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const Scratch S = carve(d_scratch, n, 0);
    CI(cudaMemsetAsync(S.err, 0, 4, st));
    return scan(fgfa::LoadStepLen{d_path_steps, d_seg_len, n_segs, S.err}, fgfa::OpSum{}, fgfa::StoreU64{d_seg_end},
                n, S.tile_totals, st);
}

int fgfa_make_windows_device(uint64_t start, uint64_t end, uint64_t size, uint64_t n_windows,
                             uint64_t* d_win_start, uint64_t* d_win_end, void* cuda_stream) {
    if (size == 0 || (n_windows && (!d_win_start || !d_win_end))) return FGFA_ERR_INVALID_ARG;
    if (n_windows == 0) return FGFA_OK;
    if ((n_windows + 255) / 256 > 0x7FFFFFFFull) return FGFA_ERR_TOO_LARGE;
    fgfa::k_make_windows<<<(unsigned)((n_windows + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(
        start, end, size, n_windows, d_win_start, d_win_end);
    CI(cudaGetLastError());
    return FGFA_OK;
}

int fgfa_interval_depth_device(const uint32_t* d_path_steps, uint32_t n, const uint32_t* d_depth,
                               const uint32_t* d_seg_len, uint32_t n_segs, const uint64_t* d_seg_end,
                               const uint64_t* d_win_start, const uint64_t* d_win_end, uint64_t m,
                               double* d_out, void* d_scratch, size_t scratch_bytes, void* cuda_stream) {
    if (!d_scratch || scratch_bytes < fgfa_interval_scratch_bytes(n, m)) return FGFA_ERR_INVALID_ARG;
    if (m == 0) return FGFA_OK;
    if (!d_win_start || !d_win_end || !d_out) return FGFA_ERR_INVALID_ARG;
    if (n && (!d_path_steps || !d_depth || !d_seg_len || !d_seg_end)) return FGFA_ERR_INVALID_ARG;
    if (m > 0xFFFFFFFFull) return FGFA_ERR_TOO_LARGE;            // the worklist holds u32 interval indices
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const Scratch S = carve(d_scratch, n, m);
    const unsigned grid = (unsigned)((m + 255) / 256);
    CI(cudaMemsetAsync(S.long_count, 0, 4, st));
    fgfa::k_interval_lower_bound<<<grid, 256, 0, st>>>(d_seg_end, n, d_win_end, m, S.fin);
    CI(cudaGetLastError());
    // the cursor of assign_depths never moves backwards: fin = running maximum of lb
    int rc = scan(fgfa::LoadU32{S.fin}, fgfa::OpMax{}, fgfa::StoreU32{S.fin}, m, S.tile_totals, st);
    if (rc) return rc;
    fgfa::IntervalParams P{};
    P.steps = d_path_steps;
    P.n = n;
    P.depth = d_depth;
    P.seg_len = d_seg_len;
    P.n_segs = n_segs;
    P.seg_end = d_seg_end;
    P.win_start = d_win_start;
    P.win_end = d_win_end;
    P.n_win = m;
    P.fin = S.fin;
    P.out = d_out;
    P.long_list = S.long_list;
    P.long_count = S.long_count;
    fgfa::k_interval_accumulate<<<grid, 256, 0, st>>>(P);
    CI(cudaGetLastError());
    // one warp per long interval; the count stays on the device (an empty list costs one idle launch)
    const unsigned long_grid = (unsigned)std::min<uint64_t>((m + 7) / 8, (uint64_t)sm_count() * 8);
    fgfa::k_interval_accumulate_long<<<long_grid, 256, 0, st>>>(P);
    CI(cudaGetLastError());
    return FGFA_OK;
}

int fgfa_interval_status(const void* d_scratch, void* cuda_stream) {
    if (!d_scratch) return FGFA_ERR_INVALID_ARG;
    uint32_t e = 0;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CI(cudaMemcpyAsync(&e, d_scratch, 4, cudaMemcpyDeviceToHost, st));
    CI(cudaStreamSynchronize(st));
    return e ? FGFA_ERR_SEG_OOB : FGFA_OK;
}

}  // extern "C"
