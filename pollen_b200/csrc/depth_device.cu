// Device plan + C ABI of the B200 node-depth engine (see include/fgfa_depth.h).
//
// Replaces the loop nest of the reference's seg_depth_with_uniq / seg_depth
// (flatgfa/src/ops/depth.rs:15-56).  Two engines:
//   window  (default for large pools) segment-major: pre-pass S1-S3 bins 256-step sub-chunks by
//           segment window, kernel W counts every window in shared memory (depth counters +
//           32-path masks), kernel B2 popcounts the mask planes          -- window_kernels.cuh
//   stream  path-major: kernel A streams the pool with one L2 reduction per step and per-path
//           seen-bitmap rows, kernel B popcounts the rows                 -- depth_kernels.cuh
//           (small pools, locality-free pools, and the multi-GPU exchange kernel X)
// No CPU compute path exists here: every entry point fails with FGFA_ERR_NO_DEVICE when CUDA
// is unavailable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/fgfa_depth.h"
#include "depth_kernels.cuh"
#include "window_kernels.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return fail(FGFA_ERR_NO_DEVICE, std::string(what) + ": " + cudaGetErrorString(e));
    if (e == cudaErrorMemoryAllocation)
        return fail(FGFA_ERR_NOMEM, std::string(what) + ": " + cudaGetErrorString(e));
    return fail(FGFA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(x)                                        \
    do {                                             \
        cudaError_t e_ = (x);                        \
        if (e_ != cudaSuccess) return cuda_fail(e_, #x); \
    } while (0)

constexpr size_t kDefaultBitmapBudget = 64ull << 20;  // stays resident in the 126 MB L2
// Register cap / grid multiple for kernel A.  Shared memory (2 x 16 KiB staging + 8 KiB of
// parked runs per CTA) limits residency to 5 CTAs per SM; a grid of twice that measured best
// on B200 (the extra CTAs start as the first ones drain, which evens out the tail).
constexpr int kBlocksPerSM = 5;
constexpr int kGridPerSM = 10;
constexpr int kPopMinBlocks = 4;       // kernel B register cap (measured, see profiles/)
// kernel W: 8 rows of 32 steps per sub-chunk, double-buffered in registers with the loads of the next sub-chunk issued
// before the ATOMS burst of the current one (the OVL form, window_kernels.cuh; measured, profiles/r2_*)
constexpr int kWinRows = 8, kWinStages = 2;
template <bool WITH_SEEN>
constexpr auto kWindowKernel = fgfa::k_window_count<kWinRows, kWinStages, WITH_SEEN, false, 0, true>;
// the window engine pays a pre-pass (3 launches) and a per-CTA window set-up: below this many
// steps the stream engine is faster (config B, 20 M steps: 0.05 ms against 0.09 ms)
constexpr uint64_t kWindowMinSteps = 64ull << 20;

}  // namespace

struct fgfa_depth_plan {
    uint32_t n_paths = 0, n_segs = 0;
    uint64_t n_steps = 0;
    std::vector<uint32_t> h_start, h_end;  // host copy of the span table
    int device = 0, sms = 0;
    uint32_t n_words = 0, words_per_row = 0, rows_per_batch = 0;
    uint32_t misalign = 0;                 // d_steps misalignment (elements) the tables are built for
    fgfa::ChunkDesc* d_chunks = nullptr;   // chunk table (spans shifted by `misalign`)
    size_t chunk_capacity = 0;
    std::vector<uint32_t> h_prefix;        // chunks before path p, [n_paths+1]
    uint32_t* d_bitmap = nullptr;          // [rows_per_batch][words_per_row], zero between runs
    bool own_bitmap = true;                // false: caller-provided (symmetric) memory
    uint32_t* d_err = nullptr;
    size_t scratch_bytes = 0;
    int uniq_bytes = 4;                    // width of the uniq counters kernel B writes
    int seen_mode = fgfa::kSeenDirect;     // how kernel A records seen-bits (see depth_kernels.cuh)
    cudaEvent_t probe_before = nullptr, probe_after = nullptr;   // one-shot measurement hook
    uint32_t next_path = 0;                // begin/feed/finish cursor
    bool uniq_started = false;
    bool bitmap_dirty = false;             // seen-bits recorded that no kernel B has consumed yet
    int feed_mode = -1;                    // -1: no feed yet; 0: depth only; 1: depth + uniq (fixed per begin..finish)
    // ---- window engine (window_kernels.cuh) ----
    int engine = 0;                        // 0 = stream (kernels A/B), 1 = window (S1-S3, W, B2)
    bool win_eligible = false;
    size_t budget = 0;                     // seen-scratch budget the plan was created with
    uint32_t sub_shift = 8;                // 256-step sub-chunks
    std::vector<uint32_t> h_sub_prefix;    // sub-chunks before path p, [n_paths+1]
    uint32_t *d_sub_prefix = nullptr, *d_span_s = nullptr, *d_span_e = nullptr;
    uint32_t *d_keyrank = nullptr, *d_hist = nullptr, *d_key_total = nullptr, *d_key_begin = nullptr, *d_ticket = nullptr;
    uint2 *d_entries = nullptr, *d_entry_tmp = nullptr;
    uint32_t* d_masks = nullptr;           // [planes_per_pass][plane_pitch] path-mask planes, zero between runs
    uint64_t plane_pitch = 0;
    uint32_t planes_per_pass = 0;
    uint32_t max_blocks = 0;
    uint32_t bitmap_rows = 1;              // rows the stream engine's bitmap may hold (from the budget)
    uint32_t bitmap_alloc_rows = 0;        // rows currently allocated in d_bitmap (own_bitmap only)
};

namespace {

int build_tables(fgfa_depth_plan* pl, uint32_t misalign) {
    const uint32_t n = pl->n_paths;
    pl->h_prefix.assign((size_t)n + 1, 0u);
    uint64_t chunks = 0;
    for (uint32_t p = 0; p < n; ++p) {
        const uint64_t sp = (uint64_t)pl->h_start[p] + misalign, ep = (uint64_t)pl->h_end[p] + misalign;
        if (ep > 0xFFFFFFFFull) return fail(FGFA_ERR_TOO_LARGE, "steps pool too large for a misaligned device pointer");
        if (ep > sp) chunks += (ep - (sp & ~3ull) + fgfa::kChunk - 1) / fgfa::kChunk;
        if (chunks > 0xFFFFFFFFull) return fail(FGFA_ERR_TOO_LARGE, "too many chunks");
        pl->h_prefix[p + 1] = (uint32_t)chunks;
    }
    std::vector<fgfa::ChunkDesc> table((size_t)chunks);
    for (uint32_t p = 0; p < n; ++p) {
        const uint32_t sp = pl->h_start[p] + misalign, ep = pl->h_end[p] + misalign;
        uint64_t a = sp & ~3u;
        for (uint32_t c = pl->h_prefix[p]; c < pl->h_prefix[p + 1]; ++c, a += fgfa::kChunk)
            table[c] = fgfa::ChunkDesc{(uint32_t)a, sp, ep, p};
    }
    if (chunks > pl->chunk_capacity) {
        cudaFree(pl->d_chunks);
        pl->d_chunks = nullptr;
        CU(cudaMalloc(&pl->d_chunks, (size_t)chunks * sizeof(fgfa::ChunkDesc)));
        pl->chunk_capacity = (size_t)chunks;
    }
    if (chunks)
        CU(cudaMemcpy(pl->d_chunks, table.data(), (size_t)chunks * sizeof(fgfa::ChunkDesc), cudaMemcpyHostToDevice));
    pl->misalign = misalign;
    // ---- window engine tables: sub-chunk prefix + shifted spans ----
    pl->h_sub_prefix.assign((size_t)n + 1, 0u);
    const uint32_t sub = 1u << pl->sub_shift;
    uint64_t subs = 0;
    std::vector<uint32_t> ss(n), se(n);
    for (uint32_t p = 0; p < n; ++p) {
        const uint64_t sp = (uint64_t)pl->h_start[p] + misalign, ep = (uint64_t)pl->h_end[p] + misalign;
        ss[p] = (uint32_t)sp;
        se[p] = (uint32_t)ep;
        if (ep > sp) subs += (ep - (sp & ~31ull) + sub - 1) / sub;
        pl->h_sub_prefix[p + 1] = (uint32_t)std::min<uint64_t>(subs, 0xFFFFFFFFull);
    }
    if (pl->win_eligible && subs < 0xFFFFFFFFull && n) {
        CU(cudaMemcpy(pl->d_sub_prefix, pl->h_sub_prefix.data(), ((size_t)n + 1) * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(pl->d_span_s, ss.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(pl->d_span_e, se.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    }
    return FGFA_OK;
}

// Window engine over paths [lo, hi), which must lie inside one pass (a pass = the paths whose
// mask planes fit the scratch).  pre-pass S1-S3 + kernel W.
int launch_window(fgfa_depth_plan* pl, const uint32_t* d_steps_aligned, uint32_t lo, uint32_t hi,
                  uint32_t* d_depth, bool with_seen, cudaStream_t st) {
    const uint32_t n_sub = pl->h_sub_prefix[hi] - pl->h_sub_prefix[lo];
    if (n_sub == 0) return FGFA_OK;
    const uint32_t pass_start = (lo / pl->rows_per_batch) * pl->rows_per_batch;
    const uint32_t batch_lo = with_seen ? (lo - pass_start) >> 5 : 0u;
    const uint32_t batch_hi = with_seen ? (hi - 1 - pass_start) >> 5 : 0u;
    fgfa::BinParams B{};
    B.steps = d_steps_aligned;
    B.sub_prefix = pl->d_sub_prefix;
    B.span_s = pl->d_span_s;
    B.span_e = pl->d_span_e;
    B.path_lo = lo;
    B.path_hi = hi;
    B.mask_path_lo = pass_start + 32u * batch_lo;
    B.sub_shift = pl->sub_shift;
    B.n_segs = pl->n_segs;
    B.bin_segs = fgfa::win_bin(with_seen);
    B.n_bins = (pl->n_segs + B.bin_segs - 1) / B.bin_segs;
    B.n_batches = batch_hi - batch_lo + 1;
    B.n_keys = B.n_bins * B.n_batches;
    B.n_blocks = (n_sub + fgfa::kBinBlock - 1) / fgfa::kBinBlock;
    B.max_span = fgfa::kWinMaxSpan;
    B.keyrank = pl->d_keyrank;
    B.hist = pl->d_hist;
    B.key_total = pl->d_key_total;
    B.key_begin = pl->d_key_begin;
    B.ticket = pl->d_ticket;
    B.entries = pl->d_entries;
    B.entry_tmp = pl->d_entry_tmp;
    fgfa::k_bin_rank<<<B.n_blocks, fgfa::kBinThreads, (size_t)(B.n_keys + 1) * 4, st>>>(B);
    fgfa::k_bin_keyscan<<<1, fgfa::kScanThreads, 0, st>>>(B);
    fgfa::k_bin_scatter<<<B.n_blocks, fgfa::kBinThreads, 0, st>>>(B);
    CU(cudaGetLastError());
    fgfa::WindowParams W{};
    W.steps = d_steps_aligned;
    W.entries = pl->d_entries;
    W.key_begin = pl->d_key_begin;
    W.span_s = pl->d_span_s;
    W.span_e = pl->d_span_e;
    W.n_keys = B.n_keys;
    W.n_batches = B.n_batches;
    W.path_lo = B.mask_path_lo;
    W.n_segs = pl->n_segs;
    W.plane_pitch = pl->plane_pitch;
    W.unit = 1;
    W.zero = 0;
    W.depth = d_depth;
    W.masks = with_seen ? pl->d_masks + (size_t)batch_lo * pl->plane_pitch : nullptr;
    W.err = pl->d_err;
    W.stats = nullptr;
    const uint32_t grid = std::min<uint32_t>((uint32_t)pl->sms, std::max(1u, (n_sub + 31) / 32));
    if (pl->probe_before) CU(cudaEventRecord(pl->probe_before, st));
    if (with_seen)
        kWindowKernel<true><<<grid, fgfa::kWinThreads, fgfa::window_smem_bytes(true), st>>>(W);
    else
        kWindowKernel<false><<<grid, fgfa::kWinThreads, fgfa::window_smem_bytes(false), st>>>(W);
    CU(cudaGetLastError());
    if (pl->probe_after) CU(cudaEventRecord(pl->probe_after, st));
    pl->probe_before = pl->probe_after = nullptr;
    return FGFA_OK;
}

// kernel B2 over the first `planes` mask planes.
int launch_mask_count(fgfa_depth_plan* pl, uint32_t planes, void* d_uniq, bool accumulate, cudaStream_t st) {
    if (pl->n_segs == 0) return FGFA_OK;
    fgfa::MaskCountParams Q{};
    Q.masks = pl->d_masks;
    Q.n_planes = planes;
    Q.plane_pitch = pl->plane_pitch;
    Q.n_segs = pl->n_segs;
    Q.uniq = d_uniq;
    Q.accumulate = accumulate ? 1 : 0;
    Q.uniq_bytes = pl->uniq_bytes;
    fgfa::k_uniq_from_masks<<<(pl->n_segs + 1023) / 1024, 256, 0, st>>>(Q);
    CU(cudaGetLastError());
    return FGFA_OK;
}

// kernel A over paths [lo, hi), which must lie inside one bitmap batch.
int launch_stream(fgfa_depth_plan* pl, const uint32_t* d_steps_aligned, uint32_t lo, uint32_t hi,
                  uint32_t* d_depth, bool with_seen, cudaStream_t st) {
    const uint32_t chunks = pl->h_prefix[hi] - pl->h_prefix[lo];
    if (chunks == 0) return FGFA_OK;
    fgfa::StreamParams P{};
    P.steps = d_steps_aligned;
    P.chunks = pl->d_chunks;
    P.chunk_lo = pl->h_prefix[lo];
    P.chunk_hi = pl->h_prefix[hi];
    P.path_lo = lo;
    P.n_segs = pl->n_segs;
    P.words_per_row = pl->words_per_row;
    P.depth = d_depth;
    P.bitmap = with_seen ? pl->d_bitmap + (size_t)(lo % pl->rows_per_batch) * pl->words_per_row : nullptr;
    P.err = pl->d_err;
    const uint32_t grid = std::min<uint32_t>(chunks, (uint32_t)pl->sms * kGridPerSM);
    if (pl->probe_before) CU(cudaEventRecord(pl->probe_before, st));
    if (with_seen && pl->seen_mode == fgfa::kSeenWindow)
        fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenWindow>
            <<<grid, fgfa::kThreads, fgfa::stream_smem_bytes(fgfa::kSeenWindow), st>>>(P);
    else if (with_seen && pl->seen_mode == fgfa::kSeenDirect)
        fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenDirect>
            <<<grid, fgfa::kThreads, fgfa::stream_smem_bytes(fgfa::kSeenDirect), st>>>(P);
    else if (with_seen)
        fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenDeferred>
            <<<grid, fgfa::kThreads, fgfa::stream_smem_bytes(fgfa::kSeenDeferred), st>>>(P);
    else
        fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenNone>
            <<<grid, fgfa::kThreads, fgfa::stream_smem_bytes(fgfa::kSeenNone), st>>>(P);
    CU(cudaGetLastError());
    if (pl->probe_after) CU(cudaEventRecord(pl->probe_after, st));
    pl->probe_before = pl->probe_after = nullptr;
    return FGFA_OK;
}

// kernel B over the first `rows` rows of the bitmap scratch.
int launch_popcount(fgfa_depth_plan* pl, uint32_t rows, uint32_t* d_uniq, bool accumulate, cudaStream_t st) {
    if (pl->n_words == 0) return FGFA_OK;
    fgfa::PopcountParams Q{};
    Q.bitmap = pl->d_bitmap;
    Q.n_rows = rows;
    Q.words_per_row = pl->words_per_row;
    Q.n_words = pl->n_words;
    Q.n_segs = pl->n_segs;
    Q.uniq = d_uniq;
    Q.depth = nullptr;
    Q.accumulate = accumulate ? 1 : 0;
    Q.uniq_bytes = pl->uniq_bytes;
    const uint32_t grid = (pl->n_words + fgfa::kPopThreads - 1) / fgfa::kPopThreads;
    fgfa::k_uniq_popcount<kPopMinBlocks><<<grid, fgfa::kPopThreads, 0, st>>>(Q);
    CU(cudaGetLastError());
    return FGFA_OK;
}

// Make `engine` the plan's engine and make sure its seen scratch exists (zeroed).
int select_engine(fgfa_depth_plan* pl, int engine) {
    if (engine == 1) {
        if (!pl->win_eligible) return fail(FGFA_ERR_INVALID_ARG, "the window engine is not available for this plan");
        const size_t bytes = (size_t)pl->plane_pitch * 4 * pl->planes_per_pass;
        if (!pl->d_masks) {
            CU(cudaMalloc(&pl->d_masks, std::max<size_t>(bytes, 4)));
            CU(cudaMemset(pl->d_masks, 0, std::max<size_t>(bytes, 4)));
            pl->scratch_bytes += bytes;
        }
        pl->rows_per_batch = (uint32_t)std::min<uint64_t>(32ull * pl->planes_per_pass, 0x7FFFFFE0u);
    } else if (engine == 0) {
        if (pl->own_bitmap && pl->bitmap_alloc_rows < pl->bitmap_rows) {
            cudaFree(pl->d_bitmap);
            pl->d_bitmap = nullptr;
            const size_t bytes = std::max<size_t>((size_t)pl->words_per_row * 4 * pl->bitmap_rows, 4);
            CU(cudaMalloc(&pl->d_bitmap, bytes));
            CU(cudaMemset(pl->d_bitmap, 0, bytes));
            pl->bitmap_alloc_rows = pl->bitmap_rows;
            pl->scratch_bytes += bytes;
        }
        if (pl->own_bitmap) pl->rows_per_batch = pl->bitmap_rows;
    } else {
        return fail(FGFA_ERR_INVALID_ARG, "engine must be 0 (stream) or 1 (window)");
    }
    pl->engine = engine;
    return FGFA_OK;
}

int prepare_pointer(fgfa_depth_plan* pl, const uint32_t* d_steps, const uint32_t** aligned) {
    const uintptr_t addr = (uintptr_t)d_steps;
    if (addr & 3u) return fail(FGFA_ERR_INVALID_ARG, "d_steps must be 4-byte aligned");
    const uint32_t mis = (uint32_t)((addr >> 2) & 3u);
    if (mis != pl->misalign) {
        // Rare: the pool does not start on a 16-byte boundary.  Shift the span table so
        // that the kernels can keep issuing aligned 128-bit loads from the rounded-down base.
        CU(cudaDeviceSynchronize());
        int rc = build_tables(pl, mis);
        if (rc) return rc;
    }
    *aligned = d_steps - mis;
    return FGFA_OK;
}

}  // namespace

extern "C" {

const char* fgfa_strerror(int code) {
    switch (code) {
        case FGFA_OK: return "ok";
        case FGFA_ERR_INVALID_ARG: return "invalid argument";
        case FGFA_ERR_BAD_MAGIC: return "not a FlatGFA file (bad magic number)";
        case FGFA_ERR_TRUNCATED: return "FlatGFA image is truncated or its table of contents is inconsistent";
        case FGFA_ERR_SPAN_OOB: return "a path's steps span lies outside the steps pool";
        case FGFA_ERR_SEG_OOB: return "a step refers to a segment index outside the segs pool";
        case FGFA_ERR_CUDA: return "CUDA error";
        case FGFA_ERR_NOMEM: return "out of memory";
        case FGFA_ERR_NO_DEVICE: return "no CUDA device available (this library has no CPU path)";
        case FGFA_ERR_TOO_LARGE: return "graph too large for the format's 32-bit ids";
        case FGFA_ERR_PARSE: return "step list outside the strict grammar or unknown segment name";
        default: return "unknown error";
    }
}

const char* fgfa_last_error(void) { return g_last_error.c_str(); }

int fgfa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fgfa_depth_plan_create(fgfa_depth_plan_t** out, const uint32_t* h_span_start,
                           const uint32_t* h_span_end, uint32_t n_paths, uint32_t n_segs,
                           uint64_t n_steps, size_t bitmap_budget_bytes) {
    if (!out || (n_paths && (!h_span_start || !h_span_end))) return fail(FGFA_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n_steps > 0xFFFFFFFFull || n_segs > 0x7FFFFFFFu) return fail(FGFA_ERR_TOO_LARGE, "counts exceed u32 ids");
    for (uint32_t p = 0; p < n_paths; ++p)  // pool.rs:341-347: slicing panics outside the pool
        if (h_span_start[p] > h_span_end[p] || (uint64_t)h_span_end[p] > n_steps)
            return fail(FGFA_ERR_SPAN_OOB, "path " + std::to_string(p) + " has a steps span outside the pool");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    fgfa_depth_plan* pl = new (std::nothrow) fgfa_depth_plan();
    if (!pl) return fail(FGFA_ERR_NOMEM, "plan allocation failed");
    pl->n_paths = n_paths;
    pl->n_segs = n_segs;
    pl->n_steps = n_steps;
    pl->h_start.assign(h_span_start, h_span_start + n_paths);
    pl->h_end.assign(h_span_end, h_span_end + n_paths);
    pl->device = dev;
    auto bail = [&](int rc) { fgfa_depth_plan_destroy(pl); return rc; };
    {
        cudaError_t e = cudaDeviceGetAttribute(&pl->sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaDeviceGetAttribute"));
    }
    pl->n_words = (n_segs + 31) / 32;
    pl->words_per_row = (pl->n_words + 31) & ~31u;  // 128-byte row pitch
    const size_t row_bytes = (size_t)pl->words_per_row * 4;
    size_t budget = bitmap_budget_bytes ? bitmap_budget_bytes : kDefaultBitmapBudget;
    if (!bitmap_budget_bytes)
        if (const char* env = std::getenv("FGFA_BITMAP_BUDGET_MB"))
            if (std::atoll(env) > 0) budget = (size_t)std::atoll(env) << 20;
    uint64_t rows = row_bytes ? std::max<uint64_t>(1, budget / row_bytes) : 1;
    rows = std::min<uint64_t>(rows, std::max<uint32_t>(1u, n_paths));
    pl->rows_per_batch = (uint32_t)rows;   // (select_engine() below has the last word)
    pl->budget = budget;
    pl->bitmap_rows = (uint32_t)rows;
#define CUB_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return bail(cuda_fail(e_, #x)); } while (0)
    CUB_(cudaMalloc(&pl->d_err, 4));
    CUB_(cudaMemset(pl->d_err, 0, 4));
    // ---- window engine: eligible when every launch's key table fits S1's shared memory ----
    int engine = 0;
    {
        const uint32_t bins = (n_segs + fgfa::win_bin(true) - 1) / fgfa::win_bin(true);
        uint64_t subs = 0;
        const uint32_t sub = 1u << pl->sub_shift;
        for (uint32_t p = 0; p < n_paths; ++p)
            if (h_span_end[p] > h_span_start[p]) subs += ((uint64_t)h_span_end[p] + 3 - (h_span_start[p] & ~31u) + sub - 1) / sub;
        pl->plane_pitch = ((uint64_t)n_segs + 31) & ~31ull;
        const uint64_t plane_bytes = pl->plane_pitch * 4;
        uint64_t planes = plane_bytes ? std::max<uint64_t>(1, budget / plane_bytes) : 1;
        planes = std::min<uint64_t>(planes, ((uint64_t)std::max(n_paths, 1u) + 31) / 32);
        if (bins) planes = std::min<uint64_t>(planes, (fgfa::kMaxKeys - 2) / bins);
        pl->win_eligible = n_paths > 0 && n_paths < 0x7FFFFFFFu && n_segs > 0 && bins + 2 <= fgfa::kMaxKeys && planes >= 1 &&
                           subs + n_paths < 0xFFFFFFF0ull;
        engine = pl->win_eligible && n_steps >= kWindowMinSteps ? 1 : 0;
        if (const char* env = std::getenv("FGFA_ENGINE")) {
            if (!std::strcmp(env, "stream")) engine = 0;
            else if (!std::strcmp(env, "window") && pl->win_eligible) engine = 1;
        }
        if (pl->win_eligible) {
            pl->planes_per_pass = (uint32_t)planes;
            pl->max_blocks = (uint32_t)((subs + n_paths + fgfa::kBinBlock - 1) / fgfa::kBinBlock + 1);
            const uint64_t keys = (uint64_t)bins * planes + 2;
            CUB_(cudaMalloc(&pl->d_sub_prefix, ((size_t)n_paths + 1) * 4));
            CUB_(cudaMalloc(&pl->d_span_s, (size_t)n_paths * 4));
            CUB_(cudaMalloc(&pl->d_span_e, (size_t)n_paths * 4));
            CUB_(cudaMalloc(&pl->d_keyrank, (size_t)(subs + n_paths + 1) * 4));
            CUB_(cudaMalloc(&pl->d_entries, (size_t)(subs + n_paths + 1) * 8));
            CUB_(cudaMalloc(&pl->d_entry_tmp, (size_t)(subs + n_paths + 1) * 8));
            CUB_(cudaMalloc(&pl->d_hist, (size_t)keys * pl->max_blocks * 4));
            CUB_(cudaMalloc(&pl->d_key_total, (size_t)keys * 4));
            CUB_(cudaMemset(pl->d_key_total, 0, (size_t)keys * 4));      // S1 accumulates into it, S2 clears it again
            CUB_(cudaMalloc(&pl->d_key_begin, (size_t)(keys + 1) * 4));
            CUB_(cudaMalloc(&pl->d_ticket, 8));
            CUB_(cudaMemset(pl->d_ticket, 0, 8));
            pl->scratch_bytes += (size_t)(subs + n_paths + 1) * 20 + (size_t)keys * pl->max_blocks * 4;
            cudaFuncSetAttribute(kWindowKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fgfa::window_smem_bytes(true));
            cudaFuncSetAttribute(kWindowKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fgfa::window_smem_bytes(false));
            cudaFuncSetAttribute(fgfa::k_bin_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fgfa::kMaxKeys * 4));
        }
    }
    {
        int rc = select_engine(pl, engine);
        if (rc) return bail(rc);
    }
#undef CUB_
    // kernel A wants 6 CTAs x 32 KiB of shared memory per SM: ask for the large carve-out
    cudaFuncSetAttribute(fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenWindow>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenWindow>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fgfa::stream_smem_bytes(fgfa::kSeenWindow));
    cudaFuncSetAttribute(fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenDirect>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenDeferred>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(fgfa::k_step_stream_merged<kBlocksPerSM, fgfa::kSeenNone>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    // Seen-bit strategy.  Default: parked runs issued by ordinal (kSeenDeferred; 0.66 ms on C
    // against 0.72 for the direct form).  Paths that are much longer than the graph has
    // segments must loop (config E: tandem repeats), so their chunks keep hitting the same
    // bitmap sectors and the shared-memory window pays off (1.30 -> 1.09 ms on E, but 0.95 ms
    // on C).  FGFA_SEEN_MODE=direct|deferred|window overrides.
    {
        uint64_t longest = 0;
        for (uint32_t p = 0; p < n_paths; ++p) longest = std::max<uint64_t>(longest, (uint64_t)h_span_end[p] - h_span_start[p]);
        pl->seen_mode = (longest > 4ull * std::max<uint32_t>(n_segs, 1u)) ? fgfa::kSeenWindow : fgfa::kSeenDeferred;
        if (const char* env = std::getenv("FGFA_SEEN_MODE")) {
            if (!std::strcmp(env, "window")) pl->seen_mode = fgfa::kSeenWindow;
            else if (!std::strcmp(env, "direct")) pl->seen_mode = fgfa::kSeenDirect;
            else if (!std::strcmp(env, "deferred")) pl->seen_mode = fgfa::kSeenDeferred;
        }
    }
    int rc = build_tables(pl, 0);
    if (rc) return bail(rc);
    pl->scratch_bytes += pl->chunk_capacity * sizeof(fgfa::ChunkDesc) + 4;
    *out = pl;
    return FGFA_OK;
}

void fgfa_depth_plan_destroy(fgfa_depth_plan_t* pl) {
    if (!pl) return;
    cudaFree(pl->d_chunks);
    if (pl->own_bitmap) cudaFree(pl->d_bitmap);
    cudaFree(pl->d_err);
    cudaFree(pl->d_sub_prefix); cudaFree(pl->d_span_s); cudaFree(pl->d_span_e);
    cudaFree(pl->d_keyrank); cudaFree(pl->d_entries); cudaFree(pl->d_entry_tmp); cudaFree(pl->d_hist);
    cudaFree(pl->d_key_total); cudaFree(pl->d_key_begin); cudaFree(pl->d_ticket);
    cudaFree(pl->d_masks);
    delete pl;
}

int fgfa_depth_plan_begin(fgfa_depth_plan_t* pl, uint32_t* d_depth, void* cuda_stream) {
    if (!pl || (!d_depth && pl->n_segs)) return fail(FGFA_ERR_INVALID_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (pl->n_segs) CU(cudaMemsetAsync(d_depth, 0, (size_t)pl->n_segs * 4, st));  // depth.rs:17 vec![0; n]
    if (pl->bitmap_dirty && pl->engine == 1) {  // an earlier begin..finish run was abandoned half way
        CU(cudaMemsetAsync(pl->d_masks, 0, (size_t)pl->plane_pitch * 4 * pl->planes_per_pass, st));
        pl->bitmap_dirty = false;
    } else if (pl->bitmap_dirty && pl->own_bitmap) {
        CU(cudaMemsetAsync(pl->d_bitmap, 0, (size_t)pl->words_per_row * 4 * pl->rows_per_batch, st));
        pl->bitmap_dirty = false;
    }
    pl->next_path = 0;
    pl->uniq_started = false;
    pl->feed_mode = -1;
    return FGFA_OK;
}

int fgfa_depth_plan_feed(fgfa_depth_plan_t* pl, const uint32_t* d_steps, uint32_t path_lo,
                         uint32_t path_hi, uint32_t* d_depth, uint32_t* d_uniq, void* cuda_stream) {
    if (!pl || (!d_depth && pl->n_segs)) return fail(FGFA_ERR_INVALID_ARG, "null argument");
    if (path_lo != pl->next_path || path_hi < path_lo || path_hi > pl->n_paths)
        return fail(FGFA_ERR_INVALID_ARG, "paths must be fed contiguously and in order");
    if (!d_steps && pl->n_steps) return fail(FGFA_ERR_INVALID_ARG, "d_steps is null");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const uint32_t* base = d_steps;
    if (pl->n_steps) {
        int rc = prepare_pointer(pl, d_steps, &base);
        if (rc) return rc;
    }
    const bool with_seen = d_uniq != nullptr;
    if (pl->feed_mode >= 0 && pl->feed_mode != (int)with_seen)
        return fail(FGFA_ERR_INVALID_ARG, "d_uniq must be given to every feed of a run or to none");
    pl->feed_mode = (int)with_seen;
    uint32_t lo = path_lo;
    while (lo < path_hi) {
        const uint32_t batch_end = std::min<uint64_t>(pl->n_paths, ((uint64_t)lo / pl->rows_per_batch + 1) * pl->rows_per_batch);
        const uint32_t hi = std::min(path_hi, batch_end);
        int rc = pl->engine == 1 ? launch_window(pl, base, lo, hi, d_depth, with_seen, st)
                                 : launch_stream(pl, base, lo, hi, d_depth, with_seen, st);
        if (rc) return rc;
        if (with_seen) pl->bitmap_dirty = true;
        if (with_seen && hi == batch_end) {  // this batch of seen rows / mask planes is complete: fold it into uniq
            const uint32_t batch_start = (lo / pl->rows_per_batch) * pl->rows_per_batch;
            rc = pl->engine == 1 ? launch_mask_count(pl, (batch_end - batch_start + 31) / 32, d_uniq, pl->uniq_started, st)
                                 : launch_popcount(pl, batch_end - batch_start, d_uniq, pl->uniq_started, st);
            if (rc) return rc;
            pl->uniq_started = true;
            pl->bitmap_dirty = false;
        }
        lo = hi;
    }
    pl->next_path = path_hi;
    return FGFA_OK;
}

int fgfa_depth_plan_finish(fgfa_depth_plan_t* pl, uint32_t* d_uniq, void* cuda_stream) {
    if (!pl) return fail(FGFA_ERR_INVALID_ARG, "null plan");
    if (pl->next_path != pl->n_paths) return fail(FGFA_ERR_INVALID_ARG, "not all paths were fed");
    if (pl->feed_mode >= 0 && pl->feed_mode != (int)(d_uniq != nullptr))
        return fail(FGFA_ERR_INVALID_ARG, "d_uniq must match the feeds of this run");
    if (d_uniq && !pl->uniq_started && pl->n_segs)  // no paths at all: uniq is all zero
        CU(cudaMemsetAsync(d_uniq, 0, (size_t)pl->n_segs * pl->uniq_bytes, (cudaStream_t)cuda_stream));
    return FGFA_OK;
}

int fgfa_depth_plan_run(fgfa_depth_plan_t* pl, const uint32_t* d_steps, uint32_t* d_depth,
                        uint32_t* d_uniq, void* cuda_stream) {
    int rc = fgfa_depth_plan_begin(pl, d_depth, cuda_stream);
    if (rc) return rc;
    rc = fgfa_depth_plan_feed(pl, d_steps, 0, pl->n_paths, d_depth, d_uniq, cuda_stream);
    if (rc) return rc;
    return fgfa_depth_plan_finish(pl, d_uniq, cuda_stream);
}

int fgfa_depth_plan_status(fgfa_depth_plan_t* pl, void* cuda_stream) {
    if (!pl) return fail(FGFA_ERR_INVALID_ARG, "null plan");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    uint32_t flag = 0;
    CU(cudaMemcpyAsync(&flag, pl->d_err, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (flag) {
        CU(cudaMemsetAsync(pl->d_err, 0, 4, st));
        CU(cudaStreamSynchronize(st));
        return fail(FGFA_ERR_SEG_OOB, "a step refers to a segment index >= n_segs");
    }
    return FGFA_OK;
}

int fgfa_depth_plan_use_bitmap(fgfa_depth_plan_t* pl, void* d_bitmap, size_t bytes) {
    if (!pl || !d_bitmap) return fail(FGFA_ERR_INVALID_ARG, "null argument");
    const size_t need = (size_t)pl->words_per_row * 4 * std::max<uint32_t>(pl->n_paths, 1u);
    if (bytes < need || ((uintptr_t)d_bitmap & 127u))
        return fail(FGFA_ERR_INVALID_ARG, "external bitmap must be 128-byte aligned and hold one row per path");
    if (pl->own_bitmap) cudaFree(pl->d_bitmap);
    pl->d_bitmap = static_cast<uint32_t*>(d_bitmap);
    pl->own_bitmap = false;
    pl->engine = 0;                                             // seen-bitmap rows are the stream engine's
    pl->rows_per_batch = std::max<uint32_t>(pl->n_paths, 1u);   // one batch: every path has its row
    return FGFA_OK;
}

int fgfa_depth_plan_run_stream_only(fgfa_depth_plan_t* pl, const uint32_t* d_steps, uint32_t* d_depth,
                                    void* cuda_stream) {
    if (!pl || (!d_depth && pl->n_segs)) return fail(FGFA_ERR_INVALID_ARG, "null argument");
    if (pl->engine != 0 || pl->rows_per_batch < pl->n_paths)
        return fail(FGFA_ERR_INVALID_ARG, "needs the stream engine with a bitmap that holds every path (fgfa_depth_plan_use_bitmap)");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (pl->n_segs) CU(cudaMemsetAsync(d_depth, 0, (size_t)pl->n_segs * 4, st));
    if (pl->n_paths == 0 || pl->n_steps == 0) return FGFA_OK;
    const uint32_t* base = d_steps;
    int rc = prepare_pointer(pl, d_steps, &base);
    if (rc) return rc;
    return launch_stream(pl, base, 0, pl->n_paths, d_depth, true, st);
}

int fgfa_exchange_uniq_depth(int n_ranks, int rank, const void* const* bitmaps, const uint32_t* rows,
                             const void* const* partial_depths, void* const* final_depths,
                             void* const* final_uniqs, uint32_t n_segs, void* multicast_base,
                             uint64_t off_partial, uint64_t off_final_depth, uint64_t off_final_uniq,
                             void* cuda_stream) {
    if (n_ranks < 1 || n_ranks > fgfa::kMaxRanks || rank < 0 || rank >= n_ranks || !bitmaps || !rows ||
        !partial_depths || !final_depths || !final_uniqs)
        return fail(FGFA_ERR_INVALID_ARG, "bad exchange arguments");
    uint64_t total_rows = 0;
    for (int q = 0; q < n_ranks; ++q) total_rows += rows[q];
    if (total_rows > 255) return fail(FGFA_ERR_INVALID_ARG, "the fused exchange carries u8 uniq counters (<= 255 paths)");
    const uint32_t n_words = (n_segs + 31) / 32;
    const uint32_t words_per_row = (n_words + 31) & ~31u;
    fgfa::ExchangeParams X{};
    uint32_t k = 0;
    for (int q = 0; q < n_ranks; ++q) {
        for (uint32_t r = 0; r < rows[q]; ++r)
            X.row_ptr[k++] = static_cast<const uint32_t*>(bitmaps[q]) + (size_t)r * words_per_row;
        X.partial_depth[q] = static_cast<const uint32_t*>(partial_depths[q]);
        X.final_depth[q] = static_cast<uint32_t*>(final_depths[q]);
        X.final_uniq[q] = static_cast<uint8_t*>(final_uniqs[q]);
    }
    X.mc_base = static_cast<uint8_t*>(multicast_base);
    X.off_partial = off_partial;
    X.off_final_depth = off_final_depth;
    X.off_final_uniq = off_final_uniq;
    {   // FGFA_MC_REDUCE=1: in-switch depth sum (multimem.ld_reduce); default: peer loads + multicast stores
        const char* env = std::getenv("FGFA_MC_REDUCE");
        X.mc_reduce = (env && env[0] == '1') ? 1 : 0;
    }
    if (multicast_base && ((off_partial | off_final_depth | off_final_uniq) & 15u))
        return fail(FGFA_ERR_INVALID_ARG, "multicast regions must be 16-byte aligned");
    X.n_ranks = n_ranks;
    X.n_rows = k;
    X.n_segs = n_segs;
    const uint32_t per = ((n_words + n_ranks - 1) / n_ranks + 31) & ~31u;   // slices of whole 128-byte lines
    X.w_lo = (uint32_t)std::min<uint64_t>((uint64_t)per * rank, n_words);
    X.w_hi = (uint32_t)std::min<uint64_t>((uint64_t)per * (rank + 1), n_words);
    if (X.w_hi > X.w_lo) {
        const uint32_t words = X.w_hi - X.w_lo;
        X.uniq_blocks = (words + fgfa::kXThreads - 1) / fgfa::kXThreads;
        uint32_t depth_blocks = (words * 8 + fgfa::kXThreads - 1) / fgfa::kXThreads;   // 4 segments per thread
        if (const char* role = std::getenv("FGFA_X_ROLE")) {   // measurement only: run one role of kernel X
            if (!std::strcmp(role, "uniq")) depth_blocks = 0;
            else if (!std::strcmp(role, "depth")) { X.w_lo += 0; X.uniq_blocks = 0; }
        }
        fgfa::k_uniq_exchange<<<X.uniq_blocks + depth_blocks, fgfa::kXThreads, 0, (cudaStream_t)cuda_stream>>>(X);
        CU(cudaGetLastError());
    }
    return FGFA_OK;
}

static uint32_t exchange_per(int n_ranks, uint32_t n_words) {       // bitmap words per slice: whole 128-byte lines
    return ((n_words + n_ranks - 1) / n_ranks + 31) & ~31u;
}

size_t fgfa_exchange_recv_bytes(int n_ranks, uint32_t n_segs) {
    if (n_ranks < 1) return 0;
    const uint32_t n_words = (n_segs + 31) / 32;
    return (size_t)fgfa::push_slot_bytes(std::max(exchange_per(n_ranks, n_words), 32u)) * (size_t)n_ranks;
}

int fgfa_exchange_push(int n_ranks, int rank, void* bitmap, uint32_t rows, const void* partial_depth,
                       const void* partial_uniq_u8, void* const* recv_bufs, uint32_t n_segs, void* cuda_stream) {
    if (n_ranks < 1 || n_ranks > fgfa::kMaxRanks || rank < 0 || rank >= n_ranks || !recv_bufs ||
        (!partial_uniq_u8 && rows && !bitmap) || (n_segs && !partial_depth) || ((uintptr_t)partial_uniq_u8 & 15u))
        return fail(FGFA_ERR_INVALID_ARG, "bad exchange arguments");
    if (rows > 255) return fail(FGFA_ERR_INVALID_ARG, "the fused exchange carries u8 uniq counters (<= 255 paths)");
    const uint32_t n_words = (n_segs + 31) / 32;
    if (n_words == 0) return FGFA_OK;
    fgfa::PushParams P{};
    P.bitmap = static_cast<uint32_t*>(bitmap);
    P.partial_depth = static_cast<const uint32_t*>(partial_depth);
    P.partial_uniq = static_cast<const uint8_t*>(partial_uniq_u8);
    for (int q = 0; q < n_ranks; ++q) {
        if (!recv_bufs[q] || ((uintptr_t)recv_bufs[q] & 15u)) return fail(FGFA_ERR_INVALID_ARG, "receive buffers must be 16-byte aligned");
        P.recv[q] = static_cast<uint8_t*>(recv_bufs[q]);
    }
    P.n_ranks = n_ranks;
    P.rank = rank;
    P.n_rows = rows;
    P.words_per_row = (n_words + 31) & ~31u;
    P.n_words = n_words;
    P.n_segs = n_segs;
    P.per = std::max(exchange_per(n_ranks, n_words), 32u);
    P.uniq_blocks = (n_words + fgfa::kXThreads - 1) / fgfa::kXThreads;
    const uint32_t depth_blocks = (n_words * 8 + fgfa::kXThreads - 1) / fgfa::kXThreads;
    fgfa::k_push_partials<<<P.uniq_blocks + depth_blocks, fgfa::kXThreads, 0, (cudaStream_t)cuda_stream>>>(P);
    CU(cudaGetLastError());
    return FGFA_OK;
}

int fgfa_exchange_reduce(int n_ranks, int rank, const void* recv_buf, void* const* final_depths,
                         void* const* final_uniqs, uint32_t n_segs, void* multicast_base,
                         uint64_t off_final_depth, uint64_t off_final_uniq, void* cuda_stream) {
    if (n_ranks < 1 || n_ranks > fgfa::kMaxRanks || rank < 0 || rank >= n_ranks || !recv_buf || !final_depths || !final_uniqs)
        return fail(FGFA_ERR_INVALID_ARG, "bad exchange arguments");
    if (multicast_base && ((off_final_depth | off_final_uniq) & 15u))
        return fail(FGFA_ERR_INVALID_ARG, "multicast regions must be 16-byte aligned");
    const uint32_t n_words = (n_segs + 31) / 32;
    if (n_words == 0) return FGFA_OK;
    fgfa::ReduceParams R{};
    R.recv = static_cast<const uint8_t*>(recv_buf);
    for (int q = 0; q < n_ranks; ++q) {
        R.final_depth[q] = static_cast<uint32_t*>(final_depths[q]);
        R.final_uniq[q] = static_cast<uint8_t*>(final_uniqs[q]);
    }
    R.n_ranks = n_ranks;
    R.rank = rank;
    R.n_words = n_words;
    R.n_segs = n_segs;
    R.per = std::max(exchange_per(n_ranks, n_words), 32u);
    R.mc_base = static_cast<uint8_t*>(multicast_base);
    R.off_final_depth = off_final_depth;
    R.off_final_uniq = off_final_uniq;
    const uint32_t w_lo = std::min<uint64_t>((uint64_t)R.per * rank, n_words), w_hi = std::min<uint64_t>((uint64_t)w_lo + R.per, n_words);
    if (w_hi > w_lo) {
        const uint64_t slice_segs = (uint64_t)(w_hi - w_lo) * 32;
        const uint64_t threads = slice_segs / 4 + slice_segs / 16;
        fgfa::k_reduce_slices<<<(uint32_t)((threads + fgfa::kXThreads - 1) / fgfa::kXThreads), fgfa::kXThreads, 0, (cudaStream_t)cuda_stream>>>(R);
        CU(cudaGetLastError());
    }
    return FGFA_OK;
}

int fgfa_depth_plan_set_engine(fgfa_depth_plan_t* pl, int engine) {
    if (!pl) return fail(FGFA_ERR_INVALID_ARG, "null plan");
    if (!pl->own_bitmap && engine != 0) return fail(FGFA_ERR_INVALID_ARG, "a plan with an external bitmap runs the stream engine");
    if (pl->next_path != 0 && pl->next_path != pl->n_paths) return fail(FGFA_ERR_INVALID_ARG, "a begin..finish run is in progress");
    CU(cudaDeviceSynchronize());
    return select_engine(pl, engine);
}

int fgfa_depth_plan_engine(const fgfa_depth_plan_t* pl) { return pl ? pl->engine : -1; }

int fgfa_depth_plan_autotune(fgfa_depth_plan_t* pl, const uint32_t* d_steps, void* cuda_stream) {
    if (!pl) return fail(FGFA_ERR_INVALID_ARG, "null plan");
    if (!pl->win_eligible || !pl->own_bitmap || pl->n_steps == 0) return FGFA_OK;
    if (!d_steps) return fail(FGFA_ERR_INVALID_ARG, "d_steps is null");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const uint32_t* base = d_steps;
    int rc = prepare_pointer(pl, d_steps, &base);
    if (rc) return rc;
    const uint32_t n_sub = pl->h_sub_prefix[pl->n_paths];
    if (n_sub == 0) return FGFA_OK;
    const uint32_t samples = std::min<uint32_t>(n_sub, 16384u);
    fgfa::BinParams B{};
    B.steps = base;
    B.sub_prefix = pl->d_sub_prefix;
    B.span_s = pl->d_span_s;
    B.span_e = pl->d_span_e;
    B.path_lo = 0;
    B.path_hi = pl->n_paths;
    B.sub_shift = pl->sub_shift;
    B.n_segs = pl->n_segs;
    B.max_span = fgfa::kWinMaxSpan;
    B.ticket = pl->d_ticket;
    CU(cudaMemsetAsync(pl->d_ticket, 0, 8, st));
    fgfa::k_sample_spans<<<(samples + 255) / 256, 256, 0, st>>>(B, samples, n_sub / samples);
    CU(cudaGetLastError());
    uint32_t h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, pl->d_ticket, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemsetAsync(pl->d_ticket, 0, 8, st));
    CU(cudaStreamSynchronize(st));
    // h[1] of h[0] sampled sub-chunks jump further than a window can hold: such pools gain nothing
    // from shared-memory windows (config U), and the stream engine handles them better.
    const bool scattered = h[0] > 0 && 2ull * h[1] > h[0];
    const int want = (!scattered && pl->n_steps >= kWindowMinSteps) ? 1 : 0;
    const char* env = std::getenv("FGFA_ENGINE");
    if (env && (!std::strcmp(env, "stream") || !std::strcmp(env, "window"))) return FGFA_OK;   // forced
    if (want == pl->engine) return FGFA_OK;
    return fgfa_depth_plan_set_engine(pl, want);
}

int fgfa_depth_plan_set_uniq_width(fgfa_depth_plan_t* pl, int bytes) {
    if (!pl || (bytes != 1 && bytes != 4)) return fail(FGFA_ERR_INVALID_ARG, "uniq width must be 1 or 4 bytes");
    if (bytes == 1 && pl->n_paths > 255) return fail(FGFA_ERR_INVALID_ARG, "u8 uniq counters need <= 255 paths");
    pl->uniq_bytes = bytes;
    return FGFA_OK;
}

int fgfa_depth_plan_set_probe(fgfa_depth_plan_t* pl, void* before, void* after) {
    if (!pl) return fail(FGFA_ERR_INVALID_ARG, "null plan");
    pl->probe_before = (cudaEvent_t)before;
    pl->probe_after = (cudaEvent_t)after;
    return FGFA_OK;
}

int fgfa_depth_plan_path_sums(fgfa_depth_plan_t* pl, const uint32_t* d_steps, const uint32_t* d_depth,
                              const uint32_t* d_seg_len, void* d_scratch, uint64_t* d_sums,
                              void* cuda_stream) {
    if (!pl || !d_sums || (pl->n_segs && (!d_depth || !d_seg_len || !d_scratch)))
        return fail(FGFA_ERR_INVALID_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CU(cudaMemsetAsync(d_sums, 0, (size_t)pl->n_paths * 16, st));
    const uint32_t chunks = pl->h_prefix[pl->n_paths];
    if (chunks == 0) return FGFA_OK;
    const uint32_t* base = d_steps;
    int rc = prepare_pointer(pl, d_steps, &base);
    if (rc) return rc;
    uint2* dl = static_cast<uint2*>(d_scratch);
    fgfa::k_interleave_depth_len<<<(pl->n_segs + 255) / 256, 256, 0, st>>>(d_depth, d_seg_len, dl, pl->n_segs);
    CU(cudaGetLastError());
    fgfa::MeasureParams M{};
    M.steps = base;
    M.chunks = pl->d_chunks;
    M.chunk_lo = 0;
    M.chunk_hi = chunks;
    M.n_segs = pl->n_segs;
    M.depth_len = dl;
    M.sums = reinterpret_cast<unsigned long long*>(d_sums);
    M.err = pl->d_err;
    const uint32_t grid = std::min<uint32_t>(chunks, (uint32_t)pl->sms * 8);
    fgfa::k_path_measure<<<grid, fgfa::kThreads, 0, st>>>(M);
    CU(cudaGetLastError());
    return FGFA_OK;
}

uint32_t fgfa_depth_plan_launches(const fgfa_depth_plan_t* pl, int with_uniq) {
    if (!pl || pl->n_paths == 0) return 0;
    const uint32_t batches = (pl->n_paths + pl->rows_per_batch - 1) / pl->rows_per_batch;
    if (pl->engine == 1) return with_uniq ? 5 * batches : 4 * batches;   // S1, S2, S3, W (+ B2)
    return with_uniq ? 2 * batches : batches;
}

size_t fgfa_depth_plan_scratch_bytes(const fgfa_depth_plan_t* pl) { return pl ? pl->scratch_bytes : 0; }

int fgfa_depth_device(const uint32_t* d_steps, uint64_t n_steps, const uint32_t* d_span_start,
                      const uint32_t* d_span_end, uint32_t n_paths, uint32_t n_segs,
                      uint32_t* d_depth, uint32_t* d_uniq, void* cuda_stream) {
    if (n_paths && (!d_span_start || !d_span_end)) return fail(FGFA_ERR_INVALID_ARG, "null span table");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    std::vector<uint32_t> s(n_paths), e(n_paths);
    if (n_paths) {
        CU(cudaMemcpyAsync(s.data(), d_span_start, (size_t)n_paths * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(e.data(), d_span_end, (size_t)n_paths * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    fgfa_depth_plan_t* pl = nullptr;
    int rc = fgfa_depth_plan_create(&pl, s.data(), e.data(), n_paths, n_segs, n_steps, 0);
    if (rc) return rc;
    rc = fgfa_depth_plan_run(pl, d_steps, d_depth, d_uniq, cuda_stream);
    if (rc == FGFA_OK) rc = fgfa_depth_plan_status(pl, cuda_stream);
    fgfa_depth_plan_destroy(pl);
    return rc;
}

}  // extern "C"
