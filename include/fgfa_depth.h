/*
 * fgfa_depth.h -- thin C ABI of the B200 node-depth engine.
 *
 * This is the boundary a Rust `extern "C"` block in the reference's `flatgfa` crate
 * would bind to replace the body of
 *     flatgfa/src/ops/depth.rs:15-39   pub fn seg_depth_with_uniq(gfa) -> (Vec<usize>, Vec<usize>)
 *     flatgfa/src/ops/depth.rs:45-56   pub fn seg_depth(gfa) -> Vec<usize>
 * (callers: flatgfa/src/cli/cmds.rs:239, flatgfa-sh/src/eval/instr.rs:30,
 * flatgfa/src/ops/window_depth.rs:177).  Plain pointers and sizes only; no C++ or
 * torch types.  The reference has no such export today (flatgfa-c/src/lib.rs exposes
 * only per-step accessors), see INTEGRATION.md for the binding a maintainer would add.
 *
 * Data conventions (all little-endian, as in the .flatgfa format):
 *   steps       u32 Handle words, segment index << 1 | orientation  (flatgfa.rs:186-198)
 *   span_start/ per-path half-open range [start,end) into `steps`   (flatgfa.rs:99-112,
 *   span_end    pool.rs:80-124); arbitrary order/overlap is honoured
 *   depth/uniq  one counter per segment, indexed by segment pool index (depth.rs:17-18)
 *
 * Errors: every function returns FGFA_OK (0) or a negative code; nothing aborts or
 * throws across this boundary.  Where the reference would panic (span outside the
 * pool, segment id >= n_segs: pool.rs:341-347, depth.rs:29) the call fails with
 * FGFA_ERR_SPAN_OOB / FGFA_ERR_SEG_OOB and the outputs are unspecified.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point returns
 * FGFA_ERR_NO_DEVICE.
 */
#ifndef FGFA_DEPTH_H
#define FGFA_DEPTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    FGFA_OK = 0,
    FGFA_ERR_INVALID_ARG = -1, /* null pointer, misaligned device pointer, ... */
    FGFA_ERR_BAD_MAGIC = -2,   /* .flatgfa image: magic != 0xB1011054 (file.rs:9,172-173) */
    FGFA_ERR_TRUNCATED = -3,   /* .flatgfa image shorter than its table of contents says */
    FGFA_ERR_SPAN_OOB = -4,    /* a path's steps span lies outside the steps pool */
    FGFA_ERR_SEG_OOB = -5,     /* a step names a segment index >= n_segs */
    FGFA_ERR_CUDA = -6,        /* CUDA runtime failure; see fgfa_last_error() */
    FGFA_ERR_NOMEM = -7,
    FGFA_ERR_NO_DEVICE = -8,   /* no usable CUDA device: the product has no CPU path */
    FGFA_ERR_TOO_LARGE = -9,   /* counts do not fit the format's u32 ids (pool.rs:9-11) */
    FGFA_ERR_PARSE = -10       /* step-list text outside the strict grammar, or an unknown segment name */
};

/* Static description of an error code. */
const char* fgfa_strerror(int code);
/* Detail of the last failure on the calling thread (CUDA error string etc.). */
const char* fgfa_last_error(void);

/* ---- device-resident API -------------------------------------------------------- */

/* A plan owns everything that depends only on the graph's shape: the chunk table
 * derived from the path spans, the per-path `seen` bitmap scratch (the GPU form of
 * depth.rs:23's BitVec), launch geometry.  Create once, run many times; runs enqueue
 * only asynchronous work on the caller's stream (CUDA-graph capturable). */
typedef struct fgfa_depth_plan fgfa_depth_plan_t;

/* h_span_start/h_span_end: HOST arrays of n_paths entries.  bitmap_budget_bytes caps
 * the seen-bitmap scratch (paths are processed in batches that fit); 0 = default. */
int fgfa_depth_plan_create(fgfa_depth_plan_t** out, const uint32_t* h_span_start,
                           const uint32_t* h_span_end, uint32_t n_paths, uint32_t n_segs,
                           uint64_t n_steps, size_t bitmap_budget_bytes);
void fgfa_depth_plan_destroy(fgfa_depth_plan_t* plan);

/* seg_depth_with_uniq on device buffers.  d_steps: n_steps Handle words (4-byte
 * aligned; 16-byte aligned for full speed).  d_depth, d_uniq: n_segs u32 each, fully
 * overwritten (zero-initialisation is part of the run, as `vec![0; n]` is part of the
 * reference's, depth.rs:17-18).  d_uniq may be NULL: then this is seg_depth
 * (depth.rs:45-56).  u32 counters cannot overflow: a depth is at most n_steps < 2^32. */
int fgfa_depth_plan_run(fgfa_depth_plan_t* plan, const uint32_t* d_steps, uint32_t* d_depth,
                        uint32_t* d_uniq, void* cuda_stream);

/* Same, restricted to paths [path_lo, path_hi) and WITHOUT zeroing d_depth first or
 * running the uniq reduction: lets a host pipeline overlap uploads with compute.
 * Finish with fgfa_depth_plan_finish().  Paths must be fed in increasing order. */
int fgfa_depth_plan_begin(fgfa_depth_plan_t* plan, uint32_t* d_depth, void* cuda_stream);
int fgfa_depth_plan_feed(fgfa_depth_plan_t* plan, const uint32_t* d_steps, uint32_t path_lo,
                         uint32_t path_hi, uint32_t* d_depth, uint32_t* d_uniq, void* cuda_stream);
int fgfa_depth_plan_finish(fgfa_depth_plan_t* plan, uint32_t* d_uniq, void* cuda_stream);

/* Engines.  A plan runs one of two implementations of the same loop nest (depth.rs:25-35):
 *   FGFA_ENGINE_WINDOW  segment-major: a data-dependent pre-pass bins 256-step sub-chunks by
 *                       segment window, then every window is counted in shared memory (depth
 *                       counters + 32-path masks).  Default for pools of >= 64 Mi steps.
 *   FGFA_ENGINE_STREAM  path-major: one L2 reduction per step, per-path seen-bitmap rows.  Default
 *                       for small pools; always used with fgfa_depth_plan_use_bitmap().
 * Both give the same (bit-exact) results; they differ in speed only.  set_engine synchronises the
 * device and (re)allocates the seen scratch; FGFA_ERR_INVALID_ARG if the engine is not available
 * for this plan (the window engine needs n_segs / 16384 * planes <= 12000 keys).  autotune samples
 * the resident pool once (synchronises the stream): a pool whose consecutive steps jump further
 * than a shared-memory window can hold gains nothing from windows and is given the stream engine.
 * The environment variable FGFA_ENGINE=stream|window overrides both the default and autotune. */
enum { FGFA_ENGINE_STREAM = 0, FGFA_ENGINE_WINDOW = 1 };
int fgfa_depth_plan_set_engine(fgfa_depth_plan_t* plan, int engine);
int fgfa_depth_plan_engine(const fgfa_depth_plan_t* plan);
int fgfa_depth_plan_autotune(fgfa_depth_plan_t* plan, const uint32_t* d_steps, void* cuda_stream);

/* Synchronise the stream and report the sticky device status of the runs since the
 * last call: FGFA_OK, FGFA_ERR_SEG_OOB or FGFA_ERR_CUDA. */
int fgfa_depth_plan_status(fgfa_depth_plan_t* plan, void* cuda_stream);

/* ---- fused popcount + exchange over NVLink peer memory (multi-GPU) ------------------
 * The plan's seen-bitmap can live in caller-provided (symmetric, peer-mapped) memory:
 * `bytes` >= words_per_row*4*n_paths with words_per_row = ceil(n_segs/32) rounded up to 32,
 * 128-byte aligned, zero on first use; the caller re-zeroes it after every exchange. */
int fgfa_depth_plan_use_bitmap(fgfa_depth_plan_t* plan, void* d_bitmap, size_t bytes);
/* Zero d_depth and run only the step-stream kernel: partial depth + seen-bits, no popcount. */
int fgfa_depth_plan_run_stream_only(fgfa_depth_plan_t* plan, const uint32_t* d_steps, uint32_t* d_depth,
                                    void* cuda_stream);
/* One launch on rank `rank`: for this rank's slice of the segment axis, sum all ranks'
 * partial depths, popcount all ranks' bitmap rows (rows[q] rows of pitch words_per_row at
 * bitmaps[q]) and store final depth (u32) / uniq (u8) into every rank's result buffers.
 * All pointers are device pointers valid in THIS process (peer mappings).  The caller
 * orders the launch between two inter-rank barriers on the same stream.  <= 255 paths.
 * multicast_base (optional, else NULL): NVLS multicast mapping of a symmetric buffer that
 * holds partial depth / final depth / final uniq at the given byte offsets on every rank;
 * with it the depth sum is an in-switch `multimem.ld_reduce` and the all-gather a
 * `multimem.st`, so a rank moves one slice in and one slice out instead of N-1 of each. */
int fgfa_exchange_uniq_depth(int n_ranks, int rank, const void* const* bitmaps, const uint32_t* rows,
                             const void* const* partial_depths, void* const* final_depths,
                             void* const* final_uniqs, uint32_t n_segs, void* multicast_base,
                             uint64_t off_partial, uint64_t off_final_depth, uint64_t off_final_uniq,
                             void* cuda_stream);

/* The same exchange as PUSH + local reduce (kernels P and R; faster than the pull form above from 4 GPUs on:
 * partials travel as posted peer stores, the bitmaps never leave their GPU).  Slices and receive buffers:
 * fgfa_exchange_recv_bytes(n_ranks, n_segs) bytes per rank, any 256-byte aligned peer-mapped memory.
 *   fgfa_exchange_push    popcount this rank's `rows` bitmap rows (and clear them), and store the u8 counts
 *                         and this rank's partial depth, slice by slice, into slot `rank` of the slice owner's
 *                         receive buffer (recv_bufs[owner]).  With partial_uniq_u8 != NULL (u8 counts that a
 *                         plan with fgfa_depth_plan_set_uniq_width(1) has already produced, e.g. by the window
 *                         engine; the buffer must be readable up to a multiple of 32 bytes) the counts are
 *                         forwarded instead and bitmap / rows are ignored.
 *   -- inter-rank barrier (caller) --
 *   fgfa_exchange_reduce  add the n_ranks slots of this rank's slice and store final depth (u32) / uniq (u8)
 *                         into every rank's result buffers (multicast_base as for fgfa_exchange_uniq_depth).
 *   -- inter-rank barrier (caller) --
 * <= 255 paths in the whole graph. */
size_t fgfa_exchange_recv_bytes(int n_ranks, uint32_t n_segs);
int fgfa_exchange_push(int n_ranks, int rank, void* bitmap, uint32_t rows, const void* partial_depth,
                       const void* partial_uniq_u8, void* const* recv_bufs, uint32_t n_segs, void* cuda_stream);
int fgfa_exchange_reduce(int n_ranks, int rank, const void* recv_buf, void* const* final_depths,
                         void* const* final_uniqs, uint32_t n_segs, void* multicast_base,
                         uint64_t off_final_depth, uint64_t off_final_uniq, void* cuda_stream);

/* ---- multi-GPU, one process driving N devices ------------------------------------------
 * The reference's host is one compiled process (flatgfa/src/cli/main.rs:57-188, cmds.rs:234-245), so
 * this is the form a Rust `fgfa --gpus N` binds.  Whole paths are partitioned over the devices by step
 * count (longest first onto the least loaded device; `Path::step_count`, flatgfa.rs:114-118): depth is a
 * sum over steps and uniq a sum over paths of indicator vectors (depth.rs:25-35), so both shard exactly
 * as long as no path is split.  Every device counts its packed shard with a depth plan; the partial
 * [depth | uniq] arrays are combined by
 *   FGFA_EXCHANGE_NCCL  one ncclAllReduce(sum) per device inside one NCCL group.  uniq travels as u8
 *                       (four to a word) when the graph has <= 255 paths.  libnccl.so.2 is loaded on
 *                       first use; this library has no link-time dependency on it.
 *   FGFA_EXCHANGE_PEER  kernel X (fgfa_exchange_uniq_depth): popcount fused with a reduce-scatter /
 *                       all-gather over peer-mapped memory, ordered with CUDA events.  <= 255 paths,
 *                       devices must be able to map each other's memory.
 * create() owns the communicators, streams, plans and buffers; destroy() releases them (the
 * init / destroy pair SURVEY.md section 8b asks for).  After run() every device holds the complete
 * result.  Handles are not thread-safe; all calls return FGFA_OK or a negative code. */
typedef struct fgfa_depth_multi fgfa_depth_multi_t;
enum { FGFA_EXCHANGE_NCCL = 0, FGFA_EXCHANGE_PEER = 1 };
/* devices: n_devices CUDA ordinals (distinct for NCCL).  h_span_start/end: the whole graph's spans. */
int fgfa_depth_multi_create(fgfa_depth_multi_t** out, const int* devices, int n_devices,
                            const uint32_t* h_span_start, const uint32_t* h_span_end, uint32_t n_paths,
                            uint32_t n_segs, uint64_t n_steps, int exchange);
void fgfa_depth_multi_destroy(fgfa_depth_multi_t* m);
/* path_device[p] = index (into `devices`) of the device that owns path p; device_steps[i] = steps on
 * device i.  Either may be NULL. */
int fgfa_depth_multi_partition(const fgfa_depth_multi_t* m, uint32_t* path_device, uint64_t* device_steps);
/* Make the shards resident: every path's steps go to its device, packed back to back in ascending
 * path order (asynchronous copies; pin h_steps for full PCIe speed). */
int fgfa_depth_multi_upload(fgfa_depth_multi_t* m, const uint32_t* h_steps);
/* Device i's shard buffer, for callers that fill it themselves (then no upload() is needed). */
int fgfa_depth_multi_device_steps(fgfa_depth_multi_t* m, int index, uint32_t** d_steps, uint64_t* n_steps);
/* Enqueue one query on all devices: zero, count, exchange.  with_uniq = 0 is seg_depth (NCCL form). */
int fgfa_depth_multi_run(fgfa_depth_multi_t* m, int with_uniq);
/* Wait for all devices; FGFA_ERR_SEG_OOB if any shard saw a segment id >= n_segs. */
int fgfa_depth_multi_sync(fgfa_depth_multi_t* m);
/* sync + copy the result (device 0's replica) out as u64 counters; uniq_out may be NULL. */
int fgfa_depth_multi_download(fgfa_depth_multi_t* m, uint64_t* depth_out, uint64_t* uniq_out);
/* upload + run + download. */
int fgfa_depth_multi_run_host(fgfa_depth_multi_t* m, const uint32_t* h_steps, uint64_t* depth_out, uint64_t* uniq_out);
/* Device i's replica of the result: n_segs u32 depths; uniq as u8 (uniq_bytes = 1) or u32 (4). */
int fgfa_depth_multi_result_device(fgfa_depth_multi_t* m, int index, const uint32_t** d_depth, const void** d_uniq,
                                   int* uniq_bytes);
const char* fgfa_depth_multi_last_error(void);
/* The partition alone (pure host code): path_part[p] = part (0..n_parts-1) that owns path p. */
int fgfa_lpt_partition(const uint32_t* h_span_start, const uint32_t* h_span_end, uint32_t n_paths, int n_parts,
                       uint32_t* path_part);

/* Width of the uniq counters the plan's runs write to d_uniq: 4 (default, u32) or 1 (u8;
 * only for plans of <= 255 paths, since uniq <= n_paths).  The narrow form exists for the
 * multi-GPU exchange: [depth u32 | uniq u8] is 25 MB instead of 40 MB at 5 M segments, and
 * packed bytes can be summed as u32 words because no global uniq exceeds 255 either. */
int fgfa_depth_plan_set_uniq_width(fgfa_depth_plan_t* plan, int bytes);

/* Path-depth mode (`path_depth` + `measure_path`, depth.rs:88-131) on device buffers.
 * d_depth: the n_segs u32 depths of a previous run over ALL paths (depth.rs:93-99);
 * d_seg_len: n_segs u32 sequence lengths (`Segment::len`, flatgfa.rs:84-89);
 * d_scratch: 8*n_segs bytes; d_sums: 2*n_paths u64, receives for every path p
 * sums[2p] = sum(depth[seg]*len(seg)) and sums[2p+1] = sum(len(seg)) over its steps
 * (wrapping u64, like usize).  mean depth = sums[2p] / sums[2p+1] as f64 (depth.rs:129). */
int fgfa_depth_plan_path_sums(fgfa_depth_plan_t* plan, const uint32_t* d_steps, const uint32_t* d_depth,
                              const uint32_t* d_seg_len, void* d_scratch, uint64_t* d_sums,
                              void* cuda_stream);

/* Measurement hook: the next fgfa_depth_plan_run/feed records `before` immediately
 * before and `after` immediately after its step-stream kernel launches (CUDA events
 * owned by the caller, timing enabled).  Pass NULLs to clear.  One-shot: cleared after
 * the run that used it. */
int fgfa_depth_plan_set_probe(fgfa_depth_plan_t* plan, void* cuda_event_before, void* cuda_event_after);

/* Kernel launches one fgfa_depth_plan_run enqueues (for launch accounting). */
uint32_t fgfa_depth_plan_launches(const fgfa_depth_plan_t* plan, int with_uniq);
/* Bytes of device scratch the plan holds. */
size_t fgfa_depth_plan_scratch_bytes(const fgfa_depth_plan_t* plan);

/* One-shot form with the span table on the device too (creates and destroys a plan;
 * synchronises). */
int fgfa_depth_device(const uint32_t* d_steps, uint64_t n_steps, const uint32_t* d_span_start,
                      const uint32_t* d_span_end, uint32_t n_paths, uint32_t n_segs,
                      uint32_t* d_depth, uint32_t* d_uniq, void* cuda_stream);

/* ---- host-buffer API (uploads, runs, downloads; synchronous) -------------------- */

/* seg_depth_with_uniq over host arrays.  depth_out/uniq_out: n_segs u64 (= Rust usize)
 * each; uniq_out may be NULL (seg_depth).  h_steps may be pageable or pinned. */
int fgfa_seg_depth_with_uniq_steps(const uint32_t* h_steps, uint64_t n_steps,
                                   const uint32_t* h_span_start, const uint32_t* h_span_end,
                                   uint32_t n_paths, uint32_t n_segs, uint64_t* depth_out,
                                   uint64_t* uniq_out);

/* path_depth over host arrays (depth.rs:88-113): node depth over ALL paths, then for each
 * queried path id its length in base pairs and its mean depth.  h_seg_len: n_segs u32
 * sequence lengths.  path_ids may be NULL = all paths in order (`gfa.paths.ids()`).
 * length_out, weighted_out (sum of depth*len, exact) and mean_out have n_query entries;
 * weighted_out may be NULL. */
int fgfa_path_depth_steps(const uint32_t* h_steps, uint64_t n_steps, const uint32_t* h_span_start,
                          const uint32_t* h_span_end, uint32_t n_paths, const uint32_t* h_seg_len,
                          uint32_t n_segs, const uint32_t* path_ids, uint32_t n_query,
                          uint64_t* length_out, uint64_t* weighted_out, double* mean_out);

/* ---- interval / window depth along one path ------------------------
 * GPU form of flatgfa/src/ops/window_depth.rs: `path_length` (:69-77), `weighted_depths`
 * (:84-103), `assign_depths` (:118-153), `interval_depth` (:176-180), `window_depth` (:183-197)
 * and `bed_depth` (:203-211).  Results are f64 and bit-identical to the reference's: every
 * interval is summed in step order with the same operation sequence.
 *
 * Device level.  `d_path_steps` = the first Handle of the ONE path the intervals lie on
 * (any 4-byte-aligned device address), `n_path_steps` its step count; `d_depth` = node
 * depth over all paths (seg_depth, window_depth.rs:177) as u32; `d_seg_len` = Segment::len
 * per segment.  Scratch: fgfa_interval_scratch_bytes(n_path_steps, n_intervals) bytes.
 * Calls only enqueue work on `cuda_stream`; fgfa_interval_status() synchronises and reports
 * a step that named a segment >= n_segs (FGFA_ERR_SEG_OOB). */
size_t fgfa_interval_scratch_bytes(uint64_t n_path_steps, uint64_t n_intervals);
/* W1: d_seg_end[j] = end offset in base pairs of step j (inclusive prefix sum of the
 * segment lengths); d_seg_end[n-1] is `path_length`.  d_seg_end: n_path_steps u64. */
int fgfa_path_offsets_device(const uint32_t* d_path_steps, uint32_t n_path_steps,
                             const uint32_t* d_seg_len, uint32_t n_segs, uint64_t* d_seg_end,
                             void* d_scratch, size_t scratch_bytes, void* cuda_stream);
/* `Windows` (window_depth.rs:20-52): [start + w*size, min(start + (w+1)*size, end)). */
int fgfa_make_windows_device(uint64_t start, uint64_t end, uint64_t size, uint64_t n_windows,
                             uint64_t* d_win_start, uint64_t* d_win_end, void* cuda_stream);
/* W2 + W3: `assign_depths`.  d_out: n_intervals f64. */
int fgfa_interval_depth_device(const uint32_t* d_path_steps, uint32_t n_path_steps,
                               const uint32_t* d_depth, const uint32_t* d_seg_len, uint32_t n_segs,
                               const uint64_t* d_seg_end, const uint64_t* d_win_start,
                               const uint64_t* d_win_end, uint64_t n_intervals, double* d_out,
                               void* d_scratch, size_t scratch_bytes, void* cuda_stream);
int fgfa_interval_status(const void* d_scratch, void* cuda_stream);

/* Host level (uploads, runs, downloads; synchronous).  Node depth is taken over ALL paths
 * of the graph, the intervals lie along path `path` (a path pool index).
 * bed_depth / interval_depth (window_depth.rs:176-180, :203-211): caller-allocated
 * depth_out with n_intervals entries. */
int fgfa_interval_depth_steps(const uint32_t* h_steps, uint64_t n_steps, const uint32_t* h_span_start,
                              const uint32_t* h_span_end, uint32_t n_paths, const uint32_t* h_seg_len,
                              uint32_t n_segs, uint32_t path, const uint64_t* h_win_start,
                              const uint64_t* h_win_end, uint64_t n_intervals, double* depth_out);
/* window_depth (window_depth.rs:183-197): equally sized windows over the whole path.
 * *depth_out is allocated by the library (release with fgfa_free) and holds *n_windows_out
 * values; window w is [w*window_size, min((w+1)*window_size, *path_length_out)).
 * window_size == 0 is FGFA_ERR_INVALID_ARG (the reference loops forever). */
int fgfa_window_depth_steps(const uint32_t* h_steps, uint64_t n_steps, const uint32_t* h_span_start,
                            const uint32_t* h_span_end, uint32_t n_paths, const uint32_t* h_seg_len,
                            uint32_t n_segs, uint32_t path, uint64_t window_size, double** depth_out,
                            uint64_t* n_windows_out, uint64_t* path_length_out);
void fgfa_free(void* p);

/* The host-buffer entry points keep their device/pinned staging and the last plan in a
 * process-wide workspace between calls; this frees it (it is re-created on demand).
 * Setting FGFA_WORKSPACE=0 in the environment frees it after every call instead. */
void fgfa_release_workspace(void);

/* seg_depth_with_uniq / seg_depth over a .flatgfa image in host memory (e.g. the mmap
 * the reference's `file::view` takes, file.rs:185-213).  Outputs: n_segs u64 each,
 * where n_segs is what fgfa_flatgfa_counts() reports. */
int fgfa_flatgfa_counts(const void* flatgfa_bytes, size_t len, uint64_t* n_segs,
                        uint64_t* n_paths, uint64_t* n_steps);
int fgfa_seg_depth_with_uniq(const void* flatgfa_bytes, size_t len, uint64_t* depth_out,
                             uint64_t* uniq_out);
int fgfa_seg_depth(const void* flatgfa_bytes, size_t len, uint64_t* depth_out);

/* ---- GPU step-list tokenizer -------------------------------------
 * The text of GFA path step lists (`1+,23-,4+`) -> Handle words on the device: the GPU form
 * of `StepsParser` (gfaline.rs:201-263) + `NameMap::get` (namemap.rs:27-33) + `Handle::new`
 * (flatgfa.rs:192-198), the inner loop of `Parser::add_path` (parse.rs:149-156).
 * `h_text` is any host buffer (typically the mmapped GFA file); field f is the steps field
 * of the f-th P line: bytes [field_off[f], field_off[f]+field_len[f]).  create() uploads the
 * text and counts the steps of every field; parse() resolves names with the reference's
 * NameMap contents (names 1..sequential_max map to name-1; `other_names[i]` -> `other_ids[i]`)
 * and writes the steps of all fields back to back (field order) to h_steps_out (may be NULL:
 * the steps stay on the device, see fgfa_tokenizer_device_steps).  Only the strict grammar
 * token (',' token)*, token = digit+ ('+'|'-'), is accepted: FGFA_ERR_PARSE means "re-parse
 * this input with the host parser", which reproduces the reference's quirks and errors. */
typedef struct fgfa_tokenizer fgfa_tokenizer_t;
int fgfa_tokenizer_create(fgfa_tokenizer_t** out, const uint8_t* h_text, uint64_t n_bytes,
                          const uint64_t* field_off, const uint64_t* field_len, uint32_t n_fields);
int fgfa_tokenizer_spans(const fgfa_tokenizer_t* t, uint32_t* span_start, uint32_t* span_end, uint64_t* n_steps);
int fgfa_tokenizer_parse(fgfa_tokenizer_t* t, uint64_t sequential_max, const uint64_t* other_names,
                         const uint32_t* other_ids, uint32_t n_others, uint32_t* h_steps_out);
const uint32_t* fgfa_tokenizer_device_steps(const fgfa_tokenizer_t* t);
void fgfa_tokenizer_destroy(fgfa_tokenizer_t* t);

/* Number of CUDA devices visible (0 if none / no driver). */
int fgfa_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FGFA_DEPTH_H */
