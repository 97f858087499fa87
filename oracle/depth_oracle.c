/*
 * ORACLE — test infrastructure, NOT product code.
 *
 * A plain-C, single-threaded restatement of the reference's node-depth path, used
 * only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs as the checker and the CPU baseline.  Nothing under pollen_b200/
 * may include, link or call this file.
 *
 * Parity pinning: the Rust reference cannot be built here (no cargo/rustc, crates
 * not vendored), so this restatement is pinned against (a) the reference's in-repo
 * golden table for the path (slow_odgi/README.md:147-175; the second documented table,
 * flatgfa-sh/README.md:31-36, needs note5.gfa which the reference fetches from odgi at
 * test time and is therefore not usable offline)
 * and (b) outputs of the reference's own Python implementation `slow_odgi depth`
 * (slow_odgi/slow_odgi/depth.py:6-16) run in the build container on the reference's
 * fixtures and on seeded random graphs; those outputs are committed under
 * tests/golden/ together with the generating script (oracle/make_golden.py).
 *
 * Every function cites the reference lines it follows (paths relative to the
 * reference checkout).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* flatgfa/src/file.rs:9 */
#define ORACLE_MAGIC 0xB1011054ull

/* flatgfa/src/flatgfa.rs:201-203: Handle::segment() = word >> 1 (low bit = orientation). */
static inline uint32_t handle_segment(uint32_t h) { return h >> 1; }

/*
 * flatgfa/src/ops/depth.rs:15-39  seg_depth_with_uniq.
 * `spans` holds (start,end) u32 pairs, one per path, half-open ranges into `steps`
 * (flatgfa/src/pool.rs:80-124, 341-347).  `depth`/`uniq` have n_segs entries
 * (usize == u64 on the reference's only supported target).
 * Returns 0, or -1 where the reference would panic on an out-of-bounds index
 * (span outside the pool: pool.rs:341-347; segment id >= n_segs: depth.rs:29).
 */
int oracle_seg_depth_with_uniq(const uint32_t* steps, uint64_t n_steps, const uint32_t* spans,
                               uint32_t n_paths, uint32_t n_segs, uint64_t* depth,
                               uint64_t* uniq) {
    /* depth.rs:17-18: vec![0; segs.len()] */
    memset(depth, 0, (size_t)n_segs * sizeof(uint64_t));
    memset(uniq, 0, (size_t)n_segs * sizeof(uint64_t));
    /* depth.rs:23: BitVec::from_elem(segs.len(), false) */
    size_t n_words = ((size_t)n_segs + 63) / 64;
    uint64_t* seen = (uint64_t*)calloc(n_words ? n_words : 1, sizeof(uint64_t));
    if (!seen) return -2;
    for (uint32_t p = 0; p < n_paths; ++p) { /* depth.rs:25 */
        uint32_t start = spans[2 * p], end = spans[2 * p + 1];
        if (start > end || (uint64_t)end > n_steps) { free(seen); return -1; }
        memset(seen, 0, n_words * sizeof(uint64_t)); /* depth.rs:26: seen.clear() */
        for (uint32_t i = start; i < end; ++i) {     /* depth.rs:27 */
            uint32_t seg = handle_segment(steps[i]); /* depth.rs:28 */
            if (seg >= n_segs) { free(seen); return -1; }
            depth[seg] += 1;                         /* depth.rs:29 */
            uint64_t bit = 1ull << (seg & 63);
            if (!(seen[seg >> 6] & bit)) {           /* depth.rs:30 */
                uniq[seg] += 1;                      /* depth.rs:32 */
                seen[seg >> 6] |= bit;               /* depth.rs:33 */
            }
        }
    }
    free(seen);
    return 0;
}

/* flatgfa/src/ops/depth.rs:45-56  seg_depth (no unique depth). */
int oracle_seg_depth(const uint32_t* steps, uint64_t n_steps, const uint32_t* spans,
                     uint32_t n_paths, uint32_t n_segs, uint64_t* depth) {
    memset(depth, 0, (size_t)n_segs * sizeof(uint64_t));
    for (uint32_t p = 0; p < n_paths; ++p) {
        uint32_t start = spans[2 * p], end = spans[2 * p + 1];
        if (start > end || (uint64_t)end > n_steps) return -1;
        for (uint32_t i = start; i < end; ++i) {
            uint32_t seg = handle_segment(steps[i]);
            if (seg >= n_segs) return -1;
            depth[seg] += 1;
        }
    }
    return 0;
}

/*
 * flatgfa/src/file.rs:14-38 (Toc, Size), 163-213 (slice_prefix, read_toc, view).
 * Locates the pools the depth path touches inside a .flatgfa image.  Pool order and
 * element sizes: header u8, segs 24 B, paths 24 B, links 16 B, steps 4 B, ... ;
 * each pool occupies capacity*sizeof(T) bytes of which the first len are valid.
 */
typedef struct {
    uint64_t n_segs, n_paths, n_steps;
    uint64_t segs_off, paths_off, steps_off; /* byte offsets into the image */
} oracle_view_t;

static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

int oracle_view(const uint8_t* data, uint64_t len, oracle_view_t* out) {
    if (len < 184) return -1;                     /* file.rs:170-171: Toc::ref_from_prefix */
    if (rd64(data) != ORACLE_MAGIC) return -2;    /* file.rs:172-173 */
    static const uint64_t elem[11] = {1, 24, 24, 16, 4, 1, 8, 4, 1, 1, 1};
    uint64_t off = 184, offs[11], lens[11];
    for (int i = 0; i < 11; ++i) {                /* file.rs:188-198 */
        uint64_t l = rd64(data + 8 + 16 * i), cap = rd64(data + 16 + 16 * i);
        if (cap < l) return -3;                   /* file.rs:165: capacity - len underflow */
        offs[i] = off;
        lens[i] = l;
        if (off + cap * elem[i] < off || off + cap * elem[i] > len) return -4; /* file.rs:164 unwrap */
        off += cap * elem[i];
    }
    out->n_segs = lens[1];  out->segs_off = offs[1];
    out->n_paths = lens[2]; out->paths_off = offs[2];
    out->n_steps = lens[4]; out->steps_off = offs[4];
    return 0;
}

/*
 * seg_depth_with_uniq over a .flatgfa image: extracts each Path's `steps` span
 * (bytes 8..16 of the 24-byte packed record, flatgfa/src/flatgfa.rs:99-112) and the
 * unaligned steps pool, then runs the loop above.
 */
int oracle_file_seg_depth_with_uniq(const uint8_t* data, uint64_t len, uint64_t* depth,
                                    uint64_t* uniq) {
    oracle_view_t v;
    int rc = oracle_view(data, len, &v);
    if (rc) return rc;
    uint32_t* steps = (uint32_t*)malloc((size_t)(v.n_steps ? v.n_steps : 1) * 4);
    uint32_t* spans = (uint32_t*)malloc((size_t)(v.n_paths ? v.n_paths : 1) * 8);
    if (!steps || !spans) { free(steps); free(spans); return -5; }
    memcpy(steps, data + v.steps_off, (size_t)v.n_steps * 4);
    for (uint64_t p = 0; p < v.n_paths; ++p) {
        spans[2 * p] = rd32(data + v.paths_off + 24 * p + 8);
        spans[2 * p + 1] = rd32(data + v.paths_off + 24 * p + 12);
    }
    rc = oracle_seg_depth_with_uniq(steps, v.n_steps, spans, (uint32_t)v.n_paths,
                                    (uint32_t)v.n_segs, depth, uniq);
    free(steps);
    free(spans);
    return rc;
}

/*
 * flatgfa/src/ops/depth.rs:61-82  SegDepth::emit: header line, then one row per
 * segment in pool order with `seg.name as u32` (truncating cast, depth.rs:71).
 * `names` are the usize names read at byte offset 24*i of the segs pool
 * (flatgfa/src/flatgfa.rs:71-82).  Returns bytes written, or -1 if `cap` is too small.
 */
int64_t oracle_emit_seg_depth(const uint64_t* names, const uint64_t* depth, const uint64_t* uniq,
                              uint32_t n_segs, char* buf, uint64_t cap) {
    static const char hdr[] = "#node.id\tdepth\tdepth.uniq\n"; /* depth.rs:69 */
    uint64_t n = 0;
    if (cap < sizeof(hdr)) return -1;
    memcpy(buf, hdr, sizeof(hdr) - 1);
    n = sizeof(hdr) - 1;
    for (uint32_t i = 0; i < n_segs; ++i) { /* depth.rs:70 */
        if (cap - n < 64) return -1;
        n += (uint64_t)snprintf(buf + n, cap - n, "%u\t%llu\t%llu\n", (uint32_t)names[i],
                                (unsigned long long)depth[i], (unsigned long long)uniq[i]);
    }
    return (int64_t)n;
}

/* Segment names out of a .flatgfa image, for oracle_emit_seg_depth. */
int oracle_file_seg_names(const uint8_t* data, uint64_t len, uint64_t* names) {
    oracle_view_t v;
    int rc = oracle_view(data, len, &v);
    if (rc) return rc;
    for (uint64_t i = 0; i < v.n_segs; ++i) names[i] = rd64(data + v.segs_off + 24 * i);
    return 0;
}

/*
 * flatgfa/src/ops/depth.rs:88-113 path_depth + :116-131 measure_path ("next" row).
 * Node depth over ALL paths (depth.rs:93-99), then for each queried path id the length in
 * base pairs and the mean depth weighted by segment length.  `seg_len[i]` is Segment::len
 * (flatgfa.rs:84-89).  usize arithmetic wraps (release build), the divide is f64.
 * PARITY UNPINNED for this function: the reference repository holds no runnable golden
 * for path depth (slow_odgi has no path mode; the tables in flatgfa-sh/README.md:57-59,
 * 267-270 need note5.gfa / k.gfa, which are fetched from odgi at test time).
 */
int oracle_path_depth(const uint32_t* steps, uint64_t n_steps, const uint32_t* spans, uint32_t n_paths,
                      const uint32_t* seg_len, uint32_t n_segs, const uint32_t* path_ids,
                      uint32_t n_query, uint64_t* lengths, double* means) {
    uint64_t* seg_depths = (uint64_t*)malloc((size_t)(n_segs ? n_segs : 1) * sizeof(uint64_t));
    if (!seg_depths) return -2;
    int rc = oracle_seg_depth(steps, n_steps, spans, n_paths, n_segs, seg_depths);
    if (rc) { free(seg_depths); return rc; }
    for (uint32_t q = 0; q < n_query; ++q) {           /* depth.rs:104-108 */
        uint32_t p = path_ids ? path_ids[q] : q;
        if (p >= n_paths) { free(seg_depths); return -1; }
        uint64_t depth = 0, length = 0;                /* depth.rs:121-122 */
        for (uint32_t i = spans[2 * p]; i < spans[2 * p + 1]; ++i) {
            uint32_t seg = handle_segment(steps[i]);
            uint64_t len = seg_len[seg];               /* depth.rs:125 */
            depth += seg_depths[seg] * len;            /* depth.rs:126 */
            length += len;                             /* depth.rs:127 */
        }
        lengths[q] = length;
        means[q] = (double)depth / (double)length;     /* depth.rs:129 */
    }
    free(seg_depths);
    return 0;
}

/* depth.rs:192-197 format_float: "{:.digits$}", trim trailing '0', then a trailing '.'. */
int oracle_format_float(double x, int digits, char* out, size_t cap) {
    if (x != x) return snprintf(out, cap, "NaN");      /* Rust prints f64 NaN as "NaN" */
    if (x > 1.7976931348623157e308) return snprintf(out, cap, "inf");
    if (x < -1.7976931348623157e308) return snprintf(out, cap, "-inf");
    int n = snprintf(out, cap, "%.*f", digits, x);
    if (n < 0 || (size_t)n >= cap) return -1;
    while (n > 0 && out[n - 1] == '0') out[--n] = 0;
    while (n > 0 && out[n - 1] == '.') out[--n] = 0;
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Interval ("window") depth, SURVEY §8f-3: flatgfa/src/ops/window_depth.rs and flatbed.rs.
 * PARITY UNPINNED for these functions: the reference holds no runnable golden for them
 * (the table in flatgfa-sh/README.md:282-294 needs note5.gfa, fetched from odgi at test
 * time, plus bedtools).  tests/ cross-check this restatement against an independent
 * pure-Python restatement written from the same Rust source, and reproduce the README's
 * table on a hand-made graph with the same shape (every segment of the path at depth 2).
 * ------------------------------------------------------------------------------------------ */

/* window_depth.rs:61-63 Windows::len = (end - start).div_ceil(size). */
uint64_t oracle_window_count(uint64_t start, uint64_t end, uint64_t size) {
    uint64_t span = end - start;
    return span / size + (span % size ? 1 : 0);
}

/* window_depth.rs:41-52 Windows::emit_bed: the (start, end) pairs of the windows. */
void oracle_make_windows(uint64_t start, uint64_t end, uint64_t size, uint64_t* win_start, uint64_t* win_end) {
    uint64_t pos = start, w = 0;
    while (pos < end) {                                /* :46 */
        uint64_t stop = pos + size < end ? pos + size : end;   /* :47 */
        win_start[w] = pos;
        win_end[w] = stop;
        ++w;
        pos = stop;                                    /* :49 */
    }
}

/* window_depth.rs:69-77 path_length. */
uint64_t oracle_path_length(const uint32_t* steps, const uint32_t* spans, uint32_t path, const uint32_t* seg_len) {
    uint64_t total = 0;
    for (uint32_t i = spans[2 * path]; i < spans[2 * path + 1]; ++i) total += seg_len[handle_segment(steps[i])];
    return total;
}

/*
 * window_depth.rs:176-180 interval_depth = seg_depth over ALL paths (:177), weighted_depths
 * along `path` (:84-103) fed to assign_depths (:118-153).  The loop below is the reference's:
 * one cursor over the intervals, advanced while the steps stream by.  All f64 operations are
 * the reference's, in its order; build with -ffp-contract=off.
 */
int oracle_interval_depth(const uint32_t* steps, uint64_t n_steps, const uint32_t* spans, uint32_t n_paths,
                          const uint32_t* seg_len, uint32_t n_segs, uint32_t path, const uint64_t* win_start,
                          const uint64_t* win_end, uint64_t n_win, double* out) {
    if (path >= n_paths) return -1;
    uint64_t* depth = (uint64_t*)malloc((size_t)(n_segs ? n_segs : 1) * sizeof(uint64_t));
    if (!depth) return -2;
    int rc = oracle_seg_depth(steps, n_steps, spans, n_paths, n_segs, depth);   /* :177 */
    if (rc) { free(depth); return rc; }
    for (uint64_t w = 0; w < n_win; ++w) out[w] = 0.0;          /* :119 */
    uint64_t cur = 0;                                           /* :122 */
    uint64_t pos = 0;                                           /* :89 */
    for (uint32_t i = spans[2 * path]; i < spans[2 * path + 1]; ++i) {   /* :123, :90 */
        uint32_t seg = handle_segment(steps[i]);
        uint64_t len = seg_len[seg];
        uint64_t r0 = pos, r1 = pos + len;                      /* :92-94 */
        pos = r1;
        double seg_depth = (double)(depth[seg] * len);          /* :95, :97 */
        while (cur < n_win) {                                   /* :125 */
            uint64_t w0 = win_start[cur], w1 = win_end[cur];    /* :126-127 */
            uint64_t a = w0 > r0 ? w0 : r0;                     /* :128, overlap :110-112 */
            uint64_t b = w1 < r1 ? w1 : r1;
            if (b > a) {                                        /* :131 */
                double amt = (double)(b - a) / (double)(r1 - r0);          /* :133 */
                out[cur] += (seg_depth * amt) / (double)(w1 - w0);         /* :134-135 */
            }
            if (w1 > r1) break;                                 /* :140-142 */
            ++cur;                                              /* :145 */
        }
    }
    free(depth);
    return 0;
}

/*
 * flatbed.rs:126-152 BEDParser::parse_mem + parse_line.  Lines come from MemchrSplit
 * (memfile.rs:51-63: an unterminated last line is dropped); `#` lines are skipped;
 * name = up to the first tab (gfaline.rs:129-142); start = leading digits of the rest
 * (atoi FromRadix10, wrapping); one byte skipped; end likewise.  Output: per entry the name's
 * (offset, length) in `buf` and start/end.  Returns the number of entries, -1 where the
 * reference panics, -3 if `cap` is too small.
 */
static size_t oracle_radix10(const uint8_t* s, size_t n, uint64_t* v) {
    size_t i = 0;
    *v = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') { *v = *v * 10 + (uint64_t)(s[i] - '0'); ++i; }
    return i;
}
int64_t oracle_parse_bed(const uint8_t* buf, uint64_t len, uint64_t* name_off, uint64_t* name_len,
                         uint64_t* start, uint64_t* end, uint64_t cap) {
    uint64_t pos = 0, n = 0;
    while (pos < len) {
        const uint8_t* nl = (const uint8_t*)memchr(buf + pos, '\n', len - pos);
        if (!nl) break;
        const uint8_t* line = buf + pos;
        size_t ll = (size_t)(nl - line);
        pos = (uint64_t)(nl - buf) + 1;
        if (ll && line[0] == '#') continue;                     /* :143-145 */
        const uint8_t* tab = (const uint8_t*)memchr(line, '\t', ll);
        size_t nm = tab ? (size_t)(tab - line) : ll;
        const uint8_t* rest = tab ? tab + 1 : line + ll;
        size_t rl = tab ? ll - nm - 1 : 0;
        uint64_t s, e;
        size_t used = oracle_radix10(rest, rl, &s);             /* :148 */
        if (!used) return -1;
        rest += used; rl -= used;
        if (rl == 0) return -1;                                 /* :149 `&rest[1..]` */
        used = oracle_radix10(rest + 1, rl - 1, &e);
        if (!used) return -1;
        if (n >= cap) return -3;
        name_off[n] = (uint64_t)(line - buf); name_len[n] = nm; start[n] = s; end[n] = e;
        ++n;
    }
    return (int64_t)n;
}

/* window_depth.rs:163-174 IntervalDepth::emit: "{name}\t{start}\t{end}\t{format_float(depth, 4)}\n".
 * Names are (pointer into `names`, length) pairs.  Returns bytes written or -1. */
int64_t oracle_emit_interval_depth(const uint8_t* names, const uint64_t* name_off, const uint64_t* name_len,
                                   const uint64_t* start, const uint64_t* end, const double* depth, uint64_t n,
                                   char* buf, uint64_t cap) {
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; ++i) {
        char num[400];
        if (oracle_format_float(depth[i], 4, num, sizeof num) < 0) return -1;
        if (cap - w < name_len[i] + 64 + strlen(num)) return -1;
        memcpy(buf + w, names + name_off[i], name_len[i]);
        w += name_len[i];
        w += (uint64_t)snprintf(buf + w, cap - w, "\t%llu\t%llu\t%s\n", (unsigned long long)start[i],
                                (unsigned long long)end[i], num);
    }
    return (int64_t)w;
}

/* ------------------------------------------------------------------------------------------
 * A path-parallel CPU variant, reported BESIDE the baseline and labelled as such: it is NOT
 * the reference's algorithm (depth.rs:15-39 is one thread sharing one `seen` bitmap across
 * paths).  SURVEY §8d asks for it as an optional fairer CPU number: every thread takes whole
 * paths from a shared counter (longest first), counts into private u32 arrays with a private
 * bitmap, and the private arrays are summed per segment slice at the end.  Same results as
 * oracle_seg_depth_with_uniq (checked in tests/test_oracle_golden.py).
 * ------------------------------------------------------------------------------------------ */
#include <pthread.h>

typedef struct {
    const uint32_t* steps;
    uint64_t n_steps;
    const uint32_t* spans;
    const uint32_t* order;       /* path indices, longest first */
    uint32_t n_paths, n_segs, n_threads, tid;
    volatile uint32_t* next;     /* shared work counter */
    uint32_t* depth32;           /* [n_threads][n_segs] */
    uint32_t* uniq32;
    uint64_t* depth;             /* outputs */
    uint64_t* uniq;
    pthread_barrier_t* barrier;
    volatile int* failed;
} mt_arg_t;

static void* mt_worker(void* p) {
    mt_arg_t* a = (mt_arg_t*)p;
    uint32_t* d = a->depth32 + (size_t)a->tid * a->n_segs;
    uint32_t* u = a->uniq32 + (size_t)a->tid * a->n_segs;
    size_t words = ((size_t)a->n_segs + 63) / 64;
    uint64_t* seen = (uint64_t*)calloc(words ? words : 1, 8);
    if (!seen) *a->failed = -2;
    for (;;) {
        uint32_t k = __atomic_fetch_add(a->next, 1u, __ATOMIC_RELAXED);
        if (k >= a->n_paths || *a->failed) break;
        uint32_t path = a->order[k];
        uint32_t s = a->spans[2 * path], e = a->spans[2 * path + 1];
        if (s > e || e > a->n_steps) { *a->failed = -1; break; }
        memset(seen, 0, words * 8);
        for (uint32_t i = s; i < e; ++i) {
            uint32_t seg = handle_segment(a->steps[i]);
            if (seg >= a->n_segs) { *a->failed = -1; break; }
            d[seg] += 1;
            uint64_t bit = 1ull << (seg & 63);
            if (!(seen[seg >> 6] & bit)) { seen[seg >> 6] |= bit; u[seg] += 1; }
        }
    }
    free(seen);
    pthread_barrier_wait(a->barrier);
    /* reduction: this thread's slice of the segment axis */
    size_t lo = (size_t)a->n_segs * a->tid / a->n_threads, hi = (size_t)a->n_segs * (a->tid + 1) / a->n_threads;
    for (size_t i = lo; i < hi; ++i) {
        uint64_t sd = 0, su = 0;
        for (uint32_t t = 0; t < a->n_threads; ++t) {
            sd += a->depth32[(size_t)t * a->n_segs + i];
            su += a->uniq32[(size_t)t * a->n_segs + i];
        }
        a->depth[i] = sd;
        a->uniq[i] = su;
    }
    return NULL;
}

static const uint32_t* g_sort_spans;
static int by_length_desc(const void* x, const void* y) {
    uint32_t a = *(const uint32_t*)x, b = *(const uint32_t*)y;
    uint32_t la = g_sort_spans[2 * a + 1] - g_sort_spans[2 * a], lb = g_sort_spans[2 * b + 1] - g_sort_spans[2 * b];
    return la < lb ? 1 : la > lb ? -1 : (a > b) - (a < b);
}

int oracle_seg_depth_with_uniq_parallel(const uint32_t* steps, uint64_t n_steps, const uint32_t* spans,
                                        uint32_t n_paths, uint32_t n_segs, uint64_t* depth, uint64_t* uniq,
                                        uint32_t n_threads) {
    if (n_threads == 0) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    for (uint32_t p = 0; p < n_paths; ++p)
        if (spans[2 * p] > spans[2 * p + 1] || spans[2 * p + 1] > n_steps) return -1;
    size_t cells = (size_t)n_threads * (n_segs ? n_segs : 1);
    uint32_t* d32 = (uint32_t*)calloc(cells, 4);
    uint32_t* u32 = (uint32_t*)calloc(cells, 4);
    uint32_t* order = (uint32_t*)malloc((size_t)(n_paths ? n_paths : 1) * 4);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    mt_arg_t* args = (mt_arg_t*)malloc(sizeof(mt_arg_t) * n_threads);
    if (!d32 || !u32 || !order || !th || !args) { free(d32); free(u32); free(order); free(th); free(args); return -2; }
    for (uint32_t p = 0; p < n_paths; ++p) order[p] = p;
    g_sort_spans = spans;
    qsort(order, n_paths, 4, by_length_desc);
    pthread_barrier_t barrier;
    pthread_barrier_init(&barrier, NULL, n_threads);
    volatile uint32_t next = 0;
    volatile int failed = 0;
    for (uint32_t t = 0; t < n_threads; ++t) {
        mt_arg_t a = {steps, n_steps, spans, order, n_paths, n_segs, n_threads, t, &next, d32, u32, depth, uniq, &barrier, &failed};
        args[t] = a;
        pthread_create(&th[t], NULL, mt_worker, &args[t]);
    }
    for (uint32_t t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&barrier);
    free(d32); free(u32); free(order); free(th); free(args);
    return failed;
}
