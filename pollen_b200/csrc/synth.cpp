// Synthetic pangenome step-pool generators for the `fgfa depth` benchmarks.
//
// These build the `steps` pool (Handle = segment index << 1 | orientation bit,
// reference flatgfa/src/flatgfa.rs:186-198) and the per-path `steps` spans
// (flatgfa/src/flatgfa.rs:99-112) of the graph shapes named in BASELINE.json /
// SURVEY.md §8(d).  They are bench/test infrastructure: deterministic, seeded
// per path, multi-threaded over paths, host only.  Nothing here is on the
// product's depth path.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

constexpr uint32_t kRevThresh = 3277;     // 5 % of 65536
constexpr uint32_t kSkip0 = 52429;        // 80 %
constexpr uint32_t kSkip1 = 62259;        // +15 %
constexpr uint32_t kSkipN = 64881;        // +4 %; the remaining 1 % is a back-jump
constexpr uint32_t kHotWindow = 4096;

inline uint32_t mod_sub(uint32_t cur, uint32_t d, uint32_t n) {
    d %= n;
    return cur >= d ? cur - d : cur + n - d;
}

// One haplotype-walk transition.
inline uint32_t walk_next(uint32_t cur, uint64_t r, uint32_t n_segs) {
    uint32_t c = (uint32_t)(r >> 16) & 0xFFFF;
    uint32_t hi = (uint32_t)(r >> 32);
    if (c < kSkip0) return (cur + 1) % n_segs;
    if (c < kSkip1) return (cur + 2) % n_segs;
    if (c < kSkipN) return (uint32_t)(((uint64_t)cur + 1 + 2 + hi % 63) % n_segs);
    return mod_sub(cur, 1 + hi % 4096, n_segs);
}

void gen_walk(uint32_t* out, uint64_t len, uint32_t n_segs, uint64_t seed) {
    SplitMix64 rng(seed);
    uint32_t cur = (uint32_t)(rng.next() % n_segs);
    for (uint64_t i = 0; i < len; ++i) {
        uint64_t r = rng.next();
        uint32_t rev = ((uint32_t)r & 0xFFFF) < kRevThresh;
        out[i] = (cur << 1) | rev;
        cur = walk_next(cur, r, n_segs);
    }
}

// Config E: alternating haplotype-walk phases (mean 4096 steps) and tandem-repeat
// phases (mean 16384 steps) that loop inside one of four shared 4096-segment
// windows, so ~80 % of the steps land on hot segments.
void gen_skewed(uint32_t* out, uint64_t len, uint32_t n_segs, uint64_t seed, uint32_t path_idx,
                uint64_t graph_seed) {
    SplitMix64 rng(seed);
    SplitMix64 wrng(graph_seed ^ 0x5EEDB1011054ull);
    uint32_t bases[4];
    uint32_t win = n_segs < kHotWindow ? n_segs : kHotWindow;
    for (int k = 0; k < 4; ++k) bases[k] = (uint32_t)(wrng.next() % (n_segs - win + 1));
    uint32_t base = bases[path_idx & 3];
    uint32_t cur = (uint32_t)(rng.next() % n_segs);
    uint64_t i = 0;
    while (i < len) {
        uint64_t walk_len = 2048 + rng.next() % 4097;
        for (uint64_t j = 0; j < walk_len && i < len; ++j, ++i) {
            uint64_t r = rng.next();
            uint32_t rev = ((uint32_t)r & 0xFFFF) < kRevThresh;
            out[i] = (cur << 1) | rev;
            cur = walk_next(cur, r, n_segs);
        }
        uint64_t loop_len = 8192 + rng.next() % 16385;
        uint32_t lcur = base + (uint32_t)(rng.next() % win);
        for (uint64_t j = 0; j < loop_len && i < len; ++j, ++i) {
            uint64_t r = rng.next();
            uint32_t rev = ((uint32_t)r & 0xFFFF) < kRevThresh;
            out[i] = (lcur << 1) | rev;
            lcur = base + ((lcur - base + 1) % win);
        }
    }
}

// Kind 3: what an HPRC-style graph looks like after `odgi sort`: every haplotype walks the
// node ids almost monotonically, taking one side of small bubbles (skip 1 segment w.p. 10 %,
// skip 2-8 w.p. 1 %) and re-entering at a random place when it runs off the end.  Not a
// BASELINE.json config; used to show how the kernels behave on dense, sorted walks.
void gen_sorted_haplotype(uint32_t* out, uint64_t len, uint32_t n_segs, uint64_t seed) {
    SplitMix64 rng(seed);
    uint32_t cur = (uint32_t)(rng.next() % n_segs);
    for (uint64_t i = 0; i < len; ++i) {
        uint64_t r = rng.next();
        uint32_t rev = ((uint32_t)r & 0xFFFF) < kRevThresh;
        out[i] = (cur << 1) | rev;
        uint32_t c = (uint32_t)(r >> 16) & 0xFFFF, hi = (uint32_t)(r >> 32);
        uint32_t adv = c < 58327 ? 1u : c < 64881 ? 2u : 3u + hi % 7u;   // 89 % / 10 % / 1 %
        cur += adv;
        if (cur >= n_segs) cur = (uint32_t)(rng.next() % n_segs);
    }
}

void gen_uniform(uint32_t* out, uint64_t len, uint32_t n_segs, uint64_t seed) {
    SplitMix64 rng(seed);
    for (uint64_t i = 0; i < len; ++i) {
        uint64_t r = rng.next();
        out[i] = ((uint32_t)((r >> 16) % n_segs) << 1) | (((uint32_t)r & 0xFFFF) < kRevThresh);
    }
}

}  // namespace

extern "C" {

// Path lengths that sum to exactly n_steps: equal shares with an optional
// +/- jitter_pct spread (config C uses 20 so that LPT sharding is non-trivial).
// Writes n_paths contiguous spans [start,end) in pool order.  Returns 0, or -1
// if n_steps does not fit the format's u32 ids (flatgfa/src/pool.rs:9-11).
int fgfa_synth_spans(uint32_t n_paths, uint64_t n_steps, uint32_t jitter_pct, uint64_t seed,
                     uint32_t* span_start, uint32_t* span_end) {
    if (n_steps > 0xFFFFFFFFull || n_paths == 0) return -1;
    std::vector<double> w(n_paths);
    SplitMix64 rng(seed ^ 0xC0FFEEull);
    double tot = 0;
    for (uint32_t p = 0; p < n_paths; ++p) {
        double u = (double)(rng.next() >> 11) / 9007199254740992.0;  // [0,1)
        w[p] = 1.0 + (jitter_pct / 100.0) * (2.0 * u - 1.0);
        tot += w[p];
    }
    uint64_t acc = 0;
    double cum = 0;
    for (uint32_t p = 0; p < n_paths; ++p) {
        cum += w[p];
        uint64_t end = (p + 1 == n_paths) ? n_steps : (uint64_t)((cum / tot) * (double)n_steps);
        if (end < acc) end = acc;
        if (end > n_steps) end = n_steps;
        span_start[p] = (uint32_t)acc;
        span_end[p] = (uint32_t)end;
        acc = end;
    }
    return 0;
}

// kind 0: haplotype walk (configs B/C); 1: skewed looping paths (config E);
// 2: uniform-random segment ids (adversarial, worst L2 locality); 3: sorted haplotype walks.
int fgfa_synth_steps(int kind, uint32_t n_segs, uint32_t n_paths, const uint32_t* span_start,
                     const uint32_t* span_end, uint64_t seed, uint32_t* steps_out, int n_threads) {
    if (n_segs == 0 || n_segs > 0x7FFFFFFFu) return -1;
    if (n_threads < 1) n_threads = 1;
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        for (;;) {
            uint32_t p = next.fetch_add(1);
            if (p >= n_paths) break;
            uint64_t len = (uint64_t)span_end[p] - span_start[p];
            uint32_t* out = steps_out + span_start[p];
            uint64_t pseed = seed + p;
            if (kind == 0) gen_walk(out, len, n_segs, pseed);
            else if (kind == 1) gen_skewed(out, len, n_segs, pseed, p, seed);
            else if (kind == 3) gen_sorted_haplotype(out, len, n_segs, pseed);
            else gen_uniform(out, len, n_segs, pseed);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return 0;
}

// One path of a graph, bit-identical to what fgfa_synth_steps writes for path `path_idx`
// of a graph generated with `graph_seed` (paths are seeded individually): lets a rank
// materialise only its shard.
int fgfa_synth_path(int kind, uint32_t n_segs, uint64_t len, uint64_t graph_seed, uint32_t path_idx,
                    uint32_t* out) {
    if (n_segs == 0 || n_segs > 0x7FFFFFFFu) return -1;
    const uint64_t pseed = graph_seed + path_idx;
    if (kind == 0) gen_walk(out, len, n_segs, pseed);
    else if (kind == 1) gen_skewed(out, len, n_segs, pseed, path_idx, graph_seed);
    else if (kind == 3) gen_sorted_haplotype(out, len, n_segs, pseed);
    else gen_uniform(out, len, n_segs, pseed);
    return 0;
}

}  // extern "C"
