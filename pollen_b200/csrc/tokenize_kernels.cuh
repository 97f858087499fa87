// sm_100a kernels for GFA path step lists: text `1+,23-,4+` -> Handle words.
//
// GPU form of the reference's `StepsParser` (flatgfa/src/gfaline.rs:201-263) composed with
// `NameMap::get` (flatgfa/src/namemap.rs:27-33) and `Handle::new` (flatgfa/src/flatgfa.rs:192-198),
// i.e. the inner loop of `Parser::add_path` (flatgfa/src/parse.rs:149-156), which dominates
// `fgfa -I big.gfa ...` on a pangenome.
//
// Only the strict grammar  field := token (',' token)* ; token := digit+ ('+'|'-')  is accepted
// here; anything else raises the error flag and the caller re-parses that input with the host
// parser, which reproduces the reference's quirks (dropped trailing number, swallowed stray
// byte) and error messages.
//
// Work decomposition: a *tile* is up to kTokTile bytes of ONE field (a path's step list).
//   kernel T1 (k_steps_count)  counts the commas of every tile;
//   host                       turns them into per-field step counts and per-tile output bases;
//   kernel T2 (k_steps_parse)  re-reads each tile, finds every token start, parses the number,
//                              maps the segment name to its pool index and writes the Handle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fgfa {

constexpr int kTokThreads = 128;
constexpr int kTokBytesPerThread = 32;
constexpr int kTokTile = kTokThreads * kTokBytesPerThread;   // 4096 bytes of text per CTA
constexpr int kTokSlack = 32;                                 // look-ahead for a token that crosses the tile end

struct TokTile {
    uint64_t off;        // byte offset of the tile in the text buffer
    uint32_t len;        // bytes in this tile (<= kTokTile)
    uint32_t field;      // field (path) index
};

struct TokParams {
    const uint8_t* __restrict__ text;
    uint64_t n_bytes;
    const TokTile* __restrict__ tiles;
    uint32_t n_tiles;
    const uint64_t* __restrict__ field_end;   // [n_fields] byte offset one past the field
    const uint64_t* __restrict__ field_off;   // [n_fields] byte offset of the field
    uint32_t* __restrict__ tile_commas;       // T1 out: commas per tile
    const uint64_t* __restrict__ tile_base;   // T2 in: index in `steps` of the path's first step + commas in the path's earlier tiles
    uint32_t* __restrict__ steps;             // T2 out
    uint64_t sequential_max;                  // namemap.rs:11-12
    const uint64_t* __restrict__ hash_keys;   // open addressing, key+1 stored, 0 = empty (may be null)
    const uint32_t* __restrict__ hash_vals;
    uint32_t hash_mask;                       // capacity - 1 (capacity is a power of two), 0 = no table
    uint32_t* __restrict__ err;               // bit 0 grammar, bit 1 unknown segment name, bit 2 index too large
};

// Stage a tile (+ look-ahead) into shared memory with coalesced byte-granular loads.
__device__ __forceinline__ void tok_stage(uint8_t* s, const TokParams& P, const TokTile& t, uint64_t fend) {
    const uint64_t want_end = min(t.off + t.len + kTokSlack, fend);
    const uint32_t n = (uint32_t)(want_end - t.off);
    for (uint32_t i = threadIdx.x; i < (uint32_t)(kTokTile + kTokSlack); i += kTokThreads)
        s[i] = i < n ? P.text[t.off + i] : (uint8_t)0;
}

__device__ __forceinline__ uint32_t count_commas32(const uint8_t* p, uint32_t n) {
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < kTokBytesPerThread; ++i) c += ((uint32_t)i < n && p[i] == ',') ? 1u : 0u;
    return c;
}

__global__ void __launch_bounds__(kTokThreads) k_steps_count(TokParams P) {
    __shared__ uint8_t s[kTokTile + kTokSlack];
    __shared__ uint32_t s_warp[kTokThreads / 32];
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const TokTile t = P.tiles[tile];
        tok_stage(s, P, t, P.field_end[t.field]);
        __syncthreads();
        const uint32_t r0 = threadIdx.x * kTokBytesPerThread;
        uint32_t c = r0 < t.len ? count_commas32(s + r0, t.len - r0) : 0u;
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < kTokThreads / 32; ++w) tot += s_warp[w];
            P.tile_commas[tile] = tot;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ bool tok_lookup(const TokParams& P, uint64_t name, uint32_t& id) {
    if (name >= 1 && name <= P.sequential_max) { id = (uint32_t)(name - 1); return true; }   // namemap.rs:28-29
    if (P.hash_mask == 0 && P.hash_keys == nullptr) return false;
    uint64_t h = (name + 1) * 0x9E3779B97F4A7C15ull;
    for (uint32_t probe = 0; probe <= P.hash_mask; ++probe) {
        const uint32_t slot = (uint32_t)((h >> 32) + probe) & P.hash_mask;
        const uint64_t k = P.hash_keys[slot];
        if (k == name + 1) { id = P.hash_vals[slot]; return true; }
        if (k == 0) return false;
    }
    return false;
}

__global__ void __launch_bounds__(kTokThreads) k_steps_parse(TokParams P) {
    __shared__ uint8_t s[kTokTile + kTokSlack];
    __shared__ uint32_t s_warp[kTokThreads / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const TokTile t = P.tiles[tile];
        const uint64_t fbeg = P.field_off[t.field], fend = P.field_end[t.field];
        tok_stage(s, P, t, fend);
        __syncthreads();
        const uint32_t r0 = threadIdx.x * kTokBytesPerThread;
        const uint32_t mine = r0 < t.len ? count_commas32(s + r0, t.len - r0) : 0u;
        // exclusive prefix of comma counts over the block
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = incl - mine;
        for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
        // a token that starts at byte p is the k-th step of its path, k = commas of the field
        // before p = (commas in earlier tiles, folded into tile_base by the host) + (commas of
        // this tile before p)
        const uint64_t base = P.tile_base[tile];
        const bool tile_starts_token = (t.off == fbeg) || (P.text[t.off - 1] == ',');
        if (r0 < t.len) {
            const uint32_t n = min((uint32_t)kTokBytesPerThread, t.len - r0);
            uint32_t seen = before;                     // commas strictly before the current byte
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t p = r0 + i;
                const bool starts = (p == 0) ? tile_starts_token : (s[p - 1] == ',');
                if (starts) {
                    // ---- one token: digit+ sign, then ',' or the end of the field ----
                    uint64_t name = 0;
                    uint32_t q = p, digits = 0;
                    const uint32_t avail = (uint32_t)min((uint64_t)(kTokTile + kTokSlack), fend - t.off);
                    while (q < avail && s[q] >= '0' && s[q] <= '9') { name = name * 10 + (uint64_t)(s[q] - '0'); ++q; ++digits; }
                    const uint8_t sign = q < avail ? s[q] : (uint8_t)0;
                    const bool ends_ok = (t.off + q + 1 == fend) || (q + 1 < avail && s[q + 1] == ',');
                    if (digits == 0 || digits > 19 || (sign != '+' && sign != '-') || !ends_ok) {
                        atomicOr(P.err, 1u);
                    } else {
                        uint32_t id;
                        if (!tok_lookup(P, name, id)) atomicOr(P.err, 2u);
                        else if (id & 0x80000000u) atomicOr(P.err, 4u);           // flatgfa.rs:194
                        else {
                            P.steps[base + seen] = (id << 1) | (sign == '-' ? 1u : 0u);
                        }
                    }
                }
                if (s[p] == ',') {
                    ++seen;
                    if (t.off + p + 1 == fend) atomicOr(P.err, 1u);   // trailing comma: no token follows
                } else if (!((s[p] >= '0' && s[p] <= '9') || s[p] == '+' || s[p] == '-')) atomicOr(P.err, 1u);
            }
        }
        __syncthreads();
    }
}

}  // namespace fgfa
