// Host side + C ABI of the GPU step-list tokenizer (see tokenize_kernels.cuh and
// include/fgfa_depth.h).  Replaces, for well-formed input, the loop of
// `Parser::add_path` (flatgfa/src/parse.rs:149-156) over `StepsParser` (gfaline.rs:201-263).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/fgfa_depth.h"
#include "tokenize_kernels.cuh"

struct fgfa_tokenizer {
    uint64_t n_bytes = 0;
    uint32_t n_fields = 0;
    std::vector<uint64_t> off, len;
    std::vector<fgfa::TokTile> tiles;
    std::vector<uint32_t> counts;          // steps per field
    uint64_t n_steps = 0;
    uint8_t* d_text = nullptr;
    fgfa::TokTile* d_tiles = nullptr;
    uint64_t *d_field_off = nullptr, *d_field_end = nullptr, *d_tile_base = nullptr;
    uint32_t *d_tile_commas = nullptr, *d_steps = nullptr, *d_err = nullptr;
    uint64_t* d_hash_keys = nullptr;
    uint32_t* d_hash_vals = nullptr;
    int sms = 1;
};

namespace {
int rc_of(cudaError_t e) {
    if (e == cudaSuccess) return FGFA_OK;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return FGFA_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return FGFA_ERR_NOMEM;
    return FGFA_ERR_CUDA;
}
#define CT(x) do { int rc_ = rc_of(x); if (rc_) return rc_; } while (0)

fgfa::TokParams params_of(const fgfa_tokenizer* t) {
    fgfa::TokParams P{};
    P.text = t->d_text;
    P.n_bytes = t->n_bytes;
    P.tiles = t->d_tiles;
    P.n_tiles = (uint32_t)t->tiles.size();
    P.field_off = t->d_field_off;
    P.field_end = t->d_field_end;
    P.tile_commas = t->d_tile_commas;
    P.tile_base = t->d_tile_base;
    P.steps = t->d_steps;
    P.err = t->d_err;
    return P;
}
}  // namespace

extern "C" {

void fgfa_tokenizer_destroy(fgfa_tokenizer_t* t) {
    if (!t) return;
    cudaFree(t->d_text); cudaFree(t->d_tiles); cudaFree(t->d_field_off); cudaFree(t->d_field_end);
    cudaFree(t->d_tile_base); cudaFree(t->d_tile_commas); cudaFree(t->d_steps); cudaFree(t->d_err);
    cudaFree(t->d_hash_keys); cudaFree(t->d_hash_vals);
    delete t;
}

int fgfa_tokenizer_create(fgfa_tokenizer_t** out, const uint8_t* h_text, uint64_t n_bytes,
                          const uint64_t* field_off, const uint64_t* field_len, uint32_t n_fields) {
    if (!out || (n_bytes && !h_text) || (n_fields && (!field_off || !field_len))) return FGFA_ERR_INVALID_ARG;
    *out = nullptr;
    if (fgfa_device_count() <= 0) return FGFA_ERR_NO_DEVICE;
    for (uint32_t f = 0; f < n_fields; ++f)
        if (field_off[f] > n_bytes || field_len[f] > n_bytes - field_off[f]) return FGFA_ERR_INVALID_ARG;
    fgfa_tokenizer* t = new (std::nothrow) fgfa_tokenizer();
    if (!t) return FGFA_ERR_NOMEM;
    struct Guard { fgfa_tokenizer* t; ~Guard() { if (t) fgfa_tokenizer_destroy(t); } } guard{t};
    t->n_bytes = n_bytes;
    t->n_fields = n_fields;
    t->off.assign(field_off, field_off + n_fields);
    t->len.assign(field_len, field_len + n_fields);
    int dev = 0;
    CT(cudaGetDevice(&dev));
    CT(cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, dev));
    std::vector<uint64_t> fend(n_fields);
    for (uint32_t f = 0; f < n_fields; ++f) {
        fend[f] = field_off[f] + field_len[f];
        for (uint64_t o = 0; o < field_len[f]; o += fgfa::kTokTile) {
            if (t->tiles.size() >= 0xFFFFFFFEull) return FGFA_ERR_TOO_LARGE;
            t->tiles.push_back(fgfa::TokTile{field_off[f] + o, (uint32_t)std::min<uint64_t>(fgfa::kTokTile, field_len[f] - o), f});
        }
    }
    const size_t n_tiles = t->tiles.size();
    CT(cudaMalloc(&t->d_text, std::max<uint64_t>(n_bytes, 16)));
    CT(cudaMalloc(&t->d_tiles, std::max<size_t>(n_tiles, 1) * sizeof(fgfa::TokTile)));
    CT(cudaMalloc(&t->d_field_off, std::max<size_t>(n_fields, 1) * 8));
    CT(cudaMalloc(&t->d_field_end, std::max<size_t>(n_fields, 1) * 8));
    CT(cudaMalloc(&t->d_tile_base, std::max<size_t>(n_tiles, 1) * 8));
    CT(cudaMalloc(&t->d_tile_commas, std::max<size_t>(n_tiles, 1) * 4));
    CT(cudaMalloc(&t->d_err, 4));
    CT(cudaMemset(t->d_err, 0, 4));
    if (n_bytes) CT(cudaMemcpy(t->d_text, h_text, n_bytes, cudaMemcpyHostToDevice));
    if (n_tiles) CT(cudaMemcpy(t->d_tiles, t->tiles.data(), n_tiles * sizeof(fgfa::TokTile), cudaMemcpyHostToDevice));
    if (n_fields) {
        CT(cudaMemcpy(t->d_field_off, field_off, (size_t)n_fields * 8, cudaMemcpyHostToDevice));
        CT(cudaMemcpy(t->d_field_end, fend.data(), (size_t)n_fields * 8, cudaMemcpyHostToDevice));
    }
    // T1: commas per tile -> steps per field (commas + 1 for a non-empty field) and tile bases
    std::vector<uint32_t> commas(n_tiles, 0);
    if (n_tiles) {
        const uint32_t grid = (uint32_t)std::min<size_t>(n_tiles, (size_t)t->sms * 16);
        fgfa::k_steps_count<<<grid, fgfa::kTokThreads>>>(params_of(t));
        CT(cudaGetLastError());
        CT(cudaMemcpy(commas.data(), t->d_tile_commas, n_tiles * 4, cudaMemcpyDeviceToHost));
    }
    t->counts.assign(n_fields, 0);
    std::vector<uint64_t> tile_base(n_tiles, 0);
    uint64_t total = 0;
    size_t k = 0;
    for (uint32_t f = 0; f < n_fields; ++f) {
        uint64_t in_field = 0;
        for (; k < n_tiles && t->tiles[k].field == f; ++k) {
            tile_base[k] = total + in_field;
            in_field += commas[k];
        }
        const uint64_t steps = field_len[f] ? in_field + 1 : 0;
        if (total + steps > 0xFFFFFFFFull) return FGFA_ERR_TOO_LARGE;   // pool.rs:51-53 "id too large"
        t->counts[f] = (uint32_t)steps;
        total += steps;
    }
    t->n_steps = total;
    CT(cudaMalloc(&t->d_steps, std::max<uint64_t>(total, 4) * 4));
    if (n_tiles) CT(cudaMemcpy(t->d_tile_base, tile_base.data(), n_tiles * 8, cudaMemcpyHostToDevice));
    guard.t = nullptr;
    *out = t;
    return FGFA_OK;
}

int fgfa_tokenizer_spans(const fgfa_tokenizer_t* t, uint32_t* span_start, uint32_t* span_end, uint64_t* n_steps) {
    if (!t) return FGFA_ERR_INVALID_ARG;
    uint64_t acc = 0;
    for (uint32_t f = 0; f < t->n_fields; ++f) {
        if (span_start) span_start[f] = (uint32_t)acc;
        acc += t->counts[f];
        if (span_end) span_end[f] = (uint32_t)acc;
    }
    if (n_steps) *n_steps = t->n_steps;
    return FGFA_OK;
}

int fgfa_tokenizer_parse(fgfa_tokenizer_t* t, uint64_t sequential_max, const uint64_t* other_names,
                         const uint32_t* other_ids, uint32_t n_others, uint32_t* h_steps_out) {
    if (!t || (n_others && (!other_names || !other_ids))) return FGFA_ERR_INVALID_ARG;
    fgfa::TokParams P = params_of(t);
    P.sequential_max = sequential_max;
    if (n_others) {
        uint32_t cap = 16;
        while (cap < 2ull * n_others) cap <<= 1;
        std::vector<uint64_t> keys(cap, 0);
        std::vector<uint32_t> vals(cap, 0);
        for (uint32_t i = 0; i < n_others; ++i) {
            const uint64_t h = (other_names[i] + 1) * 0x9E3779B97F4A7C15ull;
            for (uint32_t probe = 0;; ++probe) {
                const uint32_t slot = (uint32_t)((h >> 32) + probe) & (cap - 1);
                if (keys[slot] == 0 || keys[slot] == other_names[i] + 1) { keys[slot] = other_names[i] + 1; vals[slot] = other_ids[i]; break; }
            }
        }
        cudaFree(t->d_hash_keys); cudaFree(t->d_hash_vals);
        t->d_hash_keys = nullptr; t->d_hash_vals = nullptr;
        CT(cudaMalloc(&t->d_hash_keys, (size_t)cap * 8));
        CT(cudaMalloc(&t->d_hash_vals, (size_t)cap * 4));
        CT(cudaMemcpy(t->d_hash_keys, keys.data(), (size_t)cap * 8, cudaMemcpyHostToDevice));
        CT(cudaMemcpy(t->d_hash_vals, vals.data(), (size_t)cap * 4, cudaMemcpyHostToDevice));
        P.hash_keys = t->d_hash_keys;
        P.hash_vals = t->d_hash_vals;
        P.hash_mask = cap - 1;
    }
    CT(cudaMemset(t->d_err, 0, 4));
    if (P.n_tiles) {
        const uint32_t grid = (uint32_t)std::min<size_t>(P.n_tiles, (size_t)t->sms * 16);
        fgfa::k_steps_parse<<<grid, fgfa::kTokThreads>>>(P);
        CT(cudaGetLastError());
    }
    uint32_t err = 0;
    CT(cudaMemcpy(&err, t->d_err, 4, cudaMemcpyDeviceToHost));
    if (err) return FGFA_ERR_PARSE;
    if (h_steps_out && t->n_steps) CT(cudaMemcpy(h_steps_out, t->d_steps, t->n_steps * 4, cudaMemcpyDeviceToHost));
    return FGFA_OK;
}

const uint32_t* fgfa_tokenizer_device_steps(const fgfa_tokenizer_t* t) { return t ? t->d_steps : nullptr; }

}  // extern "C"
