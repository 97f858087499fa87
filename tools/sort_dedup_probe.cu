// Measurement probe for the north-star's alternative formulation of depth.uniq
// (BASELINE.json: "a CUB-style segmented radix sort plus adjacent-difference dedup"):
// per-path segmented radix sort of the segment ids (cub::DeviceSegmentedRadixSort, a
// library call -- this is a probe, not product code), then one thread per sorted element
// adds 1 to uniq[seg] where the element differs from its predecessor inside its path.
// Prints the time of both stages and checks uniq against the oracle.
//   sort_dedup_probe <B|C>
#include <cub/cub.cuh>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

extern "C" {
int fgfa_synth_spans(uint32_t, uint64_t, uint32_t, uint64_t, uint32_t*, uint32_t*);
int fgfa_synth_steps(int, uint32_t, uint32_t, const uint32_t*, const uint32_t*, uint64_t, uint32_t*, int);
int oracle_seg_depth_with_uniq(const uint32_t*, uint64_t, const uint32_t*, uint32_t, uint32_t, uint64_t*, uint64_t*);
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void k_adjacent_dedup(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ seg_of_elem_start,
                                 const uint32_t* __restrict__ starts, const uint32_t* __restrict__ ends, uint32_t n_paths,
                                 uint64_t n, uint32_t* __restrict__ uniq) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t seg = sorted[i] >> 1;
    // first element of a path: binary search the span table (n_paths is small)
    uint32_t lo = 0, hi = n_paths;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (starts[mid] <= i) lo = mid; else hi = mid; }
    const bool first = (i == starts[lo]);
    if (first || (sorted[i - 1] >> 1) != seg) atomicAdd(uniq + seg, 1u);
}

int main(int argc, char** argv) {
    std::string which = argc > 1 ? argv[1] : "B";
    uint32_t n_segs = which == "C" ? 5000000 : 1000000, n_paths = which == "C" ? 90 : 16;
    uint64_t n_steps = which == "C" ? 400000000ull : 20000000ull;
    uint32_t jitter = which == "C" ? 20 : 0;
    std::vector<uint32_t> ss(n_paths), se(n_paths), steps(n_steps);
    fgfa_synth_spans(n_paths, n_steps, jitter, 0xB1011054ull, ss.data(), se.data());
    fgfa_synth_steps(0, n_segs, n_paths, ss.data(), se.data(), 0xB1011054ull, steps.data(), 16);
    uint32_t *d_in, *d_out, *d_ss, *d_se, *d_uniq;
    CK(cudaMalloc(&d_in, n_steps * 4)); CK(cudaMalloc(&d_out, n_steps * 4));
    CK(cudaMalloc(&d_ss, n_paths * 4)); CK(cudaMalloc(&d_se, n_paths * 4)); CK(cudaMalloc(&d_uniq, (size_t)n_segs * 4));
    CK(cudaMemcpy(d_in, steps.data(), n_steps * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ss, ss.data(), n_paths * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_se, se.data(), n_paths * 4, cudaMemcpyHostToDevice));
    int end_bit = 1; while ((1ull << end_bit) < 2ull * n_segs) ++end_bit;
    size_t temp_bytes = 0;
    CK(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, temp_bytes, d_in, d_out, (int64_t)n_steps, (int)n_paths, d_ss, d_se, 1, end_bit));
    void* d_temp; CK(cudaMalloc(&d_temp, temp_bytes));
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    float best_sort = 1e30f, best_dedup = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaMemset(d_uniq, 0, (size_t)n_segs * 4));
        CK(cudaEventRecord(e0));
        CK(cub::DeviceSegmentedRadixSort::SortKeys(d_temp, temp_bytes, d_in, d_out, (int64_t)n_steps, (int)n_paths, d_ss, d_se, 1, end_bit));
        CK(cudaEventRecord(e1));
        k_adjacent_dedup<<<(unsigned)((n_steps + 255) / 256), 256>>>(d_out, nullptr, d_ss, d_se, n_paths, n_steps, d_uniq);
        CK(cudaEventRecord(e2));
        CK(cudaEventSynchronize(e2));
        float a, b; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
        best_sort = std::min(best_sort, a); best_dedup = std::min(best_dedup, b);
    }
    std::vector<uint32_t> g(n_segs), spans(2 * n_paths);
    CK(cudaMemcpy(g.data(), d_uniq, (size_t)n_segs * 4, cudaMemcpyDeviceToHost));
    for (uint32_t p = 0; p < n_paths; ++p) { spans[2 * p] = ss[p]; spans[2 * p + 1] = se[p]; }
    std::vector<uint64_t> od(n_segs), ou(n_segs);
    oracle_seg_depth_with_uniq(steps.data(), n_steps, spans.data(), n_paths, n_segs, od.data(), ou.data());
    uint64_t bad = 0; for (uint32_t i = 0; i < n_segs; ++i) bad += g[i] != ou[i];
    printf("{\"config\": \"%s\", \"uniq_by_segmented_radix_sort_ms\": %.3f, \"adjacent_dedup_ms\": %.3f, \"temp_bytes\": %zu, \"uniq_mismatches\": %llu}\n",
           which.c_str(), best_sort, best_dedup, temp_bytes, (unsigned long long)bad);
    return 0;
}
