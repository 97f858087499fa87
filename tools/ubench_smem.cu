// Shared-memory operation throughput on sm_100a (development tool): what one step of a
// shared-memory privatised histogram may cost.  Every CTA has 1024 threads, one CTA per SM;
// each warp issues `iters` instructions of the given kind; reported: SM cycles per WARP
// instruction (all 32 warps issuing concurrently), i.e. the SM-level throughput.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kWords = 40960;   // 160 KiB of u32 counters (the pair ops use the first 20480 as counters, the rest as masks)
constexpr int kBytes = 40960;   // + 40 KiB of seen bytes

enum Op { kAtomAdd, kStoreU8, kAtomAddStoreU8, kLoad32, kStore32, kAtomOrBitmap, kAtomAddU16, kAtomAddRet, kStoreU8Cond,
          kAtomAddReg, kAtomOrReg, kAtomOrHeads7, kAtomOrHeads7Distinct, kAtomOrHeads3,
          kAtomAddOpaque, kAtomOrRet, kAtomAddOrPair, kAtomAddOrPairRet, kAtomAdd64, kAtomAddOr64 };

// lane offsets: pattern 0 = consecutive segments, 1 = stride ~2.46 (config C walk), 2 = random in window
__device__ __forceinline__ uint32_t lane_off(int pattern, uint32_t lane, uint32_t it) {
    if (pattern == 0) return lane;
    if (pattern == 1) return (lane * 631u) >> 8;            // 2.465 * lane
    uint32_t x = (lane + 1u) * 2654435761u + it * 40503u;   // pseudo-random
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    return x % 30000u;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k_smem(int pattern, int iters, unsigned long long* out, uint32_t* sink) {
    extern __shared__ uint32_t sm[];
    uint32_t* cnt = sm;
    uint8_t* seen = reinterpret_cast<uint8_t*>(sm + kWords);
    for (int i = threadIdx.x; i < kWords + kBytes / 4; i += 1024) sm[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t base = warp * 997u;
    uint32_t acc = 0;
    const uint32_t opaque_one = *sink | 1u;       // *sink is 0
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        uint32_t loc = base + lane_off(pattern, lane, it);
        if (loc >= 40000u) loc -= 40000u;
        if ((OP == kAtomAddOrPair || OP == kAtomAddOrPairRet) && loc >= 20000u) loc -= 20000u;
        if (OP == kAtomAdd || OP == kAtomAddStoreU8) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + loc);
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
        }
        if (OP == kAtomAddRet) acc += atomicAdd(cnt + loc, 1u);
        if (OP == kStoreU8 || OP == kAtomAddStoreU8) seen[loc] = 1;
        if (OP == kStoreU8Cond) { if (seen[loc] == 0) seen[loc] = 1; }
        if (OP == kLoad32) acc += cnt[loc];
        if (OP == kStore32) cnt[loc] = it;
        if (OP == kAtomOrBitmap) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + (loc >> 5));
            asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(1u << (loc & 31)) : "memory");
        }
        if (OP == kAtomAddReg) {                      // what kernel W issues: register operand (ATOMS.ADD, not POPC.INC)
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + loc);
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"((uint32_t)iters >> 30 | 1u) : "memory");
        }
        if (OP == kAtomOrReg) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + loc);
            asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(1u << (it & 31)) : "memory");
        }
        if (OP == kAtomOrHeads7 || OP == kAtomOrHeads7Distinct || OP == kAtomOrHeads3) {
            // run heads of a row: few active lanes; same-word lanes (bit rows) or distinct words
            const bool head = OP == kAtomOrHeads3 ? (lane % 11u == 0u) : (lane % 5u == 0u);
            if (head) {
                const uint32_t w = OP == kAtomOrHeads7Distinct ? loc : (loc >> 5);
                const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + w);
                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(0x1Fu << (loc & 31)) : "memory");
            }
        }
        if (OP == kAtomAddOpaque) {                   // ATOMS.ADD RZ, [a], R: operand ptxas cannot fold
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + loc);
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(opaque_one) : "memory");
        }
        if (OP == kAtomOrRet) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + loc);
            uint32_t old;
            asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(1u << (it & 31)) : "memory");
            acc += old;
        }
        if (OP == kAtomAddOrPair) {                   // kernel W's pair: counter word + mask word
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + loc);
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(opaque_one) : "memory");
            asm volatile("red.shared.or.b32 [%0+81920], %1;" ::"r"(a), "r"(1u << (it & 31)) : "memory");
        }
        if (OP == kAtomAddOrPairRet) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + loc);
            uint32_t o1, o2;
            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o1) : "r"(a), "r"(opaque_one) : "memory");
            asm volatile("atom.shared.or.b32 %0, [%1+81920], %2;" : "=r"(o2) : "r"(a), "r"(1u << (it & 31)) : "memory");
            acc += o1 + o2;
        }
        if (OP == kAtomAdd64 || OP == kAtomAddOr64) {   // lane l owns the aligned counter pair (2l, 2l+1) of a 64-segment row
            const uint32_t pair = ((base & ~1u) + 2u * lane) % 20000u & ~1u;
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + pair);
            const unsigned long long v = (unsigned long long)opaque_one | ((unsigned long long)(it & 1) << 32);
            asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
            if (OP == kAtomAddOr64) asm volatile("red.shared.or.b64 [%0+81920], %1;" ::"r"(a), "l"(v << 3) : "memory");
        }
        if (OP == kAtomAddU16) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(cnt + (loc >> 1));
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(1u << ((loc & 1) << 4)) : "memory");
        }
        base += 79u;
        if (base >= 40000u) base -= 40000u;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) *sink = acc;
}

template <int OP>
void run(const char* name, int sms, unsigned long long* d_out, uint32_t* d_sink) {
    const size_t smem = (size_t)kWords * 4 + kBytes;
    CK(cudaFuncSetAttribute(k_smem<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4096;
    for (int pattern = 0; pattern < 3; ++pattern) {
        k_smem<OP><<<sms, 1024, smem>>>(pattern, iters, d_out, d_sink);
        k_smem<OP><<<sms, 1024, smem>>>(pattern, iters, d_out, d_sink);
        CK(cudaDeviceSynchronize());
        unsigned long long h[256];
        CK(cudaMemcpy(h, d_out, sms * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i];
        avg /= sms;
        printf("%-26s pattern %d (%s): %.2f cycles per warp instruction (SM level, 32 warps)\n", name, pattern,
               pattern == 0 ? "consecutive" : pattern == 1 ? "stride 2.46" : "random     ", avg / ((double)iters * 32));
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned long long* d_out;
    uint32_t* d_sink;
    CK(cudaMalloc(&d_out, 256 * 8));
    CK(cudaMalloc(&d_sink, 4));
    CK(cudaMemset(d_sink, 0, 4));
    run<kLoad32>("LDS.32", sms, d_out, d_sink);
    run<kStore32>("STS.32", sms, d_out, d_sink);
    run<kStoreU8>("STS.U8", sms, d_out, d_sink);
    run<kStoreU8Cond>("LDS.U8 + cond STS.U8", sms, d_out, d_sink);
    run<kAtomAdd>("red.shared.add.u32", sms, d_out, d_sink);
    run<kAtomAddRet>("atom.shared.add.u32 (ret)", sms, d_out, d_sink);
    run<kAtomAddStoreU8>("red.add.u32 + STS.U8", sms, d_out, d_sink);
    run<kAtomOrBitmap>("red.shared.or bitmap", sms, d_out, d_sink);
    run<kAtomAddU16>("red.add packed u16", sms, d_out, d_sink);
    run<kAtomAddReg>("red.add.u32 reg operand", sms, d_out, d_sink);
    run<kAtomOrReg>("red.or.b32 32 lanes", sms, d_out, d_sink);
    run<kAtomOrHeads7>("red.or 7 head lanes, bit rows", sms, d_out, d_sink);
    run<kAtomOrHeads7Distinct>("red.or 7 lanes, distinct words", sms, d_out, d_sink);
    run<kAtomOrHeads3>("red.or 3 head lanes, bit rows", sms, d_out, d_sink);
    run<kAtomAddOpaque>("ATOMS.ADD RZ (reg operand)", sms, d_out, d_sink);
    run<kAtomOrRet>("ATOMS.OR with return", sms, d_out, d_sink);
    run<kAtomAddOrPair>("ATOMS.ADD RZ + ATOMS.OR RZ", sms, d_out, d_sink);
    run<kAtomAddOrPairRet>("ATOMS.ADD ret + ATOMS.OR ret", sms, d_out, d_sink);
    run<kAtomAdd64>("ATOMS.ADD.64 lane pairs", sms, d_out, d_sink);
    run<kAtomAddOr64>("ATOMS.ADD.64 + ATOMS.OR.64", sms, d_out, d_sink);
    return 0;
}
