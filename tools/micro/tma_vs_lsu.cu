// Micro-benchmark (development evidence for DESIGN.md 4.1): does shared-memory staging by the bulk-copy engine
// (cp.async.bulk global->shared) take cycles from the LSU data pipe that LDS gathers saturate?
//   mode 0: gathers only                      (every warp: conflict-free LDS.32 in a dependent-free loop)
//   mode 1: gathers + bulk copies             (one elected thread streams rows into a second buffer, mbarrier completion)
//   mode 2: gathers + LDG.128/STS.128 staging (one warp copies the same bytes through registers)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_vs_lsu tma_vs_lsu.cu && ./tma_vs_lsu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ inline uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kThreads = 512;
constexpr int kGatherWords = 16 * 1024;      // 64 KB
constexpr int kStageBytes = 64 * 1024;
constexpr int kRow = 4096;                   // bytes per bulk copy

__global__ void __launch_bounds__(kThreads, 1) k_bench(int mode, int iters, const uint8_t *__restrict__ src, size_t src_bytes,
                                                      unsigned *out, unsigned long long *copied, unsigned long long *gclk) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *g = reinterpret_cast<uint32_t *>(smem);
    uint8_t *stage = smem + kGatherWords * 4;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kGatherWords; i += kThreads) g[i] = i * 2654435761u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t *my_src = src + ((size_t)blockIdx.x * (1 << 20)) % (src_bytes - (2 << 20));
    unsigned acc = 0;
    unsigned long long moved = 0;
    const bool copier_thread = (mode == 1 && tid == kThreads - 32);
    const bool copier_warp = (mode == 2 && warp == kThreads / 32 - 1);
    if (copier_thread) {
        unsigned parity = 0;
        size_t off = 0;
        for (int it = 0; it < iters; it++) {
            // 16 rows of 4 KB per round = the whole staging buffer
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(kStageBytes) : "memory");
            for (int r = 0; r < kStageBytes / kRow; r++) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(stage + r * kRow)), "l"(my_src + off), "r"(kRow), "r"(smem_u32(&bar)) : "memory");
                off = (off + kRow) & ((1 << 20) - 1);
            }
            asm volatile(
                "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(
                    smem_u32(&bar)), "r"(parity) : "memory");
            parity ^= 1;
            moved += kStageBytes;
        }
    } else if (copier_warp) {
        size_t off = 0;
        for (int it = 0; it < iters; it++) {
            for (int r = 0; r < kStageBytes / 512 / 4; r++) {      // 4 x 512 B per pass: four loads in flight per lane
                uint4 v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) v[q] = __ldg(reinterpret_cast<const uint4 *>(my_src + off + q * 512) + lane);
#pragma unroll
                for (int q = 0; q < 4; q++) reinterpret_cast<uint4 *>(stage + (r * 4 + q) * 512)[lane] = v[q];
                off = (off + 2048) & ((1 << 20) - 1);
            }
            moved += kStageBytes;
        }
    } else if (warp < kThreads / 32 - 1) {
        // gather (warps 0..14 in every mode): 16 independent conflict-free LDS.32 per iteration (bank = lane)
        int base = (warp * 977) & (kGatherWords - 1);
        const long long t0 = clock64();
        for (int it = 0; it < iters * 8; it++) {
#pragma unroll
            for (int q = 0; q < 16; q++) acc += g[(base + q * 32 * 3 + lane) & (kGatherWords - 1)];
            base = (base + 1567) & (kGatherWords - 1);
        }
        const long long t1 = clock64();
        if (lane == 0) atomicMax(gclk, (unsigned long long)(t1 - t0));
    }
    if (mode && (copier_thread || (copier_warp && lane == 0))) atomicAdd(copied, moved);
    // keep the staged bytes alive
    if (tid < 32) acc += reinterpret_cast<uint32_t *>(stage)[tid];
    atomicAdd(out, acc);
}

int main() {
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const size_t src_bytes = (size_t)512 << 20;
    uint8_t *src;
    unsigned *out;
    unsigned long long *copied, *gclk;
    cudaMalloc(&src, src_bytes);
    cudaMemset(src, 1, src_bytes);
    cudaMalloc(&out, 4);
    cudaMalloc(&copied, 8);
    cudaMalloc(&gclk, 8);
    const size_t smem = kGatherWords * 4 + kStageBytes;
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 400;
    for (int mode = 0; mode < 3; mode++) {
        cudaMemset(copied, 0, 8);
        k_bench<<<prop.multiProcessorCount, kThreads, smem>>>(mode, iters, src, src_bytes, out, copied, gclk);
        cudaDeviceSynchronize();
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaMemset(copied, 0, 8);
        cudaMemset(gclk, 0, 8);
        cudaEventRecord(a);
        k_bench<<<prop.multiProcessorCount, kThreads, smem>>>(mode, iters, src, src_bytes, out, copied, gclk);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        unsigned long long c;
        cudaMemcpy(&c, copied, 8, cudaMemcpyDeviceToHost);
        const double lds_eff = (double)prop.multiProcessorCount * (kThreads / 32 - 1) * iters * 8 * 16;   // warp-level LDS.32 = wavefronts
        unsigned long long gc;
        cudaMemcpy(&gc, gclk, 8, cudaMemcpyDeviceToHost);
        printf("mode %d: slowest gather warp %llu clk = %.3f wavefronts/clk/SM | ", mode, gc, (double)(kThreads / 32 - 1) * iters * 8 * 16 / (double)gc);
        printf("mode %d: %.3f ms  gather wavefronts/clk/SM %.3f (at 1.965 GHz)  staged %.1f GB/s per SM-aggregate %.1f GB/s (%s)\n", mode, ms,
               lds_eff / prop.multiProcessorCount / (ms * 1e-3 * 1.965e9), c / (ms * 1e-3) / 1e9 / prop.multiProcessorCount,
               c / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
