// Micro-benchmark (development evidence for DESIGN.md 4.1): can the weight-table gathers of k_warp_fused leave the LSU
// data pipe through the texture path?  Every warp runs the kernel's mix -- 16 conflict-free LDS.32 "pixel" gathers + 2
// 16-byte "weight" gathers with a random phase per lane -- with the weights taken
//   mode 0: from a 32 KB table in shared memory (LDS.128 x 2: bank conflicts like the real kernel)
//   mode 1: through a texture object over the same table in global memory (tex1Dfetch<uint4> x 2), table out of shared memory
//   mode 2: through ld.global.nc (LDG.128 x 2, L1-cached)
//   mode 3: no weight gathers at all (the bound of what moving them could gain)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tex_weights tex_weights.cu && ./tex_weights
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kThreads = 896;                 // like k_warp_fused: 7 groups of 128
constexpr int kPixWords = 40 * 1024;          // 160 KB of "footprints"
constexpr int kTabBytes = 32 * 1024;

__global__ void __launch_bounds__(kThreads, 1) k_mix(int mode, int iters, cudaTextureObject_t tex, const uint4 *__restrict__ gtab,
                                                    unsigned *out, unsigned long long *clk) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *pix = reinterpret_cast<uint32_t *>(smem);
    uint4 *stab = reinterpret_cast<uint4 *>(smem + kPixWords * 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kPixWords; i += kThreads) pix[i] = i * 2654435761u;
    if (mode == 0) for (int i = tid; i < kTabBytes / 16; i += kThreads) stab[i] = gtab[i];
    __syncthreads();
    unsigned acc = 0, rng = tid * 747796405u + 2891336453u;
    int base = (warp * 1409) & (kPixWords - 1);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        rng = rng * 1664525u + 1013904223u;
        const int slot = (rng >> 12) & 1023;                 // phase (ax, ay) of this lane's pixel
        uint4 wa = make_uint4(0, 0, 0, 0), wb = wa;
        if (mode == 0) { wa = stab[slot]; wb = stab[1024 + slot]; }
        else if (mode == 1) { wa = tex1Dfetch<uint4>(tex, slot); wb = tex1Dfetch<uint4>(tex, 1024 + slot); }
        else if (mode == 2) { wa = __ldg(gtab + slot); wb = __ldg(gtab + 1024 + slot); }
        unsigned s = 0;
#pragma unroll
        for (int q = 0; q < 16; q++) s += pix[(base + q * 96 + lane) & (kPixWords - 1)];
        acc += s * (wa.x ^ wb.y) + (wa.z ^ wb.w ^ wa.y ^ wb.x ^ wa.w ^ wb.z);
        base = (base + 1567) & (kPixWords - 1);
    }
    const long long t1 = clock64();
    if (lane == 0) atomicMax(clk, (unsigned long long)(t1 - t0));
    atomicAdd(out, acc);
}

int main() {
    cudaSetDevice(0);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    uint4 *gtab;
    cudaMalloc(&gtab, kTabBytes);
    cudaMemset(gtab, 3, kTabBytes);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = gtab;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
    rd.res.linear.sizeInBytes = kTabBytes;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    unsigned *out;
    unsigned long long *clk;
    cudaMalloc(&out, 4);
    cudaMalloc(&clk, 8);
    const int iters = 4000;
    for (int mode = 0; mode < 4; mode++) {
        const size_t smem = kPixWords * 4 + (mode == 0 ? kTabBytes : 0);
        cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(clk, 0, 8);
            k_mix<<<prop.multiProcessorCount, kThreads, smem>>>(mode, iters, tex, gtab, out, clk);
            cudaDeviceSynchronize();
        }
        unsigned long long c;
        cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        printf("mode %d: %.1f clk per warp-iteration per SM (slowest warp %llu clk / %d iterations / %d warps)  %s\n", mode,
               (double)c / iters / (kThreads / 32) * 1.0, c, iters, kThreads / 32, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
