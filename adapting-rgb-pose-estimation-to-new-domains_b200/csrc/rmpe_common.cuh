// Shared declarations of the sm_100a kernels behind include/rmpe_b200.h.
// Compiled with -fmad=false: every float/double expression below is evaluated with one IEEE
// rounding per written operation, which is what the reference's NumPy / OpenCV / SciPy code
// does; fused multiply-adds appear only where written explicitly (fma()/__fma_rn).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rmpe_b200.h"

namespace rmpe {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

// Per-kernel device timing (rmpe_profile_* in the C ABI): when enabled, a ProfScope brackets one
// kernel launch with two CUDA events on the launching stream; rmpe_profile_get() resolves them.
// Disabled (the default) it costs one predictable branch.
bool prof_enabled();
void prof_mark(const char *name, cudaStream_t st, bool begin);
struct ProfScope {
    const char *name;
    cudaStream_t st;
    bool on;
    ProfScope(const char *n, cudaStream_t s) : name(n), st(s), on(prof_enabled()) {
        if (on) prof_mark(name, st, true);
    }
    ~ProfScope() {
        if (on) prof_mark(name, st, false);
    }
};

#define RMPE_CUDA_TRY(expr)                                                               \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            rmpe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                            __FILE__, __LINE__);                                          \
            return RMPE_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define RMPE_REQUIRE(cond, msg)                                                           \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            rmpe::set_error("bad argument: %s (%s)", msg, #cond);                         \
            return RMPE_E_BADARG;                                                         \
        }                                                                                 \
    } while (0)

// ------------------------------------------------------------------------------------------
// library state (rmpe_host.cu)
// ------------------------------------------------------------------------------------------
struct DeviceTables {
    const int16_t *bicubic_i16;   // [32][32][4][4]  OpenCV initInterTab2D(INTER_CUBIC, fixpt)
    const uint32_t *bicubic_dp4a; // [32][32][4][2]  per tap row: {hi bytes (s8 x4), lo bytes (u8 x4)}
    int sm_count;
    int32_t *counters;            // kCounterRing work counters of the persistent kernels (one per stream), one 128-byte line each
};
constexpr int kCounterRing = 1024;    // work-counter slots of the persistent kernels: one per launching stream (rmpe_gt.cu)
constexpr int kCounterStride = 32;    // int32 per counter line
// Programmatic dependent launch: a kernel launched with launch_pdl may be scheduled while its predecessor on the stream
// still runs (once every CTA of the predecessor has called pdl_trigger or exited); it must call pdl_wait before it reads
// anything the predecessor wrote.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

bool is_initialised();
const DeviceTables &tables();

constexpr int kOutW = RMPE_OUT_W;
constexpr int kOutH = RMPE_OUT_H;
constexpr int kGrid = RMPE_GRID;
constexpr int kCells = kGrid * kGrid;          // 2116
constexpr int kCellVec = kCells / 4;           // 529 float4 per label plane
constexpr int kParts = RMPE_NUM_PARTS;
constexpr int kLimbs = RMPE_NUM_LIMBS;
constexpr int kLayers = RMPE_NUM_LAYERS;
constexpr int kMaxPersonsGt = 64;

// py_rmpe_config.py:30-33 (0-based from/to part of limb k; PAF channels 2k, 2k+1)
__constant__ const int8_t c_limb_from[kLimbs] = {1, 8, 9, 1, 11, 12, 1, 2, 3, 2, 1, 5, 6, 5, 1, 0, 0, 14, 15};
__constant__ const int8_t c_limb_to[kLimbs] = {8, 9, 10, 11, 12, 13, 2, 3, 4, 16, 5, 6, 7, 17, 0, 14, 15, 16, 17};
// flip partner of each part (leftParts <-> rightParts, py_rmpe_config.py:5-9), identity otherwise
__constant__ const int8_t c_flip_partner[kParts] = {0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16};

// eval/eval_coco2014_multi_modes.py:28-35, made 0-based: parts (a,b) of decode limb k and the
// PAF channel of its x component (y component = +1)
__constant__ const int8_t c_dec_a[kLimbs] = {1, 1, 2, 3, 5, 6, 1, 8, 9, 1, 11, 12, 1, 0, 14, 0, 15, 2, 5};
__constant__ const int8_t c_dec_b[kLimbs] = {2, 5, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 0, 14, 16, 15, 17, 16, 17};
__constant__ const int8_t c_dec_paf[kLimbs] = {12, 20, 14, 16, 22, 24, 0, 2, 4, 6, 8, 10, 28, 30, 34, 32, 36, 18, 26};

// ------------------------------------------------------------------------------------------
// exact-arithmetic helpers
// ------------------------------------------------------------------------------------------
// OpenCV interpolateCubic (imgwarp.cpp / resize.cpp), float32, A = -0.75
__host__ __device__ inline void cubic_coeffs(float t, float c[4]) {
    const float A = -0.75f;
    float t1 = t + 1.0f;
    c[0] = ((A * t1 - 5.0f * A) * t1 + 8.0f * A) * t1 - 4.0f * A;
    c[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
    float u = 1.0f - t;
    c[2] = ((A + 2.0f) * u - (A + 3.0f)) * u * u + 1.0f;
    c[3] = 1.0f - c[0] - c[1] - c[2];
}

// cv::warpAffine's inverse of the forward matrix (imgwarp.cpp), f64, no contraction
__device__ inline bool invert_affine(const double *M, double iM[6]) {
    double D = __dsub_rn(__dmul_rn(M[0], M[4]), __dmul_rn(M[1], M[3]));
    bool ok = (D != 0.0);
    D = ok ? __ddiv_rn(1.0, D) : 0.0;
    double A11 = __dmul_rn(M[4], D), A22 = __dmul_rn(M[0], D);
    iM[0] = A11;
    iM[1] = __dmul_rn(M[1], -D);
    iM[3] = __dmul_rn(M[3], -D);
    iM[4] = A22;
    iM[2] = __dsub_rn(__dmul_rn(-iM[0], M[2]), __dmul_rn(iM[1], M[5]));
    iM[5] = __dsub_rn(__dmul_rn(-iM[3], M[2]), __dmul_rn(iM[4], M[5]));
    return ok;
}

// fixed-point column / row terms of WarpAffineInvoker (AB_BITS=10, round_delta=16)
__device__ inline int warp_col_term(double m, int x) {
    return __double2int_rn(__dmul_rn(__dmul_rn(m, (double)x), 1024.0));
}
__device__ inline int warp_row_term(double m, double b, int y) {
    return __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m, (double)y), b), 1024.0)) + 16;
}
__device__ inline int sat_short(int v) { return max(-32768, min(32767, v)); }

__device__ inline int dp4a_us(unsigned a_u8x4, unsigned b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
// two-way dot product of signed 16-bit weights with the low (.lo) / high (.hi) two unsigned bytes of b: exact int32
__device__ inline int dp2a_lo_su(unsigned a_s16x2, unsigned b_u8x4, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_u8x4), "r"(c));
    return d;
}
__device__ inline int dp2a_hi_su(unsigned a_s16x2, unsigned b_u8x4, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_u8x4), "r"(c));
    return d;
}
__device__ inline int dp4a_uu(unsigned a_u8x4, unsigned b_u8x4, int c) {
    unsigned d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_u8x4), "r"(c));
    return (int)d;
}

// ------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA, non-tensor) primitives for sm_100a
// ------------------------------------------------------------------------------------------
__device__ inline uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ inline void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ inline void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ inline void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ inline void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ inline void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16
__device__ inline void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ inline void bulk_s2g(void *gmem_dst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ inline void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ inline void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ inline void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace rmpe
