// Inference decode on sm_100a, from the network's output blobs to (candidate, subset).
//
// Reference path replaced (eval/eval_coco2014_multi_modes.py, relative to the reference root):
//   :277-278 / :79-92   cv2.resize of the blobs (+ 4-scale f64 average)  -> k_resize_h / k_resize_v (heat),
//                                                                           paf_point() on the fly (PAF)
//   :283-304 / :97-118  gaussian_filter(sigma=3) + 4-neighbour peaks      -> k_smooth_peaks, k_peaks_finalize
//   :310-360 / :126-176 PAF line integral, criteria, sort, greedy         -> k_limbs
//   :364-415 / :180-231 person assembly + prune                            -> k_assemble
//
// All float arithmetic follows the order of operations of OpenCV 4.13 (resize.cpp, SSE baseline:
// no FMA) and SciPy (ni_filters.c symmetric correlate, f64 accumulate), see oracle/decode_oracle.py.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "rmpe_common.cuh"

namespace rmpe {

constexpr int kMaxPeaksCap = 1024;
constexpr int kMaxCandCap = 4096;
constexpr int kMaxSubsetCap = 128;
constexpr int kChunkFrames = 64;
constexpr int kHeatC = 19;
constexpr int kPafC = 38;

// scipy _gaussian_kernel1d(sigma=3, radius=12) as produced by numpy on x86-64 (checked against the
// live scipy weights in tests/test_oracle_pin.py); index 0..12 = offsets -12..0, symmetric.
__constant__ double c_gauss[13] = {
    0x1.763a210dfb306p-15, 0x1.4fbe39149e277p-13, 0x1.0d8a5ad43c165p-11, 0x1.8345966f69518p-10,
    0x1.f1e9915139406p-9,  0x1.1e6bccad344bap-7,  0x1.26defcaeb0202p-6,  0x1.0fa58939b528fp-5,
    0x1.bfde9c12bec92p-5,  0x1.4a614d1afd337p-4,  0x1.b42a57d56c0bep-4,  0x1.01a25f86eb137p-3,
    0x1.105a329f98197p-3};

// ------------------------------------------------------------------------------------------
// cv2.resize INTER_CUBIC coordinate of one destination index (resize.cpp):
//   scale = 1/inv_scale (double); f = float((d+0.5)*scale-0.5); s = floor(f); t = f - s
// ------------------------------------------------------------------------------------------
__device__ inline double resize_scale(int dst_full, int src_len, double inv_fx) {
    double inv = (inv_fx > 0.0) ? inv_fx : __ddiv_rn((double)dst_full, (double)src_len);
    return __ddiv_rn(1.0, inv);
}
__device__ inline int resize_axis(int d, double scale, float co[4]) {
    float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
    float fl = floorf(f);
    cubic_coeffs(__fsub_rn(f, fl), co);
    return (int)fl;
}
__device__ inline int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
// horizontal pass: ((S0*a0 + S1*a1) + S2*a2) + S3*a3
__device__ inline float tap_ltr(float s0, float s1, float s2, float s3, const float a[4]) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s0, a[0]), __fmul_rn(s1, a[1])), __fmul_rn(s2, a[2])),
                     __fmul_rn(s3, a[3]));
}
// vertical pass, vector body: S0*b0 + (S1*b1 + (S2*b2 + S3*b3))
__device__ inline float tap_rtl(float s0, float s1, float s2, float s3, const float b[4]) {
    return __fadd_rn(__fmul_rn(s0, b[0]),
                     __fadd_rn(__fmul_rn(s1, b[1]), __fadd_rn(__fmul_rn(s2, b[2]), __fmul_rn(s3, b[3]))));
}
// the last (W*C mod 4) floats of a destination row are produced by the scalar (left-to-right) tail loop
__device__ inline bool in_row_tail(int x, int c, int W, int C) {
    int n = W * C;
    return (x * C + c) >= n - (n & 3);
}

// ------------------------------------------------------------------------------------------
// separable resize passes over planar / strided float maps (heat path)
// ------------------------------------------------------------------------------------------
struct RJob {
    const float *src;
    float *dst;        // plain store target (planar, col stride 1) or null
    void *acc;         // accumulate target (float for single-scale store, double for the average)
    long long src_cs, src_rs, src_xs;
    long long dst_cs, dst_rs;
    int n_rows, n_cols;    // extent to compute
    int src_len;           // source length along the resized axis (clamp bound)
    int dst_full;          // full destination length along the resized axis
    double inv_fx;         // explicit scale factor (fx=fy=stride call) or 0
    int tail_w, tail_c;    // vertical pass: destination row width / channels of the cv2 call
    int ndiv;              // >0: acc[...] (+)= double(v / ndiv)
    int first;             // first scale of the average: store, do not add
};
struct RJobs {
    RJob j[kChunkFrames];
};

__global__ void __launch_bounds__(128) k_resize_h(const __grid_constant__ RJobs jobs) {
    const RJob &J = jobs.j[blockIdx.z];
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int r = blockIdx.y;
    if (x >= J.n_cols || r >= J.n_rows) return;
    float a[4];
    const int sx = resize_axis(x, resize_scale(J.dst_full, J.src_len, J.inv_fx), a);
    const long long o0 = clampi(sx - 1, 0, J.src_len - 1) * J.src_xs, o1 = clampi(sx, 0, J.src_len - 1) * J.src_xs;
    const long long o2 = clampi(sx + 1, 0, J.src_len - 1) * J.src_xs, o3 = clampi(sx + 2, 0, J.src_len - 1) * J.src_xs;
#pragma unroll 6
    for (int c = 0; c < kParts; c++) {
        const float *row = J.src + c * J.src_cs + r * J.src_rs;
        J.dst[c * J.dst_cs + r * J.dst_rs + x] = tap_ltr(row[o0], row[o1], row[o2], row[o3], a);
    }
}

__global__ void __launch_bounds__(128) k_resize_v(const __grid_constant__ RJobs jobs) {
    const RJob &J = jobs.j[blockIdx.z];
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= J.n_cols || y >= J.n_rows) return;
    float b[4];
    const int sy = resize_axis(y, resize_scale(J.dst_full, J.src_len, J.inv_fx), b);
    const long long o0 = clampi(sy - 1, 0, J.src_len - 1) * J.src_rs, o1 = clampi(sy, 0, J.src_len - 1) * J.src_rs;
    const long long o2 = clampi(sy + 1, 0, J.src_len - 1) * J.src_rs, o3 = clampi(sy + 2, 0, J.src_len - 1) * J.src_rs;
#pragma unroll 6
    for (int c = 0; c < kParts; c++) {
        const float *col = J.src + c * J.src_cs + x * J.src_xs;
        float s0 = col[o0], s1 = col[o1], s2 = col[o2], s3 = col[o3];
        float v = in_row_tail(x, c, J.tail_w, J.tail_c) ? tap_ltr(s0, s1, s2, s3, b) : tap_rtl(s0, s1, s2, s3, b);
        long long o = c * J.dst_cs + y * J.dst_rs + x;
        if (J.ndiv > 0) {
            double add = (double)__fdiv_rn(v, (float)J.ndiv);
            double *acc = reinterpret_cast<double *>(J.acc);
            acc[o] = J.first ? __dadd_rn(0.0, add) : __dadd_rn(acc[o], add);
        } else {
            J.dst[o] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// PAF value at an integer point of the up-sampled (single) / scale-averaged (multi) field,
// evaluated on the fly from the blob(s) with the exact per-pixel arithmetic of cv2.resize.
// ------------------------------------------------------------------------------------------
// one bicubic resize (src h x w, NHWC with C channels) evaluated at destination (y, x), channel c.  scl_x / scl_y are
// resize_scale() of the two axes: two f64 divisions each, a constant of the frame -- the kernels that evaluate many
// points of one frame compute them once (PtScales) instead of per point.
__device__ inline float resize_point_blob_s(const float *__restrict__ blob, int h, int w, int C, int c, int y, int x,
                                            int W, double scl_x, double scl_y) {
    float a[4], b[4];
    int sx = resize_axis(x, scl_x, a);
    int sy = resize_axis(y, scl_y, b);
    float hp[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float *row = blob + ((size_t)clampi(sy - 1 + j, 0, h - 1) * w) * C + c;
        hp[j] = tap_ltr(row[(size_t)clampi(sx - 1, 0, w - 1) * C], row[(size_t)clampi(sx, 0, w - 1) * C],
                        row[(size_t)clampi(sx + 1, 0, w - 1) * C], row[(size_t)clampi(sx + 2, 0, w - 1) * C], a);
    }
    return in_row_tail(x, c, W, C) ? tap_ltr(hp[0], hp[1], hp[2], hp[3], b) : tap_rtl(hp[0], hp[1], hp[2], hp[3], b);
}
__device__ inline float resize_point_blob(const float *__restrict__ blob, int h, int w, int C, int c, int y, int x,
                                          int H, int W, double inv_f) {
    return resize_point_blob_s(blob, h, w, C, c, y, x, W, resize_scale(W, w, inv_f), resize_scale(H, h, inv_f));
}

// multi-scale chain for one scale: blob -> x stride (fx=fy=stride) -> crop (Hc,Wc) -> (H,W).  s2x / s2y: scales of the
// second resize, s1: scale of the first (1 / stride)
__device__ inline float resize_chain_point_s(const float *__restrict__ blob, int hs, int ws, int C, int c, int y, int x,
                                             int W, int Hc, int Wc, int stride, double s2x, double s2y, double s1) {
    float a[4], b[4];
    int sx = resize_axis(x, s2x, a);
    int sy = resize_axis(y, s2y, b);
    float hp[4];
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        int r = clampi(sy - 1 + j, 0, Hc - 1);
        float t[4];
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            int q = clampi(sx - 1 + k, 0, Wc - 1);
            t[k] = resize_point_blob_s(blob, hs, ws, C, c, r, q, ws * stride, s1, s1);
        }
        hp[j] = tap_ltr(t[0], t[1], t[2], t[3], a);
    }
    return in_row_tail(x, c, W, C) ? tap_ltr(hp[0], hp[1], hp[2], hp[3], b) : tap_rtl(hp[0], hp[1], hp[2], hp[3], b);
}
__device__ inline float resize_chain_point(const float *__restrict__ blob, int hs, int ws, int C, int c, int y, int x,
                                           int H, int W, int Hc, int Wc, int stride) {
    return resize_chain_point_s(blob, hs, ws, C, c, y, x, W, Hc, Wc, stride, resize_scale(W, Wc, 0.0), resize_scale(H, Hc, 0.0),
                                resize_scale(ws * stride, ws, (double)stride));
}

// resize scales of a frame: per scale the (second) resize to (H, W); s1 = the x stride up-sampling of the chain
struct PtScales {
    double sy[RMPE_MAX_SCALES], sx[RMPE_MAX_SCALES], s1;
};
__device__ inline void pt_scales_one(PtScales &ps, const RmpeFrameDesc &f, int stride, int s) {
    if (f.n_scales == 1) {
        ps.sy[0] = resize_scale(f.height, f.grid_h[0], 0.0);
        ps.sx[0] = resize_scale(f.width, f.grid_w[0], 0.0);
    } else {
        ps.sy[s] = resize_scale(f.height, f.grid_h[s] * stride - f.pad_down[s], 0.0);
        ps.sx[s] = resize_scale(f.width, f.grid_w[s] * stride - f.pad_right[s], 0.0);
    }
    if (s == 0) ps.s1 = resize_scale(0, 1, (double)stride);
}

__device__ inline double blob_point_s(const RmpeFrameDesc &f, const float *__restrict__ blobs, const int64_t *off, int C,
                                      int stride, int c, int y, int x, const PtScales &ps) {
    if (f.n_scales == 1)
        return (double)resize_point_blob_s(blobs + off[0], f.grid_h[0], f.grid_w[0], C, c, y, x, f.width, ps.sx[0], ps.sy[0]);
    double acc = 0.0;
    for (int s = 0; s < f.n_scales; s++) {
        int Hc = f.grid_h[s] * stride - f.pad_down[s], Wc = f.grid_w[s] * stride - f.pad_right[s];
        float v = resize_chain_point_s(blobs + off[s], f.grid_h[s], f.grid_w[s], C, c, y, x, f.width, Hc, Wc, stride,
                                       ps.sx[s], ps.sy[s], ps.s1);
        acc = __dadd_rn(acc, (double)__fdiv_rn(v, (float)f.n_scales));
    }
    return acc;
}
__device__ inline double paf_point_s(const RmpeFrameDesc &f, const float *__restrict__ paf, int stride, int c, int y, int x,
                                     const PtScales &ps) {
    return blob_point_s(f, paf, f.paf_offset, kPafC, stride, c, y, x, ps);
}
__device__ inline double paf_point(const RmpeFrameDesc &f, const float *__restrict__ paf, int stride, int c, int y,
                                   int x) {
    PtScales ps;
    for (int s = 0; s < f.n_scales; s++) pt_scales_one(ps, f, stride, s);
    return paf_point_s(f, paf, stride, c, y, x, ps);
}

__global__ void k_paf_points(RmpeFrameDesc f, const float *paf, int stride, int n, const int32_t *cyx, double *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = paf_point(f, paf, stride, cyx[3 * i], cyx[3 * i + 1], cyx[3 * i + 2]);
}

// ------------------------------------------------------------------------------------------
// k_smooth_peaks: per (tile, part, frame): U tile + 13-px halo -> shared memory; scipy's
// separable 25-tap Gaussian (axis 0, store in T, axis 1, store in T; f64 accumulation in the
// reference's order; 'reflect' boundary); zero-padded 4-neighbour >= test and > thre1;
// peaks leave through a warp-aggregated atomic slot grab (order restored in k_peaks_finalize).
// ------------------------------------------------------------------------------------------
constexpr int kST = 32;          // tile edge
constexpr int kSR = 12;          // gaussian radius
constexpr int kSU = kST + 2 * kSR + 2;  // 58: U region edge
constexpr int kSA = kST + 2;            // 34: rows/cols that need smoothed values (tile + 1)
constexpr int kSmoothThreads = 256;

struct SmoothJob {
    const void *U;     // planar [18][H][W] of T
    int H, W;
    int frame;         // global frame index (for the counters / lists)
    void *S_out;       // optional: smoothed map out (debug), planar like U
};
struct SmoothJobs {
    SmoothJob j[kChunkFrames];
};

__device__ inline int reflect_idx(int i, int n) {
    int p = 2 * n;
    int m = i % p;
    if (m < 0) m += p;
    return m < n ? m : p - 1 - m;
}

template <typename T>
__global__ void __launch_bounds__(kSmoothThreads) k_smooth_peaks(const __grid_constant__ SmoothJobs jobs, T thre1,
                                                                  int max_peaks, int32_t *__restrict__ raw_key,
                                                                  double *__restrict__ raw_score,
                                                                  int32_t *__restrict__ raw_count,
                                                                  int32_t *__restrict__ status) {
    const SmoothJob &J = jobs.j[blockIdx.z];
    const int H = J.H, W = J.W;
    const int tiles_x = (W + kST - 1) / kST, tiles_y = (H + kST - 1) / kST;
    if ((int)blockIdx.x >= tiles_x * tiles_y) return;
    const int part = blockIdx.y;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int y0 = ty * kST, x0 = tx * kST;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) uint8_t sm_raw[];
    T *sU = reinterpret_cast<T *>(sm_raw);   // [kSU][kSU]
    T *sA = sU + kSU * kSU;                  // [kSA][kSU]
    T *sS = sA + kSA * kSU;                  // [kSA][kSA]

    // region of U this tile can touch (clipped); reflect() of any needed index lands inside it
    const int ry0 = max(0, y0 - kSR - 1), ry1 = min(H - 1, y0 + kST + kSR);
    const int rx0 = max(0, x0 - kSR - 1), rx1 = min(W - 1, x0 + kST + kSR);
    const int rh = ry1 - ry0 + 1, rw = rx1 - rx0 + 1;
    const T *U = reinterpret_cast<const T *>(J.U) + (size_t)part * H * W;
    for (int i = tid; i < rh * rw; i += kSmoothThreads) {
        int r = i / rw, c = i - r * rw;
        sU[r * kSU + c] = U[(size_t)(ry0 + r) * W + rx0 + c];
    }
    __syncthreads();

    // rows / cols that need smoothed values: tile +-1, clipped
    const int ay0 = max(0, y0 - 1), ay1 = min(H - 1, y0 + kST);
    const int ax0 = max(0, x0 - 1), ax1 = min(W - 1, x0 + kST);
    const int ah = ay1 - ay0 + 1, aw = ax1 - ax0 + 1;

    // axis 0 (along y) for every region column
    for (int i = tid; i < ah * rw; i += kSmoothThreads) {
        int r = i / rw, c = i - r * rw;
        int y = ay0 + r;
        double tmp = __dmul_rn((double)sU[(y - ry0) * kSU + c], c_gauss[12]);
#pragma unroll
        for (int j = -kSR; j < 0; j++) {
            int ya = reflect_idx(y + j, H) - ry0, yb = reflect_idx(y - j, H) - ry0;
            double pair = __dadd_rn((double)sU[ya * kSU + c], (double)sU[yb * kSU + c]);
            tmp = __dadd_rn(tmp, __dmul_rn(pair, c_gauss[12 + j]));
        }
        sA[r * kSU + c] = (T)tmp;
    }
    __syncthreads();
    // axis 1 (along x)
    for (int i = tid; i < ah * aw; i += kSmoothThreads) {
        int r = i / aw, c = i - r * aw;
        int x = ax0 + c;
        double tmp = __dmul_rn((double)sA[r * kSU + (x - rx0)], c_gauss[12]);
#pragma unroll
        for (int j = -kSR; j < 0; j++) {
            int xa = reflect_idx(x + j, W) - rx0, xb = reflect_idx(x - j, W) - rx0;
            double pair = __dadd_rn((double)sA[r * kSU + xa], (double)sA[r * kSU + xb]);
            tmp = __dadd_rn(tmp, __dmul_rn(pair, c_gauss[12 + j]));
        }
        sS[r * kSA + c] = (T)tmp;
    }
    __syncthreads();

    // 4-neighbour test with zero padding outside the image
    const int th = min(kST, H - y0), tw = min(kST, W - x0);
    for (int base = 0; base < th * kST; base += kSmoothThreads) {
        int i = base + tid;
        int ly = i / kST, lx = i - ly * kST;
        bool peak = false;
        int y = y0 + ly, x = x0 + lx;
        if (ly < th && lx < tw) {
            int r = y - ay0, c = x - ax0;
            T s = sS[r * kSA + c];
            T up = (y > 0) ? sS[(r - 1) * kSA + c] : (T)0;
            T dn = (y < H - 1) ? sS[(r + 1) * kSA + c] : (T)0;
            T lf = (x > 0) ? sS[r * kSA + c - 1] : (T)0;
            T rt = (x < W - 1) ? sS[r * kSA + c + 1] : (T)0;
            peak = (s >= up) && (s >= dn) && (s >= lf) && (s >= rt) && (s > thre1);
            if (J.S_out) reinterpret_cast<T *>(J.S_out)[((size_t)part * H + y) * W + x] = s;
        }
        unsigned bal = __ballot_sync(0xffffffffu, peak);
        if (bal) {
            int lane = tid & 31;
            int leader = __ffs(bal) - 1;
            int slot0 = 0;
            if (lane == leader) slot0 = atomicAdd(raw_count + J.frame * kParts + part, __popc(bal));
            slot0 = __shfl_sync(0xffffffffu, slot0, leader);
            if (peak) {
                int slot = slot0 + __popc(bal & ((1u << lane) - 1));
                if (slot < max_peaks) {
                    size_t o = ((size_t)J.frame * kParts + part) * max_peaks + slot;
                    raw_key[o] = y * W + x;
                    raw_score[o] = (double)sU[(y - ry0) * kSU + (x - rx0)];
                } else {
                    atomicOr(status + J.frame, RMPE_ST_PEAK_OVERFLOW);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fast peak path for single-scale frames: screen in float32, decide in the reference's arithmetic.
//
// Up-sampling (cv2.resize, replicate border) and smoothing (scipy gaussian, reflect border) are
// both linear and separable, so the smoothed map is  S = Ky * blob * Kx^T  with composite
// per-axis operators  K = G * R  that have ~(24 h/H + 5) non-zeros per row.
//   k_axis_tables  builds K (f64 accumulation, stored f32) and the first source index per row.
//   k_screen_plan  decides from bounds on S which (32 x 126 tile, part) pairs can hold a peak at all
//                  (tile bound from the largest blob values, then per-row and per-column bounds) and
//                  which 32-column groups of such a tile; the survivors become work items.
//   k_screen_pairs  evaluates S~ = Ky * blob * Kx^T for the live column groups of an item straight
//                  from the NHWC blob (the 57x larger up-sampled map is never written) and emits every
//                  pixel that could be a peak once a rigorous bound delta on |S~ - S| is allowed for:
//                  S~ > thre1 - delta and S~ >= neighbour~ - 2 delta for the four neighbours.
//   k_peak_verify  re-evaluates each such pixel and its four neighbours EXACTLY -- cv2's float32
//                  tap order, scipy's f64 accumulation order and per-axis float32 store -- and
//                  applies the reference's test; survivors go to the same raw lists that
//                  k_smooth_peaks fills, so ordering/ids are restored by k_peaks_finalize.
// ------------------------------------------------------------------------------------------
constexpr int kScrTW = 126;          // interior columns of a screening tile
constexpr int kScrTH = 32;           // interior rows
constexpr int kScrCols = kScrTW + 2; // with the 1-pixel ring the 4-neighbour test needs
constexpr int kScrRows = kScrTH + 2;
constexpr int kScrThreads = 256;
constexpr int kScrMaxSrcRows = 16;   // single scale: staged blob rows per tile (dense vertical operator)
constexpr int kScrMaxKW = 12;        // single scale: non-zeros per row of Kx (a 10-wide variant covers stride 8)
constexpr int kMsKW = 12;            // multi scale: non-zeros per row of Kx_s
constexpr int kMsMaxRows = 24;       // multi scale: staged blob rows per scale and tile
constexpr int kMsKWBig = 16;         // small frames (scale 2 of a 240-row image is up-sampled by only 2.6): wider variant
constexpr int kMsMaxRowsBig = 32;
constexpr int kPlanSlots = 4;        // parts of a tile refined together by k_screen_plan
constexpr int kPlanMaxRefined = 8;   // more surviving parts than this in a tile: no column bounds (all groups stay live)
constexpr int kPlanMaxCols = 96;     // staged columns + padding per scale that the refinement holds

struct AxisJob {
    int dst, src, kw;
    int mid, stride;   // mid > 0: chain  src -> x stride -> crop to mid -> dst  (process_multi_scale)
    float wscale;      // folded into the operator (1 / n_scales on the vertical axis)
    float *K;          // [dst][kw]
    int *lo;           // [dst]
};
struct AxisJobs {
    AxisJob j[2 * kChunkFrames];
};

// weights of destination index dp over the source axis, cv2.resize coefficient arithmetic
template <typename F>
__device__ inline void axis_taps(const AxisJob &J, int dp, F &&emit) {
    if (J.mid == 0) {
        float co[4];
        const int s0 = resize_axis(dp, resize_scale(J.dst, J.src, 0.0), co);
#pragma unroll
        for (int k = 0; k < 4; k++) emit(clampi(s0 - 1 + k, 0, J.src - 1), (double)co[k]);
    } else {
        float c2[4];
        const int s2 = resize_axis(dp, resize_scale(J.dst, J.mid, 0.0), c2);
        const double scale1 = resize_scale(J.src * J.stride, J.src, (double)J.stride);
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            float c1[4];
            const int s1 = resize_axis(clampi(s2 - 1 + k, 0, J.mid - 1), scale1, c1);
#pragma unroll
            for (int q = 0; q < 4; q++) emit(clampi(s1 - 1 + q, 0, J.src - 1), (double)c2[k] * (double)c1[q]);
        }
    }
}

__global__ void __launch_bounds__(128) k_axis_tables(const __grid_constant__ AxisJobs jobs, int32_t *__restrict__ err) {
    const AxisJob &J = jobs.j[blockIdx.y];
    const int d = blockIdx.x * 128 + threadIdx.x;
    __shared__ double s_acc[128][25];     // one accumulator row per thread (odd pitch: conflict-free)
    if (d >= J.dst) return;
    double *acc = s_acc[threadIdx.x];
    int lo = INT_MAX, hi = INT_MIN;
    for (int t = -kSR; t <= kSR; t++)
        axis_taps(J, reflect_idx(d + t, J.dst), [&](int idx, double) { lo = min(lo, idx); hi = max(hi, idx); });
    const int kw = min(J.kw, 24);
    if (hi - lo + 1 > kw || J.kw > 24) atomicOr(err, 1);
    for (int i = 0; i < kw; i++) acc[i] = 0.0;
    for (int t = -kSR; t <= kSR; t++) {
        const double g = c_gauss[12 - abs(t)];
        axis_taps(J, reflect_idx(d + t, J.dst), [&](int idx, double wgt) {
            if (idx - lo < kw) acc[idx - lo] = fma(g, wgt, acc[idx - lo]);
        });
    }
    J.lo[d] = lo;
    double pos = 0.0, neg = 0.0;
    for (int i = 0; i < kw; i++) {
        const double v = acc[i] * (double)J.wscale;
        J.K[(size_t)d * J.kw + i] = (float)v;
        if (v > 0.0) pos += v; else neg -= v;
    }
    // largest positive / negative row mass of the operator (float bits; the table block is zeroed before):
    // bounds S by P * max(blob,0) + N * max(-blob,0) in k_screen_plan
    atomicMax(&J.lo[J.dst], __float_as_int((float)(pos * 1.000001)));
    atomicMax(&J.lo[J.dst + 1], __float_as_int((float)(neg * 1.000001)));
}

// ------------------------------------------------------------------------------------------
// Screening kernels.  S = sum_s Ky_s * blob_s * Kx_s^T (one scale for process_single_scale; for
// process_multi_scale each K_s is the composite of the Gaussian, the resize to (H,W), the crop and the
// x8 resize, with 1/n_scales folded into Ky_s).
//   k_screen_plan   per (tile, frame): max(blob,0) and max(-blob,0) of every part over the blob region the tile
//                   depends on.  With P / N the positive / negative mass of Ky (x) Kx (row masses left behind
//                   the tables by k_axis_tables),  S <= sum_s P_s max(B_s,0) + N_s max(-B_s,0);  a part whose
//                   bound + delta <= thre1 cannot hold a peak in the tile and is dropped, the others become
//                   work items of up to `group` parts of one tile.  Heat maps are near zero away from
//                   keypoints (and network noise stays below thre1 / P), so most of the 18 x tiles pairs
//                   disappear here.
//   k_screen_pairs  persistent CTAs over the work items (a tile with many active parts does not serialise them
//                   in one CTA): operators of the tile once per item; per part: blob region per scale through
//                   cp.async (the next part's copy in flight), horizontal pass, vertical pass accumulated over
//                   the scales in registers, conservative 4-neighbour test.
// Rounding budget: the reference spends <= 14 float32 roundings per cv2.resize (two per scale in the
// multi-scale chain) plus two stores per Gaussian axis, S~ ~40; all partial sums are bounded by
// A = sum_s (P_s + N_s) max|B_s|, so |S~ - S| <= 70 * 2^-24 * A = 4.2e-6 A; kScreenDelta = 1.2e-5 A.
// ------------------------------------------------------------------------------------------
constexpr int kMsJobsPerLaunch = kChunkFrames;
constexpr float kScreenDelta = 1.2e-5f;
constexpr int kPlanThreads = 256;
constexpr int kPlanStride = 13 * kHeatC;     // 247 of the 256 threads scan the blob: thread t stays on channel t % 19

struct MsScale {
    const float *heat;
    const float *Ky, *Kx;
    const int *loy, *lox;
    int h, w, kwy, kwx;
};
struct MsJob {
    MsScale sc[RMPE_MAX_SCALES];
    int H, W, n_scales, frame, tiles_x, tiles;
};
struct MsJobs {
    MsJob j[kMsJobsPerLaunch];
};
struct ActEntry {          // one work item of k_screen_pairs: up to `group` (<= 16) active parts of one tile
    int job, tile;
    unsigned parts;         // bit p = part p is screened by this item
    int slot;               // index of its 18 bounds in act_A
    unsigned long long live;   // nibble k = column groups (32 tile columns each) of the item's k-th part that can hold a peak
};

// blob region of every scale that a tile's 34 x 128 smoothed values depend on: s_rng[sc] = {r0, r1, c0, c1};
// also the first source row of each tile row (s_loy) and this thread's first source column (mylox)
__device__ __forceinline__ void tile_ranges(const MsJob &J, int y0, int x0, int tid, int col, int half,
                                            int (*s_rng)[4], int (*s_loy)[kScrRows], int mylox[RMPE_MAX_SCALES]) {
    if (tid < 4 * RMPE_MAX_SCALES) s_rng[tid >> 2][tid & 3] = (tid & 1) ? INT_MIN : INT_MAX;
    __syncthreads();
    const int xg = clampi(x0 - 1 + col, 0, J.W - 1);
#pragma unroll
    for (int sc = 0; sc < RMPE_MAX_SCALES; sc++) {
        mylox[sc] = 0;
        if (sc < J.n_scales) {
            const MsScale &S = J.sc[sc];
            if (tid < kScrRows) {
                const int lo = S.loy[clampi(y0 - 1 + tid, 0, J.H - 1)];
                s_loy[sc][tid] = lo;
                atomicMin(&s_rng[sc][0], lo);
                atomicMax(&s_rng[sc][1], min(lo + S.kwy - 1, S.h - 1));
            }
            mylox[sc] = S.lox[xg];
            if (half == 0) {
                atomicMin(&s_rng[sc][2], mylox[sc]);
                atomicMax(&s_rng[sc][3], min(mylox[sc] + S.kwx - 1, S.w - 1));
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPlanThreads) k_screen_plan(const __grid_constant__ MsJobs jobs, float thre1, int act_cap,
                                                              int group, ActEntry *__restrict__ act,
                                                              float *__restrict__ act_A, int32_t *__restrict__ act_count,
                                                              const int32_t *__restrict__ tab_err,
                                                              int32_t *__restrict__ status, int refine) {
    pdl_trigger();
    const MsJob &J = jobs.j[blockIdx.y];
    if ((int)blockIdx.x >= J.tiles) return;
    if (*tab_err) {   // an operator row did not fit its table (cannot happen with plan_frame's bound): never screen on it
        if (threadIdx.x == 0) atomicOr(status + J.frame, RMPE_ST_PEAK_OVERFLOW);
        return;
    }
    const int NS = J.n_scales;
    const int ty = blockIdx.x / J.tiles_x, tx = blockIdx.x - ty * J.tiles_x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int y0 = ty * kScrTH, x0 = tx * kScrTW;
    __shared__ int s_rng[RMPE_MAX_SCALES][4];
    __shared__ int s_loy[RMPE_MAX_SCALES][kScrRows];
    __shared__ int s_bpos[RMPE_MAX_SCALES][kHeatC + 1], s_bneg[RMPE_MAX_SCALES][kHeatC + 1];
    int mylox[RMPE_MAX_SCALES];
    if (tid < RMPE_MAX_SCALES * (kHeatC + 1)) { (&s_bpos[0][0])[tid] = 0; (&s_bneg[0][0])[tid] = 0; }
    tile_ranges(J, y0, x0, tid, tid & (kScrCols - 1), tid >> 7, s_rng, s_loy, mylox);
    for (int sc = 0; sc < NS; sc++) {
        const MsScale &S = J.sc[sc];
        const int r0 = s_rng[sc][0], c0 = s_rng[sc][2];
        const int nrows = s_rng[sc][1] - r0 + 1, rowlen = (s_rng[sc][3] - c0 + 1) * kHeatC;
        // coalesced row segments of the NHWC blob; element e of a segment belongs to channel e % 19.  13 x 19 = 247
        // threads walk a segment with stride 247, so a thread stays on ONE channel and keeps its maxima in registers;
        // the 13 threads of a channel meet in shared memory once per scale.
        float vpos = 0.f, vneg = 0.f;
        if (tid < kPlanStride) {
            const int ch = tid % kHeatC;
            const float *base = S.heat + ((size_t)r0 * S.w + c0) * kHeatC;
            const size_t rstride = (size_t)S.w * kHeatC;
            int r = 0;
            for (; r + 1 < nrows; r += 2) {                       // two rows in flight
                const float *row0 = base + (size_t)r * rstride, *row1 = row0 + rstride;
                for (int e = tid; e < rowlen; e += kPlanStride) {
                    const float v0 = row0[e], v1 = row1[e];
                    vpos = fmaxf(vpos, fmaxf(v0, v1));
                    vneg = fmaxf(vneg, fmaxf(-v0, -v1));
                }
            }
            if (r < nrows) {
                const float *row0 = base + (size_t)r * rstride;
                for (int e = tid; e < rowlen; e += kPlanStride) {
                    const float v0 = row0[e];
                    vpos = fmaxf(vpos, v0);
                    vneg = fmaxf(vneg, -v0);
                }
            }
            if (vpos > 0.f) atomicMax(&s_bpos[sc][ch], __float_as_int(vpos));
            if (vneg > 0.f) atomicMax(&s_bneg[sc][ch], __float_as_int(vneg));
        }
    }
    __syncthreads();
    // S = sum_s sum_ij Ky_s[i] Kx_s[j] B_s[i][j]: the positive entries of Ky (x) Kx weigh max(B,0), the negative ones
    // max(-B,0); P = Py+ Px+ + Py- Px-, N = Py+ Px- + Py- Px+ with the row masses k_axis_tables left behind the tables
    __shared__ float s_A[kParts];
    __shared__ int s_parts[kParts], s_nact;      // parts still active, compacted
    __shared__ unsigned char s_live[kParts];     // their live column groups
    __shared__ int s_rb[kParts];
    if (tid < 32) {
        bool on = false;
        if (tid < kParts) {
            float bound = 0.f, atot = 0.f;
            for (int sc = 0; sc < NS; sc++) {
                const MsScale &S = J.sc[sc];
                const float yp = __int_as_float(S.loy[J.H]), yn = __int_as_float(S.loy[J.H + 1]);
                const float xp = __int_as_float(S.lox[J.W]), xn = __int_as_float(S.lox[J.W + 1]);
                const float P = yp * xp + yn * xn, N = yp * xn + yn * xp;
                const float bp = __int_as_float(s_bpos[sc][tid]), bn = __int_as_float(s_bneg[sc][tid]);
                bound += P * bp + N * bn;
                atot += (P + N) * fmaxf(bp, bn);
            }
            bound *= 1.0001f; atot *= 1.0001f;      // atot bounds every partial sum: scales the rounding allowance
            s_A[tid] = atot;
            on = bound + kScreenDelta * atot > thre1;
            s_live[tid] = on ? 0xF : 0;
            s_rb[tid] = 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) s_parts[__popc(m & ((1u << tid) - 1))] = tid;
        if (tid == 0) s_nact = __popc(m);
    }
    __syncthreads();
    int n_on = s_nact;
    if (n_on == 0) return;
    // ---- refinement of the parts that passed.  The bound above lets a peak switch on every tile within the reach of
    //      the operators (+-45 pixels: ~10 tiles per peak where ~2.6 hold values above thre1).  Two one-dimensional
    //      bounds from the same blob values follow:
    //        rows:    |S[r][c]| <= sum_s sum_i |Ky_s[r][i]| zy_s[i],  zy_s[i] = max(xp B+ + xn B-, xn B+ + xp B-) over staged row i
    //        columns: |S[r][c]| <= sum_s sum_j |Kx_s[c][j]| zx_s[j],  zx_s[j] likewise over staged column j with the masses of Ky
    //      (B+- = largest positive / negative blob value, xp / xn, yp / yn = largest row masses of the operators).
    //      A part stays active if some row bound and some column bound exceed thre1; the column bounds also say WHICH
    //      groups of 32 tile columns (with their two neighbour columns) can hold a peak: k_screen_pairs evaluates those only.
    bool can_refine = refine != 0;
    for (int sc = 0; sc < NS; sc++)
        can_refine = can_refine && (s_rng[sc][1] - s_rng[sc][0] + 1 <= kMsMaxRowsBig) &&
                     (s_rng[sc][3] - s_rng[sc][2] + 1 + kMsKWBig <= kPlanMaxCols) && J.sc[sc].kwx <= kMsKWBig;
    if (can_refine) {
        __shared__ float s_zx[kPlanSlots][RMPE_MAX_SCALES][kPlanMaxCols];
        __shared__ float s_cbp[kPlanSlots][kScrCols];
        __shared__ float s_zy[kParts][RMPE_MAX_SCALES][kMsMaxRowsBig];
        // (0) staged rows of the active parts: zy of one (part, scale, row) per thread; the lines were read by the scan above
        for (int sc = 0; sc < NS; sc++) {
            const MsScale &S = J.sc[sc];
            const int r0 = s_rng[sc][0], c0 = s_rng[sc][2];
            const int nrows = s_rng[sc][1] - r0 + 1, ncols = s_rng[sc][3] - c0 + 1;
            const float xp = __int_as_float(S.lox[J.W]), xn = __int_as_float(S.lox[J.W + 1]);
            for (int task = tid; task < n_on * nrows; task += kPlanThreads) {
                const int slot = task / nrows, i = task - slot * nrows;
                const float *base = S.heat + ((size_t)(r0 + i) * S.w + c0) * kHeatC + s_parts[slot];
                float vp = 0.f, vn = 0.f;
                for (int j = 0; j < ncols; j++) {
                    const float v = base[j * kHeatC];
                    vp = fmaxf(vp, v); vn = fmaxf(vn, -v);
                }
                s_zy[slot][sc][i] = fmaxf(xp * vp + xn * vn, xn * vp + xp * vn);
            }
        }
        __syncthreads();
        // (1) row bounds of every active part: one thread per (part, tile row)
        for (int task = tid; task < n_on * kScrRows; task += kPlanThreads) {
            const int slot = task / kScrRows, t = task - slot * kScrRows;
            const int part = s_parts[slot];
            const int y = clampi(y0 - 1 + t, 0, J.H - 1);
            float rb = 0.f;
            for (int sc = 0; sc < NS; sc++) {
                const MsScale &S = J.sc[sc];
                const int lo = s_loy[sc][t] - s_rng[sc][0], nrows = s_rng[sc][1] - s_rng[sc][0] + 1;
                const float *ky = S.Ky + (size_t)y * S.kwy;
                for (int k = 0; k < S.kwy; k++)
                    if (lo + k < nrows) rb = fmaf(fabsf(ky[k]), s_zy[slot][sc][lo + k], rb);
            }
            atomicMax(&s_rb[part], __float_as_int(rb));
        }
        __syncthreads();
        if (tid < 32) {                             // parts whose rows reach thre1, compacted again
            bool on = false;
            int part = 0;
            if (tid < n_on) {
                part = s_parts[tid];
                on = !(__int_as_float(s_rb[part]) * 1.0001f + kScreenDelta * s_A[part] <= thre1);
                if (!on) s_live[part] = 0;
            }
            const unsigned m = __ballot_sync(0xffffffffu, on);
            __syncwarp();
            if (on) s_parts[__popc(m & ((1u << tid) - 1))] = part;
            if (tid == 0) s_nact = __popc(m);
        }
        __syncthreads();
        n_on = s_nact;
        // a tile where most parts survive is dense (or crowded): its column groups are all live more often than not and
        // the column bounds would cost a batch of four parts each
        if (n_on > kPlanMaxRefined) n_on = 0;
        for (int b0 = 0; b0 < n_on; b0 += kPlanSlots) {
            const int nb = min(kPlanSlots, n_on - b0);
            __syncthreads();                       // the previous batch is done with the tables
            // (2) staged columns: one thread per (part, scale, column); the lines were read by the scan above
            for (int sc = 0; sc < NS; sc++) {
                const MsScale &S = J.sc[sc];
                const int r0 = s_rng[sc][0], c0 = s_rng[sc][2];
                const int nrows = s_rng[sc][1] - r0 + 1, ncols = s_rng[sc][3] - c0 + 1, np = ncols + kMsKWBig;
                const float yp = __int_as_float(S.loy[J.H]), yn = __int_as_float(S.loy[J.H + 1]);
                const size_t rstride = (size_t)S.w * kHeatC;
                for (int task = tid; task < nb * np; task += kPlanThreads) {
                    const int slot = task / np, j = task - slot * np;
                    float z = 0.f;
                    if (j < ncols) {
                        const float *base = S.heat + ((size_t)r0 * S.w + c0 + j) * kHeatC + s_parts[b0 + slot];
                        float vp = 0.f, vn = 0.f;
                        for (int i = 0; i < nrows; i++) {
                            const float v = base[i * rstride];
                            vp = fmaxf(vp, v); vn = fmaxf(vn, -v);
                        }
                        z = fmaxf(yp * vp + yn * vn, yn * vp + yp * vn);
                    }
                    s_zx[slot][sc][j] = z;          // zero beyond the region: the taps there weigh nothing
                }
            }
            __syncthreads();
            // (3) column bounds: thread = (tile column, parity of the part's slot)
            {
                const int c = tid & (kScrCols - 1), par = tid >> 7;
                const int xg = clampi(x0 - 1 + c, 0, J.W - 1);
                float cb[kPlanSlots / 2];
#pragma unroll
                for (int q = 0; q < kPlanSlots / 2; q++) cb[q] = 0.f;
                if (par < nb)
                    for (int sc = 0; sc < NS; sc++) {
                        const MsScale &S = J.sc[sc];
                        const float *kx = S.Kx + (size_t)xg * S.kwx;
                        const int off = mylox[sc] - s_rng[sc][2];
                        for (int j = 0; j < S.kwx; j++) {
                            const float w = fabsf(kx[j]);
#pragma unroll
                            for (int q = 0; q < kPlanSlots / 2; q++) cb[q] = fmaf(w, s_zx[2 * q + par][sc][off + j], cb[q]);
                        }
                    }
#pragma unroll
                for (int q = 0; q < kPlanSlots / 2; q++) s_cbp[2 * q + par][c] = cb[q];
            }
            __syncthreads();
            // (4) live column groups of a part: one warp per part
            if (warp < nb) {
                const int part = s_parts[b0 + warp];
                const float lim_b = thre1 - kScreenDelta * s_A[part];
                unsigned live = 0;
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float v = s_cbp[warp][32 * g + lane];
                    if (lane == 0 && g > 0) v = fmaxf(v, s_cbp[warp][32 * g - 1]);     // a hot pixel in the first or last column
                    if (lane == 31 && g < 3) v = fmaxf(v, s_cbp[warp][32 * g + 32]);   // compares with a value of the next group
                    if (__any_sync(0xffffffffu, !(v * 1.0001f <= lim_b))) live |= 1u << g;   // NaN / Inf keep the group
                }
                if (lane == 0) s_live[part] = (unsigned char)live;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        // active parts (a peak needs S > thre1), cut into work items of at most `group` parts
        unsigned m = 0;
        unsigned long long lv = 0;
        int n = 0;
        for (int p = 0; p <= kParts; p++) {
            const bool on = p < kParts && s_live[p];
            if (on) { m |= 1u << p; lv |= (unsigned long long)s_live[p] << (4 * n); n++; }
            if ((n == group || p == kParts) && m) {
                const int slot = atomicAdd(act_count, 1);
                if (slot < act_cap) {
                    act[slot] = ActEntry{(int)blockIdx.y, (int)blockIdx.x, m, slot, lv};
                    for (int q = 0; q < kParts; q++) act_A[(size_t)slot * kParts + q] = s_A[q];
                } else {
                    atomicOr(status + J.frame, RMPE_ST_PEAK_OVERFLOW);
                }
                m = 0; lv = 0; n = 0;
            }
        }
    }
}

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// vertical pass of R tile rows (row0 ..; rows beyond the tile repeat its last row) over NR staged rows
template <int NR, int MAXROWS, int R>
__device__ __forceinline__ void ms_vertical(const float *__restrict__ sT, const float *__restrict__ sKy, int col, int row0,
                                            int nrows, float acc[R]) {
    float tcol[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) tcol[i] = (i < nrows) ? sT[i * kScrCols + col] : 0.f;
#pragma unroll
    for (int q = 0; q < R; q++) {
        const float4 *ky = reinterpret_cast<const float4 *>(sKy + min(row0 + q, kScrRows - 1) * MAXROWS);
        float a = acc[q];
#pragma unroll
        for (int i4 = 0; i4 < NR / 4; i4++) {
            const float4 k4 = ky[i4];
            a = fmaf(k4.x, tcol[4 * i4 + 0], a);
            a = fmaf(k4.y, tcol[4 * i4 + 1], a);
            a = fmaf(k4.z, tcol[4 * i4 + 2], a);
            a = fmaf(k4.w, tcol[4 * i4 + 3], a);
        }
        acc[q] = a;
    }
}

// ms_vertical over the staged rows [i0, end) only, NR of them (a multiple of four, >= end - i0); the window is moved down
// if it would leave the MAXROWS-wide rows of Ky
template <int NR, int MAXROWS, int R>
__device__ __forceinline__ void band_vertical(const float *__restrict__ sT, const float *__restrict__ sKy, int i0, int end,
                                              int col, int row0, float acc[R]) {
    static_assert(NR <= MAXROWS && NR % 4 == 0 && MAXROWS % 4 == 0, "float4 rows of Ky");
    i0 = min(i0, MAXROWS - NR);
    ms_vertical<NR, MAXROWS, R>(sT + i0 * kScrCols, sKy + i0, col, row0, end - i0, acc);
}

// S of R tile rows of one column, accumulated over the scales, -> sS
template <int MAXROWS, int R>
__device__ __forceinline__ void slice_vertical(const float *__restrict__ sTall, const int *__restrict__ s_toff,
                                               const float *__restrict__ sKyAll, int (*s_slab)[14][2], int slab, int NS,
                                               int col, int row0, float *__restrict__ sS) {
    float acc[R];
#pragma unroll
    for (int q = 0; q < R; q++) acc[q] = 0.f;
    for (int sc = 0; sc < NS; sc++) {
        const float *sT = sTall + s_toff[sc];
        const float *sKy = sKyAll + sc * kScrRows * MAXROWS;
        // only the band of staged rows these tile rows touch (the rest of Ky is zero: same sums)
        const int i0 = s_slab[sc][slab][0], end = s_slab[sc][slab][1], len = end - i0;
        if (len <= 8) band_vertical<8, MAXROWS, R>(sT, sKy, i0, end, col, row0, acc);
        else if (len <= 12) band_vertical<12, MAXROWS, R>(sT, sKy, i0, end, col, row0, acc);
        else if (MAXROWS <= 16 || len <= 16) band_vertical<16, MAXROWS, R>(sT, sKy, i0, end, col, row0, acc);
        else if (len <= 20) band_vertical<(MAXROWS >= 20 ? 20 : MAXROWS), MAXROWS, R>(sT, sKy, i0, end, col, row0, acc);
        else if (MAXROWS <= 24 || len <= 24) band_vertical<(MAXROWS >= 24 ? 24 : MAXROWS), MAXROWS, R>(sT, sKy, i0, end, col, row0, acc);
        else band_vertical<MAXROWS, MAXROWS, R>(sT, sKy, i0, end, col, row0, acc);
    }
#pragma unroll
    for (int q = 0; q < R; q++)
        if (row0 + q < kScrRows) sS[(row0 + q) * kScrCols + col] = acc[q];
}

template <int KWX, int MAXROWS>
__global__ void __launch_bounds__(kScrThreads, MAXROWS <= 16 ? 3 : 2) k_screen_pairs(const __grid_constant__ MsJobs jobs, float thre1, int act_cap,
                                                              const ActEntry *__restrict__ act,
                                                              const float *__restrict__ act_A,
                                                              const int32_t *__restrict__ act_count,
                                                              int32_t *__restrict__ next_item, int cand_cap,
                                                              int32_t *__restrict__ cand_key,
                                                              int32_t *__restrict__ cand_fp,
                                                              int32_t *__restrict__ cand_count,
                                                              int32_t *__restrict__ status, int cull) {
    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col = tid & (kScrCols - 1), half = tid >> 7;   // set-up of an item: 128 columns x 2
    extern __shared__ __align__(16) uint8_t sm_raw[];
    __shared__ int s_rng[RMPE_MAX_SCALES][4];
    __shared__ int s_loy[RMPE_MAX_SCALES][kScrRows];
    __shared__ int s_boff[RMPE_MAX_SCALES + 1], s_toff[RMPE_MAX_SCALES + 1];
    // staged rows [first, end) that a slice of the 34 tile rows touches, for 2 / 4 / 8 slices (entries 0-1, 2-5, 6-13)
    __shared__ int s_slab[RMPE_MAX_SCALES][14][2];
    __shared__ short s_loff[RMPE_MAX_SCALES][kScrCols];        // first staged column of every tile column
    __shared__ float s_A[kParts];                              // the item's bounds on the partial sums (rounding allowance)
    __shared__ int s_item;
    const int n_act = min(*act_count, act_cap);

    // work items are handed out dynamically (one atomic per item): their cost is proportional to the number of
    // active parts of the tile (1..18 for multi scale) and a CTA only sees a handful of them, so a static stride
    // leaves the launch waiting for its unluckiest CTA
    for (;;) {
        __syncthreads();                                        // s_item of the previous round consumed
        if (tid == 0) s_item = atomicAdd(next_item, 1);
        __syncthreads();
        const int ai = s_item;
        if (ai >= n_act) break;
        const ActEntry E = act[ai];
        const MsJob &J = jobs.j[E.job];
        const int H = J.H, W = J.W, NS = J.n_scales;
        const int ty = E.tile / J.tiles_x, tx = E.tile - ty * J.tiles_x;
        const int y0 = ty * kScrTH, x0 = tx * kScrTW;          // first interior pixel
        int mylox[RMPE_MAX_SCALES];
        __syncthreads();                                        // the previous item's shared memory is free
        tile_ranges(J, y0, x0, tid, col, half, s_rng, s_loy, mylox);
        if (tid == 0) {
            int bo = 0, to = 0;
            for (int sc = 0; sc < NS; sc++) {
                const int nr = s_rng[sc][1] - s_rng[sc][0] + 1, nc = s_rng[sc][3] - s_rng[sc][2] + 1;
                s_boff[sc] = bo; bo += (nr * (nc + KWX) + 3) & ~3;      // rows padded with KWX zeros: no tap predicate
                s_toff[sc] = to; to += nr * kScrCols;
            }
            s_boff[NS] = bo; s_toff[NS] = to;
        }
        if (tid >= 32 && tid < 32 + 14 * NS) {
            // Ky is zero outside [loy[r], loy[r] + kwy) of its row: the vertical pass of a slice only needs that band
            const int sc = (tid - 32) / 14, e = (tid - 32) - 14 * sc, r0 = s_rng[sc][0];
            const int ns = e < 2 ? 2 : (e < 6 ? 4 : 8), k = e < 2 ? e : (e < 6 ? e - 2 : e - 6);
            const int R = (kScrRows + ns - 1) / ns;
            int lo = INT_MAX, hi = 0;
            for (int r = k * R; r < min((k + 1) * R, kScrRows); r++) {
                lo = min(lo, s_loy[sc][r] - r0);
                hi = max(hi, s_loy[sc][r] - r0 + J.sc[sc].kwy);
            }
            if (lo == INT_MAX) lo = 0;                  // a slice beyond the tile (the eighth of eight)
            s_slab[sc][e][0] = max(lo, 0) & ~3;         // Ky rows are read as float4
            s_slab[sc][e][1] = max(min(hi, s_rng[sc][1] - r0 + 1), s_slab[sc][e][0]);
        }
        if (tid >= 128 && tid < 128 + kParts) s_A[tid - 128] = act_A[(size_t)E.slot * kParts + tid - 128];
        if (half == 0)
            for (int sc = 0; sc < NS; sc++) s_loff[sc][col] = (short)(mylox[sc] - s_rng[sc][2]);
        __syncthreads();
        float *sBall = reinterpret_cast<float *>(sm_raw);                 // 2 x [sc][nrows][ncols + KWX]: this part / next part
        float *sTall = sBall + 2 * s_boff[NS];                            // [sc][nrows][128]
        float *sS = sTall + s_toff[NS];                                   // [34][128]
        float *sKyAll = sS + kScrRows * kScrCols;                         // [NS][34][MAXROWS] dense over staged rows
        float *sKxAll = sKyAll + NS * kScrRows * MAXROWS;                 // [NS][KWX][128]: tap-major, a warp's 32 columns in 32 banks
        bool bad = false;
        for (int sc = 0; sc < NS; sc++) bad = bad || (s_rng[sc][1] - s_rng[sc][0] + 1 > MAXROWS);
        if (bad) {   // host sized the launch for this never to happen
            if (tid == 0) atomicOr(status + J.frame, RMPE_ST_PEAK_OVERFLOW);
            continue;
        }
        const int xg = clampi(x0 - 1 + col, 0, W - 1);
        // operators of the tile, once per item
        for (int sc = 0; sc < NS; sc++) {
            const MsScale &S = J.sc[sc];
            const int r0 = s_rng[sc][0];
            float *sKy = sKyAll + sc * kScrRows * MAXROWS;
            for (int i = tid; i < kScrRows * MAXROWS; i += kScrThreads) {
                const int r = i / MAXROWS, q = i - r * MAXROWS;
                const int y = clampi(y0 - 1 + r, 0, H - 1);
                const int k = q - (s_loy[sc][r] - r0);
                sKy[i] = (k >= 0 && k < S.kwy) ? S.Ky[(size_t)y * S.kwy + k] : 0.f;
            }
            if (half == 0) {
                float *kx = sKxAll + (size_t)sc * KWX * kScrCols + col;
                for (int j = 0; j < KWX; j++) kx[j * kScrCols] = (j < S.kwx) ? S.Kx[(size_t)xg * S.kwx + j] : 0.f;
            }
        }
        // blob values of one part -> shared memory with cp.async (4-byte gathers out of the NHWC blob): the copy
        // of the NEXT part is in flight while this part is screened
        for (int i = tid; i < 2 * s_boff[NS]; i += kScrThreads) sBall[i] = 0.f;   // incl. the zero padding of every row
        auto stage_part = [&](int part, int buf) {
            for (int sc = 0; sc < NS; sc++) {
                const MsScale &S = J.sc[sc];
                const int r0 = s_rng[sc][0], c0 = s_rng[sc][2];
                const int nrows = s_rng[sc][1] - r0 + 1, ncols = s_rng[sc][3] - c0 + 1;
                float *sB = sBall + buf * s_boff[NS] + s_boff[sc];
                const int bp = ncols + KWX;
                for (int r = warp; r < nrows; r += kScrThreads / 32) {       // a staged row per warp: no index division
                    const float *src = S.heat + ((size_t)(r0 + r) * S.w + c0) * kHeatC + part;
                    float *dst = sB + r * bp;
                    for (int c = lane; c < ncols; c += 32) cp_async4(dst + c, src + c * kHeatC);
                }
            }
        };
        __syncthreads();                                        // zero fill done before the first copies land
        stage_part(__ffs(E.parts) - 1, 0);
        int buf = 0, part_k = 0;
      for (unsigned pm = E.parts; pm; pm &= pm - 1, buf ^= 1) {
        const int part = __ffs(pm) - 1;
        const float A = s_A[part];
        cp_async_wait_all();
        __syncthreads();                                        // this part's blob values landed; previous sT / sS are free
        if (pm & (pm - 1)) stage_part(__ffs(pm & (pm - 1)) - 1, buf ^ 1);
        // ---- the column groups (32 tile columns each) of this part that can hold a peak come with the item
        //      (k_screen_plan's column bounds); their work is spread over all eight warps: 1 / 2 / 3-4 live groups ->
        //      8 / 4 / 2 row slices per group.  A skipped group's part of sS stays stale and is never compared with a
        //      pixel that can be a peak.
        const unsigned live = cull ? (unsigned)(E.live >> (4 * part_k)) & 0xFu : 0xFu;
        part_k++;
        if (!live) continue;                                    // (uniform over the CTA)
        const int n_live = __popc(live);
        const int ns = n_live == 1 ? 8 : (n_live == 2 ? 4 : 2);       // row slices per group
        const int slab0 = ns == 2 ? 0 : (ns == 4 ? 2 : 6);
        const int gi = warp / ns, k = warp - gi * ns;
        const bool on = gi < n_live;
        const int cg = 32 * (on ? __fns(live, 0, gi + 1) : 0) + lane;  // this warp's tile columns
        // ---- horizontal passes: T_s[i][col] = sum_j Kx_s[col][j] * B_s[i][lox+j] ----
        if (on)
        for (int sc = 0; sc < NS; sc++) {
            const int nrows = s_rng[sc][1] - s_rng[sc][0] + 1, bp = s_rng[sc][3] - s_rng[sc][2] + 1 + KWX;
            const float *sB = sBall + buf * s_boff[NS] + s_boff[sc] + s_loff[sc][cg];
            float *sT = sTall + s_toff[sc];
            const float *kx = sKxAll + (size_t)sc * KWX * kScrCols + cg;
            float kxw[KWX];
#pragma unroll
            for (int j = 0; j < KWX; j++) kxw[j] = kx[j * kScrCols];
            for (int i = k; i < nrows; i += ns) {
                const float *b = sB + i * bp;
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < KWX; j++) acc = fmaf(kxw[j], b[j], acc);   // beyond the region: zero x zero padding
                sT[i * kScrCols + cg] = acc;
            }
        }
        __syncthreads();
        // ---- vertical passes, accumulated over the scales: S[r][col] = sum_s sum_i Ky_s[r][i] * T_s[i][col] ----
        if (on) {
            if (ns == 2) slice_vertical<MAXROWS, 17>(sTall, s_toff, sKyAll, s_slab, slab0 + k, NS, cg, 17 * k, sS);
            else if (ns == 4) slice_vertical<MAXROWS, 9>(sTall, s_toff, sKyAll, s_slab, slab0 + k, NS, cg, 9 * k, sS);
            else slice_vertical<MAXROWS, 5>(sTall, s_toff, sKyAll, s_slab, slab0 + k, NS, cg, 5 * k, sS);
        }
        __syncthreads();
        // ---- conservative 4-neighbour test on the interior (most rows of a tile hold nothing above thre1:
        //      one shared-memory read, one compare and one vote per row then) ----
        const float delta = kScreenDelta * A;
        const float lim = thre1 - delta, d2 = 2.f * delta;
        const int x = x0 + cg - 1;
        const bool col_ok = cg >= 1 && cg <= kScrTW && x < W;
        const int rt_n = kScrTH / ns;                           // 16, 8 or 4 interior rows per slice
#pragma unroll 1
        for (int q = 0; on && q < rt_n; q++) {
            const int r = 1 + k * rt_n + q;
            const int y = y0 + r - 1;
            const float sv = sS[r * kScrCols + cg];
            const bool hot = col_ok && y < H && sv > lim;
            if (!__any_sync(0xffffffffu, hot)) continue;
            bool cand = false;
            if (hot) {
                const float up = (y > 0) ? sS[(r - 1) * kScrCols + cg] : 0.f;
                const float dn = (y < H - 1) ? sS[(r + 1) * kScrCols + cg] : 0.f;
                const float lf = (x > 0) ? sS[r * kScrCols + cg - 1] : 0.f;
                const float rt = (x < W - 1) ? sS[r * kScrCols + cg + 1] : 0.f;
                cand = (sv >= up - d2) && (sv >= dn - d2) && (sv >= lf - d2) && (sv >= rt - d2);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, cand);
            if (bal) {
                const int leader = __ffs(bal) - 1;
                int slot0 = 0;
                if (lane == leader) slot0 = atomicAdd(cand_count, __popc(bal));
                slot0 = __shfl_sync(0xffffffffu, slot0, leader);
                if (cand) {
                    const int slot = slot0 + __popc(bal & ((1u << lane) - 1));
                    if (slot < cand_cap) {
                        cand_key[slot] = y * W + x;
                        cand_fp[slot] = J.frame * kParts + part;
                    } else {
                        atomicOr(status + J.frame, RMPE_ST_PEAK_OVERFLOW);
                    }
                }
            }
        }
      }   // parts of the item
    }
}

// one CTA per screened pixel: exact S at the pixel and its four neighbours
constexpr int kVerThreads = 128;
constexpr int kVerN = 2 * kSR + 3;   // 27: the pixel +-1, +-12
constexpr int kVerI1Cap = 64 * 64;   // cached x-stride rectangle (floats); larger ones are evaluated point by point

// heat value of the (scale-averaged) up-sampled map at an integer point, the reference's arithmetic
__device__ inline double heat_point_s(const RmpeFrameDesc &f, const float *__restrict__ heat, int stride, int c, int y, int x,
                                      const PtScales &ps) {
    return blob_point_s(f, heat, f.heat_offset, kHeatC, stride, c, y, x, ps);
}

__global__ void __launch_bounds__(kVerThreads, 6) k_peak_verify(const RmpeFrameDesc *__restrict__ frames,
                                                            const float *__restrict__ heat, int stride, double thre1,
                                                            int cand_cap, const int32_t *__restrict__ cand_key,
                                                            const int32_t *__restrict__ cand_fp,
                                                            const int32_t *__restrict__ cand_count, int max_peaks,
                                                            int32_t *__restrict__ raw_key, double *__restrict__ raw_score,
                                                            int32_t *__restrict__ raw_count, int32_t *__restrict__ status) {
    pdl_wait();
    pdl_trigger();
    // doubles hold both map dtypes: float32 maps (single scale) are rounded to float32 where the
    // reference stores float32, and comparing float32 values as doubles is the float32 comparison
    __shared__ double sU[kVerN][kVerN + 1];
    __shared__ double sA[3][kVerN + 1];
    __shared__ double sS5[5];
    __shared__ float sI1[kVerI1Cap];               // multi scale: rectangle of the x-stride map under the neighbourhood
    __shared__ int s_tap0[2][kVerN];               // first tap of the second resize per row / column
    __shared__ float s_tapc[2][kVerN][4];          // its coefficients
    __shared__ int s_box[4];
    __shared__ PtScales s_ps;                      // resize scales of the candidate's frame (f64 divisions: once per frame)
    int ps_frame = -1;
    const int total = min(*cand_count, cand_cap);
    const int tid = threadIdx.x;
    for (int ci = blockIdx.x; ci < total; ci += gridDim.x) {
        const int fp = cand_fp[ci], key = cand_key[ci];
        const int frame = fp / kParts, part = fp - frame * kParts;
        const RmpeFrameDesc f = frames[frame];
        if (frame != ps_frame) {                   // (uniform; the previous candidate ended behind a barrier)
            if (tid < f.n_scales) pt_scales_one(s_ps, f, stride, tid);
            ps_frame = frame;
            __syncthreads();
        }
        const int H = f.height, W = f.width;
        const int y = key / W, x = key - y * W;
        const bool f32map = f.n_scales == 1;
        // U on the 27x27 reflected neighbourhood
        if (f32map) {
            for (int i = tid; i < kVerN * kVerN; i += kVerThreads) {
                const int a = i / kVerN, b = i - a * kVerN;
                sU[a][b] = heat_point_s(f, heat, stride, part, reflect_idx(y - kSR - 1 + a, H), reflect_idx(x - kSR - 1 + b, W), s_ps);
            }
        } else {
            // multi scale: U = sum_s f64(chain_s / n).  The 729 points of one scale share the x-stride map they are
            // resized from: its needed rectangle is evaluated once into shared memory (same operations as
            // resize_chain_point, each intermediate value computed once instead of up to 16 times).
            for (int i = tid; i < kVerN * kVerN; i += kVerThreads) sU[i / kVerN][i % kVerN] = 0.0;
            for (int sI = 0; sI < f.n_scales; sI++) {
                const int hs = f.grid_h[sI], ws = f.grid_w[sI];
                const int Hc = hs * stride - f.pad_down[sI], Wc = ws * stride - f.pad_right[sI];
                const float *blob = heat + f.heat_offset[sI];
                __syncthreads();
                if (tid < 2 * kVerN) {          // second-resize taps of the 27 rows / 27 columns
                    const bool isrow = tid < kVerN;
                    const int a = isrow ? tid : tid - kVerN;
                    float co[4];
                    const int g = isrow ? reflect_idx(y - kSR - 1 + a, H) : reflect_idx(x - kSR - 1 + a, W);
                    const int s0 = resize_axis(g, isrow ? s_ps.sy[sI] : s_ps.sx[sI], co);
                    s_tap0[isrow ? 0 : 1][a] = s0;
#pragma unroll
                    for (int k = 0; k < 4; k++) s_tapc[isrow ? 0 : 1][a][k] = co[k];
                }
                if (tid == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MIN; s_box[2] = INT_MAX; s_box[3] = INT_MIN; }
                __syncthreads();
                if (tid < 2 * kVerN) {
                    const bool isrow = tid < kVerN;
                    const int a = isrow ? tid : tid - kVerN, lim = (isrow ? Hc : Wc) - 1;
                    const int s0 = s_tap0[isrow ? 0 : 1][a];
                    atomicMin(&s_box[isrow ? 0 : 2], clampi(s0 - 1, 0, lim));
                    atomicMax(&s_box[isrow ? 1 : 3], clampi(s0 + 2, 0, lim));
                }
                __syncthreads();
                const int r0 = s_box[0], q0 = s_box[2];
                const int nr = s_box[1] - r0 + 1, nq = s_box[3] - q0 + 1;
                const bool cached = nr * nq <= kVerI1Cap;
                if (cached)
                    for (int i = tid; i < nr * nq; i += kVerThreads) {
                        const int r = i / nq, q = i - r * nq;
                        sI1[i] = resize_point_blob_s(blob, hs, ws, kHeatC, part, r0 + r, q0 + q, ws * stride, s_ps.s1, s_ps.s1);
                    }
                __syncthreads();
                for (int i = tid; i < kVerN * kVerN; i += kVerThreads) {
                    const int a = i / kVerN, b = i - a * kVerN;
                    float v;
                    if (cached) {
                        const int sy = s_tap0[0][a], sx = s_tap0[1][b];
                        float hp[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float *row = sI1 + (clampi(sy - 1 + j, 0, Hc - 1) - r0) * nq - q0;
                            hp[j] = tap_ltr(row[clampi(sx - 1, 0, Wc - 1)], row[clampi(sx, 0, Wc - 1)], row[clampi(sx + 1, 0, Wc - 1)],
                                            row[clampi(sx + 2, 0, Wc - 1)], s_tapc[1][b]);
                        }
                        const int xg = reflect_idx(x - kSR - 1 + b, W);
                        v = in_row_tail(xg, part, W, kHeatC) ? tap_ltr(hp[0], hp[1], hp[2], hp[3], s_tapc[0][a])
                                                             : tap_rtl(hp[0], hp[1], hp[2], hp[3], s_tapc[0][a]);
                    } else {
                        v = resize_chain_point_s(blob, hs, ws, kHeatC, part, reflect_idx(y - kSR - 1 + a, H),
                                                 reflect_idx(x - kSR - 1 + b, W), W, Hc, Wc, stride, s_ps.sx[sI], s_ps.sy[sI], s_ps.s1);
                    }
                    sU[a][b] = __dadd_rn(sU[a][b], (double)__fdiv_rn(v, (float)f.n_scales));
                }
            }
        }
        __syncthreads();
        // axis 0 on rows y-1, y, y+1 (scipy order, f64 accumulate, store in the map dtype)
        for (int i = tid; i < 3 * kVerN; i += kVerThreads) {
            const int rr = i / kVerN, b = i - rr * kVerN;
            const int a = kSR + rr;   // local row of y-1+rr
            double tmp = __dmul_rn(sU[a][b], c_gauss[12]);
#pragma unroll
            for (int j = -kSR; j < 0; j++) {
                const double pair = __dadd_rn(sU[a + j][b], sU[a - j][b]);
                tmp = __dadd_rn(tmp, __dmul_rn(pair, c_gauss[12 + j]));
            }
            sA[rr][b] = f32map ? (double)(float)tmp : tmp;
        }
        __syncthreads();
        // axis 1 at (y,x), (y-1,x), (y+1,x), (y,x-1), (y,x+1)
        if (tid < 5) {
            const int rr = (tid == 1) ? 0 : (tid == 2) ? 2 : 1;
            const int b = kSR + 1 + ((tid == 3) ? -1 : (tid == 4) ? 1 : 0);
            double tmp = __dmul_rn(sA[rr][b], c_gauss[12]);
#pragma unroll
            for (int j = -kSR; j < 0; j++) {
                const double pair = __dadd_rn(sA[rr][b + j], sA[rr][b - j]);
                tmp = __dadd_rn(tmp, __dmul_rn(pair, c_gauss[12 + j]));
            }
            sS5[tid] = f32map ? (double)(float)tmp : tmp;
        }
        __syncthreads();
        if (tid == 0) {
            const double thr = f32map ? (double)(float)thre1 : thre1;
            const double sv = sS5[0];
            const double up = (y > 0) ? sS5[1] : 0.0, dn = (y < H - 1) ? sS5[2] : 0.0;
            const double lf = (x > 0) ? sS5[3] : 0.0, rt = (x < W - 1) ? sS5[4] : 0.0;
            if ((sv >= up) && (sv >= dn) && (sv >= lf) && (sv >= rt) && (sv > thr)) {
                const int slot = atomicAdd(raw_count + fp, 1);
                if (slot < max_peaks) {
                    const size_t o = (size_t)fp * max_peaks + slot;
                    raw_key[o] = key;
                    raw_score[o] = sU[kSR + 1][kSR + 1];
                } else {
                    atomicOr(status + frame, RMPE_ST_PEAK_OVERFLOW);
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// k_peaks_finalize: restore np.nonzero order (ascending y*W+x) per part, assign consecutive ids
// across parts, write candidate rows [x, y, score, id] and the per-part peak tables.
// ------------------------------------------------------------------------------------------
struct FinalizeJobs {
    int W[kChunkFrames];
    int frame[kChunkFrames];      // chunks are cut from the frames sorted by shape: a chunk's frames are not consecutive
};
__global__ void __launch_bounds__(256) k_peaks_finalize(const __grid_constant__ FinalizeJobs fj, int first_frame,
                                                        int max_peaks, const int32_t *__restrict__ raw_key,
                                                        const double *__restrict__ raw_score,
                                                        const int32_t *__restrict__ raw_count,
                                                        double *__restrict__ candidate, int32_t *__restrict__ n_peaks,
                                                        int32_t *__restrict__ pk_x, int32_t *__restrict__ pk_y,
                                                        double *__restrict__ pk_s) {
    pdl_wait();
    pdl_trigger();
    const int part = blockIdx.x, frame = fj.frame[blockIdx.y];
    (void)first_frame;
    const int W = fj.W[blockIdx.y];
    __shared__ int s_key[kMaxPeaksCap];
    __shared__ double s_sc[kMaxPeaksCap];
    const int n = min(raw_count[frame * kParts + part], max_peaks);
    int npow = 1;
    while (npow < n) npow <<= 1;
    const size_t base = ((size_t)frame * kParts + part) * max_peaks;
    for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        s_key[i] = (i < n) ? raw_key[base + i] : INT_MAX;
        s_sc[i] = (i < n) ? raw_score[base + i] : 0.0;
    }
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool up = ((i & k) == 0);
                    int a = s_key[i], b = s_key[ixj];
                    if ((a > b) == up) {
                        s_key[i] = b; s_key[ixj] = a;
                        double t = s_sc[i]; s_sc[i] = s_sc[ixj]; s_sc[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    int id0 = 0;
    for (int q = 0; q < part; q++) id0 += min(raw_count[frame * kParts + q], max_peaks);
    if (threadIdx.x == 0) n_peaks[frame * kParts + part] = n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int key = s_key[i];
        int y = key / W, x = key - y * W;
        double sc = s_sc[i];
        double *row = candidate + ((size_t)frame * kParts * max_peaks + id0 + i) * 4;
        row[0] = (double)x; row[1] = (double)y; row[2] = sc; row[3] = (double)(id0 + i);
        pk_x[base + i] = x; pk_y[base + i] = y; pk_s[base + i] = sc;
    }
}

// ------------------------------------------------------------------------------------------
// k_limbs: one CTA per (limb, frame).  All nA x nB pairs in (i-major, j-minor) order, 10-point
// PAF line integral in f64, criteria, ordered ballot/scan compaction, stable descending sort
// (bitonic on (score desc, generation index asc)), greedy one-to-one pick.
// ------------------------------------------------------------------------------------------
constexpr int kLimbThreads = 256;
constexpr int kLimbPairs = 12;     // pairs per pass: 12 x 10 sample points x 2 channels = 240 threads
constexpr int kLimbFineMax = 96;   // above this many pairs per limb a thread per pair hides the latency by itself

__global__ void __launch_bounds__(kLimbThreads) k_limbs(const RmpeFrameDesc *__restrict__ frames, int first_frame,
                                                        const float *__restrict__ paf, int stride, double thre2,
                                                        int max_peaks, int max_cand,
                                                        const int32_t *__restrict__ n_peaks,
                                                        const int32_t *__restrict__ pk_x,
                                                        const int32_t *__restrict__ pk_y,
                                                        const double *__restrict__ pk_s,
                                                        double *__restrict__ limb_cand, int32_t *__restrict__ n_limb_cand,
                                                        double *__restrict__ connections, int32_t *__restrict__ n_conn,
                                                        double *__restrict__ ws_cand, int32_t *__restrict__ status) {
    pdl_wait();
    pdl_trigger();
    const int k = blockIdx.x, frame = first_frame + blockIdx.y;
    const RmpeFrameDesc f = frames[frame];
    const int pa = c_dec_a[k], pb = c_dec_b[k], pc = c_dec_paf[k];
    const int nA = n_peaks[frame * kParts + pa], nB = n_peaks[frame * kParts + pb];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (nA == 0 || nB == 0) {   // special_k
        if (tid == 0) { n_conn[frame * kLimbs + k] = -1; if (n_limb_cand) n_limb_cand[frame * kLimbs + k] = 0; }
        return;
    }
    const size_t tA = ((size_t)frame * kParts + pa) * max_peaks, tB = ((size_t)frame * kParts + pb) * max_peaks;
    // candidate rows (i, j, score, score + sA + sB) in generation order
    double *cand = (limb_cand ? limb_cand : ws_cand) + ((size_t)frame * kLimbs + k) * max_cand * 4;

    __shared__ int s_total;
    extern __shared__ __align__(16) uint8_t sm_raw[];
    double *s_score = reinterpret_cast<double *>(sm_raw);          // [npow]
    int *s_seq = reinterpret_cast<int *>(s_score + max_cand);       // [npow]
    uint8_t *s_usedA = reinterpret_cast<uint8_t *>(s_seq + max_cand);
    uint8_t *s_usedB = s_usedA + kMaxPeaksCap;
    __shared__ PtScales s_ps;                      // resize scales of the frame (f64 divisions: once per CTA, not per point)
    if (tid < f.n_scales) pt_scales_one(s_ps, f, stride, tid);
    if (tid == 0) s_total = 0;
    __syncthreads();

    // Pairs in (i-major, j-minor) order, kLimbPairs at a time.  Phase 1 spreads the 10 sample points x 2 PAF
    // channels of every pair over the CTA (the PAF value of a point is a chain of dependent global loads: with one
    // thread per pair a frame with 3 persons kept 9 threads busy); phase 2 sums them per pair in the reference's
    // order, applies the criteria and compacts in order.
    const int total = nA * nB;
    const double halfH = __dmul_rn(0.5, (double)f.height);
    __shared__ double s_val[kLimbPairs][10][2];
    __shared__ int s_warp_cnt[kLimbThreads / 32];
    if (total <= kLimbFineMax) {
    for (int base = 0; base < total; base += kLimbPairs) {
        if (tid < kLimbPairs * 20) {
            const int p = tid / 20, rem = tid - p * 20, I = rem >> 1, ch = rem & 1;
            const int idx = base + p;
            if (idx < total) {
                const int i = idx / nB, j = idx - i * nB;
                const int ax = pk_x[tA + i], ay = pk_y[tA + i], bx = pk_x[tB + j], by = pk_y[tB + j];
                const int vx = bx - ax, vy = by - ay;
                double v = 0.0;
                if (vx != 0 || vy != 0) {
                    // np.linspace(a, b, 10): step = (b-a)/9; p_I = I*step + a, p_9 = b
                    const double stepx = __ddiv_rn((double)vx, 9.0), stepy = __ddiv_rn((double)vy, 9.0);
                    const double px = (I == 9) ? (double)bx : __dadd_rn(__dmul_rn((double)I, stepx), (double)ax);
                    const double py = (I == 9) ? (double)by : __dadd_rn(__dmul_rn((double)I, stepy), (double)ay);
                    const int xi = __double2int_rn(px), yi = __double2int_rn(py);   // round half to even
                    v = paf_point_s(f, paf, stride, pc + ch, yi, xi, s_ps);
                }
                s_val[p][I][ch] = v;
            }
        }
        __syncthreads();
        if (warp == 0) {
            const int idx = base + lane;
            bool pass = false;
            double score = 0.0, score2 = 0.0;
            int i = 0, j = 0;
            if (lane < kLimbPairs && idx < total) {
                i = idx / nB; j = idx - i * nB;
                const int ax = pk_x[tA + i], ay = pk_y[tA + i], bx = pk_x[tB + j], by = pk_y[tB + j];
                const int vx = bx - ax, vy = by - ay;
                const double norm = __dsqrt_rn((double)(vx * vx + vy * vy));
                if (norm != 0.0) {
                    const double ux = __ddiv_rn((double)vx, norm), uy = __ddiv_rn((double)vy, norm);
                    double sum = 0.0;
                    int nok = 0;
#pragma unroll 1
                    for (int I = 0; I < 10; I++) {
                        const double sI = __dadd_rn(__dmul_rn(s_val[lane][I][0], ux), __dmul_rn(s_val[lane][I][1], uy));
                        sum = __dadd_rn(sum, sI);
                        nok += (sI > thre2) ? 1 : 0;
                    }
                    double prior = __dsub_rn(__ddiv_rn(halfH, norm), 1.0);
                    prior = prior < 0.0 ? prior : 0.0;
                    score = __dadd_rn(__ddiv_rn(sum, 10.0), prior);
                    pass = (nok > 8) && (score > 0.0);
                    score2 = __dadd_rn(__dadd_rn(score, pk_s[tA + i]), pk_s[tB + j]);
                }
            }
            // ordered compaction
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            const int off = s_total;
            if (pass) {
                const int slot = off + __popc(bal & ((1u << lane) - 1));
                if (slot < max_cand) {
                    double *row = cand + (size_t)slot * 4;
                    row[0] = (double)i; row[1] = (double)j; row[2] = score; row[3] = score2;
                } else {
                    atomicOr(status + frame, RMPE_ST_CAND_OVERFLOW);
                }
            }
            __syncwarp();
            if (lane == 0) s_total = off + __popc(bal);
        }
        __syncthreads();
    }
    } else {
    // many pairs (crowded scenes): one thread per pair, 256 pairs per pass -- enough independent chains per warp
    for (int base = 0; base < total; base += kLimbThreads) {
        int idx = base + tid;
        bool pass = false;
        double score = 0.0, score2 = 0.0;
        int i = 0, j = 0;
        if (idx < total) {
            i = idx / nB; j = idx - i * nB;
            int ax = pk_x[tA + i], ay = pk_y[tA + i], bx = pk_x[tB + j], by = pk_y[tB + j];
            int vx = bx - ax, vy = by - ay;
            double norm = __dsqrt_rn((double)(vx * vx + vy * vy));
            if (norm != 0.0) {
                double ux = __ddiv_rn((double)vx, norm), uy = __ddiv_rn((double)vy, norm);
                // np.linspace(a, b, 10): step = (b-a)/9; p_I = I*step + a, p_9 = b
                double stepx = __ddiv_rn((double)vx, 9.0), stepy = __ddiv_rn((double)vy, 9.0);
                double sum = 0.0;
                int nok = 0;
#pragma unroll 1
                for (int I = 0; I < 10; I++) {
                    double px = (I == 9) ? (double)bx : __dadd_rn(__dmul_rn((double)I, stepx), (double)ax);
                    double py = (I == 9) ? (double)by : __dadd_rn(__dmul_rn((double)I, stepy), (double)ay);
                    int xi = __double2int_rn(px), yi = __double2int_rn(py);   // round half to even
                    double vxp = paf_point_s(f, paf, stride, pc, yi, xi, s_ps);
                    double vyp = paf_point_s(f, paf, stride, pc + 1, yi, xi, s_ps);
                    double s = __dadd_rn(__dmul_rn(vxp, ux), __dmul_rn(vyp, uy));
                    sum = __dadd_rn(sum, s);
                    nok += (s > thre2) ? 1 : 0;
                }
                double prior = __dsub_rn(__ddiv_rn(halfH, norm), 1.0);
                prior = prior < 0.0 ? prior : 0.0;
                score = __dadd_rn(__ddiv_rn(sum, 10.0), prior);
                pass = (nok > 8) && (score > 0.0);
                score2 = __dadd_rn(__dadd_rn(score, pk_s[tA + i]), pk_s[tB + j]);
            }
        }
        // ordered compaction
        unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_total;
        for (int w = 0; w < warp; w++) off += s_warp_cnt[w];
        if (pass) {
            int slot = off + __popc(bal & ((1u << lane) - 1));
            if (slot < max_cand) {
                double *row = cand + (size_t)slot * 4;
                row[0] = (double)i; row[1] = (double)j; row[2] = score; row[3] = score2;
            } else {
                atomicOr(status + frame, RMPE_ST_CAND_OVERFLOW);
            }
        }
        __syncthreads();
        if (tid == 0) {
            int t = s_total;
            for (int w = 0; w < kLimbThreads / 32; w++) t += s_warp_cnt[w];
            s_total = t;
        }
        __syncthreads();
    }
    }
    const int nc = min(s_total, max_cand);
    if (tid == 0 && n_limb_cand) n_limb_cand[frame * kLimbs + k] = nc;
    __threadfence_block();
    __syncthreads();

    // stable descending sort by score
    int npow = 1;
    while (npow < nc) npow <<= 1;
    for (int t = tid; t < npow; t += kLimbThreads) {
        s_score[t] = (t < nc) ? cand[(size_t)t * 4 + 2] : -1.0e300;
        s_seq[t] = t;
    }
    for (int t = tid; t < kMaxPeaksCap; t += kLimbThreads) { s_usedA[t] = 0; s_usedB[t] = 0; }
    __syncthreads();
    for (int kk = 2; kk <= npow; kk <<= 1)
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            for (int t = tid; t < npow; t += kLimbThreads) {
                int u = t ^ jj;
                if (u > t) {
                    bool up = ((t & kk) == 0);
                    double sa = s_score[t], sb = s_score[u];
                    int qa = s_seq[t], qb = s_seq[u];
                    // "a after b" in the wanted order (score desc, seq asc)
                    bool a_after_b = (sa < sb) || (sa == sb && qa > qb);
                    if (a_after_b == up) { s_score[t] = sb; s_score[u] = sa; s_seq[t] = qb; s_seq[u] = qa; }
                }
            }
            __syncthreads();
        }
    // greedy pick (sequential, tiny)
    if (tid == 0) {
        int n = 0;
        const int lim = min(nA, nB);
        double *conn = connections + ((size_t)frame * kLimbs + k) * max_peaks * 5;
        int idA0 = 0, idB0 = 0;
        for (int q = 0; q < pa; q++) idA0 += n_peaks[frame * kParts + q];
        for (int q = 0; q < pb; q++) idB0 += n_peaks[frame * kParts + q];
        for (int t = 0; t < nc && n < lim; t++) {
            const double *row = cand + (size_t)s_seq[t] * 4;
            int i = (int)row[0], j = (int)row[1];
            if (!s_usedA[i] && !s_usedB[j]) {
                s_usedA[i] = 1; s_usedB[j] = 1;
                double *o = conn + (size_t)n * 5;
                o[0] = (double)(idA0 + i); o[1] = (double)(idB0 + j); o[2] = row[2]; o[3] = (double)i; o[4] = (double)j;
                n++;
            }
        }
        n_conn[frame * kLimbs + k] = n;
    }
}

// ------------------------------------------------------------------------------------------
// k_assemble: one warp per frame; person assembly limb by limb (the connections of a limb side by side where that
// cannot change the result, in the reference's order otherwise), merge and prune.
// ------------------------------------------------------------------------------------------
constexpr int kAsmConnRows = 1024;   // connection rows of a frame held in shared memory (more: read from global)
constexpr int kAsmRowCap = kMaxSubsetCap + 1;

__host__ __device__ inline bool asm_use_map(int max_peaks) { return max_peaks <= 256; }
static size_t assemble_smem_bytes(int max_peaks) {
    return ((size_t)kAsmRowCap * 2 + (size_t)kAsmConnRows * 3 + (size_t)kParts * max_peaks) * 8 + (size_t)kParts * kAsmRowCap * 4 +
           (asm_use_map(max_peaks) ? (size_t)kParts * max_peaks * 4 : 0);
}

// Working copy of `subset` in shared memory, column-major: ids as int32 (they are small exact integers in the
// reference's float64 rows; -1 = no part), total score and part count as float64 with the reference's addition order.
// Column-major turns the row search of every connection into conflict-free loads and a row deletion into one pass.
__global__ void __launch_bounds__(32) k_assemble(int first_frame, int max_peaks, int max_persons,
                                                 const double *__restrict__ candidate,
                                                 const double *__restrict__ connections,
                                                 const int32_t *__restrict__ n_conn, const int32_t *__restrict__ n_peaks,
                                                 double *__restrict__ subset,
                                                 int32_t *__restrict__ n_subset, int32_t *__restrict__ status) {
    pdl_wait();
    const int frame = first_frame + blockIdx.x, lane = threadIdx.x;
    extern __shared__ __align__(16) double sm_asm[];
    double *s_sc = sm_asm;                                        // [kAsmRowCap] subset[:, 18]
    double *s_ct = s_sc + kAsmRowCap;                             // [kAsmRowCap] subset[:, 19]
    double *s_conn = s_ct + kAsmRowCap;                           // [kAsmConnRows][3]: idA, idB, score
    double *s_score = s_conn + kAsmConnRows * 3;                  // [18 * max_peaks] peak scores
    int *s_id = reinterpret_cast<int *>(s_score + kParts * max_peaks);   // [18][kAsmRowCap] subset[:, 0:18]
    // row that holds candidate id (or -1; entries may go stale and are checked on use): replaces the scan over the rows
    // while every id sits in at most one row (asm_use_map: kept for max_peaks <= 256)
    int *s_row = s_id + kParts * kAsmRowCap;                             // [18 * max_peaks]
    const bool use_map = asm_use_map(max_peaks);
    bool map_ok = use_map;
    __shared__ int s_off[kLimbs + 1];
    __shared__ int s_nc[kLimbs];
    const double *cand = candidate + (size_t)frame * kParts * max_peaks * 4;
    // ---- one pass over global memory: counts, connection rows and peak scores of the frame ----
    int ntot = (lane < kParts) ? n_peaks[frame * kParts + lane] : 0;
    const int my_nc = (lane < kLimbs) ? n_conn[frame * kLimbs + lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ntot += __shfl_xor_sync(0xffffffffu, ntot, o);
    int incl = max(my_nc, 0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane < kLimbs) { s_off[lane] = incl - max(my_nc, 0); s_nc[lane] = my_nc; }
    if (lane == kLimbs - 1) s_off[kLimbs] = incl;
    __syncwarp();
    for (int i = lane; i < ntot; i += 32) s_score[i] = cand[(size_t)i * 4 + 2];
    if (use_map) for (int i = lane; i < ntot; i += 32) s_row[i] = -1;
    {
        const int rows_all = min(s_off[kLimbs], kAsmConnRows);
        for (int i = lane; i < rows_all * 3; i += 32) {
            const int row = i / 3, c = i - 3 * row;
            int k = 0;
#pragma unroll
            for (int q = 1; q < kLimbs; q++) k += (s_off[q] <= row) ? 1 : 0;   // s_off is non-decreasing
            s_conn[i] = connections[((size_t)frame * kLimbs + k) * max_peaks * 5 + (size_t)(row - s_off[k]) * 5 + c];
        }
    }
    __syncwarp();
    int nrows = 0;
    int st = 0;
    for (int k = 0; k < kLimbs; k++) {
        const int nc = s_nc[k];
        if (nc < 0) continue;
        const int ia = c_dec_a[k], ib = c_dec_b[k];
        int *colA = s_id + ia * kAsmRowCap, *colB = s_id + ib * kAsmRowCap;
        const double *conn = connections + ((size_t)frame * kLimbs + k) * max_peaks * 5;
        for (int i0 = 0; i0 < nc; i0 += 32) {
        const int n_chunk = min(32, nc - i0);
        {
            // ---- the connections of a limb side by side, one per lane.  The greedy pick gives every connection of a
            // limb its own A peak and its own B peak, so a connection that extends one row (found == 1) or opens a new
            // one (found == 0) cannot change what the other connections of the limb find -- unless two of them meet in
            // the same row or one of them joins two rows (found == 2): those chunks take the reference's order below.
            int pA = -1, pB = -1;
            double sc = 0.0;
            if (lane < n_chunk) {
                const int i = i0 + lane;
                const double *row = (s_off[k] + i < kAsmConnRows) ? s_conn + (s_off[k] + i) * 3 : conn + i * 5;
                pA = (int)row[0]; pB = (int)row[1]; sc = row[2];
            }
            int found = 0, j1 = -1;
            if (map_ok) {
                if (lane < n_chunk) {
                    int ra = s_row[pA], rb = s_row[pB];
                    if (ra >= 0 && colA[ra] != pA) ra = -1;          // the peak was overwritten in that row since
                    if (rb >= 0 && colB[rb] != pB) rb = -1;
                    if (ra >= 0 && rb >= 0 && ra != rb) { found = 2; j1 = min(ra, rb); }
                    else if (ra >= 0 || rb >= 0) { found = 1; j1 = ra >= 0 ? ra : rb; }
                }
            } else {
                for (int j = 0; j < nrows; j++) {
                    const bool hit = (lane < n_chunk) && (colA[j] == pA || colB[j] == pB);
                    if (hit) { if (found == 0) j1 = j; found++; }
                }
            }
            bool bad = found >= 2;
            const unsigned one_m = __ballot_sync(0xffffffffu, found == 1);
            if (found == 1) bad = __popc(__match_any_sync(one_m, j1)) > 1;
            const unsigned new_m = __ballot_sync(0xffffffffu, lane < n_chunk && found == 0 && k < 17);
            const int n_new = __popc(new_m);
            bad = bad || (nrows + n_new > min(max_persons, kMaxSubsetCap));   // overflow bits are set in order below
            if (!__any_sync(0xffffffffu, bad)) {
                if (found == 1) {
                    if (colB[j1] != pB) {
                        colB[j1] = pB;
                        if (use_map) s_row[pB] = j1;
                        s_ct[j1] = __dadd_rn(s_ct[j1], 1.0);
                        s_sc[j1] = __dadd_rn(s_sc[j1], __dadd_rn(s_score[pB], sc));
                    }
                } else if ((new_m >> lane) & 1u) {
                    const int r = nrows + __popc(new_m & ((1u << lane) - 1));
#pragma unroll
                    for (int c = 0; c < kParts; c++) s_id[c * kAsmRowCap + r] = (c == ia) ? pA : ((c == ib) ? pB : -1);
                    if (use_map) { s_row[pA] = r; s_row[pB] = r; }
                    s_sc[r] = __dadd_rn(__dadd_rn(__dadd_rn(0.0, s_score[pA]), s_score[pB]), sc);
                    s_ct[r] = 2.0;
                }
                nrows += n_new;
                __syncwarp();
                continue;
            }
        }
        for (int i = i0; i < i0 + n_chunk; i++) {
            const bool in_smem = s_off[k] + i < kAsmConnRows;
            const double *row = in_smem ? s_conn + (s_off[k] + i) * 3 : conn + i * 5;
            const int pA = (int)row[0], pB = (int)row[1];
            const double sc = row[2];
            // rows that already hold this A peak or this B peak (eval...:186-195: the first two are recorded)
            int found = 0, j1 = -1, j2 = -1;
            for (int j0 = 0; j0 < nrows; j0 += 32) {
                const int j = j0 + lane;
                const bool hit = (j < nrows) && (colA[j] == pA || colB[j] == pB);
                unsigned bal = __ballot_sync(0xffffffffu, hit);
                const int n = __popc(bal);
                if (n) {
                    if (found == 0) { j1 = j0 + __ffs(bal) - 1; bal &= bal - 1; if (bal) j2 = j0 + __ffs(bal) - 1; }
                    else if (found == 1) j2 = j0 + __ffs(bal) - 1;
                    found += n;
                }
            }
            if (found > 2) { st |= RMPE_ST_FOUND_GT2; found = 2; }
            if (found == 1) {
                if (lane == 0 && colB[j1] != pB) {
                    colB[j1] = pB;
                    s_ct[j1] = __dadd_rn(s_ct[j1], 1.0);
                    s_sc[j1] = __dadd_rn(s_sc[j1], __dadd_rn(s_score[pB], sc));
                }
            } else if (found == 2) {
                const bool both = (lane < kParts) && (s_id[lane * kAsmRowCap + j1] >= 0) && (s_id[lane * kAsmRowCap + j2] >= 0);
                const unsigned overlap = __ballot_sync(0xffffffffu, both);
                if (overlap == 0) {
                    // merge (eval...:203-207): r1[:18] += r2[:18] + 1; r1[18:] += r2[18:]; r1[18] += score; delete r2
                    if (lane < kParts) s_id[lane * kAsmRowCap + j1] += s_id[lane * kAsmRowCap + j2] + 1;
                    if (lane == 18) s_sc[j1] = __dadd_rn(__dadd_rn(s_sc[j1], s_sc[j2]), sc);
                    if (lane == 19) s_ct[j1] = __dadd_rn(s_ct[j1], s_ct[j2]);
                    __syncwarp();
                    for (int r0 = j2; r0 < nrows - 1; r0 += 32) {          // np.delete(subset, j2, 0): shift up by one
                        const int r = r0 + lane;
                        const bool on = r < nrows - 1;
                        int idv[kParts];
#pragma unroll
                        for (int c = 0; c < kParts; c++) idv[c] = on ? s_id[c * kAsmRowCap + r + 1] : 0;
                        const double scv = on ? s_sc[r + 1] : 0.0, ctv = on ? s_ct[r + 1] : 0.0;
                        __syncwarp();
                        if (on) {
#pragma unroll
                            for (int c = 0; c < kParts; c++) s_id[c * kAsmRowCap + r] = idv[c];
                            s_sc[r] = scv; s_ct[r] = ctv;
                        }
                        __syncwarp();
                    }
                    nrows--;
                } else {
                    map_ok = false;        // the B peak now sits in two rows: only the scan finds both from here on
                    if (lane == 0) {
                        colB[j1] = pB;
                        s_ct[j1] = __dadd_rn(s_ct[j1], 1.0);
                        s_sc[j1] = __dadd_rn(s_sc[j1], __dadd_rn(s_score[pB], sc));
                    }
                }
            } else if (found == 0 && k < 17) {
                if (nrows < max_persons && nrows < kMaxSubsetCap) {
                    if (lane < kParts) s_id[lane * kAsmRowCap + nrows] = (lane == ia) ? pA : ((lane == ib) ? pB : -1);
                    if (lane == 18) {
                        const double s2 = __dadd_rn(__dadd_rn(0.0, s_score[pA]), s_score[pB]);
                        s_sc[nrows] = __dadd_rn(s2, sc);
                    }
                    if (lane == 19) s_ct[nrows] = 2.0;
                    nrows++;
                } else {
                    st |= RMPE_ST_PERSON_OVERFLOW;
                }
            }
            __syncwarp();
        }
        if (map_ok) {
            // the sequential path moved rows around (merges) or filled them without the map: rebuild it from the columns
            for (int c = 0; c < kParts; c++)
                for (int r = lane; r < nrows; r += 32) {
                    const int id = s_id[c * kAsmRowCap + r];
                    if (id >= 0) s_row[id] = r;
                }
            __syncwarp();
        }
        }
    }
    // prune: fewer than 4 parts or mean score < 0.4 (eval...:411-415); rows leave as float64
    int nout = 0;
    double *out = subset + (size_t)frame * max_persons * 20;
    for (int r0 = 0; r0 < nrows; r0 += 32) {
        const int r = r0 + lane;
        const bool keep = (r < nrows) && !((s_ct[r] < 4.0) || (__ddiv_rn(s_sc[r], s_ct[r]) < 0.4));
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            double *o = out + (size_t)(nout + __popc(bal & ((1u << lane) - 1))) * 20;
#pragma unroll
            for (int c = 0; c < kParts; c++) o[c] = (double)s_id[c * kAsmRowCap + r];
            o[18] = s_sc[r]; o[19] = s_ct[r];
        }
        nout += __popc(bal);
    }
    if (lane == 0) {
        n_subset[frame] = nout;
        if (st) atomicOr(status + frame, st);
    }
}

// ==========================================================================================
// host side
// ==========================================================================================
struct FramePlan {
    bool screen;         // frame decoded through k_heat_screen[_ms] / k_peak_verify
    int kwy[RMPE_MAX_SCALES], kwx[RMPE_MAX_SCALES];          // non-zeros per row of the composite operators
    int nrows_b[RMPE_MAX_SCALES], ncols_b[RMPE_MAX_SCALES];  // bounds on the staged blob region of a screening tile
    bool big;            // multi scale: needs the wider k_screen_pairs variant
    size_t tab_elems;    // 4-byte elements of the frame's operator tables
    size_t smem;         // dynamic shared memory of its screening kernel
    bool multi;
    size_t u_elems;      // 18*H*W (T)
    size_t p1_elems;     // floats
    size_t i1_elems;
    size_t p2_elems;
    size_t bytes;        // total workspace of the frame (256-aligned parts)
};

static size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }
constexpr int kScreenSlots = 5;     // launch groups of k_screen_pairs per chunk: single scale (2 widths), multi scale (light, heavy, wide)
constexpr int kActCap = 1 << 15;   // work items per kernel variant and chunk (0.5 MB + 2.3 MB of bounds each)

static int ceil_div_d(double a) { return (int)ceil(a - 1e-9); }

// dynamic shared memory of k_screen_pairs<KWX, MAXROWS> for one scale's share
static size_t pairs_smem_scale(int nrows_b, int ncols_b, int kwx_t, int maxrows_t) {
    return (2 * (size_t)((nrows_b * (ncols_b + kwx_t) + 3) & ~3) + (size_t)nrows_b * kScrCols + (size_t)kScrRows * maxrows_t +
            (size_t)kScrCols * kwx_t) * 4;
}

static FramePlan plan_frame(const RmpeFrameDesc &f, int stride, bool allow_screen = true) {
    FramePlan p{};
    p.multi = f.n_scales > 1;
    p.u_elems = (size_t)kParts * f.height * f.width;
    static const bool off = getenv("RMPE_DECODE_EXACT_MAPS") != nullptr;   // debugging: force the full-map path
    // composite operator supports: a 25-tap window of destination pixels spans 24*mid/dst pixels of the
    // (cropped) x-stride map, i.e. (that + 5)/stride blob cells, plus the taps of one more bicubic
    bool ok = allow_screen && !off;
    size_t smem = (size_t)kScrRows * kScrCols * 4;     // sS
    for (int s = 0; s < f.n_scales && ok; s++) {
        const int h = f.grid_h[s], w = f.grid_w[s];
        // Support of one operator row.  25 destination taps span 24*scale source pixels, floor() of positions
        // over a span L takes values at most ceil(L) apart, and a bicubic adds taps -1..+2:
        //   single resize:  ceil(24 src/dst) + 4;   chain: D = ceil(24 mid/dst) + 3 pixels of the x-stride map,
        //   then ceil(D/stride) + 4 blob cells.  (+1 spare; k_axis_tables flags any row that does not fit.)
        if (!p.multi) {
            p.kwy[s] = ceil_div_d(24.0 * h / f.height) + 5;
            p.kwx[s] = ceil_div_d(24.0 * w / f.width) + 5;
        } else {
            const double Hc = (double)h * stride - f.pad_down[s], Wc = (double)w * stride - f.pad_right[s];
            p.kwy[s] = ceil_div_d((ceil_div_d(24.0 * Hc / f.height) + 3.0) / stride) + 5;
            p.kwx[s] = ceil_div_d((ceil_div_d(24.0 * Wc / f.width) + 3.0) / stride) + 5;
        }
        // a tile's 34 rows / 128 columns move the first source index by at most ceil(33 h/H) + 1
        p.nrows_b[s] = std::min(h, ceil_div_d((double)(kScrRows - 1) * h / f.height) + 1 + p.kwy[s]);
        p.ncols_b[s] = std::min(w, ceil_div_d((double)(kScrCols - 1) * w / f.width) + 1 + p.kwx[s]);
        p.tab_elems += (size_t)f.height * (p.kwy[s] + 1) + (size_t)f.width * (p.kwx[s] + 1) + 4;   // + row masses
        if (!p.multi) {
            ok = p.kwy[s] <= 24 && p.kwx[s] <= kScrMaxKW && p.nrows_b[s] <= kScrMaxSrcRows;
            smem += pairs_smem_scale(p.nrows_b[s], p.ncols_b[s], p.kwx[s] <= 10 ? 10 : kScrMaxKW, kScrMaxSrcRows);
        } else {
            ok = p.kwy[s] <= 24 && p.kwx[s] <= kMsKWBig && p.nrows_b[s] <= kMsMaxRowsBig;
            if (p.kwx[s] > kMsKW || p.nrows_b[s] > kMsMaxRows) p.big = true;
        }
    }
    if (p.multi)
        for (int s = 0; s < f.n_scales && ok; s++)
            smem += pairs_smem_scale(p.nrows_b[s], p.ncols_b[s], p.big ? kMsKWBig : kMsKW, p.big ? kMsMaxRowsBig : kMsMaxRows);
    p.smem = smem;
    ok = ok && smem <= 160 * 1024;
    p.screen = ok;
    if (p.screen) {
        p.u_elems = 0;
        p.bytes = al256(p.tab_elems * 4);
        return p;
    }
    if (!p.multi) {
        p.p1_elems = (size_t)kParts * f.grid_h[0] * f.width;
    } else {
        for (int s = 0; s < f.n_scales; s++) {
            size_t Hc = (size_t)f.grid_h[s] * stride - f.pad_down[s], Wc = (size_t)f.grid_w[s] * stride - f.pad_right[s];
            size_t p1 = (size_t)kParts * f.grid_h[s] * Wc, i1 = (size_t)kParts * Hc * Wc, p2 = (size_t)kParts * Hc * f.width;
            if (p1 > p.p1_elems) p.p1_elems = p1;
            if (i1 > p.i1_elems) p.i1_elems = i1;
            if (p2 > p.p2_elems) p.p2_elems = p2;
        }
    }
    p.bytes = al256(p.u_elems * (p.multi ? 8 : 4)) + al256(p.p1_elems * 4) + al256(p.i1_elems * 4) + al256(p.p2_elems * 4);
    return p;
}

// resident CTAs of k_screen_pairs per SM for a dynamic shared-memory need (+ static shared memory / reserve)
static int ctas_per_sm(size_t smem) { return std::max(1, std::min(6, (int)((224 * 1024) / (std::max<size_t>(smem, 1) + 4096)))); }

static size_t fixed_ws_bytes(int batch, int max_peaks, int max_cand) {
    size_t per_list = (size_t)batch * kParts * max_peaks;
    return al256(per_list * 4) /*raw_key*/ + al256(per_list * 8) /*raw_score*/ + al256((size_t)batch * kParts * 4) /*raw_count*/ +
           al256(per_list * 4) * 2 /*pk_x, pk_y*/ + al256(per_list * 8) /*pk_s*/ +
           al256((size_t)batch * kLimbs * max_cand * 4 * 8) /*ws_cand*/ +
           al256(per_list * 2 * 4) * 2 /*cand_key, cand_fp*/ + 256 /*cand_count, tab_err*/ +
           al256((size_t)kScreenSlots * kActCap * sizeof(ActEntry)) + al256((size_t)kScreenSlots * kActCap * kParts * 4) /*work items*/ + 4096;
}

static bool frame_ok(const RmpeFrameDesc &f, int stride) {
    if (f.height <= 0 || f.width <= 0 || f.n_scales < 1 || f.n_scales > RMPE_MAX_SCALES) return false;
    for (int s = 0; s < f.n_scales; s++) {
        if (f.grid_h[s] <= 0 || f.grid_w[s] <= 0) return false;
        if (f.n_scales > 1 && (f.grid_h[s] * stride - f.pad_down[s] <= 0 || f.grid_w[s] * stride - f.pad_right[s] <= 0))
            return false;
    }
    return true;
}

// heat maps of a chunk of frames -> U (planar [18][H][W], float single-scale / double multi-scale)
static int launch_heat_up(const RmpeFrameDesc *fr, int n, const float *heat, int stride, uint8_t *const *u_ptr,
                          float *const *p1_ptr, float *const *i1_ptr, float *const *p2_ptr, const FramePlan *plans,
                          cudaStream_t st) {
    // single-scale frames: blob -> (h x W) -> (H x W)
    {
        RJobs jh{}, jv{};
        int m = 0, maxc = 0, maxr_h = 0, maxr_v = 0;
        for (int i = 0; i < n; i++) {
            const RmpeFrameDesc &f = fr[i];
            if (f.n_scales != 1 || (plans && plans[i].screen)) continue;
            int h = f.grid_h[0], w = f.grid_w[0];
            RJob &a = jh.j[m];
            a.src = heat + f.heat_offset[0]; a.dst = p1_ptr[i]; a.acc = nullptr;
            a.src_cs = 1; a.src_rs = (long long)w * kHeatC; a.src_xs = kHeatC;
            a.dst_cs = (long long)h * f.width; a.dst_rs = f.width;
            a.n_rows = h; a.n_cols = f.width; a.src_len = w; a.dst_full = f.width; a.inv_fx = 0.0;
            RJob &b = jv.j[m];
            b.src = p1_ptr[i]; b.dst = reinterpret_cast<float *>(u_ptr[i]); b.acc = nullptr;
            b.src_cs = (long long)h * f.width; b.src_rs = f.width; b.src_xs = 1;
            b.dst_cs = (long long)f.height * f.width; b.dst_rs = f.width;
            b.n_rows = f.height; b.n_cols = f.width; b.src_len = h; b.dst_full = f.height; b.inv_fx = 0.0;
            b.tail_w = f.width; b.tail_c = kHeatC; b.ndiv = 0; b.first = 0;
            maxc = max(maxc, f.width); maxr_h = max(maxr_h, h); maxr_v = max(maxr_v, f.height);
            m++;
        }
        if (m) {
            { ProfScope ps("k_resize_h", st); k_resize_h<<<dim3((maxc + 127) / 128, maxr_h, m), 128, 0, st>>>(jh); }
            { ProfScope ps("k_resize_v", st); k_resize_v<<<dim3((maxc + 127) / 128, maxr_v, m), 128, 0, st>>>(jv); }
            count_launch(2);
        }
    }
    // multi-scale frames: per scale  blob -> x stride -> crop -> (H,W), accumulated in f64
    for (int s = 0; s < RMPE_MAX_SCALES; s++) {
        RJobs j1{}, j2{}, j3{}, j4{};
        int m = 0, c1 = 0, r1 = 0, r2 = 0, c3 = 0, r4 = 0;
        for (int i = 0; i < n; i++) {
            const RmpeFrameDesc &f = fr[i];
            if (f.n_scales <= 1 || s >= f.n_scales || (plans && plans[i].screen)) continue;
            int hs = f.grid_h[s], ws = f.grid_w[s];
            int Hc = hs * stride - f.pad_down[s], Wc = ws * stride - f.pad_right[s];
            // (1) horizontal x stride: blob (hs, ws) -> P1 (hs, Wc)   [columns beyond the crop never read]
            RJob &a = j1.j[m];
            a.src = heat + f.heat_offset[s]; a.dst = p1_ptr[i];
            a.src_cs = 1; a.src_rs = (long long)ws * kHeatC; a.src_xs = kHeatC;
            a.dst_cs = (long long)hs * Wc; a.dst_rs = Wc;
            a.n_rows = hs; a.n_cols = Wc; a.src_len = ws; a.dst_full = ws * stride; a.inv_fx = (double)stride;
            // (2) vertical x stride: P1 -> I1 (Hc, Wc); full row = ws*stride*19 floats, a multiple of 4
            RJob &b = j2.j[m];
            b.src = p1_ptr[i]; b.dst = i1_ptr[i];
            b.src_cs = (long long)hs * Wc; b.src_rs = Wc; b.src_xs = 1;
            b.dst_cs = (long long)Hc * Wc; b.dst_rs = Wc;
            b.n_rows = Hc; b.n_cols = Wc; b.src_len = hs; b.dst_full = hs * stride; b.inv_fx = (double)stride;
            b.tail_w = ws * stride; b.tail_c = kHeatC; b.ndiv = 0; b.first = 0;
            // (3) horizontal to W: I1 (Hc, Wc) -> P2 (Hc, W)
            RJob &c = j3.j[m];
            c.src = i1_ptr[i]; c.dst = p2_ptr[i];
            c.src_cs = (long long)Hc * Wc; c.src_rs = Wc; c.src_xs = 1;
            c.dst_cs = (long long)Hc * f.width; c.dst_rs = f.width;
            c.n_rows = Hc; c.n_cols = f.width; c.src_len = Wc; c.dst_full = f.width; c.inv_fx = 0.0;
            // (4) vertical to H, accumulate v / n_scales into the f64 average
            RJob &d = j4.j[m];
            d.src = p2_ptr[i]; d.dst = nullptr; d.acc = u_ptr[i];
            d.src_cs = (long long)Hc * f.width; d.src_rs = f.width; d.src_xs = 1;
            d.dst_cs = (long long)f.height * f.width; d.dst_rs = f.width;
            d.n_rows = f.height; d.n_cols = f.width; d.src_len = Hc; d.dst_full = f.height; d.inv_fx = 0.0;
            d.tail_w = f.width; d.tail_c = kHeatC; d.ndiv = f.n_scales; d.first = (s == 0) ? 1 : 0;
            c1 = max(c1, Wc); r1 = max(r1, hs); r2 = max(r2, Hc); c3 = max(c3, f.width); r4 = max(r4, f.height);
            m++;
        }
        if (!m) continue;
        { ProfScope ps("k_resize_h", st); k_resize_h<<<dim3((c1 + 127) / 128, r1, m), 128, 0, st>>>(j1); }
        { ProfScope ps("k_resize_v", st); k_resize_v<<<dim3((c1 + 127) / 128, r2, m), 128, 0, st>>>(j2); }
        { ProfScope ps("k_resize_h", st); k_resize_h<<<dim3((c3 + 127) / 128, r2, m), 128, 0, st>>>(j3); }
        { ProfScope ps("k_resize_v", st); k_resize_v<<<dim3((c3 + 127) / 128, r4, m), 128, 0, st>>>(j4); }
        count_launch(4);
    }
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}

static size_t smooth_smem_bytes(bool f64) { return (size_t)(kSU * kSU + kSA * kSU + kSA * kSA) * (f64 ? 8 : 4); }

static int ensure_smooth_attr() {
    static bool done = false;
    if (done) return RMPE_OK;
    RMPE_CUDA_TRY(cudaFuncSetAttribute(k_smooth_peaks<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smooth_smem_bytes(true)));
    RMPE_CUDA_TRY(cudaFuncSetAttribute(k_smooth_peaks<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smooth_smem_bytes(false)));
    RMPE_CUDA_TRY(cudaFuncSetAttribute(k_limbs, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kMaxCandCap * 12 + 2 * kMaxPeaksCap));
    RMPE_CUDA_TRY(cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)std::max(assemble_smem_bytes(kMaxPeaksCap), assemble_smem_bytes(256))));
    RMPE_CUDA_TRY(cudaFuncSetAttribute(k_screen_pairs<10, kScrMaxSrcRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    RMPE_CUDA_TRY(cudaFuncSetAttribute(k_screen_pairs<kScrMaxKW, kScrMaxSrcRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    RMPE_CUDA_TRY(cudaFuncSetAttribute(k_screen_pairs<kMsKW, kMsMaxRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    RMPE_CUDA_TRY(cudaFuncSetAttribute(k_screen_pairs<kMsKWBig, kMsMaxRowsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    done = true;
    return RMPE_OK;
}

}  // namespace rmpe

using namespace rmpe;

extern "C" size_t rmpe_decode_workspace_bytes(int batch, const RmpeFrameDesc *frames_host, int max_peaks, int max_cand, int stride) {
    if (batch <= 0 || !frames_host || stride <= 0) return 0;
    // enough for a full chunk (kChunkFrames) of the largest screened frame -- those only keep their operator tables here,
    // a few hundred KB each -- and for up to 32 of the largest frame that needs materialised maps (tens of MB each);
    // a chunk takes frames while they fit, so any size that holds one frame works
    size_t big_screen = 0, big_mat = 0;
    for (int i = 0; i < batch; i++) {
        const FramePlan p = plan_frame(frames_host[i], stride);    // the same plan rmpe_decode_batch makes with b->stride
        size_t &b = p.screen ? big_screen : big_mat;
        if (p.bytes > b) b = p.bytes;
    }
    const size_t n_screen = batch < kChunkFrames ? batch : kChunkFrames, n_mat = batch < 32 ? batch : 32;
    return fixed_ws_bytes(batch, max_peaks, max_cand) + std::max(big_screen * n_screen, big_mat * n_mat) + 4096;
}

extern "C" int rmpe_decode_batch(const RmpeDecodeBatch *b, void *stream_) {
    if (!is_initialised()) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(b != nullptr, "descriptor is null");
    RMPE_REQUIRE(b->batch >= 0, "negative batch");
    if (b->batch == 0) return RMPE_OK;
    RMPE_REQUIRE(b->max_peaks > 0 && b->max_peaks <= kMaxPeaksCap, "max_peaks must be in [1,1024]");
    RMPE_REQUIRE(b->max_cand > 0 && b->max_cand <= kMaxCandCap && (b->max_cand & (b->max_cand - 1)) == 0,
                 "max_cand must be a power of two in [1,4096]");
    RMPE_REQUIRE(b->max_persons > 0 && b->max_persons <= kMaxSubsetCap, "max_persons must be in [1,128]");
    RMPE_REQUIRE(b->stride > 0, "stride");
    RMPE_REQUIRE(b->heat && b->paf && b->frames && b->frames_host, "blobs / frame descriptors");
    RMPE_REQUIRE(b->candidate && b->n_peaks && b->connections && b->n_conn && b->subset && b->n_subset && b->status,
                 "output pointers");
    RMPE_REQUIRE(b->workspace != nullptr, "workspace");
    for (int i = 0; i < b->batch; i++) RMPE_REQUIRE(frame_ok(b->frames_host[i], b->stride), "frame descriptor");
    cudaStream_t st = (cudaStream_t)stream_;
    int rc = ensure_smooth_attr();
    if (rc != RMPE_OK) return rc;

    const int B = b->batch, MP = b->max_peaks, MC = b->max_cand;
    uint8_t *ws = (uint8_t *)b->workspace;
    size_t off = 0;
    auto take = [&](size_t bytes) { void *p = ws + off; off += al256(bytes); return p; };
    const size_t per_list = (size_t)B * kParts * MP;
    int32_t *raw_key = (int32_t *)take(per_list * 4);
    double *raw_score = (double *)take(per_list * 8);
    int32_t *raw_count = (int32_t *)take((size_t)B * kParts * 4);
    int32_t *pk_x = (int32_t *)take(per_list * 4);
    int32_t *pk_y = (int32_t *)take(per_list * 4);
    double *pk_s = (double *)take(per_list * 8);
    double *ws_cand = (double *)take((size_t)B * kLimbs * MC * 4 * 8);
    int32_t *cand_key = (int32_t *)take(per_list * 2 * 4);
    int32_t *cand_fp = (int32_t *)take(per_list * 2 * 4);
    int32_t *cand_count = (int32_t *)take(256);      // [0] candidates, [1..5] work items per launch group, [8] table error, [16..20] next item per launch group
    int32_t *tab_err = cand_count + 8;
    const int act_cap = kActCap;
    ActEntry *act = (ActEntry *)take((size_t)kScreenSlots * kActCap * sizeof(ActEntry));
    float *act_A = (float *)take((size_t)kScreenSlots * kActCap * kParts * sizeof(float));
    RMPE_REQUIRE(off <= b->workspace_bytes, "workspace too small (see rmpe_decode_workspace_bytes)");
    const size_t frame_ws0 = off;

    RMPE_CUDA_TRY(cudaMemsetAsync(raw_count, 0, (size_t)B * kParts * 4, st));
    RMPE_CUDA_TRY(cudaMemsetAsync(b->status, 0, (size_t)B * 4, st));
    RMPE_CUDA_TRY(cudaMemsetAsync(cand_count, 0, 256, st));

    // Frames are processed sorted by shape (outputs stay indexed by the caller's frame number): frames of one shape share
    // their operator tables and their k_screen_pairs variant, so a chunk of 64 same-shaped frames builds 8 tables instead
    // of ~240 and runs one plan / pairs launch instead of three (1000 COCO-val shapes: 45 -> 14 table launches per pass).
    std::vector<int> order(B);
    for (int i = 0; i < B; i++) order[i] = i;
    if (B > kChunkFrames)
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
            const RmpeFrameDesc &fx = b->frames_host[x], &fy = b->frames_host[y];
            if (fx.n_scales != fy.n_scales) return fx.n_scales < fy.n_scales;
            if (fx.height != fy.height) return fx.height < fy.height;
            return fx.width < fy.width;
        });
    int f0 = 0;
    while (f0 < B) {
        // take frames while they fit
        int n = 0;
        size_t o = frame_ws0;
        uint8_t *u_ptr[kChunkFrames];
        float *p1_ptr[kChunkFrames], *i1_ptr[kChunkFrames], *p2_ptr[kChunkFrames];
        FramePlan plans[kChunkFrames];
        bool any_single = false, any_multi = false, any_screen = false;
        while (f0 + n < B && n < kChunkFrames) {
            FramePlan p = plan_frame(b->frames_host[order[f0 + n]], b->stride);
            if (o + p.bytes > b->workspace_bytes) break;
            plans[n] = p;
            if (p.screen) {
                u_ptr[n] = ws + o; o += p.bytes;       // the frame's operator tables
                p1_ptr[n] = i1_ptr[n] = p2_ptr[n] = nullptr;
                any_screen = true;
            } else {
                u_ptr[n] = ws + o; o += al256(p.u_elems * (p.multi ? 8 : 4));
                p1_ptr[n] = (float *)(ws + o); o += al256(p.p1_elems * 4);
                i1_ptr[n] = (float *)(ws + o); o += al256(p.i1_elems * 4);
                p2_ptr[n] = (float *)(ws + o); o += al256(p.p2_elems * 4);
                (p.multi ? any_multi : any_single) = true;
            }
            n++;
        }
        RMPE_REQUIRE(n > 0, "workspace too small for one frame (see rmpe_decode_workspace_bytes)");
        RmpeFrameDesc fr[kChunkFrames];          // the chunk's descriptors in processing order, fid[i] = caller's frame number
        int fid[kChunkFrames];
        for (int i = 0; i < n; i++) { fid[i] = order[f0 + i]; fr[i] = b->frames_host[fid[i]]; }

        if (any_screen) {
            // ---- screen in float32 straight from the blobs, decide exactly per surviving pixel ----
            AxisJobs aj{};
            // single scale (Kx <= 10 / <= 12 wide), multi scale (two CTAs per SM fit / only one fits), multi scale wide.
            // A launch allocates the largest shared-memory need among its frames: ONE frame above ~110 KB used to put a
            // whole chunk at one CTA per SM (9 % of the COCO-val shapes are that wide, i.e. nearly every chunk of 64)
            MsJobs jobs1{}, jobs2{}, jobsM{}, jobsH{}, jobsB{};
            int n1 = 0, n2 = 0, nM = 0, nH = 0, nB = 0, t1 = 0, t2 = 0, tM = 0, tH = 0, tB = 0;
            size_t sm1 = 0, sm2 = 0, smM = 0, smH = 0, smB = 0;
            int n_tab = 0, max_len = 0;
            const int cand_cap = (int)std::min<size_t>(per_list * 2, (size_t)n * kParts * MP * 2);
            if (f0 > 0) {
                RMPE_CUDA_TRY(cudaMemsetAsync(cand_count, 0, 4 * (1 + kScreenSlots), st));
                RMPE_CUDA_TRY(cudaMemsetAsync(cand_count + 16, 0, 4 * kScreenSlots, st));
            }
            // the tables of a previous call are still there only if that call (same frames, same workspace) was ONE
            // chunk: every further chunk rebuilds its tables in the same workspace region
            const bool reuse_tables = (b->flags & RMPE_DECODE_REUSE_TABLES) != 0 && f0 == 0 && n == B;
            bool tables_failed = false;
            auto flush_tables = [&]() {
                if (!n_tab) return;
                if (reuse_tables) { n_tab = 0; max_len = 0; return; }
                ProfScope ps("k_axis_tables", st);
                k_axis_tables<<<dim3((max_len + 127) / 128, n_tab), 128, 0, st>>>(aj, tab_err);
                if (cudaGetLastError() != cudaSuccess) tables_failed = true;
                count_launch();
                n_tab = 0; max_len = 0;
            };
            // a chunk of screened frames only holds nothing but their tables: one memset for all of them
            const bool tables_cleared = !reuse_tables && !any_single && !any_multi;
            if (tables_cleared) RMPE_CUDA_TRY(cudaMemsetAsync(ws + frame_ws0, 0, o - frame_ws0, st));
            for (int i = 0; i < n; i++) {
                if (!plans[i].screen) continue;
                const RmpeFrameDesc &f = fr[i];
                const FramePlan &p = plans[i];
                if (n_tab + 2 * f.n_scales > 2 * kChunkFrames) flush_tables();
                // frames of one shape share their operators: build each table once per launch of k_axis_tables
                auto table = [&](const AxisJob &want, float *&K, int *&lo) {
                    for (int q = 0; q < n_tab; q++) {
                        const AxisJob &o = aj.j[q];
                        if (o.dst == want.dst && o.src == want.src && o.kw == want.kw && o.mid == want.mid &&
                            o.stride == want.stride && o.wscale == want.wscale) { K = o.K; lo = o.lo; return; }
                    }
                    aj.j[n_tab] = want; aj.j[n_tab].K = K; aj.j[n_tab].lo = lo;
                    n_tab++;
                    max_len = std::max(max_len, want.dst);
                };
                uint8_t *tp = u_ptr[i];
                if (!reuse_tables && !tables_cleared) RMPE_CUDA_TRY(cudaMemsetAsync(tp, 0, p.bytes, st));   // row-mass maxima start at 0
                MsJob mj{};
                for (int sI = 0; sI < f.n_scales; sI++) {
                    float *Ky = (float *)tp; tp += (size_t)f.height * p.kwy[sI] * 4;
                    int *loy = (int *)tp; tp += (size_t)(f.height + 2) * 4;
                    float *Kx = (float *)tp; tp += (size_t)f.width * p.kwx[sI] * 4;
                    int *lox = (int *)tp; tp += (size_t)(f.width + 2) * 4;
                    const int h = f.grid_h[sI], w = f.grid_w[sI];
                    if (!p.multi) {
                        table(AxisJob{f.height, h, p.kwy[sI], 0, 1, 1.0f, nullptr, nullptr}, Ky, loy);
                        table(AxisJob{f.width, w, p.kwx[sI], 0, 1, 1.0f, nullptr, nullptr}, Kx, lox);
                    } else {
                        table(AxisJob{f.height, h, p.kwy[sI], h * b->stride - f.pad_down[sI], b->stride,
                                      1.0f / (float)f.n_scales, nullptr, nullptr}, Ky, loy);
                        table(AxisJob{f.width, w, p.kwx[sI], w * b->stride - f.pad_right[sI], b->stride, 1.0f, nullptr, nullptr},
                              Kx, lox);
                    }
                    mj.sc[sI] = MsScale{b->heat + f.heat_offset[sI], Ky, Kx, loy, lox, h, w, p.kwy[sI], p.kwx[sI]};
                }
                mj.H = f.height; mj.W = f.width; mj.n_scales = f.n_scales; mj.frame = fid[i];
                mj.tiles_x = (f.width + kScrTW - 1) / kScrTW;
                mj.tiles = mj.tiles_x * ((f.height + kScrTH - 1) / kScrTH);
                if (p.multi && p.big) { jobsB.j[nB++] = mj; tB = std::max(tB, mj.tiles); smB = std::max(smB, p.smem); }
                else if (p.multi && ctas_per_sm(p.smem) >= 2) { jobsM.j[nM++] = mj; tM = std::max(tM, mj.tiles); smM = std::max(smM, p.smem); }
                else if (p.multi) { jobsH.j[nH++] = mj; tH = std::max(tH, mj.tiles); smH = std::max(smH, p.smem); }
                else if (p.kwx[0] <= 10) { jobs1.j[n1++] = mj; t1 = std::max(t1, mj.tiles); sm1 = std::max(sm1, p.smem); }
                else { jobs2.j[n2++] = mj; t2 = std::max(t2, mj.tiles); sm2 = std::max(sm2, p.smem); }
            }
            flush_tables();
            if (tables_failed) { set_error("k_axis_tables launch failed"); return RMPE_E_CUDA; }
            // plan (which (tile, part) pairs can hold a peak) + pairs (screen them), per kernel variant
            const int sms = tables().sm_count;
            auto screen = [&](const MsJobs &jobs, int nj, int mt, size_t smem, int variant, int slot) -> int {
                if (!nj) return RMPE_OK;
                int32_t *cnt = cand_count + 1 + slot;      // this variant's work-item counter (zeroed with cand_count)
                int32_t *nxt = cand_count + 16 + slot;     // next item k_screen_pairs hands out
                ActEntry *lst = act + (size_t)slot * act_cap;
                float *lstA = act_A + (size_t)slot * act_cap * kParts;
                // parts per work item (measured): single-scale items are cheap to set up, multi-scale items stage the operators
                // of four scales (amortise them over up to nine active parts of the tile)
                static const int ms_group = [] {      // 256 COCO-val-shaped frames: 6 parts per item 7.23 ms, 9: 7.09, 12: 7.13, 18: 7.23 (before the plan refinement; 5.78 / 5.86 with it)
                    const char *e = getenv("RMPE_MS_GROUP");
                    int v = e ? atoi(e) : 9;
                    return (v >= 1 && v <= kParts) ? v : 9;
                }();
                static const int ss_group = [] {      // 8 ski-shaped frames: 3 parts per item 0.109 ms, 6: 0.111; 64 frames: 3: 0.402, 6: 0.388, 9: 0.388
                    const char *e = getenv("RMPE_SS_GROUP");
                    int v = e ? atoi(e) : 0;
                    return (v >= 1 && v <= kParts) ? v : 0;
                }();
                // few frames: small items spread better over the SMs; many: larger items amortise the set-up of a tile
                const int group = std::min(16, (variant >= 2) ? ms_group : (ss_group ? ss_group : (nj > 16 ? 6 : 3)));   // 16 nibbles of ActEntry::live
                // row / column refinement in k_screen_plan and column-group skipping in k_screen_pairs (RMPE_SCREEN_CULL=0: the
                // tile bound alone decides and every column group of an active tile is evaluated)
                static const int cull = [] { const char *e = getenv("RMPE_SCREEN_CULL"); return (e && atoi(e) == 0) ? 0 : 1; }();
                {
                    ProfScope ps("k_screen_plan", st);
                    k_screen_plan<<<dim3(mt, nj), kPlanThreads, 0, st>>>(jobs, (float)b->thre1, act_cap, group, lst, lstA, cnt, tab_err,
                                                                       b->status, cull);
                    RMPE_CUDA_TRY(cudaGetLastError());
                }
                {
                    ProfScope ps("k_screen_pairs", st);
                    const int per_sm = ctas_per_sm(smem);
                    const int grid = sms * per_sm;
                    if (variant == 0)
                        RMPE_CUDA_TRY(launch_pdl(k_screen_pairs<10, kScrMaxSrcRows>, dim3(grid), dim3(kScrThreads), smem, st,
                            jobs, (float)b->thre1, act_cap, lst, lstA, cnt, nxt, cand_cap, cand_key, cand_fp, cand_count, b->status, cull));
                    else if (variant == 1)
                        RMPE_CUDA_TRY(launch_pdl(k_screen_pairs<kScrMaxKW, kScrMaxSrcRows>, dim3(grid), dim3(kScrThreads), smem, st,
                            jobs, (float)b->thre1, act_cap, lst, lstA, cnt, nxt, cand_cap, cand_key, cand_fp, cand_count, b->status, cull));
                    else if (variant == 2)
                        RMPE_CUDA_TRY(launch_pdl(k_screen_pairs<kMsKW, kMsMaxRows>, dim3(grid), dim3(kScrThreads), smem, st,
                            jobs, (float)b->thre1, act_cap, lst, lstA, cnt, nxt, cand_cap, cand_key, cand_fp, cand_count, b->status, cull));
                    else
                        RMPE_CUDA_TRY(launch_pdl(k_screen_pairs<kMsKWBig, kMsMaxRowsBig>, dim3(grid), dim3(kScrThreads), smem, st,
                            jobs, (float)b->thre1, act_cap, lst, lstA, cnt, nxt, cand_cap, cand_key, cand_fp, cand_count, b->status, cull));
                }
                count_launch(2);
                return RMPE_OK;
            };
            // a failed launch must not let the verify / finalize kernels run on stale lists
            if ((rc = screen(jobs1, n1, t1, sm1, 0, 0)) != RMPE_OK) return rc;
            if ((rc = screen(jobs2, n2, t2, sm2, 1, 1)) != RMPE_OK) return rc;
            if ((rc = screen(jobsM, nM, tM, smM, 2, 2)) != RMPE_OK) return rc;
            if ((rc = screen(jobsB, nB, tB, smB, 3, 3)) != RMPE_OK) return rc;
            if ((rc = screen(jobsH, nH, tH, smH, 2, 4)) != RMPE_OK) return rc;
            {
                ProfScope ps("k_peak_verify", st);
                // one candidate per CTA and round.  Resident CTAs per SM do not decide (4 / 6 / 8 by register cap measured
                // 0.99 / 0.98 / 1.02 ms per 256 multi-scale frames): the kernel waits for instruction fetches (ncu: no_instruction)
                const int grid = std::min(cand_cap, 8 * sms);
                RMPE_CUDA_TRY(launch_pdl(k_peak_verify, dim3(grid), dim3(kVerThreads), 0, st, b->frames, b->heat, b->stride, b->thre1,
                                         cand_cap, cand_key, cand_fp, cand_count, MP, raw_key, raw_score, raw_count, b->status));
                count_launch();
            }
        }
        if (any_single || any_multi) {
            rc = launch_heat_up(fr, n, b->heat, b->stride, u_ptr, p1_ptr, i1_ptr, p2_ptr, plans, st);
            if (rc != RMPE_OK) return rc;
        }

        // smooth + peaks on the materialised maps: one launch per map dtype present in the chunk
        for (int pass = 0; pass < 2; pass++) {
            bool multi = (pass == 1);
            if (!(multi ? any_multi : any_single)) continue;
            SmoothJobs sj{};
            int m = 0, max_tiles = 0;
            for (int i = 0; i < n; i++) {
                if (plans[i].screen || (fr[i].n_scales > 1) != multi) continue;
                sj.j[m].U = u_ptr[i]; sj.j[m].H = fr[i].height; sj.j[m].W = fr[i].width;
                sj.j[m].frame = fid[i]; sj.j[m].S_out = nullptr;
                int t = ((fr[i].height + kST - 1) / kST) * ((fr[i].width + kST - 1) / kST);
                max_tiles = max(max_tiles, t);
                m++;
            }
            dim3 grid(max_tiles, kParts, m);
            ProfScope ps("k_smooth_peaks", st);
            if (multi)
                k_smooth_peaks<double><<<grid, kSmoothThreads, smooth_smem_bytes(true), st>>>(
                    sj, b->thre1, MP, raw_key, raw_score, raw_count, b->status);
            else
                k_smooth_peaks<float><<<grid, kSmoothThreads, smooth_smem_bytes(false), st>>>(
                    sj, (float)b->thre1, MP, raw_key, raw_score, raw_count, b->status);
            count_launch();
        }
        {
            FinalizeJobs fj{};
            for (int i = 0; i < n; i++) { fj.W[i] = fr[i].width; fj.frame[i] = fid[i]; }
            ProfScope ps("k_peaks_finalize", st);
            RMPE_CUDA_TRY(launch_pdl(k_peaks_finalize, dim3(kParts, n), dim3(256), 0, st, fj, f0, MP, raw_key, raw_score, raw_count,
                                     b->candidate, b->n_peaks, pk_x, pk_y, pk_s));
            count_launch();
        }
        f0 += n;
    }
    // limbs + assembly for the whole batch
    {
        size_t smem = (size_t)MC * 12 + 2 * kMaxPeaksCap;
        {
            ProfScope ps("k_limbs", st);
            RMPE_CUDA_TRY(launch_pdl(k_limbs, dim3(kLimbs, B), dim3(kLimbThreads), smem, st, b->frames, 0, b->paf, b->stride,
                                     b->thre2, MP, MC, b->n_peaks, pk_x, pk_y, pk_s, b->limb_cand, b->n_limb_cand,
                                     b->connections, b->n_conn, ws_cand, b->status));
        }
        {
            ProfScope ps("k_assemble", st);
            const size_t asm_smem = assemble_smem_bytes(MP);
            RMPE_CUDA_TRY(launch_pdl(k_assemble, dim3(B), dim3(32), asm_smem, st, 0, MP, b->max_persons, b->candidate,
                                     b->connections, b->n_conn, b->n_peaks, b->subset, b->n_subset, b->status));
        }
        count_launch(2);
    }
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}

// ------------------------------------------------------------------------------------------
// debugging / stage-parity hooks
// ------------------------------------------------------------------------------------------
extern "C" int rmpe_debug_heat_maps(const RmpeFrameDesc *f, const float *heat_dev, void *up_out_dev,
                                    void *smooth_out_dev, void *stream_) {
    if (!is_initialised()) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(f && heat_dev && up_out_dev, "null argument");
    RMPE_REQUIRE(frame_ok(*f, 8), "frame descriptor");
    cudaStream_t st = (cudaStream_t)stream_;
    int rc = ensure_smooth_attr();
    if (rc != RMPE_OK) return rc;
    FramePlan p = plan_frame(*f, 8, false);
    uint8_t *tmp = nullptr;
    size_t scratch = al256(p.p1_elems * 4) + al256(p.i1_elems * 4) + al256(p.p2_elems * 4) + al256(kParts * 8 * 4) * 3;
    RMPE_CUDA_TRY(cudaMalloc((void **)&tmp, scratch + 4096));
    uint8_t *u_ptr[1] = {(uint8_t *)up_out_dev};
    float *p1[1] = {(float *)tmp};
    float *i1[1] = {(float *)(tmp + al256(p.p1_elems * 4))};
    float *p2[1] = {(float *)(tmp + al256(p.p1_elems * 4) + al256(p.i1_elems * 4))};
    rc = launch_heat_up(f, 1, heat_dev, 8, u_ptr, p1, i1, p2, nullptr, st);
    if (rc == RMPE_OK && smooth_out_dev) {
        uint8_t *q = tmp + al256(p.p1_elems * 4) + al256(p.i1_elems * 4) + al256(p.p2_elems * 4);
        int32_t *cnt = (int32_t *)q;
        int32_t *keys = (int32_t *)(q + al256(kParts * 8 * 4));
        double *scores = (double *)(q + 2 * al256(kParts * 8 * 4));
        cudaMemsetAsync(cnt, 0, al256(kParts * 8 * 4), st);
        SmoothJobs sj{};
        sj.j[0].U = up_out_dev; sj.j[0].H = f->height; sj.j[0].W = f->width; sj.j[0].frame = 0;
        sj.j[0].S_out = smooth_out_dev;
        int tiles = ((f->height + kST - 1) / kST) * ((f->width + kST - 1) / kST);
        dim3 grid(tiles, kParts, 1);
        // thre1 = +inf: no peaks are emitted, only the smoothed map is written; status -> cnt[kParts..]
        if (p.multi)
            k_smooth_peaks<double><<<grid, kSmoothThreads, smooth_smem_bytes(true), st>>>(
                sj, 1.0e300, 1, keys, scores, cnt, cnt + kParts);
        else
            k_smooth_peaks<float><<<grid, kSmoothThreads, smooth_smem_bytes(false), st>>>(
                sj, 3.0e38f, 1, keys, scores, cnt, cnt + kParts);
        count_launch();
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (rc != RMPE_OK) return rc;
    RMPE_CUDA_TRY(e);
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}

// k_assemble on caller-made connection lists of ONE frame: the only way to reach the reference's `found > 2` IndexError
// (eval...:192-195) -- lists that come out of k_limbs are one-to-one per limb and never produce it.
extern "C" int rmpe_debug_assemble(int max_peaks, int max_persons, const double *candidate_dev, const double *connections_dev,
                                   const int32_t *n_conn_dev, const int32_t *n_peaks_dev, double *subset_dev,
                                   int32_t *n_subset_dev, int32_t *status_dev, void *stream_) {
    if (!is_initialised()) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(candidate_dev && connections_dev && n_conn_dev && n_peaks_dev && subset_dev && n_subset_dev && status_dev, "null argument");
    RMPE_REQUIRE(max_peaks > 0 && max_peaks <= kMaxPeaksCap && max_persons > 0 && max_persons <= kMaxSubsetCap, "capacities");
    int rc = ensure_smooth_attr();
    if (rc != RMPE_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream_;
    RMPE_CUDA_TRY(cudaMemsetAsync(status_dev, 0, sizeof(int32_t), st));
    const size_t asm_smem = assemble_smem_bytes(max_peaks);
    k_assemble<<<1, 32, asm_smem, st>>>(0, max_peaks, max_persons, candidate_dev, connections_dev, n_conn_dev, n_peaks_dev,
                                        subset_dev, n_subset_dev, status_dev);
    count_launch();
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}

extern "C" int rmpe_debug_paf_points(const RmpeFrameDesc *f, const float *paf_dev, int n, const int32_t *cyx_dev,
                                     double *out_dev, void *stream_) {
    if (!is_initialised()) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(f && paf_dev && cyx_dev && out_dev && n >= 0, "null argument");
    RMPE_REQUIRE(frame_ok(*f, 8), "frame descriptor");
    if (n == 0) return RMPE_OK;
    k_paf_points<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(*f, paf_dev, 8, n, cyx_dev, out_dev);
    count_launch();
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}
