// Target generation on sm_100a: affine warp of image + miss-mask (bit-exact OpenCV fixed-point
// bicubic), fused 368->46 mask resize, keypoint transform and the 57-plane heat/PAF rasteriser.
//
// Reference path replaced (paths relative to the reference root):
//   py_rmpe_server/py_rmpe_transformer.py:83-114  Transformer.transform      -> k_warp_tile / k_warp_simple,
//                                                                               k_mask46, joints in k_raster
//   py_rmpe_server/py_rmpe_heatmapper.py:32-138    Heatmapper.create_heatmaps -> k_raster
#include "rmpe_common.cuh"

namespace rmpe {

// ==========================================================================================
// shared per-pixel bicubic evaluation (generic path: any tap may fall outside the source)
// ==========================================================================================
template <int C>
__device__ inline void warp_pixel_generic(const uint8_t *__restrict__ src, int H, int W, int pitch, int X, int Y,
                                          const int16_t *__restrict__ tab, int border, int out[C]) {
    int sx = sat_short(X >> 5) - 1;
    int sy = sat_short(Y >> 5) - 1;
    if (sx >= W || sx + 4 <= 0 || sy >= H || sy + 4 <= 0) {
#pragma unroll
        for (int c = 0; c < C; c++) out[c] = border;
        return;
    }
    // 16 int16 weights of this sub-pixel phase = 32 aligned bytes
    const uint4 *wp = reinterpret_cast<const uint4 *>(tab + (((Y & 31) * 32 + (X & 31)) << 4));
    uint4 wa = __ldg(wp), wb = __ldg(wp + 1);
    const unsigned wpk[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0;
#pragma unroll
    for (int ky = 0; ky < 4; ky++) {
        int yy = sy + ky;
        bool yin = (unsigned)yy < (unsigned)H;
        const uint8_t *row = src + (size_t)(yin ? yy : 0) * pitch;
#pragma unroll
        for (int kx = 0; kx < 4; kx++) {
            int xx = sx + kx;
            bool in = yin && ((unsigned)xx < (unsigned)W);
            unsigned pk = wpk[ky * 2 + (kx >> 1)];
            int wt = (kx & 1) ? ((int)pk >> 16) : (int)(short)(pk & 0xffff);
#pragma unroll
            for (int c = 0; c < C; c++) {
                int v = border;
                if (in) v = (int)__ldg(row + xx * C + c);
                acc[c] += wt * v;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C; c++) out[c] = min(255, max(0, (acc[c] + 16384) >> 15));
}

// ==========================================================================================
// k_warp_simple: one thread per destination pixel, taps straight from global memory.
// Debug / fallback-for-huge-footprints variant; also the A/B check of k_warp_tile on the GPU.
// ==========================================================================================
struct WarpArgs {
    const uint8_t *src_img;
    const RmpeSrcDesc *desc;
    const double *M;
    uint8_t *out_img;
    int32_t *status;
    const int16_t *tab;
    const uint32_t *tab_dp4a;
    int batch;
    int chw;
};

__global__ void __launch_bounds__(256) k_warp_simple(WarpArgs a) {
    int b = blockIdx.y;
    __shared__ double s_iM[6];
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        s_ok = invert_affine(a.M + 6 * b, s_iM) ? 1 : 0;
        if (!s_ok && blockIdx.x == 0) atomicOr(a.status + b, RMPE_ST_SINGULAR);
    }
    __syncthreads();
    int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= kOutW * kOutH) return;
    int y = pix / kOutW, x = pix - y * kOutW;
    RmpeSrcDesc d = a.desc[b];
    int X = (warp_row_term(s_iM[1], s_iM[2], y) + warp_col_term(s_iM[0], x)) >> 5;
    int Y = (warp_row_term(s_iM[4], s_iM[5], y) + warp_col_term(s_iM[3], x)) >> 5;
    int o[3];
    warp_pixel_generic<3>(a.src_img + d.img_offset, d.height, d.width, d.img_pitch, X, Y, a.tab, 127, o);
    uint8_t *out = a.out_img + (size_t)b * (3 * kOutW * kOutH);
    if (a.chw) {
        out[pix] = (uint8_t)o[0];
        out[kOutW * kOutH + pix] = (uint8_t)o[1];
        out[2 * kOutW * kOutH + pix] = (uint8_t)o[2];
    } else {
        out[pix * 3 + 0] = (uint8_t)o[0];
        out[pix * 3 + 1] = (uint8_t)o[1];
        out[pix * 3 + 2] = (uint8_t)o[2];
    }
}

// ==========================================================================================
// k_warp_tile: persistent CTAs; per 32x32 destination tile the clipped source footprint is
// staged row by row into shared memory with cp.async.bulk (TMA bulk copies, double buffered on
// mbarriers), the 16-tap sum runs as dp4a over hi/lo byte-split int16 weights (exact), and the
// finished tile leaves through a shared-memory staging buffer as 16-byte row segments.
// ==========================================================================================
constexpr int kTile = 32;
constexpr int kTilesX = (kOutW + kTile - 1) / kTile;  // 12
constexpr int kTilesPerSample = kTilesX * kTilesX;    // 144
constexpr int kFootRows = 96;                         // staged source rows per tile (max)
constexpr int kFootPitch = 336;                       // bytes per staged row (multiple of 16)
constexpr int kWarpConsumers = 512;              // 16 compute warps
constexpr int kWarpThreads = kWarpConsumers + 32;  // + 1 producer warp (bulk-copy issue)
constexpr int kTabBytes = 32 * 32 * 8 * 4;            // dp4a table

struct TileGeom {      // per staged tile, written by warp 0
    int X0[kTile];     // row terms  (+16 rounding already in)
    int Y0[kTile];
    int ad[kTile];     // column terms
    int bd[kTile];
    int fx0, fy0, fx1, fy1;  // clipped footprint in source pixels (inclusive); fx1 < fx0 = empty
    int staged;              // 1 = footprint is in shared memory, 0 = take taps from global
    int g0;                  // (address of footprint row 0, col fx0) & 15
    int sample, x0, y0, tw, th;
    int ok;                  // matrix invertible
};

struct WarpSmem {
    uint32_t tab[kTabBytes / 4];
    uint8_t foot[2][kFootRows * kFootPitch];
    uint8_t outbuf[kTile * kTile * 3];
    TileGeom geom[2];
    uint64_t full[2];   // producer lanes + bulk-copy bytes -> consumers
    uint64_t empty[2];  // consumer warps -> producer
};

__device__ inline void stage_tile(const WarpArgs &a, WarpSmem &s, int buf, int item) {
    // executed by warp 0 (all 32 lanes)
    int lane = threadIdx.x & 31;
    TileGeom &g = s.geom[buf];
    int sample = item / kTilesPerSample;
    int t = item - sample * kTilesPerSample;
    int ty = t / kTilesX, tx = t - ty * kTilesX;
    int x0 = tx * kTile, y0 = ty * kTile;
    int tw = min(kTile, kOutW - x0), th = min(kTile, kOutH - y0);
    double iM[6];
    bool ok = invert_affine(a.M + 6 * sample, iM);
    g.X0[lane] = warp_row_term(iM[1], iM[2], y0 + lane);
    g.Y0[lane] = warp_row_term(iM[4], iM[5], y0 + lane);
    g.ad[lane] = warp_col_term(iM[0], x0 + lane);
    g.bd[lane] = warp_col_term(iM[3], x0 + lane);
    __syncwarp();
    RmpeSrcDesc d = a.desc[sample];
    // footprint from the 4 corners: X(x,y) = (X0[y]+ad[x])>>5 is monotone in x and in y
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int cx = (c & 1) ? tw - 1 : 0, cy = (c & 2) ? th - 1 : 0;
        int X = (g.X0[cy] + g.ad[cx]) >> 5, Y = (g.Y0[cy] + g.bd[cx]) >> 5;
        int sx = sat_short(X >> 5) - 1, sy = sat_short(Y >> 5) - 1;
        mnx = min(mnx, sx); mxx = max(mxx, sx + 3);
        mny = min(mny, sy); mxy = max(mxy, sy + 3);
    }
    int fx0 = max(mnx, 0), fx1 = min(mxx, d.width - 1);
    int fy0 = max(mny, 0), fy1 = min(mxy, d.height - 1);
    bool empty = (fx1 < fx0) || (fy1 < fy0);
    const uint8_t *base = a.src_img + d.img_offset;
    size_t addr0 = (size_t)(base + (size_t)(empty ? 0 : fy0) * d.img_pitch + 3 * (empty ? 0 : fx0));
    int rows = empty ? 0 : fy1 - fy0 + 1;
    int row_bytes = empty ? 0 : 3 * (fx1 - fx0 + 1);
    // every staged row must fit: misalignment (<16) + payload + funnel slack (4 words) + round-up
    bool fits = rows <= kFootRows && (row_bytes + 15 + 16 + 15) <= kFootPitch &&
                ((size_t)a.src_img & 15) == 0;
    unsigned my_bytes = 0;
    if (!empty && fits) {
        for (int r = lane; r < rows; r += 32) {
            size_t ga = addr0 + (size_t)r * d.img_pitch;
            size_t ga0 = ga & ~(size_t)15;
            unsigned nbytes = (unsigned)(((ga + row_bytes + 15) & ~(size_t)15) - ga0);
            bulk_g2s(&s.foot[buf][r * kFootPitch], (const void *)ga0, nbytes, &s.full[buf]);
            my_bytes += nbytes;
        }
    }
    if (lane == 0) {
        g.fx0 = fx0; g.fy0 = fy0; g.fx1 = empty ? fx0 - 1 : fx1; g.fy1 = empty ? fy0 - 1 : fy1;
        g.staged = (!empty && fits) ? 1 : 0;
        g.g0 = (int)(addr0 & 15);
        g.sample = sample; g.x0 = x0; g.y0 = y0; g.tw = tw; g.th = th;
        g.ok = ok ? 1 : 0;
    }
    __syncwarp();
    if (my_bytes) mbar_arrive_expect_tx(&s.full[buf], my_bytes);  // release: geom + copies
    else mbar_arrive(&s.full[buf]);
}

__global__ void __launch_bounds__(kWarpThreads, 2) k_warp_tile(WarpArgs a, int n_items) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    WarpSmem &s = *reinterpret_cast<WarpSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        mbar_init(&s.full[0], 32);
        mbar_init(&s.full[1], 32);
        mbar_init(&s.empty[0], kWarpConsumers / 32);
        mbar_init(&s.empty[1], kWarpConsumers / 32);
        mbar_fence_init();
    }
    // dp4a weight table -> shared memory (32 KB, once per persistent CTA)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.tab_dp4a);
        uint4 *dst = reinterpret_cast<uint4 *>(s.tab);
        for (int i = tid; i < kTabBytes / 16; i += kWarpThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    if (warp == kWarpConsumers / 32) {
        // ---------------- producer warp: geometry + bulk copies, two tiles ahead ----------------
        int k = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, k++) {
            int buf = k & 1;
            if (k >= 2) {
                mbar_wait(&s.empty[buf], ((k >> 1) - 1) & 1);
                fence_proxy_async();
            }
            stage_tile(a, s, buf, item);
        }
        return;
    }

    // ---------------- consumer warps ----------------
    int k = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, k++) {
        const int buf = k & 1;
        mbar_wait(&s.full[buf], (k >> 1) & 1);

        const TileGeom &g = s.geom[buf];
        const RmpeSrcDesc d = a.desc[g.sample];
        const uint8_t *gsrc = a.src_img + d.img_offset;
        const uint8_t *foot = s.foot[buf];
        const int staged = g.staged;
        const int fx0 = g.fx0, fy0 = g.fy0, fx1 = g.fx1, fy1 = g.fy1, g0 = g.g0;
        const int pitch = d.img_pitch;
        const int tw = g.tw, th = g.th, x0 = g.x0, y0 = g.y0, sample = g.sample;

        // warp w owns tile rows w and w+16; lane = tile column
        const int adx = g.ad[lane], bdx = g.bd[lane];
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int ly = warp + half * 16;
            if (lane < tw && ly < th) {
                int X = (g.X0[ly] + adx) >> 5;
                int Y = (g.Y0[ly] + bdx) >> 5;
                int sx = sat_short(X >> 5) - 1;
                int sy = sat_short(Y >> 5) - 1;
                int o0, o1, o2;
                if (staged && sx >= fx0 && sx + 3 <= fx1 && sy >= fy0 && sy + 3 <= fy1) {
                    const uint4 *wp = reinterpret_cast<const uint4 *>(s.tab + (((Y & 31) * 32 + (X & 31)) << 3));
                    uint4 wa = wp[0], wb = wp[1];  // rows 0,1 and rows 2,3: {hi,lo,hi,lo}
                    const unsigned whi[4] = {wa.x, wa.z, wb.x, wb.z};
                    const unsigned wlo[4] = {wa.y, wa.w, wb.y, wb.w};
                    int hB = 0, hG = 0, hR = 0, lB = 0, lG = 0, lR = 0;
                    int r0i = sy - fy0;
                    int colb = 3 * (sx - fx0);
#pragma unroll
                    for (int ky = 0; ky < 4; ky++) {
                        int r = r0i + ky;
                        int aoff = r * kFootPitch + ((g0 + r * pitch) & 15) + colb;
                        const uint32_t *wptr = reinterpret_cast<const uint32_t *>(foot + (aoff & ~3));
                        unsigned w0 = wptr[0], w1 = wptr[1], w2 = wptr[2], w3 = wptr[3];
                        unsigned sh = (aoff & 3) * 8;
                        unsigned q0 = __funnelshift_r(w0, w1, sh);  // B0 G0 R0 B1
                        unsigned q1 = __funnelshift_r(w1, w2, sh);  // G1 R1 B2 G2
                        unsigned q2 = __funnelshift_r(w2, w3, sh);  // R2 B3 G3 R3
                        unsigned pb = __byte_perm(__byte_perm(q0, q1, 0x0630), q2, 0x5210);
                        unsigned pg = __byte_perm(__byte_perm(q0, q1, 0x0741), q2, 0x6210);
                        unsigned pr = __byte_perm(__byte_perm(q0, q1, 0x0052), q2, 0x7410);
                        hB = dp4a_us(pb, whi[ky], hB); lB = dp4a_uu(pb, wlo[ky], lB);
                        hG = dp4a_us(pg, whi[ky], hG); lG = dp4a_uu(pg, wlo[ky], lG);
                        hR = dp4a_us(pr, whi[ky], hR); lR = dp4a_uu(pr, wlo[ky], lR);
                    }
                    o0 = min(255, max(0, (hB * 256 + lB + 16384) >> 15));
                    o1 = min(255, max(0, (hG * 256 + lG + 16384) >> 15));
                    o2 = min(255, max(0, (hR * 256 + lR + 16384) >> 15));
                } else {
                    int o[3];
                    warp_pixel_generic<3>(gsrc, d.height, d.width, pitch, X, Y, a.tab, 127, o);
                    o0 = o[0]; o1 = o[1]; o2 = o[2];
                }
                if (a.chw) {
                    s.outbuf[(0 * kTile + ly) * kTile + lane] = (uint8_t)o0;
                    s.outbuf[(1 * kTile + ly) * kTile + lane] = (uint8_t)o1;
                    s.outbuf[(2 * kTile + ly) * kTile + lane] = (uint8_t)o2;
                } else {
                    uint8_t *ob = s.outbuf + (ly * kTile + lane) * 3;
                    ob[0] = (uint8_t)o0; ob[1] = (uint8_t)o1; ob[2] = (uint8_t)o2;
                }
            }
        }
        if (tid == 0 && !g.ok && x0 == 0 && y0 == 0) atomicOr(a.status + sample, RMPE_ST_SINGULAR);
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[buf]);  // this warp is done with foot[buf] / geom[buf]
        // this warp's two rows leave as 16-byte segments: HWC row = 96 B (48 B for the 16-px edge
        // tile), CHW plane row = 32 B (16 B at the edge)
        uint8_t *out = a.out_img + (size_t)sample * (3 * kOutW * kOutH);
        if (!a.chw) {
            int segs = tw * 3 / 16;
            if (lane < 2 * segs) {
                int half = lane / segs, sg = lane - half * segs;
                int r = warp + half * 16;
                if (r < th) {
                    uint4 v = *reinterpret_cast<const uint4 *>(s.outbuf + r * (kTile * 3) + sg * 16);
                    *reinterpret_cast<uint4 *>(out + ((size_t)(y0 + r) * kOutW + x0) * 3 + sg * 16) = v;
                }
            }
        } else {
            int segs = tw / 16;
            if (lane < 6 * segs) {
                int q = lane / segs, sg = lane - q * segs;  // q = plane*2 + half
                int pl = q >> 1, r = warp + (q & 1) * 16;
                if (r < th) {
                    uint4 v = *reinterpret_cast<const uint4 *>(s.outbuf + (pl * kTile + r) * kTile + sg * 16);
                    *reinterpret_cast<uint4 *>(out + (size_t)pl * (kOutW * kOutH) + (size_t)(y0 + r) * kOutW + x0 +
                                               sg * 16) = v;
                }
            }
        }
        __syncwarp();  // outbuf rows of this warp are free again
    }
}

// ==========================================================================================
// k_mask46: fused warpAffine(mask, border 255) + cv2.resize(368->46, INTER_CUBIC) + /255.
// Only rows/cols 8d+2..8d+5 of the warped mask feed the resize, so each thread evaluates the
// 4x4 warped pixels of its own cell and reduces them with OpenCV's weights:
// horizontal exact int32 (-192,1216,1216,-192), vertical float32 FMA chain, rint, saturate.
// ==========================================================================================
struct MaskArgs {
    const uint8_t *src_mask;
    const RmpeSrcDesc *desc;
    const double *M;
    void *out_mask;
    const int16_t *tab;
    int f64;
};

__global__ void __launch_bounds__(128) k_mask46(MaskArgs a) {
    int b = blockIdx.y;
    __shared__ double s_iM[6];
    if (threadIdx.x == 0) invert_affine(a.M + 6 * b, s_iM);
    __syncthreads();
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= kCells) return;
    int cy = cell / kGrid, cx = cell - cy * kGrid;
    RmpeSrcDesc d = a.desc[b];
    const uint8_t *src = a.src_mask + d.mask_offset;
    int ad[4], bd[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        ad[k] = warp_col_term(s_iM[0], 8 * cx + 2 + k);
        bd[k] = warp_col_term(s_iM[3], 8 * cx + 2 + k);
    }
    const int wv[4] = {-192, 1216, 1216, -192};
    float S[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int y = 8 * cy + 2 + j;
        int X0 = warp_row_term(s_iM[1], s_iM[2], y);
        int Y0 = warp_row_term(s_iM[4], s_iM[5], y);
        int hsum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int o[1];
            warp_pixel_generic<1>(src, d.height, d.width, d.mask_pitch, (X0 + ad[k]) >> 5, (Y0 + bd[k]) >> 5, a.tab,
                                  255, o);
            hsum += wv[k] * o[0];
        }
        S[j] = (float)hsum;
    }
    const float sc = 1.0f / 4194304.0f;  // 2^-22
    float b0 = -192.0f * sc, b1 = 1216.0f * sc;
    float v = S[3] * b0;
    v = fmaf(S[2], b1, v);
    v = fmaf(S[1], b1, v);
    v = fmaf(S[0], b0, v);
    int iv = min(255, max(0, __float2int_rn(v)));
    double m = __ddiv_rn((double)iv, 255.0);
    if (a.f64) reinterpret_cast<double *>(a.out_mask)[(size_t)b * kCells + cell] = m;
    else reinterpret_cast<float *>(a.out_mask)[(size_t)b * kCells + cell] = (float)m;
}

// ==========================================================================================
// k_raster: (57,46,46) labels of one sample.  grid = (3 channel groups, batch):
//   group 0: 18 Gaussian part maps (max-merge) + background, group 1: limbs 0..9, group 2: 10..18.
// Each thread owns float4 runs of a plane (pixels 4i..4i+3) so every store is a coalesced
// 128-bit write; joints are transformed (T4) in shared memory by every CTA, group 0 writes them.
// ==========================================================================================
struct RasterArgs {
    const double *joints;
    const int32_t *n_persons;
    const double *M;
    const uint8_t *flip;
    const void *mask;     // [B][46][46] f32/f64
    void *labels;
    double *out_joints;
    int32_t *out_count;
    int32_t *status;
    int max_persons;
    int f64;
    int no_transform;
    double sigma, thre;
};

struct LimbRec {
    double x1, y1, xD, yD, norm2;
    float ux, uy;
    int minx, maxx, miny, maxy;  // cell box, max exclusive; valid iff maxx > minx
};

constexpr int kRasterThreads = 288;
constexpr int kRasterChunks = 2;  // 2 * 288 >= 529
constexpr int kHeatBatch = 32;    // persons whose separable exp tables are resident at once

template <typename T>
__device__ inline void store4(T *p, float a, float b, float c, float d, const T m[4]);
template <>
__device__ inline void store4<float>(float *p, float a, float b, float c, float d, const float m[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(a * m[0], b * m[1], c * m[2], d * m[3]);
}
template <>
__device__ inline void store4<double>(double *p, float a, float b, float c, float d, const double m[4]) {
    reinterpret_cast<double2 *>(p)[0] = make_double2((double)a * m[0], (double)b * m[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2((double)c * m[2], (double)d * m[3]);
}

// python-3 round(): half to even, on an f64 value already divided by the stride
__device__ inline int py_round(double v) { return __double2int_rn(v); }

template <typename T>
__global__ void __launch_bounds__(kRasterThreads) k_raster(RasterArgs a) {
    const int b = blockIdx.y;
    const int group = blockIdx.x;
    const int tid = threadIdx.x;
    const int P = min(a.n_persons[b], kMaxPersonsGt);

    __shared__ double s_j[kMaxPersonsGt * kParts * 3];
    __shared__ float s_ex[kHeatBatch * kGrid];
    __shared__ float s_ey[kHeatBatch * kGrid];
    __shared__ LimbRec s_rec[kMaxPersonsGt];

    // ---- T4: keypoint transform + flip swap (py_rmpe_transformer.py:100-111) ----
    for (int i = tid; i < P * kParts; i += kRasterThreads) {
        int p = i / kParts, part = i - p * kParts;
        const double *jin = a.joints + ((size_t)b * a.max_persons + p) * (kParts * 3);
        double ox, oy, ov;
        if (a.no_transform) {
            ox = jin[part * 3 + 0]; oy = jin[part * 3 + 1]; ov = jin[part * 3 + 2];
        } else {
            int sp = a.flip[b] ? c_flip_partner[part] : part;
            double x = jin[sp * 3 + 0], y = jin[sp * 3 + 1];
            ov = jin[sp * 3 + 2];
            const double *M = a.M + 6 * b;
            // numpy matmul order: t = M0*x; t = fma(M1, y, t); out = t + M2
            ox = __dadd_rn(__fma_rn(M[1], y, __dmul_rn(M[0], x)), M[2]);
            oy = __dadd_rn(__fma_rn(M[4], y, __dmul_rn(M[3], x)), M[5]);
        }
        s_j[i * 3 + 0] = ox; s_j[i * 3 + 1] = oy; s_j[i * 3 + 2] = ov;
        if (group == 0 && a.out_joints) {
            double *jo = a.out_joints + ((size_t)b * a.max_persons + p) * (kParts * 3) + part * 3;
            jo[0] = ox; jo[1] = oy; jo[2] = ov;
        }
    }
    __syncthreads();

    // mask values of this thread's pixels
    T m[kRasterChunks][4];
    const T *mk = reinterpret_cast<const T *>(a.mask) + (size_t)b * kCells;
#pragma unroll
    for (int c = 0; c < kRasterChunks; c++) {
        int i = tid + c * kRasterThreads;
#pragma unroll
        for (int q = 0; q < 4; q++) m[c][q] = (i < kCellVec) ? mk[4 * i + q] : (T)0;
    }
    T *lab = reinterpret_cast<T *>(a.labels) + (size_t)b * kLayers * kCells;

    if (group == 0) {
        // ---- H2/H3: Gaussian part maps with max merge, background = 1 - max ----
        const float inv2s2 = (float)(1.0 / (2.0 * a.sigma * a.sigma));
        float bk[kRasterChunks][4];
#pragma unroll
        for (int c = 0; c < kRasterChunks; c++)
#pragma unroll
            for (int q = 0; q < 4; q++) bk[c][q] = 0.f;
        for (int part = 0; part < kParts; part++) {
            float v[kRasterChunks][4];
#pragma unroll
            for (int c = 0; c < kRasterChunks; c++)
#pragma unroll
                for (int q = 0; q < 4; q++) v[c][q] = 0.f;
            for (int pb = 0; pb < P; pb += kHeatBatch) {
                const int np = min(kHeatBatch, P - pb);
                for (int i = tid; i < np * kGrid; i += kRasterThreads) {
                    int p = i / kGrid, cidx = i - p * kGrid;
                    const double *j = s_j + ((pb + p) * kParts + part) * 3;
                    double g = 8.0 * cidx + 3.5;   // cell centres (py_rmpe_heatmapper.py:22-23)
                    float dx = (float)(g - j[0]), dy = (float)(g - j[1]);
                    bool vis = j[2] < 2.0;
                    s_ex[i] = vis ? expf(-(dx * dx) * inv2s2) : 0.f;
                    s_ey[i] = vis ? expf(-(dy * dy) * inv2s2) : 0.f;
                }
                __syncthreads();
#pragma unroll
                for (int c = 0; c < kRasterChunks; c++) {
                    int i = tid + c * kRasterThreads;
                    if (i < kCellVec) {
                        int pix = 4 * i;
                        int y[4], x[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) { y[q] = (pix + q) / kGrid; x[q] = (pix + q) - y[q] * kGrid; }
                        for (int p = 0; p < np; p++) {
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                v[c][q] = fmaxf(v[c][q], s_ey[p * kGrid + y[q]] * s_ex[p * kGrid + x[q]]);
                        }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int c = 0; c < kRasterChunks; c++) {
                int i = tid + c * kRasterThreads;
                if (i < kCellVec) {
#pragma unroll
                    for (int q = 0; q < 4; q++) bk[c][q] = fmaxf(bk[c][q], v[c][q]);
                    store4<T>(lab + (size_t)(38 + part) * kCells + 4 * i, v[c][0], v[c][1], v[c][2], v[c][3], m[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kRasterChunks; c++) {
            int i = tid + c * kRasterThreads;
            if (i < kCellVec)
                store4<T>(lab + (size_t)56 * kCells + 4 * i, 1.f - bk[c][0], 1.f - bk[c][1], 1.f - bk[c][2],
                          1.f - bk[c][3], m[c]);
        }
        return;
    }

    // ---- H4: part-affinity fields, limbs [k0,k1) ----
    const int k0 = (group == 1) ? 0 : 10, k1 = (group == 1) ? 10 : kLimbs;
    const double thre = a.thre;
    // rows covered by this warp's pixels (for a warp-uniform reject of far-away limbs)
    for (int k = k0; k < k1; k++) {
        const int fr = c_limb_from[k], to = c_limb_to[k];
        for (int p = tid; p < P; p += kRasterThreads) {
            const double *jf = s_j + (p * kParts + fr) * 3, *jt = s_j + (p * kParts + to) * 3;
            LimbRec r;
            r.minx = r.maxx = r.miny = r.maxy = 0;
            r.x1 = jf[0]; r.y1 = jf[1];
            double x2 = jt[0], y2 = jt[1];
            r.xD = __dsub_rn(x2, r.x1); r.yD = __dsub_rn(y2, r.y1);
            // distances(): sqrt(xD**2 + yD**2); put_vector_maps: sqrt(dx*dx + dy*dy) -- same value
            r.norm2 = __dsqrt_rn(__dadd_rn(__dmul_rn(r.xD, r.xD), __dmul_rn(r.yD, r.yD)));
            r.ux = r.uy = 0.f;
            if (jf[2] < 2.0 && jt[2] < 2.0) {
                if (r.norm2 == 0.0) {
                    atomicOr(a.status + b, RMPE_ST_ZERO_LIMB);
                } else {
                    r.ux = (float)__ddiv_rn(r.xD, r.norm2);
                    r.uy = (float)__ddiv_rn(r.yD, r.norm2);
                    double mnx = r.x1 < x2 ? r.x1 : x2, mxx = r.x1 < x2 ? x2 : r.x1;
                    double mny = r.y1 < y2 ? r.y1 : y2, mxy = r.y1 < y2 ? y2 : r.y1;
                    int a0 = py_round(__ddiv_rn(__dsub_rn(mnx, thre), 8.0));
                    int b0 = py_round(__ddiv_rn(__dsub_rn(mny, thre), 8.0));
                    int a1 = py_round(__ddiv_rn(__dadd_rn(mxx, thre), 8.0));
                    int b1 = py_round(__ddiv_rn(__dadd_rn(mxy, thre), 8.0));
                    if (a1 >= 0 && b1 >= 0) {
                        r.minx = max(a0, 0); r.miny = max(b0, 0);
                        r.maxx = min(a1, kGrid); r.maxy = min(b1, kGrid);
                        if (r.maxy <= r.miny) r.maxx = r.minx;  // empty slice
                    }
                }
            }
            s_rec[p] = r;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < kRasterChunks; c++) {
            int i = tid + c * kRasterThreads;
            // warp-uniform row range of this chunk's 32 threads
            int ibase = (tid & ~31) + c * kRasterThreads;
            int wy0 = (4 * ibase) / kGrid, wy1 = min(kCells - 1, 4 * ibase + 127) / kGrid;
            float vx[4] = {0.f, 0.f, 0.f, 0.f}, vy[4] = {0.f, 0.f, 0.f, 0.f};
            int cnt[4] = {0, 0, 0, 0};
            int pix = 4 * i;
            int y[4], x[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { y[q] = (pix + q) / kGrid; x[q] = (pix + q) - y[q] * kGrid; }
            if (ibase < kCellVec) {
                for (int p = 0; p < P; p++) {
                    const LimbRec &r = s_rec[p];
                    if (r.maxx <= r.minx || r.maxy <= wy0 || r.miny > wy1) continue;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (x[q] >= r.minx && x[q] < r.maxx && y[q] >= r.miny && y[q] < r.maxy) {
                            double X = (double)(8 * x[q]), Y = (double)(8 * y[q]);
                            double dd = __dsub_rn(__dmul_rn(r.xD, __dsub_rn(r.y1, Y)),
                                                  __dmul_rn(__dsub_rn(r.x1, X), r.yD));
                            dd = __ddiv_rn(dd, r.norm2);
                            if (fabs(dd) <= thre) { vx[q] = r.ux; vy[q] = r.uy; cnt[q]++; }
                        }
                    }
                }
            }
            if (i < kCellVec) {
                store4<T>(lab + (size_t)(2 * k) * kCells + pix, vx[0], vx[1], vx[2], vx[3], m[c]);
                store4<T>(lab + (size_t)(2 * k + 1) * kCells + pix, vy[0], vy[1], vy[2], vy[3], m[c]);
                if (a.out_count)
                    *reinterpret_cast<int4 *>(a.out_count + ((size_t)b * kLimbs + k) * kCells + pix) =
                        make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
            }
        }
        __syncthreads();
    }
}

// ==========================================================================================
// host entry
// ==========================================================================================
static size_t warp_smem_bytes() { return sizeof(WarpSmem) + 128; }

}  // namespace rmpe

using namespace rmpe;

extern "C" int rmpe_gt_batch(const RmpeGtBatch *b, void *stream_) {
    if (!is_initialised()) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(b != nullptr, "batch descriptor is null");
    RMPE_REQUIRE(b->batch >= 0, "negative batch");
    if (b->batch == 0) return RMPE_OK;
    RMPE_REQUIRE(b->max_persons >= 0 && b->max_persons <= kMaxPersonsGt, "max_persons must be in [0,64]");
    RMPE_REQUIRE(b->status != nullptr, "status is required");
    const bool no_transform = (b->flags & RMPE_GT_NO_TRANSFORM) != 0;
    const bool no_warp = (b->flags & RMPE_GT_NO_WARP) != 0 || no_transform;
    RMPE_REQUIRE(b->out_mask != nullptr, "out_mask is required");
    RMPE_REQUIRE(b->n_persons != nullptr, "n_persons is required");
    RMPE_REQUIRE(b->max_persons == 0 || b->joints != nullptr, "joints is required");
    if (!no_transform) {
        RMPE_REQUIRE(b->M && b->flip && b->src_desc && b->src_mask, "M, flip, src_desc, src_mask are required");
    }
    if (!no_warp) RMPE_REQUIRE(b->src_img && b->out_img, "src_img and out_img are required");
    cudaStream_t st = (cudaStream_t)stream_;
    const DeviceTables &T = tables();
    RMPE_CUDA_TRY(cudaMemsetAsync(b->status, 0, sizeof(int32_t) * b->batch, st));

    if (!no_warp) {
        WarpArgs wa;
        wa.src_img = b->src_img; wa.desc = b->src_desc; wa.M = b->M; wa.out_img = b->out_img;
        wa.status = b->status; wa.tab = T.bicubic_i16; wa.tab_dp4a = T.bicubic_dp4a;
        wa.batch = b->batch; wa.chw = (b->flags & RMPE_GT_IMG_CHW) ? 1 : 0;
        if (b->flags & RMPE_GT_SIMPLE_KERNELS) {
            dim3 grid((kOutW * kOutH + 255) / 256, b->batch);
            ProfScope ps("k_warp_simple", st);
            k_warp_simple<<<grid, 256, 0, st>>>(wa);
        } else {
            static bool attr_set = false;
            size_t smem = warp_smem_bytes();
            if (!attr_set) {
                RMPE_CUDA_TRY(cudaFuncSetAttribute(k_warp_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_set = true;
            }
            int n_items = b->batch * kTilesPerSample;
            int grid = min(n_items, 2 * T.sm_count);
            ProfScope ps("k_warp_tile", st);
            k_warp_tile<<<grid, kWarpThreads, smem, st>>>(wa, n_items);
        }
        count_launch();
    }
    const bool warp_only = (b->flags & RMPE_GT_WARP_ONLY) != 0;
    if (!no_transform && !warp_only) {
        MaskArgs ma;
        ma.src_mask = b->src_mask; ma.desc = b->src_desc; ma.M = b->M; ma.out_mask = b->out_mask;
        ma.tab = T.bicubic_i16; ma.f64 = (b->flags & RMPE_GT_LABELS_F64) ? 1 : 0;
        dim3 grid((kCells + 127) / 128, b->batch);
        ProfScope ps("k_mask46", st);
        k_mask46<<<grid, 128, 0, st>>>(ma);
        count_launch();
    }
    if (b->out_labels && !warp_only) {
        RasterArgs ra;
        ra.joints = b->joints; ra.n_persons = b->n_persons; ra.M = b->M; ra.flip = b->flip;
        ra.mask = b->out_mask; ra.labels = b->out_labels; ra.out_joints = b->out_joints;
        ra.out_count = b->out_count; ra.status = b->status; ra.max_persons = b->max_persons;
        ra.f64 = (b->flags & RMPE_GT_LABELS_F64) ? 1 : 0; ra.no_transform = no_transform ? 1 : 0;
        ra.sigma = 7.0; ra.thre = 8.0;
        dim3 grid(3, b->batch);
        ProfScope ps("k_raster", st);
        if (ra.f64) k_raster<double><<<grid, kRasterThreads, 0, st>>>(ra);
        else k_raster<float><<<grid, kRasterThreads, 0, st>>>(ra);
        count_launch();
    }
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}
