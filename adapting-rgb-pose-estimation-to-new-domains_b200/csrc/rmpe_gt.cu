// Target generation on sm_100a: affine warp of image + miss-mask (bit-exact OpenCV fixed-point
// bicubic), fused 368->46 mask resize, keypoint transform and the 57-plane heat/PAF rasteriser.
//
// Reference path replaced (paths relative to the reference root):
//   py_rmpe_server/py_rmpe_transformer.py:83-114  Transformer.transform      -> k_warp_fused (k_warp_simple + k_mask46 with
//                                                                               RMPE_GT_SIMPLE_KERNELS), joints in the rasterisers
//   py_rmpe_server/py_rmpe_heatmapper.py:32-138    Heatmapper.create_heatmaps -> k_raster_small / k_raster_blocks / k_raster_roles
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <unordered_map>

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
// Debug / fallback-for-huge-footprints variant; also the A/B check of k_warp_fused on the GPU.
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
// k_warp_fused: warpAffine(image, border 127) + warpAffine(mask, border 255) + the 368->46 mask
// resize in ONE pass over the destination.
//
// A persistent CTA (one per SM) holds OpenCV's 32 KB int16 weight table once and runs NG independent
// groups of 128 threads; each group draws 32x32 destination tiles from a global counter (three
// passes ahead), warp w owning the cell row (destination rows 8w..8w+7) so every warp does
// identical work.  Per tile the group
//   1. (two tiles ahead, overlapped with the taps) builds the fixed-point row/column terms of
//      cv::WarpAffineInvoker in f64; (one tile ahead) the footprint bounding box,
//   2. stages the UNCLIPPED source footprint of the tile into shared memory as one 32-bit word
//      per source pixel (B,G,R,mask), out-of-image pixels already replaced by the border
//      constants (127,127,127,255) -- the 16-tap loop needs no bounds test at all; 8-byte aligned
//      sources are read with 64-bit loads, 8 pixels per thread (4-byte aligned: 32-bit, 4 pixels),
//   3. evaluates the 16 taps of the channels with dp2a.lo / dp2a.hi: a pair of OpenCV's int16 weights times the low /
//      high two pixel bytes of a word (exact: identical to OpenCV's int32 accumulation): 4 LDS.32 + 4 PRMT + 8 IDP.2A
//      per tap row on the four rows that feed the mask, 4 + 4 + 6 on the others,
//   4. reduces the warped mask to the 46x46 grid in registers/shuffles (cells are 8x8 destination
//      pixels of which rows/cols 2..5 feed cv2.resize),
//   5. packs B,G,R of 32 neighbouring pixels into 24 words with one shuffle and stores the row
//      as one coalesced 96-byte segment (planar rows for CHW).
// The kernel is bound by the L1/shared-memory data pipe (wavefronts of the tap and weight gathers),
// so the layouts are chosen for bank behaviour: the footprint pitch is 0 (mod 32) words (the bank
// of a tap is its column, whatever its row) and the weight table is stored transposed (the 16-byte
// bank group of an entry is ay & 7); DESIGN.md 4.1 has the measured wavefront budget.
// ==========================================================================================
constexpr int kTile = 32;
constexpr int kTilesX = (kOutW + kTile - 1) / kTile;  // 12
constexpr int kTilesPerSample = kTilesX * kTilesX;    // 144
constexpr int kTabBytes = 32 * 32 * 8 * 4;            // 1024 phases x 16 int16 weights
constexpr int kGroupThreads = 128;
constexpr int kGeoInts = 4 * kTile;                   // {X0,Y0}[32] then {ad,bd}[32], interleaved
constexpr int kGroupFixedBytes = 4 * kGeoInts * 4;    // geometry ring: tile k, k+1 (bounding box), k+2 (being built); 4th slot: tile queue
constexpr unsigned kBorderWord = 0xFF7F7F7Fu;         // (B,G,R,mask) = (127,127,127,255)

struct FusedArgs {
    const uint8_t *src_img;
    const uint8_t *src_mask;
    const RmpeSrcDesc *desc;
    const double *M;
    uint8_t *out_img;
    void *out_mask;
    int32_t *status;
    const int16_t *tab;
    const uint32_t *tab_dp4a;
    int batch;
    int chw;
    int mask_f64;
    int foot_cap;   // footprint capacity of one group, in pixels (32-bit words)
    int32_t *counter;   // next tile to hand out (zero at launch): tile groups take tiles as they finish theirs
    int prefetch;       // pull the next tile's source lines into L2 during the taps
};

__device__ __forceinline__ void group_bar(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kGroupThreads) : "memory");
}

// saturate four int32 to u8 and pack them as B | G<<8 | R<<16 | M<<24: two I2IP (cvt.pack.sat.u8.s32) instead of eight
// min/max and three shift-ors
__device__ __forceinline__ unsigned pack_sat_u8x4(int vB, int vG, int vR, int vM) {
    unsigned t, w;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(vM), "r"(vR), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(vG), "r"(vB), "r"(t));
    return w;
}

// 16 taps x {B,G,R[,mask]} from a staged (B,G,R,mask) footprint.  p -> tap (0,0).
template <bool kMask>
__device__ __forceinline__ void bicubic_bgrm(const uint32_t *__restrict__ p, int fpitch,
                                             const uint32_t *__restrict__ wt, int &oB, int &oG, int &oR, int &oM) {
    // the 16 int16 weights of this phase, [ky][kx] as in OpenCV's table: tap rows 0-1 in one table of 16-byte entries,
    // rows 2-3 in a second one (a quarter-warp's LDS.128 then spreads over all 8 bank groups)
    const uint4 wa = *reinterpret_cast<const uint4 *>(wt);
    const uint4 wb = *reinterpret_cast<const uint4 *>(wt + kTabBytes / 8);
    const unsigned w01[4] = {wa.x, wa.z, wb.x, wb.z};    // (w[ky][0], w[ky][1]) as s16 x2
    const unsigned w23[4] = {wa.y, wa.w, wb.y, wb.w};    // (w[ky][2], w[ky][3])
    int vB = 16384, vG = 16384, vR = 16384, vM = 16384;  // rounding term of the >>15
#pragma unroll
    for (int ky = 0; ky < 4; ky++) {
        const unsigned p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3];
        p += fpitch;
        // half a 4x4 byte transpose is enough for dp2a: it multiplies the two s16 weights of a pair with the low or the
        // high two bytes of its second operand
        const unsigned t0 = __byte_perm(p0, p1, 0x5140);   // B0 B1 G0 G1
        const unsigned t1 = __byte_perm(p2, p3, 0x5140);   // B2 B3 G2 G3
        const unsigned t2 = __byte_perm(p0, p1, 0x7362);   // R0 R1 M0 M1
        const unsigned t3 = __byte_perm(p2, p3, 0x7362);   // R2 R3 M2 M3
        vB = dp2a_lo_su(w01[ky], t0, vB); vB = dp2a_lo_su(w23[ky], t1, vB);
        vG = dp2a_hi_su(w01[ky], t0, vG); vG = dp2a_hi_su(w23[ky], t1, vG);
        vR = dp2a_lo_su(w01[ky], t2, vR); vR = dp2a_lo_su(w23[ky], t3, vR);
        if (kMask) { vM = dp2a_hi_su(w01[ky], t2, vM); vM = dp2a_hi_su(w23[ky], t3, vM); }
    }
    // unclamped; the caller saturates and packs (pack_sat_u8x4)
    oB = vB >> 15; oG = vG >> 15; oR = vR >> 15;
    oM = kMask ? vM >> 15 : 0;
}

// (B,G,R,mask) word of one source pixel with cv2's per-tap constant border
__device__ __forceinline__ unsigned bgrm_pixel(const uint8_t *__restrict__ img, const uint8_t *__restrict__ msk,
                                               int H, int W, int ipitch, int mpitch, int yy, int xx) {
    if ((unsigned)yy >= (unsigned)H || (unsigned)xx >= (unsigned)W) return kBorderWord;
    const uint8_t *ip = img + (size_t)yy * ipitch + 3 * xx;
    return (unsigned)__ldg(ip) | ((unsigned)__ldg(ip + 1) << 8) | ((unsigned)__ldg(ip + 2) << 16) |
           ((unsigned)__ldg(msk + (size_t)yy * mpitch + xx) << 24);
}

// 4 pixels (12 image bytes in q0..q2, 4 mask bytes in mq) -> 4 (B,G,R,mask) words
__device__ __forceinline__ uint4 bgrm_pack4(unsigned q0, unsigned q1, unsigned q2, unsigned mq) {
    uint4 w;
    w.x = __byte_perm(q0, mq, 0x4210);
    w.y = __byte_perm(__byte_perm(q0, q1, 0x0543), mq, 0x5210);
    w.z = __byte_perm(__byte_perm(q1, q2, 0x0432), mq, 0x6210);
    w.w = __byte_perm(q2, mq, 0x7321);
    return w;
}

// 4 pixels starting at (yy, xx) of an arbitrary-pitch source
__device__ __forceinline__ uint4 bgrm_load4(const uint8_t *__restrict__ img, const uint8_t *__restrict__ msk,
                                            int H, int W, int ipitch, int mpitch, int yy, int xx) {
    if ((unsigned)yy >= (unsigned)H) return make_uint4(kBorderWord, kBorderWord, kBorderWord, kBorderWord);
    if (xx >= 0 && xx + 3 < W) {
        // 12 image bytes + 4 mask bytes through aligned 32-bit loads
        const size_t ia = (size_t)(img + (size_t)yy * ipitch + 3 * xx);
        const uint32_t *iw = reinterpret_cast<const uint32_t *>(ia & ~(size_t)3);
        const unsigned ish = (unsigned)(ia & 3) * 8;
        const unsigned u0 = __ldg(iw), u1 = __ldg(iw + 1), u2 = __ldg(iw + 2);
        const unsigned u3 = ish ? __ldg(iw + 3) : 0u;
        const size_t ma = (size_t)(msk + (size_t)yy * mpitch + xx);
        const uint32_t *mwp = reinterpret_cast<const uint32_t *>(ma & ~(size_t)3);
        const unsigned msh = (unsigned)(ma & 3) * 8;
        const unsigned v0 = __ldg(mwp);
        const unsigned v1 = msh ? __ldg(mwp + 1) : 0u;
        return bgrm_pack4(__funnelshift_r(u0, u1, ish), __funnelshift_r(u1, u2, ish), __funnelshift_r(u2, u3, ish),
                          __funnelshift_r(v0, v1, msh));
    }
    uint4 w;
    w.x = bgrm_pixel(img, msk, H, W, ipitch, mpitch, yy, xx);
    w.y = bgrm_pixel(img, msk, H, W, ipitch, mpitch, yy, xx + 1);
    w.z = bgrm_pixel(img, msk, H, W, ipitch, mpitch, yy, xx + 2);
    w.w = bgrm_pixel(img, msk, H, W, ipitch, mpitch, yy, xx + 3);
    return w;
}

// tiles whose footprint does not fit the staging buffer (heavy down-scaling) take their taps
// straight from global memory
__device__ __noinline__ unsigned generic_bgrm(const uint8_t *__restrict__ img, const uint8_t *__restrict__ msk,
                                              int H, int W, int ipitch, int mpitch, int X, int Y,
                                              const int16_t *__restrict__ tab, bool want_mask) {
    int o[3], m1[1] = {255};
    warp_pixel_generic<3>(img, H, W, ipitch, X, Y, tab, 127, o);
    if (want_mask) warp_pixel_generic<1>(msk, H, W, mpitch, X, Y, tab, 255, m1);
    return (unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16) | ((unsigned)m1[0] << 24);
}

struct TileCtx {
    const uint32_t *foot;
    const uint32_t *tab;
    const int2 *gXY;                   // {X0[y], Y0[y]} of the tile's rows
    int adx, bdx, fpitch, bx0, by0;   // fpitch in 32-bit words
    const uint8_t *img, *msk;
    int H, W, ipitch, mpitch;
    const int16_t *tab16;
};

// one destination row of the warp: returns packed B | G<<8 | R<<16 | mask<<24
// kMode: 0 = staged (B,G,R,mask) words, 1 = all border, 2 = generic (taps from global memory)
template <int kMode, bool kMask>
__device__ __forceinline__ unsigned fused_row(const TileCtx &c, int ly) {
    if (kMode == 1) return kBorderWord;
    const int2 xy0 = c.gXY[ly];        // one 64-bit broadcast load per row
    const int X = (xy0.x + c.adx) >> 5;
    const int Y = (xy0.y + c.bdx) >> 5;
    const uint32_t *wt = c.tab + (((X & 31) * 32 + (Y & 31)) << 2);      // slot = ax * 32 + ay (see the table fill)
    if (kMode == 0) {
        int oB, oG, oR, oM;
        const uint32_t *p = c.foot + ((Y >> 5) - 1 - c.by0) * c.fpitch + ((X >> 5) - 1 - c.bx0);
        bicubic_bgrm<kMask>(p, c.fpitch, wt, oB, oG, oR, oM);
        return pack_sat_u8x4(oB, oG, oR, oM);
    }
    return generic_bgrm(c.img, c.msk, c.H, c.W, c.ipitch, c.mpitch, X, Y, c.tab16, kMask);
}

// per-warp constants of the row stores: HWC packs B,G,R of 32 neighbouring pixels into 24 words with one shuffle
struct RowStore {
    uint8_t *op;        // HWC: this lane's word of the first row of the warp; CHW: pixel `lane` of plane 0
    unsigned pk_sel;
    bool on;
};

// store destination row j (0..7) of the warp, one pixel per lane.  The image layout is a template parameter of the
// kernel: a run-time test here put a branch, a divergence check around the shuffle and a re-materialised output pointer
// into every row (12 instructions where 4 do).
template <bool kChw>
__device__ __forceinline__ void store_row(const RowStore &rs, int j, unsigned bgr) {
    if (!kChw) {
        // 24 words of the 96-byte row.  Pixel l starts at byte 3l, so lane l = 4k + m (m = 0, 1, 2) holds, together with
        // its right neighbour's pixel, exactly the aligned word 3k + m (bytes 12k + 4m ..): ONE shuffle per row.
        const unsigned nb = __shfl_down_sync(0xffffffffu, bgr, 1);
        if (rs.on) *reinterpret_cast<uint32_t *>(rs.op + j * (kOutW * 3)) = __byte_perm(bgr, nb, rs.pk_sel);
    } else if (rs.on) {
        uint8_t *o = rs.op + j * kOutW;
        o[0] = (uint8_t)bgr;
        o[kOutW * kOutH] = (uint8_t)(bgr >> 8);
        o[2 * kOutW * kOutH] = (uint8_t)(bgr >> 16);
    }
}

// the 8 destination rows (one cell row) of a warp: taps, stores, and the 8x8 -> 1 mask reduction of cv2.resize
// returns the cell's float32 accumulator (valid in lane 3 of the cell's 8-lane group)
template <int kMode, bool kWantMask, bool kChw>
__device__ __forceinline__ float cell_rows(const TileCtx &c, const RowStore &rs, int r0, float b0, float b1) {
#pragma unroll
    for (int j = 0; j < 2; j++) store_row<kChw>(rs, j, fused_row<kMode, false>(c, r0 + j));
    // rows 2..5 of the cell feed cv2.resize's 8:1 reduction of the mask: their mask bytes are collected in one word
    // (byte j-2 = row j) and meet in lane 3 of the cell with THREE shuffles per cell row (one word from each of lanes
    // 2, 4, 5) instead of two per mask row
    unsigned mrows = 0;
#pragma unroll
    for (int j = 5; j >= 2; j--) {
        const unsigned v4 = fused_row<kMode, kWantMask>(c, r0 + j);
        store_row<kChw>(rs, j, v4);
        if (kWantMask) mrows = (mrows << 8) | (v4 >> 24);
    }
    float macc = 0.f;
    if (kWantMask) {
        const unsigned m2 = __shfl_up_sync(0xffffffffu, mrows, 1);
        const unsigned m4 = __shfl_down_sync(0xffffffffu, mrows, 1);
        const unsigned m5 = __shfl_down_sync(0xffffffffu, mrows, 2);
        // horizontal pass: exact int32 sum of (-192,1216,1216,-192) x cols 8c+2..8c+5; vertical pass: float32 FMA chain
        // S5*b0 -> S4*b1 -> S3*b1 -> S2*b0 (cv2.resize's order).  Meaningful in lane 3 of a cell only.
        int S[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int inner = (int)__byte_perm(mrows, 0, 0x4440 + k) + (int)__byte_perm(m4, 0, 0x4440 + k);
            const int outer = (int)__byte_perm(m2, 0, 0x4440 + k) + (int)__byte_perm(m5, 0, 0x4440 + k);
            S[k] = 1216 * inner - 192 * outer;
        }
        macc = (float)S[3] * b0;
        macc = fmaf((float)S[2], b1, macc);
        macc = fmaf((float)S[1], b1, macc);
        macc = fmaf((float)S[0], b0, macc);
    }
#pragma unroll
    for (int j = 6; j < 8; j++) store_row<kChw>(rs, j, fused_row<kMode, false>(c, r0 + j));
    return macc;
}

template <int NG, bool kWantMask, bool kChw>
__global__ void __launch_bounds__(NG *kGroupThreads, 1) k_warp_fused(FusedArgs a, int n_items) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t *s_tab = reinterpret_cast<uint32_t *>(smem_raw);
    // the rasteriser that follows on the stream may be scheduled as soon as an SM is free (it waits for this grid's
    // completion itself before it touches the mask)
    pdl_trigger();
    const int group = threadIdx.x / kGroupThreads;
    const int t = threadIdx.x - group * kGroupThreads;
    const int warp = t >> 5, lane = t & 31;
    uint8_t *gbase = smem_raw + kTabBytes + (size_t)group * ((size_t)a.foot_cap * 4 + kGroupFixedBytes);
    uint32_t *foot = reinterpret_cast<uint32_t *>(gbase);
    int *geo = reinterpret_cast<int *>(gbase + (size_t)a.foot_cap * 4);

    {   // weight table -> shared memory, once per persistent CTA, as two tables of 16-byte entries.
        // shared-memory slot of phase (ay, ax) = ax * 32 + ay: the 16-byte bank group of an entry is then ay & 7.  Along a
        // destination row ay moves by 32 sin(theta) per pixel and ax by 32 (cos(theta) - 1): for the rotations of the
        // augmentation (|theta| <= 40 deg) ax often stays put over a quarter-warp while ay changes, which with the
        // [ay][ax] order put 8 different entries into one bank group (measured 2.75 wavefronts per quarter-warp, 2.2 now)
        // dp2a operands: OpenCV's int16 table as it is, tap rows 0-1 in the first table, tap rows 2-3 in the second
        uint4 *dst = reinterpret_cast<uint4 *>(s_tab);
        const uint4 *src = reinterpret_cast<const uint4 *>(a.tab);
        for (int e = threadIdx.x; e < 1024; e += NG * kGroupThreads) {
            const int slot = ((e & 31) << 5) | (e >> 5);
            dst[slot] = __ldg(src + 2 * e);
            dst[1024 + slot] = __ldg(src + 2 * e + 1);
        }
    }

    const int l7 = lane & 7;
    const int pk_word = 3 * (lane >> 2) + (lane & 3);         // HWC row packing (store_row): lanes 4k+3 store nothing
    const unsigned pk_sel = (lane & 3) == 0 ? 0x4210u : ((lane & 3) == 1 ? 0x5421u : 0x6542u);
    const float sc = 1.0f / 4194304.0f;                       // 2^-22
    const float b0 = -192.0f * sc, b1 = 1216.0f * sc;

    // geometry of one tile -> geo[buf]  (threads 0..63 of the group)
    auto geometry = [&](int item, int buf) {
        const int sample = item / kTilesPerSample;
        const int tt = item - sample * kTilesPerSample;
        const int ty = tt / kTilesX, tx = tt - ty * kTilesX;
        int *g = geo + buf * kGeoInts;
        double iM[6];
        const bool ok = invert_affine(a.M + 6 * sample, iM);
        if (t < kTile) {
            g[2 * t] = warp_row_term(iM[1], iM[2], ty * kTile + t);
            g[2 * t + 1] = warp_row_term(iM[4], iM[5], ty * kTile + t);
        } else {
            g[2 * kTile + 2 * (t - kTile)] = warp_col_term(iM[0], tx * kTile + t - kTile);
            g[2 * kTile + 2 * (t - kTile) + 1] = warp_col_term(iM[3], tx * kTile + t - kTile);
        }
        if (t == 0 && !ok && tt == 0) atomicOr(a.status + sample, RMPE_ST_SINGULAR);
    };

    // footprint bounding box of a tile from its 4 corners: X(x,y) = (X0[y]+ad[x])>>5 is monotone in x and y
    auto bbox = [&](const int *g, int tw, int th, int &mnx, int &mxx, int &mny, int &mxy) -> bool {
        mnx = INT_MAX; mxx = INT_MIN; mny = INT_MAX; mxy = INT_MIN;
        bool sane = true;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int cx = (c & 1) ? tw - 1 : 0, cy = (c & 2) ? th - 1 : 0;
            const int qx = (g[2 * cy] + g[2 * kTile + 2 * cx]) >> 10, qy = (g[2 * cy + 1] + g[2 * kTile + 2 * cx + 1]) >> 10;
            sane = sane && qx > -30000 && qx < 30000 && qy > -30000 && qy < 30000;  // no short saturation
            mnx = min(mnx, qx - 1); mxx = max(mxx, qx + 2);
            mny = min(mny, qy - 1); mxy = max(mxy, qy + 2);
        }
        return sane;
    };

    // Tiles are handed out dynamically (border tiles are cheap, strongly rotated ones expensive): q[k & 3] is the tile of
    // pass k; thread 0 of the group draws the tile of pass k+3 while pass k runs, so geometry (two passes ahead) and the
    // L2 prefetch (one ahead) still know their tiles, and the atomic's latency is never waited for.
    int *q = geo + 3 * kGeoInts;
    if (t == 0) {
        q[0] = atomicAdd(a.counter, 1); q[1] = atomicAdd(a.counter, 1); q[2] = atomicAdd(a.counter, 1);
    }
    group_bar(group);
    if (t < 2 * kTile) {
        if (q[0] < n_items) geometry(q[0], 0);
        if (q[1] < n_items) geometry(q[1], 1);
    }
    __syncthreads();   // weight table + first two geometries

    // bounding box of the current tile: computed once (for the L2 prefetch, one tile ahead) and carried over
    int mnx = 0, mxx = 0, mny = 0, mxy = 0;
    bool sane = false;
    if (q[0] < n_items) {
        const int tt = q[0] % kTilesPerSample;
        const int ty = tt / kTilesX, tx = tt - ty * kTilesX;
        sane = bbox(geo, min(kTile, kOutW - tx * kTile), min(kTile, kOutH - ty * kTile), mnx, mxx, mny, mxy);
    }

    for (int k = 0;; k++) {
        const int item = q[k & 3];
        if (item >= n_items) break;
        const int item1 = q[(k + 1) & 3], item2 = q[(k + 2) & 3];
        const int sample = item / kTilesPerSample;
        const int tt = item - sample * kTilesPerSample;
        const int ty = tt / kTilesX, tx = tt - ty * kTilesX;
        const int x0 = tx * kTile, y0 = ty * kTile;
        const int tw = min(kTile, kOutW - x0), th = min(kTile, kOutH - y0);
        const RmpeSrcDesc d = a.desc[sample];
        const uint8_t *img = a.src_img + d.img_offset;
        const uint8_t *msk = a.src_mask + d.mask_offset;
        const int *gcur = geo + (k % 3) * kGeoInts;      // {X0,Y0}[32] then {ad,bd}[32], interleaved

        // 4-byte aligned sources are read with plain 32-bit loads; the footprint then starts on a multiple of 4
        const bool al4 = ((((size_t)img | (size_t)msk) & 3) == 0) && (((d.img_pitch | d.mask_pitch) & 3) == 0);
        // 8-byte aligned sources: 64-bit loads, 8 pixels per thread, footprint on a multiple of 8
        const bool al8 = ((((size_t)img | (size_t)msk) & 7) == 0) && (((d.img_pitch | d.mask_pitch) & 7) == 0);
        const int bx0 = al8 ? (mnx & ~7) : (al4 ? (mnx & ~3) : mnx), by0 = mny;
        const int fw = mxx - bx0 + 1, fh = mxy - mny + 1;
        const int fwa = al8 ? ((fw + 7) & ~7) : ((fw + 3) & ~3);
        int fpitch = (fwa + 31) & ~31;     // = 0 (mod 32): the bank of a tap is its column, whatever its row
        if ((long long)fpitch * fh > a.foot_cap) fpitch = fwa;
        const bool outside = sane && (mxx < 0 || mnx >= d.width || mxy < 0 || mny >= d.height);
        const bool staged = sane && !outside && (long long)fpitch * fh <= a.foot_cap;

        // ---- stage the footprint as (B,G,R,mask) words: one thread = 4 pixels = one 16-byte store ----
        if (staged && al8) {
            // 8 threads x 8 pixels per footprint row (three 64-bit image loads + one 64-bit mask load, two 16-byte stores);
            // two rows (r, r+16) in flight per thread and pass
            for (int r = t >> 3; r < fh; r += 2 * (kGroupThreads / 8))
            for (int c = (t & 7) << 3; c < fwa; c += 64) {
                const int xx = bx0 + c;
                const int r1 = r + kGroupThreads / 8;
                const int ya = by0 + r, yb = by0 + r1;
                const bool xin = xx >= 0 && xx + 7 < d.width;
                const bool fa = xin && (unsigned)ya < (unsigned)d.height;
                const bool fb = xin && (unsigned)yb < (unsigned)d.height && r1 < fh;
                uint2 qa0, qa1, qa2, qam, qb0, qb1, qb2, qbm;
                qa0 = qa1 = qa2 = qam = qb0 = qb1 = qb2 = qbm = make_uint2(0u, 0u);
                if (fa) {
                    const uint2 *ip = reinterpret_cast<const uint2 *>(img + (size_t)ya * d.img_pitch + 3 * xx);
                    qa0 = __ldg(ip); qa1 = __ldg(ip + 1); qa2 = __ldg(ip + 2);
                    qam = __ldg(reinterpret_cast<const uint2 *>(msk + (size_t)ya * d.mask_pitch + xx));
                }
                if (fb) {
                    const uint2 *ip = reinterpret_cast<const uint2 *>(img + (size_t)yb * d.img_pitch + 3 * xx);
                    qb0 = __ldg(ip); qb1 = __ldg(ip + 1); qb2 = __ldg(ip + 2);
                    qbm = __ldg(reinterpret_cast<const uint2 *>(msk + (size_t)yb * d.mask_pitch + xx));
                }
                uint4 *dst = reinterpret_cast<uint4 *>(foot + r * fpitch + c);
                dst[0] = fa ? bgrm_pack4(qa0.x, qa0.y, qa1.x, qam.x)
                            : bgrm_load4(img, msk, d.height, d.width, d.img_pitch, d.mask_pitch, ya, xx);
                dst[1] = fa ? bgrm_pack4(qa1.y, qa2.x, qa2.y, qam.y)
                            : bgrm_load4(img, msk, d.height, d.width, d.img_pitch, d.mask_pitch, ya, xx + 4);
                if (r1 < fh) {
                    dst = reinterpret_cast<uint4 *>(foot + r1 * fpitch + c);
                    dst[0] = fb ? bgrm_pack4(qb0.x, qb0.y, qb1.x, qbm.x)
                                : bgrm_load4(img, msk, d.height, d.width, d.img_pitch, d.mask_pitch, yb, xx);
                    dst[1] = fb ? bgrm_pack4(qb1.y, qb2.x, qb2.y, qbm.y)
                                : bgrm_load4(img, msk, d.height, d.width, d.img_pitch, d.mask_pitch, yb, xx + 4);
                }
            }
        } else if (staged) {
            // 16 threads x 4 pixels per footprint row; two rows (r, r+8) in flight per thread and pass
            for (int r = t >> 4; r < fh; r += 2 * (kGroupThreads / 16))
            for (int c = (t & 15) << 2; c < fwa; c += 64) {
                const int xx = bx0 + c;
                const int r1 = r + kGroupThreads / 16;
                const int ya = by0 + r, yb = by0 + r1;
                const bool xin = al4 && xx >= 0 && xx + 3 < d.width;
                const bool fa = xin && (unsigned)ya < (unsigned)d.height;
                const bool fb = xin && (unsigned)yb < (unsigned)d.height && r1 < fh;
                unsigned qa0 = 0, qa1 = 0, qa2 = 0, qam = 0, qb0 = 0, qb1 = 0, qb2 = 0, qbm = 0;
                if (fa) {
                    const uint32_t *ip = reinterpret_cast<const uint32_t *>(img + (size_t)ya * d.img_pitch + 3 * xx);
                    qa0 = __ldg(ip); qa1 = __ldg(ip + 1); qa2 = __ldg(ip + 2);
                    qam = __ldg(reinterpret_cast<const uint32_t *>(msk + (size_t)ya * d.mask_pitch + xx));
                }
                if (fb) {
                    const uint32_t *ip = reinterpret_cast<const uint32_t *>(img + (size_t)yb * d.img_pitch + 3 * xx);
                    qb0 = __ldg(ip); qb1 = __ldg(ip + 1); qb2 = __ldg(ip + 2);
                    qbm = __ldg(reinterpret_cast<const uint32_t *>(msk + (size_t)yb * d.mask_pitch + xx));
                }
                *reinterpret_cast<uint4 *>(foot + r * fpitch + c) =
                    fa ? bgrm_pack4(qa0, qa1, qa2, qam)
                       : bgrm_load4(img, msk, d.height, d.width, d.img_pitch, d.mask_pitch, ya, xx);
                if (r1 < fh)
                    *reinterpret_cast<uint4 *>(foot + r1 * fpitch + c) =
                        fb ? bgrm_pack4(qb0, qb1, qb2, qbm)
                           : bgrm_load4(img, msk, d.height, d.width, d.img_pitch, d.mask_pitch, yb, xx);
            }
        }
        group_bar(group);     // footprint visible

        // two tiles ahead: geometry; one tile ahead: bounding box (kept for the next pass) and its lines pulled into L2
        // -- all of it overlaps this tile's taps
        if (t == 0) q[(k + 3) & 3] = atomicAdd(a.counter, 1);    // slot of pass k-1: free since its closing barrier
        if (item2 < n_items && t < 2 * kTile) geometry(item2, (k + 2) % 3);
        int nmnx = 0, nmxx = 0, nmny = 0, nmxy = 0;
        bool nsane = false;
        if (item1 < n_items) {
            const int ni = item1;
            const int ns = ni / kTilesPerSample, ntt = ni - ns * kTilesPerSample;
            const int nty = ntt / kTilesX, ntx = ntt - nty * kTilesX;
            nsane = bbox(geo + ((k + 1) % 3) * kGeoInts, min(kTile, kOutW - ntx * kTile), min(kTile, kOutH - nty * kTile),
                         nmnx, nmxx, nmny, nmxy);
            if (nsane) {
                const RmpeSrcDesc nd = a.desc[ns];
                const int pnx = max(nmnx, 0), pxx = min(nmxx, nd.width - 1);
                const int pny = max(nmny, 0), pxy = min(nmxy, nd.height - 1);
                const int rows = pxy - pny + 1;
                if (a.prefetch && pxx >= pnx && rows > 0 && rows <= 256) {
                    // per row: up to 3 image lines + 1 mask line of 128 bytes (wider footprints: first 384 bytes)
                    for (int i = t; i < rows * 4; i += kGroupThreads) {
                        const int row = pny + (i >> 2), seg = i & 3;
                        const uint8_t *pa;
                        bool on = true;
                        if (seg < 3) {
                            const uint8_t *rb = a.src_img + nd.img_offset + (size_t)row * nd.img_pitch;
                            pa = rb + 3 * pnx + seg * 128;
                            on = pa <= rb + 3 * pxx + 2 + 127;
                            if (pa > rb + 3 * pxx + 2) pa = rb + 3 * pxx + 2;
                        } else {
                            pa = a.src_mask + nd.mask_offset + (size_t)row * nd.mask_pitch + pnx;
                        }
                        if (on) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
                    }
                }
            }
        }

        // ---- taps, mask reduction, stores: warp = cell row ----
        if (8 * warp < th) {
            TileCtx c;
            const int gl = min(lane, tw - 1);     // lanes beyond the tile edge repeat its last pixel (never stored)
            c.foot = foot; c.tab = s_tab; c.gXY = reinterpret_cast<const int2 *>(gcur);
            const int2 abd = reinterpret_cast<const int2 *>(gcur + 2 * kTile)[gl];
            c.adx = abd.x; c.bdx = abd.y; c.fpitch = fpitch; c.bx0 = bx0; c.by0 = by0;
            c.img = img; c.msk = msk; c.H = d.height; c.W = d.width; c.ipitch = d.img_pitch; c.mpitch = d.mask_pitch;
            c.tab16 = a.tab;
            const bool lane_on = lane < tw;
            const int r0 = 8 * warp;
            uint8_t *out_sample = a.out_img + (size_t)sample * (3 * kOutW * kOutH);
            RowStore rs;
            rs.pk_sel = pk_sel;
            rs.op = kChw ? out_sample + (size_t)(y0 + r0) * kOutW + x0 + lane
                          : out_sample + ((size_t)(y0 + r0) * kOutW + x0) * 3 + 4 * pk_word;
            rs.on = kChw ? lane_on : ((lane & 3) != 3 && 4 * pk_word < 3 * tw);
            float macc;
            if (staged) macc = cell_rows<0, kWantMask, kChw>(c, rs, r0, b0, b1);
            else if (outside) macc = cell_rows<1, kWantMask, kChw>(c, rs, r0, b0, b1);
            else macc = cell_rows<2, kWantMask, kChw>(c, rs, r0, b0, b1);
            if (kWantMask && l7 == 3 && lane_on) {
                // rint, saturate; then /255.  (py_rmpe_transformer.py:92,95)
                const int iv = min(255, max(0, __float2int_rn(macc)));
                const double m = __ddiv_rn((double)iv, 255.0);
                const size_t o = (size_t)sample * kCells + (size_t)((y0 >> 3) + warp) * kGrid + (x0 >> 3) + (lane >> 3);
                if (a.mask_f64) reinterpret_cast<double *>(a.out_mask)[o] = m;
                else reinterpret_cast<float *>(a.out_mask)[o] = (float)m;
            }
        }
        mnx = nmnx; mxx = nmxx; mny = nmny; mxy = nmxy; sane = nsane;
        group_bar(group);     // footprint free, next geometry visible
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
// Rasteriser (H1-H5 + T4): (57,46,46) labels of one sample from its joints and its 46x46 mask.
//   k_raster_small  <= 4 persons (COCO crops), planar labels: one CTA per sample, every table built once, a thread
//                   owns one float4 run of pixels for all 57 planes, no barrier in the plane loop.
//   k_raster_roles  any person count (crowded scenes) and / or the Keras-ready NHWC tensors: two CTAs per sample,
//                   one per ROLE (18 Gaussian maps + background | 19 limbs), same barrier-free plane loops.
// Joints are transformed (T4) into shared memory by every CTA; the first CTA of a sample writes them out.
// ==========================================================================================
struct RasterArgs {
    const double *joints;
    const int32_t *n_persons;
    const double *M;
    const uint8_t *flip;
    const void *mask;     // [B][46][46] f32/f64
    void *labels;         // planar [B][57][46][46] or null
    double *out_joints;
    int32_t *out_count;
    int32_t *status;
    int max_persons;
    int f64;
    int no_transform;
    double sigma, thre;
    // k_raster_roles only
    void *y1, *y2, *x1, *x2;   // NHWC tensors (training/ds_generators.py:52-63), each may be null
    int part_batch;            // parts whose exp tables are resident at once (18: one pass)
    int band_px;               // NHWC: pixels staged per pass (544, 272 or 136)
    int paf_average;           // RMPE_GT_PAF_AVERAGE
    int batch;                 // k_raster_blocks: samples of the launch (1-D grid)
};

struct LimbRec {
    double x1, y1, xD, yD, norm2;
    double tn, en;               // thre * norm2 and its half-ulp allowance (band_on); en = 0: use the division
    float ux, uy;
    int minx, maxx, miny, maxy;  // cell box, max exclusive; valid iff maxx > minx  (16-byte aligned: one 128-bit load)
};
static_assert(sizeof(LimbRec) == 80 && offsetof(LimbRec, minx) == 64, "LimbRec layout");

// thre * norm2 and the allowance of band_on for a power-of-two thre (8.0 in the reference); 0 otherwise
__device__ __forceinline__ void band_prepare(LimbRec &r, double thre) {
    const bool pow2 = thre > 0.0 && (__double2hiint(thre) & 0x000FFFFF) == 0 && __double2loint(thre) == 0;
    r.tn = __dmul_rn(thre, r.norm2);                 // exact for a power of two
    r.en = pow2 ? __dmul_rn(r.tn, 1.1102230246251565e-16) : 0.0;   // * 2^-53: half an ulp of thre, times norm2 (exact)
    if (!(r.tn < 1e300) || !(r.en > 1e-300)) r.en = 0.0;
}

// py_rmpe_heatmapper.py:116-118: abs(dd / norm) <= thre with dd / norm rounded to f64.  RN(q) <= thre  <=>  q <= thre (1 +
// 2^-53) (the midpoint above a power of two ties to even, i.e. down)  <=>  |dd| - thre*norm <= thre*norm*2^-53.  Both products
// are exact for a power-of-two thre and the subtraction is exact wherever the answer is in doubt (|dd| within a factor 2 of
// thre*norm, Sterbenz), so the test needs no division; any other thre takes the division.
__device__ __forceinline__ bool band_on(const LimbRec &r, double dd, double thre) {
    if (r.en > 0.0) return __dsub_rn(fabs(dd), r.tn) <= r.en;
    return fabs(__ddiv_rn(dd, r.norm2)) <= thre;
}

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

template <typename T>
__device__ __forceinline__ void store_pair(T *p, T a, T b);
template <>
__device__ __forceinline__ void store_pair<float>(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
template <>
__device__ __forceinline__ void store_pair<double>(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

// python-3 round(): half to even, on an f64 value already divided by the stride
__device__ inline int py_round(double v) { return __double2int_rn(v); }

// persons of sample b that the kernels look at: n_persons is caller data, never trusted beyond the person stride
__device__ __forceinline__ int raster_persons(const RasterArgs &a, int b, int cap, bool &clamped) {
    const int n = a.n_persons[b];
    const int hi = min(a.max_persons, cap);
    clamped = n < 0 || n > hi;
    return max(0, min(n, hi));
}

// ---- T4: keypoint transform + flip swap (py_rmpe_transformer.py:100-111) into shared memory ----
__device__ __forceinline__ void raster_joints(const RasterArgs &a, int b, int P, double *s_j, int tid, int nthreads) {
    for (int i = tid; i < P * kParts; i += nthreads) {
        const int p = i / kParts, part = i - p * kParts;
        const double *jin = a.joints + ((size_t)b * a.max_persons + p) * (kParts * 3);
        double ox, oy, ov;
        if (a.no_transform) {
            ox = jin[part * 3 + 0]; oy = jin[part * 3 + 1]; ov = jin[part * 3 + 2];
        } else {
            const int sp = a.flip[b] ? c_flip_partner[part] : part;
            const double x = jin[sp * 3 + 0], y = jin[sp * 3 + 1];
            ov = jin[sp * 3 + 2];
            const double *M = a.M + 6 * b;
            // numpy matmul order: t = M0*x; t = fma(M1, y, t); out = t + M2
            ox = __dadd_rn(__fma_rn(M[1], y, __dmul_rn(M[0], x)), M[2]);
            oy = __dadd_rn(__fma_rn(M[4], y, __dmul_rn(M[3], x)), M[5]);
        }
        s_j[i * 3 + 0] = ox; s_j[i * 3 + 1] = oy; s_j[i * 3 + 2] = ov;
    }
}
__device__ __forceinline__ void raster_joints_out(const RasterArgs &a, int b, int P, const double *s_j, int tid, int nthreads) {
    if (!a.out_joints) return;
    double *jo = a.out_joints + (size_t)b * a.max_persons * (kParts * 3);
    for (int i = tid; i < P * kParts * 3; i += nthreads) jo[i] = s_j[i];
}

// ---- H1/H2 tables of one (part, person, cell): cell centres 8i + 3.5 (py_rmpe_heatmapper.py:22-23, 51-57) ----
__device__ __forceinline__ void raster_exp_entry(const double *j, int cidx, float inv2s2, float &ex, float &ey) {
    const double g = 8.0 * cidx + 3.5;
    const float dx = (float)(g - j[0]), dy = (float)(g - j[1]);
    const bool vis = j[2] < 2.0;
    ex = vis ? expf(-(dx * dx) * inv2s2) : 0.f;
    ey = vis ? expf(-(dy * dy) * inv2s2) : 0.f;
}

// ---- H4 record of one (limb, person) (py_rmpe_heatmapper.py:69-114); returns true for a zero-length limb ----
// [row_lo, row_hi): the grid rows the calling CTA rasterises; a record whose box misses them is left empty and skips the
// square root and the divisions (the zero-length test needs neither: sqrt(s) == 0 <=> s == 0)
__device__ __forceinline__ bool raster_limb_rec(const double *jf, const double *jt, double thre, LimbRec &r,
                                                int row_lo = 0, int row_hi = kGrid) {
    bool zero = false;
    r.minx = r.maxx = r.miny = r.maxy = 0;
    r.x1 = jf[0]; r.y1 = jf[1];
    const double x2 = jt[0], y2 = jt[1];
    r.xD = __dsub_rn(x2, r.x1); r.yD = __dsub_rn(y2, r.y1);
    const double n2 = __dadd_rn(__dmul_rn(r.xD, r.xD), __dmul_rn(r.yD, r.yD));
    r.norm2 = 0.0;
    r.ux = r.uy = 0.f;
    r.tn = r.en = 0.0;
    if (jf[2] < 2.0 && jt[2] < 2.0) {
        if (n2 == 0.0) {
            zero = true;
        } else {
            const double mnx = r.x1 < x2 ? r.x1 : x2, mxx = r.x1 < x2 ? x2 : r.x1;
            const double mny = r.y1 < y2 ? r.y1 : y2, mxy = r.y1 < y2 ? y2 : r.y1;
            // round((v -+ thre) / 8): the division by the stride is an exact scaling, written as a multiplication
            const int a0 = py_round(__dmul_rn(__dsub_rn(mnx, thre), 0.125));
            const int b0 = py_round(__dmul_rn(__dsub_rn(mny, thre), 0.125));
            const int a1 = py_round(__dmul_rn(__dadd_rn(mxx, thre), 0.125));
            const int b1 = py_round(__dmul_rn(__dadd_rn(mxy, thre), 0.125));
            if (a1 >= 0 && b1 >= 0) {
                const int minx = max(a0, 0), miny = max(b0, 0), maxx = min(a1, kGrid), maxy = min(b1, kGrid);
                if (maxx > minx && maxy > miny && miny < row_hi && maxy > row_lo) {
                    r.minx = minx; r.miny = miny; r.maxx = maxx; r.maxy = maxy;
                    // distances(): sqrt(xD**2 + yD**2); put_vector_maps: sqrt(dx*dx + dy*dy) -- same value
                    r.norm2 = __dsqrt_rn(n2);
                    r.ux = (float)__ddiv_rn(r.xD, r.norm2);
                    r.uy = (float)__ddiv_rn(r.yD, r.norm2);
                    band_prepare(r, thre);
                }
            }
        }
    }
    return zero;
}

// distance numerator of a grid cell's top-left corner (8x, 8y) to the limb's line, f64 un-fused (distances(), :141-155)
__device__ __forceinline__ double raster_dd(const LimbRec &r, int x, int y) {
    const double X = (double)(8 * x), Y = (double)(8 * y);
    return __dsub_rn(__dmul_rn(r.xD, __dsub_rn(r.y1, Y)), __dmul_rn(__dsub_rn(r.x1, X), r.yD));
}

// ==========================================================================================
// k_raster_small: <= 4 persons per sample (COCO crops), planar labels.
// grid = (plane groups, batch); a thread owns ONE float4 run of pixels and writes it for all planes of its group.
// Everything the planes need is built once per CTA -- transformed joints, the separable exp tables of the group's
// parts x persons, the limb records of the group's limbs x persons -- so the plane loop has no barrier and every
// store is an independent coalesced 128-bit write.  One group (all 57 planes) is the default: splitting a sample over
// 2 or 4 CTAs (heat | PAF; parts 0-8 | parts 9-17 + background | limbs 0-9 | limbs 10-18, RMPE_RASTER_GROUPS) evens out
// the 256 CTAs over 148 SMs but repeats the per-CTA load/table latency chain: measured 46 / 49 / 57 us per 256 samples.
// ==========================================================================================
constexpr int kRsMaxP = 4;
constexpr int kRsGroups = 1;         // default plane groups per sample (gridDim.x): 1, 2 or 4

template <typename T>
__global__ void __launch_bounds__(544) k_raster_small(RasterArgs a) {
    const int b = blockIdx.y;
    const int grp = blockIdx.x, n_grp = gridDim.x;
    const int tid = threadIdx.x;
    const int kRsThreads = blockDim.x;
    bool clamped;
    const int P = raster_persons(a, b, kRsMaxP, clamped);
    // planes of this group: parts [part_lo, part_hi) are stored, the background needs the max over all 18 parts
    int part_lo = 0, part_hi = kParts, limb_lo = 0, limb_hi = kLimbs;
    bool want_bkg = true;
    if (n_grp == 2) {
        if (grp == 0) { limb_hi = 0; } else { part_hi = 0; want_bkg = false; }
    } else if (n_grp == 4) {
        if (grp == 0) { part_hi = 9; want_bkg = false; limb_hi = 0; }
        else if (grp == 1) { part_lo = 9; limb_hi = 0; }
        else if (grp == 2) { part_hi = 0; want_bkg = false; limb_hi = 10; }
        else { part_hi = 0; want_bkg = false; limb_lo = 10; }
    }
    const int tab_lo = want_bkg ? 0 : part_lo, tab_hi = want_bkg ? kParts : part_hi;

    // dynamic shared memory sized for the launch's max_persons (raster_small_smem): 24 KB at 3 persons
    extern __shared__ __align__(16) uint8_t rs_smem[];
    const int PM = a.max_persons;
    LimbRec *s_rec = reinterpret_cast<LimbRec *>(rs_smem);                     // [19][PM]
    double *s_j = reinterpret_cast<double *>(s_rec + kLimbs * PM);             // [PM][18][3]
    float *s_ex = reinterpret_cast<float *>(s_j + PM * kParts * 3);            // [18][P][46]
    float *s_ey = s_ex + kParts * PM * kGrid;
    __shared__ int s_zero;
    if (tid == 0) s_zero = 0;

    raster_joints(a, b, P, s_j, tid, kRsThreads);
    __syncthreads();

    const float inv2s2 = (float)(1.0 / (2.0 * a.sigma * a.sigma));
    for (int i = tab_lo * P * kGrid + tid; i < tab_hi * P * kGrid; i += kRsThreads) {
        const int cidx = i % kGrid, pp = i / kGrid;           // pp = part * P + p
        const int part = pp / P, p = pp - part * P;
        raster_exp_entry(s_j + (p * kParts + part) * 3, cidx, inv2s2, s_ex[i], s_ey[i]);
    }
    // ---- H4 limb records: [limb][person] ----
    const double thre = a.thre;
    for (int i = limb_lo * P + tid; i < limb_hi * P; i += kRsThreads) {
        const int k = i / P, p = i - k * P;
        LimbRec r;
        if (raster_limb_rec(s_j + (p * kParts + c_limb_from[k]) * 3, s_j + (p * kParts + c_limb_to[k]) * 3, thre, r)) s_zero = 1;
        s_rec[i] = r;
    }
    __syncthreads();

    // Launched with programmatic stream serialization behind k_warp_fused: everything above (joints, tables, limb records)
    // reads only the caller's inputs and touches no global memory, so it overlaps the tail of the warp kernel; the 46x46
    // mask is that kernel's output and status / out_joints are shared with it, so wait before the first of them.
    pdl_wait();
    if (grp == 0) {
        raster_joints_out(a, b, P, s_j, tid, kRsThreads);
        if (tid == 0 && (s_zero || clamped))
            atomicOr(a.status + b, (s_zero ? RMPE_ST_ZERO_LIMB : 0) | (clamped ? RMPE_ST_PERSONS_CLAMPED : 0));
    }
    const int run = tid;
    if (run >= kCellVec) return;
    const int pix = 4 * run;
    int y[4], x[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { y[q] = (pix + q) / kGrid; x[q] = (pix + q) - y[q] * kGrid; }
    T m[4];
    {
        const T *mk = reinterpret_cast<const T *>(a.mask) + (size_t)b * kCells + pix;
#pragma unroll
        for (int q = 0; q < 4; q++) m[q] = mk[q];
    }
    T *lab = reinterpret_cast<T *>(a.labels) + (size_t)b * kLayers * kCells + pix;

    // ---- H2/H3: Gaussian part maps with max merge, background = 1 - max ----
    float bk[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int part = tab_lo; part < tab_hi; part++) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int p = 0; p < P; p++) {
            // a run starts on an even column: its four cells are two aligned pairs (the second one on the next grid row
            // when the run wraps at column 44), each pair sharing its row's ey
            const float *ex = s_ex + (part * P + p) * kGrid, *ey = s_ey + (part * P + p) * kGrid;
            const float2 ea = *reinterpret_cast<const float2 *>(ex + x[0]);
            const float2 eb = *reinterpret_cast<const float2 *>(ex + x[2]);
            const float ya = ey[y[0]], yb = ey[y[2]];
            v[0] = fmaxf(v[0], ya * ea.x); v[1] = fmaxf(v[1], ya * ea.y);
            v[2] = fmaxf(v[2], yb * eb.x); v[3] = fmaxf(v[3], yb * eb.y);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) bk[q] = fmaxf(bk[q], v[q]);
        if (part >= part_lo) store4<T>(lab + (size_t)(38 + part) * kCells, v[0], v[1], v[2], v[3], m);
    }
    if (want_bkg) store4<T>(lab + (size_t)56 * kCells, 1.f - bk[0], 1.f - bk[1], 1.f - bk[2], 1.f - bk[3], m);

    // ---- H4: part-affinity fields ----
    const int ymin = y[0], ymax = y[3];
    for (int k = limb_lo; k < limb_hi; k++) {
        float vx[4] = {0.f, 0.f, 0.f, 0.f}, vy[4] = {0.f, 0.f, 0.f, 0.f};
        int cnt[4] = {0, 0, 0, 0};
        for (int p = 0; p < P; p++) {
            const LimbRec &r = s_rec[k * P + p];
            const int4 box = *reinterpret_cast<const int4 *>(&r.minx);      // minx, maxx, miny, maxy
            if (box.y <= box.x || box.w <= ymin || box.z > ymax) continue;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (x[q] >= box.x && x[q] < box.y && y[q] >= box.z && y[q] < box.w) {
                    if (band_on(r, raster_dd(r, x[q], y[q]), thre)) { vx[q] = r.ux; vy[q] = r.uy; cnt[q]++; }
                }
            }
        }
        store4<T>(lab + (size_t)(2 * k) * kCells, vx[0], vx[1], vx[2], vx[3], m);
        store4<T>(lab + (size_t)(2 * k + 1) * kCells, vy[0], vy[1], vy[2], vy[3], m);
        if (a.out_count)
            *reinterpret_cast<int4 *>(a.out_count + ((size_t)b * kLimbs + k) * kCells + pix) =
                make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
    }
}

// ==========================================================================================
// k_raster_roles: crowded scenes (up to 64 persons) and the NHWC (Keras) outputs.
// grid = (2 roles, batch): role 0 = 18 Gaussian part maps + background (+ joints out), role 1 = 19 limbs.  Each CTA builds
// what its planes need ONCE -- role 0 the separable exp tables [part][person][46] (all 18 parts when they fit, else in
// batches of `part_batch` parts, one barrier pair per batch), role 1 the 19 x P limb records -- and then runs barrier-free
// plane loops like k_raster_small: 20 persons = 132 KB of tables + 30 KB of records instead of 36 table rebuilds and
// 38 barriers per sample.
//   planar labels (kNhwc = false): a thread owns one float4 run of pixels, every store a coalesced 128-bit write.
//   kNhwc: the tensors DataIteratorBase.gen builds on the host (training/ds_generators.py:52-63) leave the kernel directly:
//     y2 = labels[38:57] as (46,46,19), y1 = labels[0:38] as (46,46,38), x2 / x1 = the mask repeated.  A thread owns one
//     PIXEL of a band of `band_px` pixels and puts its 19 / 38 channel values into a [pixel][channel] band buffer in shared
//     memory (odd word stride / 64-bit pairs: conflict-free); the band is one contiguous piece of the NHWC tensor and
//     leaves as flat 128-bit copies.  Optional planar outputs (labels, count) are still written, as scalar coalesced stores.
// ==========================================================================================
constexpr int kRrThreads = 544;
constexpr int kHeatCh = kParts + 1;    // 19 channels of y2 / x2
constexpr int kPafCh = 2 * kLimbs;     // 38 channels of y1 / x1

template <typename T>
__device__ __forceinline__ void flat_copy16(T *dst, const T *src_smem, int n_elems, int tid) {
    // n_elems * sizeof(T) is a multiple of 16 and dst is 16-byte aligned (checked on the host)
    const int n16 = (int)((size_t)n_elems * sizeof(T) / 16);
    const uint4 *s = reinterpret_cast<const uint4 *>(src_smem);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (int i = tid; i < n16; i += kRrThreads) d[i] = s[i];
}
// x[e] = mask of pixel e / C for the band's elements (mask repeated over the channels)
template <typename T, int C>
__device__ __forceinline__ void flat_mask_repeat(T *dst, const T *s_m, int n_px, int tid) {
    constexpr int V = 16 / sizeof(T);
    const int n16 = n_px * C / V;     // n_px * C * sizeof(T) is a multiple of 16
    for (int i = tid; i < n16; i += kRrThreads) {
        T v[V];
#pragma unroll
        for (int q = 0; q < V; q++) v[q] = s_m[(i * V + q) / C];
        *reinterpret_cast<uint4 *>(dst + (size_t)i * V) = *reinterpret_cast<const uint4 *>(v);
    }
}

template <typename T, bool kNhwc>
__global__ void __launch_bounds__(kRrThreads) k_raster_roles(RasterArgs a) {
    const int b = blockIdx.y, role = blockIdx.x;
    const int tid = threadIdx.x;
    bool clamped;
    const int P = raster_persons(a, b, kMaxPersonsGt, clamped);
    const int PM = a.max_persons;
    extern __shared__ __align__(16) uint8_t rr_smem[];
    double *s_j = reinterpret_cast<double *>(rr_smem);                      // [PM][18][3]
    uint8_t *rest = rr_smem + (((size_t)PM * kParts * 3 * 8 + 15) & ~(size_t)15);
    __shared__ int s_zero;
    if (tid == 0) s_zero = 0;
    raster_joints(a, b, P, s_j, tid, kRrThreads);
    __syncthreads();
    const T *mk = reinterpret_cast<const T *>(a.mask) + (size_t)b * kCells;

    if (role == 0) {
        // ================= 18 Gaussian part maps + background =================
        const int PB = a.part_batch;
        float *s_ex = reinterpret_cast<float *>(rest);                      // [PB][P][46]
        float *s_ey = s_ex + (size_t)PB * PM * kGrid;
        const float inv2s2 = (float)(1.0 / (2.0 * a.sigma * a.sigma));
        auto build = [&](int pb0, int npb) {
            for (int i = tid; i < npb * P * kGrid; i += kRrThreads) {
                const int cidx = i % kGrid, pp = i / kGrid;
                const int part = pb0 + pp / P, p = pp % P;
                raster_exp_entry(s_j + (p * kParts + part) * 3, cidx, inv2s2, s_ex[i], s_ey[i]);
            }
        };
        build(0, min(PB, kParts));
        __syncthreads();
        pdl_wait();       // see k_raster_small: nothing above touches global memory that another kernel writes
        raster_joints_out(a, b, P, s_j, tid, kRrThreads);
        if (tid == 0 && clamped) atomicOr(a.status + b, RMPE_ST_PERSONS_CLAMPED);
        if (!kNhwc) {
            const int pix = 4 * tid;
            const bool on = tid < kCellVec;
            int y[4], x[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { y[q] = (pix + q) / kGrid; x[q] = (pix + q) - y[q] * kGrid; }
            T m[4] = {(T)0, (T)0, (T)0, (T)0};
            if (on) {
#pragma unroll
                for (int q = 0; q < 4; q++) m[q] = mk[pix + q];
            }
            T *lab = reinterpret_cast<T *>(a.labels) + (size_t)b * kLayers * kCells + pix;
            float bk[4] = {0.f, 0.f, 0.f, 0.f};
            for (int pb0 = 0; pb0 < kParts; pb0 += PB) {
                const int npb = min(PB, kParts - pb0);
                if (pb0) { __syncthreads(); build(pb0, npb); __syncthreads(); }
                if (!on) continue;
                for (int pl = 0; pl < npb; pl++) {
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    const float *ex = s_ex + (size_t)pl * P * kGrid, *ey = s_ey + (size_t)pl * P * kGrid;
#pragma unroll 4
                    for (int p = 0; p < P; p++, ex += kGrid, ey += kGrid) {
                        const float2 ea = *reinterpret_cast<const float2 *>(ex + x[0]);
                        const float2 eb = *reinterpret_cast<const float2 *>(ex + x[2]);
                        const float ya = ey[y[0]], yb = ey[y[2]];
                        v[0] = fmaxf(v[0], ya * ea.x); v[1] = fmaxf(v[1], ya * ea.y);
                        v[2] = fmaxf(v[2], yb * eb.x); v[3] = fmaxf(v[3], yb * eb.y);
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) bk[q] = fmaxf(bk[q], v[q]);
                    store4<T>(lab + (size_t)(38 + pb0 + pl) * kCells, v[0], v[1], v[2], v[3], m);
                }
            }
            if (on) store4<T>(lab + (size_t)56 * kCells, 1.f - bk[0], 1.f - bk[1], 1.f - bk[2], 1.f - bk[3], m);
        } else {
            // NHWC: all 18 parts' tables are resident (part_batch == 18, the host guarantees it)
            const int BP = a.band_px;
            T *s_band = reinterpret_cast<T *>(reinterpret_cast<uint8_t *>(s_ey + (size_t)PB * PM * kGrid));   // [BP][19]
            T *s_m = s_band + (size_t)BP * kHeatCh;
            T *lab = a.labels ? reinterpret_cast<T *>(a.labels) + (size_t)b * kLayers * kCells : nullptr;
            for (int band0 = 0; band0 < kCells; band0 += BP) {
                const int n_px = min(BP, kCells - band0);
                if (tid < n_px) {
                    const int pix = band0 + tid;
                    const int y = pix / kGrid, x = pix - y * kGrid;
                    const T m = mk[pix];
                    s_m[tid] = m;
                    float bk = 0.f;
                    T *o = s_band + (size_t)tid * kHeatCh;
                    for (int part = 0; part < kParts; part++) {
                        float v = 0.f;
                        const float *ex = s_ex + (size_t)part * P * kGrid + x, *ey = s_ey + (size_t)part * P * kGrid + y;
#pragma unroll 4
                        for (int p = 0; p < P; p++) v = fmaxf(v, ey[p * kGrid] * ex[p * kGrid]);
                        bk = fmaxf(bk, v);
                        const T out = (T)v * m;
                        o[part] = out;
                        if (lab) lab[(size_t)(38 + part) * kCells + pix] = out;
                    }
                    const T outb = (T)(1.f - bk) * m;
                    o[kParts] = outb;
                    if (lab) lab[(size_t)56 * kCells + pix] = outb;
                }
                __syncthreads();
                const size_t go = ((size_t)b * kCells + band0) * kHeatCh;
                if (a.y2) flat_copy16<T>(reinterpret_cast<T *>(a.y2) + go, s_band, n_px * kHeatCh, tid);
                if (a.x2) flat_mask_repeat<T, kHeatCh>(reinterpret_cast<T *>(a.x2) + go, s_m, n_px, tid);
                __syncthreads();
            }
        }
        return;
    }

    // ================= 19 limbs: part-affinity fields =================
    LimbRec *s_rec = reinterpret_cast<LimbRec *>(rest);                      // [19][P]
    const double thre = a.thre;
    for (int i = tid; i < kLimbs * P; i += kRrThreads) {
        const int k = i / P, p = i - k * P;
        LimbRec r;
        if (raster_limb_rec(s_j + (p * kParts + c_limb_from[k]) * 3, s_j + (p * kParts + c_limb_to[k]) * 3, thre, r)) s_zero = 1;
        s_rec[i] = r;
    }
    __syncthreads();
    pdl_wait();
    if (tid == 0 && s_zero) atomicOr(a.status + b, RMPE_ST_ZERO_LIMB);
    const bool avg = a.paf_average != 0;
    // one pixel against the records of limb k: the reference's overwrite in person order (last band that covers the pixel
    // wins), or the averaging variant (f64 sums in person order, divided by the count)
    auto limb_pixel = [&](int k, int x, int y, float &vx, float &vy, int &cnt) {
        vx = 0.f; vy = 0.f; cnt = 0;
        double sx = 0.0, sy = 0.0;
        const LimbRec *rk = s_rec + k * P;
        for (int p = 0; p < P; p++) {
            const LimbRec &r = rk[p];
            const int4 box = *reinterpret_cast<const int4 *>(&r.minx);
            if (x >= box.x && x < box.y && y >= box.z && y < box.w) {
                if (band_on(r, raster_dd(r, x, y), thre)) {
                    vx = r.ux; vy = r.uy; cnt++;
                    if (avg) { sx = __dadd_rn(sx, __ddiv_rn(r.xD, r.norm2)); sy = __dadd_rn(sy, __ddiv_rn(r.yD, r.norm2)); }
                }
            }
        }
        if (avg && cnt > 0) { vx = (float)__ddiv_rn(sx, (double)cnt); vy = (float)__ddiv_rn(sy, (double)cnt); }
    };
    if (!kNhwc) {
        if (tid >= kCellVec) return;
        const int pix = 4 * tid;
        int y[4], x[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { y[q] = (pix + q) / kGrid; x[q] = (pix + q) - y[q] * kGrid; }
        T m[4];
#pragma unroll
        for (int q = 0; q < 4; q++) m[q] = mk[pix + q];
        T *lab = reinterpret_cast<T *>(a.labels) + (size_t)b * kLayers * kCells + pix;
        const int ymin = y[0], ymax = y[3];
        for (int k = 0; k < kLimbs; k++) {
            float vx[4] = {0.f, 0.f, 0.f, 0.f}, vy[4] = {0.f, 0.f, 0.f, 0.f};
            int cnt[4] = {0, 0, 0, 0};
            if (!avg) {
                const LimbRec *rk = s_rec + k * P;
                for (int p = 0; p < P; p++) {
                    const LimbRec &r = rk[p];
                    const int4 box = *reinterpret_cast<const int4 *>(&r.minx);      // minx, maxx, miny, maxy
                    if (box.y <= box.x || box.w <= ymin || box.z > ymax) continue;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (x[q] >= box.x && x[q] < box.y && y[q] >= box.z && y[q] < box.w) {
                            if (band_on(r, raster_dd(r, x[q], y[q]), thre)) { vx[q] = r.ux; vy[q] = r.uy; cnt[q]++; }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) limb_pixel(k, x[q], y[q], vx[q], vy[q], cnt[q]);
            }
            store4<T>(lab + (size_t)(2 * k) * kCells, vx[0], vx[1], vx[2], vx[3], m);
            store4<T>(lab + (size_t)(2 * k + 1) * kCells, vy[0], vy[1], vy[2], vy[3], m);
            if (a.out_count)
                *reinterpret_cast<int4 *>(a.out_count + ((size_t)b * kLimbs + k) * kCells + pix) =
                    make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
        }
    } else {
        const int BP = a.band_px;
        T *s_band = reinterpret_cast<T *>(rest + (((size_t)kLimbs * PM * sizeof(LimbRec) + 15) & ~(size_t)15));   // [BP][38]
        T *s_m = s_band + (size_t)BP * kPafCh;
        T *lab = a.labels ? reinterpret_cast<T *>(a.labels) + (size_t)b * kLayers * kCells : nullptr;
        for (int band0 = 0; band0 < kCells; band0 += BP) {
            const int n_px = min(BP, kCells - band0);
            if (tid < n_px) {
                const int pix = band0 + tid;
                const int y = pix / kGrid, x = pix - y * kGrid;
                const T m = mk[pix];
                s_m[tid] = m;
                T *o = s_band + (size_t)tid * kPafCh;
                for (int k = 0; k < kLimbs; k++) {
                    float vx, vy;
                    int cnt;
                    limb_pixel(k, x, y, vx, vy, cnt);
                    const T ox = (T)vx * m, oy = (T)vy * m;
                    store_pair<T>(o + 2 * k, ox, oy);      // one 64- / 128-bit store: conflict-free at the 38-word pixel stride
                    if (lab) { lab[(size_t)(2 * k) * kCells + pix] = ox; lab[(size_t)(2 * k + 1) * kCells + pix] = oy; }
                    if (a.out_count) a.out_count[((size_t)b * kLimbs + k) * kCells + pix] = cnt;
                }
            }
            __syncthreads();
            const size_t go = ((size_t)b * kCells + band0) * kPafCh;
            if (a.y1) flat_copy16<T>(reinterpret_cast<T *>(a.y1) + go, s_band, n_px * kPafCh, tid);
            if (a.x1) flat_mask_repeat<T, kPafCh>(reinterpret_cast<T *>(a.x1) + go, s_m, n_px, tid);
            __syncthreads();
        }
    }
}

// ==========================================================================================
// k_raster_blocks: crowded scenes (5..64 persons per sample), planar labels.
// grid = (8, batch), 704 threads.  Crowded batches are small (64 samples per GPU in BASELINE configs[4]), so a sample is cut
// into eight CTAs that give every SM warps to switch between: with one thread per 4-cell run and one CTA per (sample, role)
// the slowest CTA's serial instruction stream set the kernel time (83 us per 64 samples at 20 persons).
//   CTAs 0-3, heat: one quarter of the grid each (24 x 24 cells = 6 x 6 BLOCKS of 4x4 cells); a thread owns ONE (block,
//     part) item.  max_p exp(-d_p^2 / 2 sigma^2) = exp(-min_p d_p^2 / 2 sigma^2) (exp is monotone, so this IS the
//     reference's max merge, py_rmpe_heatmapper.py:59-62): the person loop is a min over separable squared distances --
//     per person two LDS.128 (dx^2 of the block's 4 columns, dy^2 of its 4 rows) feed 16 FADD and, two persons at a time,
//     16 three-input FMNMX -- and expf runs once per (part, cell) instead of twice per (part, person, column | row).  The
//     tables hold no expf, so every tile builds its own (24 columns + 24 rows per part and person).  The background needs
//     the max over all 18 parts: the part threads of a block meet through shared-memory atomicMax on the (non-negative)
//     float bits.
//   CTAs 4-7, limbs: five limbs each, whole grid.  SCATTER, then GATHER: a warp takes one (limb, person) record and walks its
//     cell box in 4 x 8 patches, one cell per lane (dense: no lane waits for another lane's person); a cell inside the band
//     sets bit p of the cell's 64-bit person mask in shared memory (atomicOr).  The reference overwrites in person order and
//     counts hits (py_rmpe_heatmapper.py:116-126), so the winner of a cell is its HIGHEST set bit and the count the number
//     of set bits.  The band test |dd| <= thre * norm is decided in float wherever the float value of the (linear) numerator
//     is further from the limit than its rounding bound, and by the exact f64 test (band_on) otherwise.  The gather pass
//     then writes the planes as coalesced 128-bit runs like k_raster_small.
// ==========================================================================================
#ifndef RMPE_RB_TILE_Y
#define RMPE_RB_TILE_Y 6
#endif
constexpr int kRbTile = 6;                           // blocks per tile row (24 cells)
constexpr int kRbTileY = RMPE_RB_TILE_Y;             // block rows per heat tile (6: four tiles per sample, 3: eight)
constexpr int kRbHeatTiles = 2 * (12 / kRbTileY);
constexpr int kRbBlocks = kRbTile * kRbTileY;        // 36
constexpr int kRbThreads = ((kRbBlocks * kLimbs + 31) / 32) * 32;    // 704
constexpr int kRbTabW = 4 * kRbTile + 4 * kRbTileY;  // floats per (part, person): dx^2 of the tile's 24 columns, dy^2 of its rows
constexpr int kRbLimbsPerCta = 5;                    // 5 + 5 + 5 + 4

static size_t raster_blocks_smem(int pm, int part_batch) {
    const size_t sj = ((size_t)pm * kParts * 3 * 8 + 15) & ~(size_t)15;
    const size_t heat = (size_t)part_batch * pm * kRbTabW * 4 + (size_t)kRbBlocks * 16 * 4;
    const size_t limb = (size_t)kRbLimbsPerCta * pm * sizeof(LimbRec) + (size_t)kRbLimbsPerCta * kCells * 8;
    return sj + std::max(heat, limb);
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

#ifndef RMPE_RB_MINB
#define RMPE_RB_MINB 2
#endif
template <typename T>
__global__ void __launch_bounds__(kRbThreads, RMPE_RB_MINB) k_raster_blocks(RasterArgs a) {
    // 1-D grid: the heat CTAs of all samples first, the (lighter) limb CTAs behind them fill the tail of the last wave
    const int role = blockIdx.x >= kRbHeatTiles * a.batch ? 1 : 0;
    const int id = blockIdx.x - role * kRbHeatTiles * a.batch;
    const int b = role ? id >> 2 : id / kRbHeatTiles, sub = role ? id & 3 : id - b * kRbHeatTiles;
    const int tid = threadIdx.x;
    bool clamped;
    const int P = raster_persons(a, b, kMaxPersonsGt, clamped);
    const int PM = a.max_persons;
    extern __shared__ __align__(16) uint8_t rb_smem[];
    double *s_j = reinterpret_cast<double *>(rb_smem);                      // [PM][18][3]
    uint8_t *rest = rb_smem + (((size_t)PM * kParts * 3 * 8 + 15) & ~(size_t)15);
    __shared__ int s_zero;
    if (tid == 0) s_zero = 0;
    raster_joints(a, b, P, s_j, tid, kRbThreads);

    if (role == 0) {
        // ================= 18 Gaussian part maps + background: tile `sub` =================
        const int tx0 = 24 * (sub & 1), ty0 = 4 * kRbTileY * (sub >> 1);      // first cell column / row of the tile
        const int slot = tid / kRbBlocks, blk = tid - slot * kRbBlocks;   // slot = part of this thread's item
        const int by = blk / kRbTile, bx = blk - by * kRbTile;
        const int x0 = tx0 + 4 * bx, y0 = ty0 + 4 * by;
        // the block's cells: pairs (x0, x0+1), (x0+2, x0+3) of rows y0..y0+3; the last block column / row hangs over the grid
        const bool pair1 = x0 + 2 < kGrid;
        const int nrows = min(4, kGrid - y0);
        const T *mk = reinterpret_cast<const T *>(a.mask) + (size_t)b * kCells + y0 * kGrid + x0;
        T *lab = reinterpret_cast<T *>(a.labels) + (size_t)b * kLayers * kCells + y0 * kGrid + x0;
        // v * mask of the block's cells -> plane (the mask comes through L1 every time: it would cost 16 registers to keep)
        auto store_plane = [&](int plane, const float *v) {
            T *o = lab + (size_t)plane * kCells;
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (r < nrows) {
                    store_pair<T>(o + r * kGrid, (T)v[4 * r] * mk[r * kGrid], (T)v[4 * r + 1] * mk[r * kGrid + 1]);
                    if (pair1) store_pair<T>(o + r * kGrid + 2, (T)v[4 * r + 2] * mk[r * kGrid + 2], (T)v[4 * r + 3] * mk[r * kGrid + 3]);
                }
        };
        const int PB = a.part_batch;                                                 // parts resident at a time (18: one pass)
        float *s_tab = reinterpret_cast<float *>(rest);                              // [PB][PM][kRbTabW]
        int *s_bk = reinterpret_cast<int *>(s_tab + (size_t)PB * PM * kRbTabW);     // [16][36] float bits
        for (int i = tid; i < kRbBlocks * 16; i += kRbThreads) s_bk[i] = 0;
        const float inv2s2 = (float)(1.0 / (2.0 * a.sigma * a.sigma));
        const float kInf = __int_as_float(0x7f800000);
        __syncthreads();
        for (int part0 = 0; part0 < kParts; part0 += PB) {
            const int npb = min(PB, kParts - part0);
            if (part0) __syncthreads();
            // squared distances of the resident parts to the cell centres 8 i + 3.5 (py_rmpe_heatmapper.py:22-23, 51-57):
            // f64 difference, rounded once to float, squared in float; +inf for a joint that is not there.  A thread
            // fills one (part, column) of the table for every person.
            for (int e = tid; e < npb * kRbTabW; e += kRbThreads) {
                const int pl = e / kRbTabW, col = e - pl * kRbTabW;
                const bool isx = col < 4 * kRbTile;
                const double gpos = 8.0 * (isx ? tx0 + col : ty0 + col - 4 * kRbTile) + 3.5;   // columns first, then the tile's rows
                const double *j = s_j + (part0 + pl) * 3 + (isx ? 0 : 1);
                float *t = s_tab + (size_t)pl * PM * kRbTabW + col;
                for (int p = 0; p < P; p++, j += kParts * 3, t += kRbTabW) {
                    const float d = (float)__dsub_rn(gpos, j[0]);
                    *t = j[isx ? 2 : 1] < 2.0 ? d * d : kInf;
                }
            }
            __syncthreads();
            if (part0 == 0) {
                // see k_raster_small: nothing above touches global memory that another kernel writes
                pdl_wait();
                if (sub == 0) {
                    raster_joints_out(a, b, P, s_j, tid, kRbThreads);
                    if (tid == 0 && clamped) atomicOr(a.status + b, RMPE_ST_PERSONS_CLAMPED);
                }
            }
            if (slot < npb) {
                float d2[16];
#pragma unroll
                for (int i = 0; i < 16; i++) d2[i] = kInf;
                const float *tp = s_tab + (size_t)slot * PM * kRbTabW + 4 * bx;
                const int yo = 4 * kRbTile + 4 * by - 4 * bx;
                int p = 0;
                for (; p + 1 < P; p += 2, tp += 2 * kRbTabW) {
                    const float4 dxa = *reinterpret_cast<const float4 *>(tp), dya = *reinterpret_cast<const float4 *>(tp + yo);
                    const float4 dxb = *reinterpret_cast<const float4 *>(tp + kRbTabW), dyb = *reinterpret_cast<const float4 *>(tp + kRbTabW + yo);
                    const float xa[4] = {dxa.x, dxa.y, dxa.z, dxa.w}, ya[4] = {dya.x, dya.y, dya.z, dya.w};
                    const float xb[4] = {dxb.x, dxb.y, dxb.z, dxb.w}, yb[4] = {dyb.x, dyb.y, dyb.z, dyb.w};
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int c = 0; c < 4; c++) d2[4 * r + c] = fmin3(d2[4 * r + c], xa[c] + ya[r], xb[c] + yb[r]);
                }
                if (p < P) {
                    const float4 dxa = *reinterpret_cast<const float4 *>(tp), dya = *reinterpret_cast<const float4 *>(tp + yo);
                    const float xa[4] = {dxa.x, dxa.y, dxa.z, dxa.w}, ya[4] = {dya.x, dya.y, dya.z, dya.w};
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int c = 0; c < 4; c++) d2[4 * r + c] = fminf(d2[4 * r + c], xa[c] + ya[r]);
                }
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    d2[i] = expf(-d2[i] * inv2s2);
                    // background = 1 - max over all 18 parts (py_rmpe_heatmapper.py:38): the part threads of a block meet here
                    atomicMax(s_bk + i * kRbBlocks + blk, __float_as_int(d2[i]));
                }
                store_plane(38 + part0 + slot, d2);
            }
        }
        __syncthreads();
        if (slot == 0) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 1.f - __int_as_float(s_bk[i * kRbBlocks + blk]);
            store_plane(56, v);
        }
        return;
    }

    // ================= part-affinity fields: limbs [k0, k0 + nl) =================
    const int k0 = kRbLimbsPerCta * sub, nl = min(kRbLimbsPerCta, kLimbs - k0);
    LimbRec *s_rec = reinterpret_cast<LimbRec *>(rest);                                                          // [nl][P]
    unsigned long long *s_hit = reinterpret_cast<unsigned long long *>(s_rec + (size_t)kRbLimbsPerCta * PM);     // [nl][2116] person masks
    for (int i = tid; i < nl * kCells; i += kRbThreads) s_hit[i] = 0ull;
    __syncthreads();         // joints
    const double thre = a.thre;
    for (int i = tid; i < nl * P; i += kRbThreads) {
        const int kl = i / P, p = i - kl * P;
        LimbRec r;
        if (raster_limb_rec(s_j + (p * kParts + c_limb_from[k0 + kl]) * 3, s_j + (p * kParts + c_limb_to[k0 + kl]) * 3, thre, r)) s_zero = 1;
        s_rec[i] = r;
    }
    __syncthreads();
    {
        // ---- scatter: one warp per record, one cell per lane ----
        const int warp = tid >> 5, lane = tid & 31;
        const int lx = lane & 7, ly = lane >> 3;
        for (int i = warp; i < nl * P; i += kRbThreads / 32) {
            const LimbRec &r = s_rec[i];
            const int4 box = *reinterpret_cast<const int4 *>(&r.minx);      // minx, maxx, miny, maxy
            if (box.y <= box.x) continue;
            const int kl = i / P, p = i - kl * P;
            unsigned long long *hit = s_hit + (size_t)kl * kCells;
            const unsigned long long bit = 1ull << p;
            // numerator dd = xD (y1 - Y) - (x1 - X) yD at the cell corner (X, Y) = (8x, 8y), in float with its rounding bound
            // E <= 2^-24 (|xD| |y1| + |yD| |x1| + 3 (|a| + |b|) + |dd|) <= 3e-5 (|xD| + |yD|) + 3e-7 (|a| + |b|) for Y, X <= 368
            const float fxD = (float)r.xD, fyD = (float)r.yD, fx1 = (float)r.x1, fy1 = (float)r.y1;
            const float tn_lo = (float)r.tn * 0.999999f, tn_hi = (float)r.tn * 1.000001f;
            const float e0 = 0.5f + 3e-5f * (fabsf(fxD) + fabsf(fyD));
            for (int px = box.x; px < box.y; px += 8) {
                const int x = px + lx;
                const float fb = (fx1 - (float)(8 * x)) * fyD;
                for (int py = box.z; py < box.w; py += 4) {
                    const int y = py + ly;
                    if (x < box.y && y < box.w) {
                        const float fa = fxD * (fy1 - (float)(8 * y));
                        const float ad = fabsf(fa - fb);
                        const float e = e0 + 1e-5f * (fabsf(fa) + fabsf(fb));
                        bool in = ad < tn_lo - e;
                        if (!in && !(ad > tn_hi + e)) in = band_on(r, raster_dd(r, x, y), thre);
                        if (in) atomicOr(hit + y * kGrid + x, bit);
                    }
                }
            }
        }
    }
    __syncthreads();
    pdl_wait();
    if (sub == 0 && tid == 0 && s_zero) atomicOr(a.status + b, RMPE_ST_ZERO_LIMB);
    // ---- gather: a thread owns float4 runs of cells; winner = highest person bit, count = number of bits ----
    const T *mk = reinterpret_cast<const T *>(a.mask) + (size_t)b * kCells;
    T *lab = reinterpret_cast<T *>(a.labels) + (size_t)b * kLayers * kCells;
    for (int i = tid; i < nl * kCellVec; i += kRbThreads) {
        const int kl = i / kCellVec, run = i - kl * kCellVec;
        const int pix = 4 * run, k = k0 + kl;
        const unsigned long long *h = s_hit + (size_t)kl * kCells + pix;
        const ulonglong2 h01 = *reinterpret_cast<const ulonglong2 *>(h), h23 = *reinterpret_cast<const ulonglong2 *>(h + 2);
        const unsigned long long hm[4] = {h01.x, h01.y, h23.x, h23.y};
        T m[4];
#pragma unroll
        for (int q = 0; q < 4; q++) m[q] = mk[pix + q];
        float vx[4], vy[4];
        int cnt[4];
        const LimbRec *rk = s_rec + kl * P;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            vx[q] = 0.f; vy[q] = 0.f;
            cnt[q] = __popcll(hm[q]);
            if (hm[q]) {
                const LimbRec &w = rk[63 - __clzll((long long)hm[q])];
                vx[q] = w.ux; vy[q] = w.uy;
            }
        }
        store4<T>(lab + (size_t)(2 * k) * kCells + pix, vx[0], vx[1], vx[2], vx[3], m);
        store4<T>(lab + (size_t)(2 * k + 1) * kCells + pix, vy[0], vy[1], vy[2], vy[3], m);
        if (a.out_count)
            *reinterpret_cast<int4 *>(a.out_count + ((size_t)b * kLimbs + k) * kCells + pix) = make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
    }
}

// ==========================================================================================
// k_keras_batch: DataIteratorBase.gen's per-sample transposes / repeats (training/ds_generators.py:47-63)
// as one pass: a CTA takes two grid rows (92 pixels) of one sample, reads the 57 label planes as
// coalesced row segments into shared memory and writes the four NHWC tensors as contiguous runs.
// ==========================================================================================
constexpr int kKbPix = 2 * kGrid;     // 92 pixels per CTA
constexpr int kKbThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kKbThreads) k_keras_batch(RmpeKerasBatch a) {
    __shared__ T s_lab[kLayers][kKbPix + 1];
    __shared__ T s_m[kKbPix];
    const int b = blockIdx.y, p0 = blockIdx.x * kKbPix, tid = threadIdx.x;
    const T *lab = reinterpret_cast<const T *>(a.labels) + (size_t)b * kLayers * kCells + p0;
    if (a.vec_label || a.heat_label)
        for (int i = tid; i < kLayers * kKbPix; i += kKbThreads) {
            const int c = i / kKbPix, p = i - c * kKbPix;
            s_lab[c][p] = lab[(size_t)c * kCells + p];
        }
    if (tid < kKbPix) s_m[tid] = reinterpret_cast<const T *>(a.mask)[(size_t)b * kCells + p0 + tid];
    __syncthreads();
    const size_t o38 = ((size_t)b * kCells + p0) * 38, o19 = ((size_t)b * kCells + p0) * 19;
    T *y1 = reinterpret_cast<T *>(a.vec_label), *y2 = reinterpret_cast<T *>(a.heat_label);
    T *x1 = reinterpret_cast<T *>(a.vec_weights), *x2 = reinterpret_cast<T *>(a.heat_weights);
    for (int i = tid; i < kKbPix * 38; i += kKbThreads) {
        const int p = i / 38, c = i - p * 38;
        if (y1) y1[o38 + i] = s_lab[c][p];
        if (x1) x1[o38 + i] = s_m[p];
    }
    for (int i = tid; i < kKbPix * 19; i += kKbThreads) {
        const int p = i / 19, c = i - p * 19;
        if (y2) y2[o19 + i] = s_lab[38 + c][p];
        if (x2) x2[o19 + i] = s_m[p];
    }
}

// ==========================================================================================
// host entry
// ==========================================================================================
constexpr int kMaxSmemOptin = 227 * 1024;

// footprint capacity (pixels) of one tile group when NG groups share a CTA
static int fused_foot_cap(int ng) {
    int bytes = (kMaxSmemOptin - kTabBytes) / ng - kGroupFixedBytes;
    return (bytes / 4) & ~31;   // keeps every group's base 128-byte aligned
}
static size_t fused_smem_bytes(int ng) { return (size_t)kTabBytes + (size_t)ng * ((size_t)fused_foot_cap(ng) * 4 + kGroupFixedBytes); }

// Work counter of a launch: one slot per stream.  Launches on one stream are ordered (a counter is zeroed and used by
// one launch at a time), launches on different streams never share a slot, so any number of them may be in flight.
// Streams beyond the table's size share its last slots by hash (kCounterRing streams is far more than the library or a
// sane caller creates); a destroyed stream's slot is simply reused by the next stream with that handle.
static int32_t *counter_for_stream(cudaStream_t st) {
    static std::mutex mu;
    static std::unordered_map<cudaStream_t, int> slots;
    std::lock_guard<std::mutex> lk(mu);
    auto it = slots.find(st);
    int slot;
    if (it != slots.end()) slot = it->second;
    else {
        slot = (int)slots.size() < kCounterRing ? (int)slots.size() : (int)(((size_t)st >> 4) % kCounterRing);
        slots.emplace(st, slot);
    }
    return tables().counters + (size_t)slot * kCounterStride;
}

template <int NG>
static int launch_fused(const FusedArgs &fa_, bool want_mask, int n_items, int sm_count, cudaStream_t st) {
    FusedArgs fa = fa_;
    fa.foot_cap = fused_foot_cap(NG);
    static const int prefetch = [] {
        // off by default: with tiles handed out dynamically, neighbouring tiles of a sample are staged at about the same
        // time by other groups and their halos bring the lines into L2 anyway (measured 0.3015 ms with, 0.2952 ms without)
        const char *e = getenv("RMPE_WARP_PREFETCH");
        return e ? atoi(e) : 0;
    }();
    fa.prefetch = prefetch;
    fa.counter = counter_for_stream(st);
    RMPE_CUDA_TRY(cudaMemsetAsync(fa.counter, 0, sizeof(int32_t), st));
    const size_t smem = fused_smem_bytes(NG);
    static std::once_flag attr_once;
    static cudaError_t attr_rc = cudaSuccess;
    std::call_once(attr_once, [smem] {
        cudaError_t e = cudaFuncSetAttribute(k_warp_fused<NG, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_fused<NG, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_fused<NG, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_fused<NG, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_rc = e;
    });
    RMPE_CUDA_TRY(attr_rc);
    const int grid = min((n_items + NG - 1) / NG, sm_count);
    const dim3 block(NG * kGroupThreads);
    if (fa.chw) {
        if (want_mask) k_warp_fused<NG, true, true><<<grid, block, smem, st>>>(fa, n_items);
        else k_warp_fused<NG, false, true><<<grid, block, smem, st>>>(fa, n_items);
    } else {
        if (want_mask) k_warp_fused<NG, true, false><<<grid, block, smem, st>>>(fa, n_items);
        else k_warp_fused<NG, false, false><<<grid, block, smem, st>>>(fa, n_items);
    }
    return RMPE_OK;
}

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

    const bool warp_only = (b->flags & RMPE_GT_WARP_ONLY) != 0;
    const bool simple = (b->flags & RMPE_GT_SIMPLE_KERNELS) != 0;
    const bool want_mask = !no_transform && !warp_only;
    bool mask_done = false;
    if (!no_warp) {
        RMPE_REQUIRE(((size_t)b->out_img & 3) == 0, "out_img must be 4-byte aligned");
        if (simple) {
            WarpArgs wa;
            wa.src_img = b->src_img; wa.desc = b->src_desc; wa.M = b->M; wa.out_img = b->out_img;
            wa.status = b->status; wa.tab = T.bicubic_i16; wa.tab_dp4a = T.bicubic_dp4a;
            wa.batch = b->batch; wa.chw = (b->flags & RMPE_GT_IMG_CHW) ? 1 : 0;
            dim3 grid((kOutW * kOutH + 255) / 256, b->batch);
            ProfScope ps("k_warp_simple", st);
            k_warp_simple<<<grid, 256, 0, st>>>(wa);
        } else {
            FusedArgs fa;
            fa.src_img = b->src_img; fa.src_mask = b->src_mask; fa.desc = b->src_desc; fa.M = b->M;
            fa.out_img = b->out_img; fa.out_mask = b->out_mask; fa.status = b->status;
            fa.tab = T.bicubic_i16; fa.tab_dp4a = T.bicubic_dp4a; fa.batch = b->batch;
            fa.chw = (b->flags & RMPE_GT_IMG_CHW) ? 1 : 0;
            fa.mask_f64 = (b->flags & RMPE_GT_LABELS_F64) ? 1 : 0;
            fa.foot_cap = 0;
            const int n_items = b->batch * kTilesPerSample;
            // 7 tile groups (896 threads, 72 registers, no spills) per SM; RMPE_WARP_GROUPS = 4 | 6 | 8 for A/B tests (measured
            // on B200, GT batch 256: 7 groups 0.296 ms, 6 groups 0.31, 8 groups 0.32 -- 8 spill at 64 registers and leave a
            // smaller footprint buffer)
            static const int ng = [] {
                const char *e = getenv("RMPE_WARP_GROUPS");
                int v = e ? atoi(e) : 0;
                return (v == 4 || v == 6 || v == 7 || v == 8) ? v : 0;
            }();
            ProfScope ps("k_warp_fused", st);
            int rc = ng == 4 ? launch_fused<4>(fa, want_mask, n_items, T.sm_count, st)
                   : ng == 6 ? launch_fused<6>(fa, want_mask, n_items, T.sm_count, st)
                   : ng == 8 ? launch_fused<8>(fa, want_mask, n_items, T.sm_count, st)
                             : launch_fused<7>(fa, want_mask, n_items, T.sm_count, st);
            if (rc != RMPE_OK) return rc;
            mask_done = want_mask;
        }
        count_launch();
    }
    if (want_mask && !mask_done) {
        MaskArgs ma;
        ma.src_mask = b->src_mask; ma.desc = b->src_desc; ma.M = b->M; ma.out_mask = b->out_mask;
        ma.tab = T.bicubic_i16; ma.f64 = (b->flags & RMPE_GT_LABELS_F64) ? 1 : 0;
        dim3 grid((kCells + 127) / 128, b->batch);
        ProfScope ps("k_mask46", st);
        k_mask46<<<grid, 128, 0, st>>>(ma);
        count_launch();
    }
    const bool nhwc = b->out_vec_label || b->out_heat_label || b->out_vec_weights || b->out_heat_weights;
    if ((b->out_labels || nhwc) && !warp_only) {
        RasterArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.joints = b->joints; ra.n_persons = b->n_persons; ra.M = b->M; ra.flip = b->flip;
        ra.mask = b->out_mask; ra.labels = b->out_labels; ra.out_joints = b->out_joints;
        ra.out_count = b->out_count; ra.status = b->status; ra.max_persons = b->max_persons;
        ra.f64 = (b->flags & RMPE_GT_LABELS_F64) ? 1 : 0; ra.no_transform = no_transform ? 1 : 0;
        // Heatmapper(sigma=7., thre=8.) (py_rmpe_heatmapper.py:10-14); 0 = the reference's defaults
        ra.sigma = b->sigma > 0.0 ? b->sigma : 7.0;
        ra.thre = b->thre > 0.0 ? b->thre : 8.0;
        ra.paf_average = (b->flags & RMPE_GT_PAF_AVERAGE) ? 1 : 0;
        ra.y1 = b->out_vec_label; ra.y2 = b->out_heat_label; ra.x1 = b->out_vec_weights; ra.x2 = b->out_heat_weights;
        RMPE_REQUIRE((((size_t)ra.y1 | (size_t)ra.y2 | (size_t)ra.x1 | (size_t)ra.x2 | (size_t)ra.labels) & 15) == 0,
                     "label outputs must be 16-byte aligned");
        const int pm = b->max_persons;
        const size_t esz = ra.f64 ? 8 : 4;
        bool two_pass = false;       // NHWC for more persons than the one-pass kernel holds tables for
        static const int small_off = [] { const char *e = getenv("RMPE_RASTER_SMALL"); return (e && atoi(e) == 0) ? 1 : 0; }();   // A/B: k_raster_blocks for every person count
        if (!nhwc && pm <= kRsMaxP && !ra.paf_average && !small_off) {       // RMPE_GT_SIMPLE_KERNELS selects the warp kernels only
            static const int groups = [] {
                const char *e = getenv("RMPE_RASTER_GROUPS");
                int v = e ? atoi(e) : kRsGroups;
                return (v == 1 || v == 2 || v == 4) ? v : kRsGroups;
            }();
            dim3 grid(groups, b->batch);
            const int kRsThreads = (kCellVec + 31) & ~31;
            const size_t smem = (size_t)kLimbs * pm * sizeof(LimbRec) + (size_t)pm * kParts * 3 * 8 + 2 * (size_t)kParts * pm * kGrid * 4;
            ProfScope ps("k_raster", st);
            // programmatic stream serialization behind k_warp_fused (see the griddepcontrol.wait in the kernel)
            if (ra.f64) RMPE_CUDA_TRY(launch_pdl(k_raster_small<double>, grid, dim3(kRsThreads), smem, st, ra));
            else RMPE_CUDA_TRY(launch_pdl(k_raster_small<float>, grid, dim3(kRsThreads), smem, st, ra));
            count_launch();
        } else {
            // k_raster_roles: shared memory = joints + max(role 0: exp tables of `part_batch` parts [+ band], role 1: limb
            // records [+ band]).  All 18 parts resident when that fits (no barrier in the plane loop); when the grid is more
            // than one wave of CTAs a smaller batch that lets two CTAs share an SM is preferred.
            static const int smem_target_kb = [] {
                const char *e = getenv("RMPE_RASTER_SMEM_KB");
                int v = e ? atoi(e) : 104;
                return (v >= 16 && v <= 220) ? v : 104;
            }();
            const size_t cap = (size_t)220 * 1024;
            const size_t sj = ((size_t)pm * kParts * 3 * 8 + 15) & ~(size_t)15;
            const size_t per_part = 2 * (size_t)pm * kGrid * 4;
            const size_t recs = ((size_t)kLimbs * pm * sizeof(LimbRec) + 15) & ~(size_t)15;
            auto fit_parts = [&](size_t budget) { return per_part ? (int)std::min<size_t>(kParts, budget > sj ? (budget - sj) / per_part : 0) : kParts; };
            size_t smem = 0;
            bool use_nhwc = nhwc;
            if (use_nhwc) {
                ra.part_batch = kParts;
                ra.band_px = 0;
                const int bands[3] = {544, 272, 136};
                auto need = [&](int bp) {
                    return sj + std::max(per_part * kParts + (size_t)bp * (kHeatCh + 1) * esz, recs + (size_t)bp * (kPafCh + 1) * esz);
                };
                for (int i = 0; i < 3 && !ra.band_px; i++) if (need(bands[i]) <= (size_t)smem_target_kb * 1024) ra.band_px = bands[i];
                for (int i = 0; i < 3 && !ra.band_px; i++) if (need(bands[i]) <= cap) ra.band_px = bands[i];
                if (ra.band_px) smem = need(ra.band_px);
                else { use_nhwc = false; two_pass = true; }
            }
            if (!use_nhwc) {
                RMPE_REQUIRE(ra.labels != nullptr, "out_labels is required (planar labels; also the scratch of the NHWC outputs "
                                                   "when max_persons is too large for the one-pass kernel)");
                const bool one_wave = 2 * b->batch <= T.sm_count;
                int pb = fit_parts(one_wave ? cap : (size_t)smem_target_kb * 1024);
                if (pb < 1) pb = fit_parts(cap);
                RMPE_REQUIRE(pb >= 1, "max_persons too large for the rasteriser's shared memory");
                ra.part_batch = pb;
                smem = sj + std::max(per_part * pb, recs);
                ra.y1 = ra.y2 = ra.x1 = ra.x2 = nullptr;
            }
            static const int crowd_blocks = [] {
                const char *e = getenv("RMPE_RASTER_CROWD");      // "roles": the previous crowded kernel, for A/B tests
                return (e && strcmp(e, "roles") == 0) ? 0 : 1;
            }();
            if (!use_nhwc && !ra.paf_average && crowd_blocks && raster_blocks_smem(pm, 1) <= cap) {
                // k_raster_blocks: 4x4-cell blocks, one (block, plane) item per thread.  All 18 parts' distance tables are
                // resident when two CTAs still share an SM (up to ~27 persons); fewer parts at a time beyond that.
                int pb = kParts;
                while (pb > 1 && raster_blocks_smem(pm, pb) > (size_t)smem_target_kb * 1024) pb--;
                if (raster_blocks_smem(pm, pb) > (size_t)smem_target_kb * 1024) { pb = kParts; while (pb > 1 && raster_blocks_smem(pm, pb) > cap) pb--; }
                ra.part_batch = pb;
                const size_t bsmem = raster_blocks_smem(pm, pb);
                static std::once_flag blocks_attr;
                static cudaError_t blocks_attr_rc = cudaSuccess;
                std::call_once(blocks_attr, [] {
                    const int mx = 220 * 1024;
                    cudaError_t e = cudaFuncSetAttribute(k_raster_blocks<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
                    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_raster_blocks<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
                    blocks_attr_rc = e;
                });
                RMPE_CUDA_TRY(blocks_attr_rc);
                dim3 bgrid((kRbHeatTiles + 4) * b->batch);
                ra.batch = b->batch;
                ProfScope ps("k_raster", st);
                if (ra.f64) RMPE_CUDA_TRY(launch_pdl(k_raster_blocks<double>, bgrid, dim3(kRbThreads), bsmem, st, ra));
                else RMPE_CUDA_TRY(launch_pdl(k_raster_blocks<float>, bgrid, dim3(kRbThreads), bsmem, st, ra));
                count_launch();
                RMPE_CUDA_TRY(cudaGetLastError());
                return RMPE_OK;
            }
            static std::once_flag roles_attr;
            static cudaError_t roles_attr_rc = cudaSuccess;
            std::call_once(roles_attr, [] {
                const int mx = 220 * 1024;      // = cap: the kernel also has a few bytes of static shared memory
                cudaError_t e = cudaFuncSetAttribute(k_raster_roles<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(k_raster_roles<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(k_raster_roles<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(k_raster_roles<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
                roles_attr_rc = e;
            });
            RMPE_CUDA_TRY(roles_attr_rc);
            dim3 grid(2, b->batch);
            {
                ProfScope ps("k_raster", st);
                if (ra.f64) {
                    if (use_nhwc) RMPE_CUDA_TRY(launch_pdl(k_raster_roles<double, true>, grid, dim3(kRrThreads), smem, st, ra));
                    else RMPE_CUDA_TRY(launch_pdl(k_raster_roles<double, false>, grid, dim3(kRrThreads), smem, st, ra));
                } else {
                    if (use_nhwc) RMPE_CUDA_TRY(launch_pdl(k_raster_roles<float, true>, grid, dim3(kRrThreads), smem, st, ra));
                    else RMPE_CUDA_TRY(launch_pdl(k_raster_roles<float, false>, grid, dim3(kRrThreads), smem, st, ra));
                }
                count_launch();
            }
            if (two_pass) {
                RmpeKerasBatch kb;
                kb.batch = b->batch; kb.flags = b->flags & RMPE_GT_LABELS_F64;
                kb.labels = b->out_labels; kb.mask = b->out_mask;
                kb.vec_weights = b->out_vec_weights; kb.heat_weights = b->out_heat_weights;
                kb.vec_label = b->out_vec_label; kb.heat_label = b->out_heat_label;
                const int rc = rmpe_keras_batch(&kb, st);
                if (rc != RMPE_OK) return rc;
            }
        }
    }
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}

extern "C" int rmpe_keras_batch(const RmpeKerasBatch *b, void *stream_) {
    if (!is_initialised()) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(b != nullptr && b->batch >= 0, "descriptor");
    if (b->batch == 0) return RMPE_OK;
    RMPE_REQUIRE(b->mask != nullptr, "mask is required");
    RMPE_REQUIRE(b->labels != nullptr || (!b->vec_label && !b->heat_label), "labels is required for y1 / y2");
    cudaStream_t st = (cudaStream_t)stream_;
    dim3 grid(kCells / kKbPix, b->batch);
    ProfScope ps("k_keras_batch", st);
    if (b->flags & RMPE_GT_LABELS_F64) k_keras_batch<double><<<grid, kKbThreads, 0, st>>>(*b);
    else k_keras_batch<float><<<grid, kKbThreads, 0, st>>>(*b);
    count_launch();
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}
