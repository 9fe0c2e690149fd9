// Host side of the C ABI: lifetime, error strings, interpolation tables, the augmentation
// matrix / RNG (T1), the host-buffer wrappers and padRightDownCorner.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "rmpe_common.cuh"

namespace rmpe {

static thread_local char t_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// ---- per-kernel event timing ----
namespace {
struct ProfEntry {
    std::string name;
    double ms = 0.0;
    int64_t launches = 0;
};
struct ProfPending {
    int entry;
    cudaEvent_t e0, e1;
};
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
std::vector<ProfEntry> g_prof;
std::vector<ProfPending> g_prof_pending;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t g_prof_open = nullptr;
int g_prof_open_entry = -1;

cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void prof_resolve() {  // g_prof_mu held
    for (auto &p : g_prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            g_prof[p.entry].ms += ms;
            g_prof[p.entry].launches += 1;
        }
        g_prof_pool.push_back(p.e0);
        g_prof_pool.push_back(p.e1);
    }
    g_prof_pending.clear();
}
}  // namespace

bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
void prof_mark(const char *name, cudaStream_t st, bool begin) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (begin) {
        int idx = -1;
        for (size_t i = 0; i < g_prof.size(); i++)
            if (g_prof[i].name == name) { idx = (int)i; break; }
        if (idx < 0) { g_prof.push_back(ProfEntry{name, 0.0, 0}); idx = (int)g_prof.size() - 1; }
        g_prof_open = prof_event();
        g_prof_open_entry = idx;
        cudaEventRecord(g_prof_open, st);
    } else if (g_prof_open) {
        cudaEvent_t e1 = prof_event();
        cudaEventRecord(e1, st);
        g_prof_pending.push_back(ProfPending{g_prof_open_entry, g_prof_open, e1});
        g_prof_open = nullptr;
        if (g_prof_pending.size() >= 4096) prof_resolve();
    }
}

struct Arena {
    uint8_t *dev = nullptr;
    size_t bytes = 0;
    size_t used = 0;
    int reserve(size_t want) {
        if (want <= bytes) return RMPE_OK;
        if (dev) cudaFree(dev);
        dev = nullptr;
        bytes = 0;
        size_t sz = want + (want >> 2) + (1 << 20);
        cudaError_t e = cudaMalloc((void **)&dev, sz);
        if (e != cudaSuccess) {
            set_error("arena cudaMalloc(%zu) failed: %s", sz, cudaGetErrorString(e));
            return RMPE_E_NOMEM;
        }
        bytes = sz;
        return RMPE_OK;
    }
    void reset() { used = 0; }
    void *take(size_t n) {
        size_t off = (used + 255) & ~(size_t)255;
        used = off + n + 16;  // 16 bytes of slack behind every block (bulk-copy over-read)
        return dev + off;
    }
    static size_t need(size_t n) { return ((n + 16 + 255) & ~(size_t)255) + 256; }
};

struct State {
    bool init = false;
    int device = -1;
    int generation = 0;           // bumped by every rmpe_init: per-thread contexts of an earlier life are rebuilt
    DeviceTables tab{};
    void *tab_mem = nullptr;
    std::mutex mu;                // lifetime only (init / shutdown / context registry): never held across a data call
};
static State g;

// Everything a *_host call needs besides its arguments -- device arena, main stream, the three streams of the chunk
// pipeline -- belongs to the CALLING THREAD: Keras worker threads (ds_generators.py:209 under fit_generator) call side
// by side without serialising on a library lock.  The contexts are registered so that rmpe_shutdown can free them.
struct HostCtx {
    Arena arena;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
    int generation = -1;
    bool registered = false;
    void release() {
        if (arena.dev) cudaFree(arena.dev);
        arena = Arena{};
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
        for (int i = 0; i < 3; i++) { if (pipe[i]) cudaStreamDestroy(pipe[i]); pipe[i] = nullptr; }
        generation = -1;
    }
    ~HostCtx();
};
static std::vector<HostCtx *> g_ctxs;     // guarded by g.mu
HostCtx::~HostCtx() {
    std::lock_guard<std::mutex> lk(g.mu);
    g_ctxs.erase(std::remove(g_ctxs.begin(), g_ctxs.end(), this), g_ctxs.end());
    if (g.init && generation == g.generation) { cudaSetDevice(g.device); release(); }
}
static thread_local HostCtx t_ctx;

// the calling thread's context, ready for use on the library's device
static int host_ctx(HostCtx **out) {
    HostCtx &c = t_ctx;
    int gen;
    {
        std::lock_guard<std::mutex> lk(g.mu);
        if (!g.init) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
        gen = g.generation;
        if (!c.registered) { g_ctxs.push_back(&c); c.registered = true; }
    }
    RMPE_CUDA_TRY(cudaSetDevice(g.device));
    if (c.generation != gen) {
        c.arena = Arena{}; c.stream = nullptr; c.pipe[0] = c.pipe[1] = c.pipe[2] = nullptr;   // an earlier life's handles died with its shutdown
        RMPE_CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        for (int i = 0; i < 3; i++) RMPE_CUDA_TRY(cudaStreamCreateWithFlags(&c.pipe[i], cudaStreamNonBlocking));
        c.generation = gen;
    }
    *out = &c;
    return RMPE_OK;
}
// error exit of a *_host call: no copy that touches the caller's buffers may still be in flight when it returns
static int host_fail(HostCtx &c, int rc) {
    cudaStreamSynchronize(c.stream);
    for (int i = 0; i < 3; i++) cudaStreamSynchronize(c.pipe[i]);
    return rc;
}
#define RMPE_HOST_TRY(expr)                                                               \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            rmpe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                            __FILE__, __LINE__);                                          \
            return host_fail(*ctx, RMPE_E_CUDA);                                          \
        }                                                                                 \
    } while (0)

bool is_initialised() { return g.init; }
const DeviceTables &tables() { return g.tab; }

// OpenCV initInterTab2D(INTER_CUBIC, fixpt=true): [ay][ax][ky][kx] int16, every 4x4 sums to 32768
static void build_bicubic_tables(std::vector<int16_t> &t16, std::vector<uint32_t> &tdp) {
    float tab1[32][4];
    for (int i = 0; i < 32; i++) cubic_coeffs((float)i * (1.0f / 32.0f), tab1[i]);
    t16.assign(32 * 32 * 16, 0);
    tdp.assign(32 * 32 * 8, 0);
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            int it[4][4];
            int isum = 0;
            for (int k1 = 0; k1 < 4; k1++) {
                float vy = tab1[i][k1];
                for (int k2 = 0; k2 < 4; k2++) {
                    float v = vy * tab1[j][k2];
                    long iv = lrintf(v * 32768.0f);
                    if (iv > 32767) iv = 32767;
                    if (iv < -32768) iv = -32768;
                    it[k1][k2] = (int)iv;
                    isum += (int)iv;
                }
            }
            if (isum != 32768) {
                int diff = isum - 32768;
                int Mk1 = 2, Mk2 = 2, mk1 = 2, mk2 = 2;
                for (int k1 = 2; k1 < 4; k1++)
                    for (int k2 = 2; k2 < 4; k2++) {
                        if (it[k1][k2] < it[mk1][mk2]) { mk1 = k1; mk2 = k2; }
                        else if (it[k1][k2] > it[Mk1][Mk2]) { Mk1 = k1; Mk2 = k2; }
                    }
                if (diff < 0) it[Mk1][Mk2] = (int16_t)(it[Mk1][Mk2] - diff);
                else it[mk1][mk2] = (int16_t)(it[mk1][mk2] - diff);
            }
            int16_t *o = &t16[(i * 32 + j) * 16];
            uint32_t *d = &tdp[(i * 32 + j) * 8];
            for (int k1 = 0; k1 < 4; k1++) {
                uint32_t hi = 0, lo = 0;
                for (int k2 = 0; k2 < 4; k2++) {
                    int w = it[k1][k2];
                    o[k1 * 4 + k2] = (int16_t)w;
                    int wh = w >> 8;          // arithmetic: w = 256*wh + wl, wl in [0,255]
                    int wl = w & 255;
                    hi |= (uint32_t)(wh & 255) << (8 * k2);
                    lo |= (uint32_t)wl << (8 * k2);
                }
                d[k1 * 2 + 0] = hi;
                d[k1 * 2 + 1] = lo;
            }
        }
}

}  // namespace rmpe

using namespace rmpe;

extern "C" int rmpe_init(int device) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.init) {
        if (device != g.device) { set_error("already initialised on device %d", g.device); return RMPE_E_BADARG; }
        return RMPE_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s): this library has no CPU path", cudaGetErrorString(e));
        return RMPE_E_CUDA;
    }
    RMPE_REQUIRE(device >= 0 && device < n, "device index out of range");
    RMPE_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    RMPE_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return RMPE_E_CUDA;
    }
    std::vector<int16_t> t16;
    std::vector<uint32_t> tdp;
    build_bicubic_tables(t16, tdp);
    size_t b16 = t16.size() * sizeof(int16_t), bdp = tdp.size() * sizeof(uint32_t);
    RMPE_CUDA_TRY(cudaMalloc(&g.tab_mem, b16 + bdp));
    RMPE_CUDA_TRY(cudaMemcpy(g.tab_mem, t16.data(), b16, cudaMemcpyHostToDevice));
    RMPE_CUDA_TRY(cudaMemcpy((uint8_t *)g.tab_mem + b16, tdp.data(), bdp, cudaMemcpyHostToDevice));
    g.tab.bicubic_i16 = (const int16_t *)g.tab_mem;
    g.tab.bicubic_dp4a = (const uint32_t *)((uint8_t *)g.tab_mem + b16);
    g.tab.sm_count = prop.multiProcessorCount;
    RMPE_CUDA_TRY(cudaMalloc(&g.tab.counters, (size_t)kCounterRing * kCounterStride * sizeof(int32_t)));
    g.device = device;
    g.generation++;
    g.init = true;
    return RMPE_OK;
}

extern "C" void rmpe_shutdown(void) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.init) return;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    if (g.tab_mem) cudaFree(g.tab_mem);
    if (g.tab.counters) cudaFree(g.tab.counters);
    for (HostCtx *c : g_ctxs) c->release();       // arenas and streams of every thread that made a *_host call
    g.init = false; g.device = -1; g.tab = DeviceTables{}; g.tab_mem = nullptr;
}

extern "C" const char *rmpe_last_error(void) { return t_error; }
extern "C" int rmpe_abi_version(void) { return RMPE_ABI_VERSION; }
extern "C" int rmpe_device(void) { return g.init ? g.device : -1; }
extern "C" int64_t rmpe_launch_count(void) { return g_launches.load(); }

extern "C" int rmpe_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on.store(on != 0);
    return RMPE_OK;
}
extern "C" int rmpe_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_resolve();
    g_prof.clear();
    return RMPE_OK;
}
extern "C" int rmpe_profile_count(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_resolve();
    return (int)g_prof.size();
}
extern "C" int rmpe_profile_get(int index, char *name_out, int name_cap, double *total_ms, int64_t *launches) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_resolve();
    RMPE_REQUIRE(index >= 0 && index < (int)g_prof.size(), "profile index out of range");
    RMPE_REQUIRE(name_out && name_cap > 0 && total_ms && launches, "null argument");
    snprintf(name_out, (size_t)name_cap, "%s", g_prof[index].name.c_str());
    *total_ms = g_prof[index].ms;
    *launches = g_prof[index].launches;
    return RMPE_OK;
}

// test hook: copies the host-built int16 table out (32*32*16 entries); no GPU needed
extern "C" int rmpe_debug_bicubic_table(int16_t *out) {
    std::vector<int16_t> t16;
    std::vector<uint32_t> tdp;
    build_bicubic_tables(t16, tdp);
    memcpy(out, t16.data(), t16.size() * sizeof(int16_t));
    return RMPE_OK;
}

// ------------------------------------------------------------------------------------------
// T1: AugmentSelection.affine (py_rmpe_transformer.py:39-78), closed form of the five-matrix
// product with the rounding sequence of the numpy chain (pinned in
// tests/test_host_logic.py against the reference's own chain)
// ------------------------------------------------------------------------------------------
extern "C" int rmpe_aug_affine(int n, const uint8_t *flip, const double *degree, const int32_t *crop_xy,
                               const double *scale, const double *center_xy, const double *scale_self,
                               double *M_out) {
    RMPE_REQUIRE(n >= 0 && flip && degree && crop_xy && scale && center_xy && scale_self && M_out, "null argument");
    for (int i = 0; i < n; i++) {
        volatile double rad = degree[i] / 180. * M_PI;
        double A = scale[i] * cos(rad);
        double B = scale[i] * sin(rad);
        double s = 0.6 / scale_self[i] * scale[i];
        double cx = center_xy[2 * i] + (double)crop_xy[2 * i];
        double cy = center_xy[2 * i + 1] + (double)crop_xy[2 * i + 1];
        double f = flip[i] ? -1.0 : 1.0;
        volatile double fs = f * s;
        volatile double m00 = fs * A, m01 = fs * B, m10 = s * (-B), m11 = s * A;
        volatile double p0 = m00 * (-cx), p1 = m10 * (-cx);
        double m02 = fma(m01, -cy, p0) + 184.0;
        double m12 = fma(m11, -cy, p1) + 184.0;
        double *M = M_out + 6 * i;
        M[0] = m00; M[1] = m01; M[2] = m02; M[3] = m10; M[4] = m11; M[5] = m12;
    }
    return RMPE_OK;
}

// ------------------------------------------------------------------------------------------
// AugmentSelection.random (py_rmpe_transformer.py:19-27) with CPython's Mersenne Twister
// ------------------------------------------------------------------------------------------
namespace {
struct MT {
    uint32_t mt[624];
    int idx;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253U * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void init_by_array(const uint32_t *key, int len) {
        init_genrand(19650218U);
        int i = 1, j = 0;
        int k = 624 > len ? 624 : len;
        for (; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525U)) + key[j] + (uint32_t)j;
            i++; j++;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
            if (j >= len) j = 0;
        }
        for (k = 623; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941U)) - (uint32_t)i;
            i++;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000U;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int kk = 0; kk < 624; kk++) {
                uint32_t y = (mt[kk] & 0x80000000U) | (mt[(kk + 1) % 624] & 0x7fffffffU);
                mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680U;
        y ^= (y << 15) & 0xefc60000U;
        y ^= (y >> 18);
        return y;
    }
    double random() {
        uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    double uniform(double lo, double hi) { return lo + (hi - lo) * random(); }
};
}  // namespace

extern "C" int rmpe_aug_random(int n, const uint64_t *seeds, uint8_t *flip, double *degree, int32_t *crop_xy,
                               double *scale) {
    RMPE_REQUIRE(n >= 0 && seeds && flip && degree && crop_xy && scale, "null argument");
    for (int i = 0; i < n; i++) {
        MT r;
        uint32_t key[2] = {(uint32_t)(seeds[i] & 0xffffffffU), (uint32_t)(seeds[i] >> 32)};
        r.init_by_array(key, key[1] ? 2 : 1);
        flip[i] = r.uniform(0., 1.) > 0.5 ? 1 : 0;
        degree[i] = r.uniform(-1., 1.) * 40.;
        // scale_prob = 1: "scale improbability"; the condition is drawn first, the value only if it fires
        if (r.uniform(0., 1.) > 1.0) scale[i] = (1.1 - 0.5) * r.uniform(0., 1.) + 0.5;
        else scale[i] = 1.;
        crop_xy[2 * i] = (int32_t)(r.uniform(-1., 1.) * 40.);
        crop_xy[2 * i + 1] = (int32_t)(r.uniform(-1., 1.) * 40.);
    }
    return RMPE_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer GT wrapper
// ------------------------------------------------------------------------------------------
extern "C" int rmpe_gt_batch_host(const RmpeGtBatchHost *h) {
    if (!g.init) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(h != nullptr, "descriptor is null");
    RMPE_REQUIRE(h->batch >= 0, "negative batch");
    if (h->batch == 0) return RMPE_OK;
    HostCtx *ctx = nullptr;
    int rc = host_ctx(&ctx);
    if (rc != RMPE_OK) return rc;
    const int B = h->batch;
    const bool no_transform = (h->flags & RMPE_GT_NO_TRANSFORM) != 0;
    const bool no_warp = (h->flags & RMPE_GT_NO_WARP) != 0 || no_transform;
    const bool f64 = (h->flags & RMPE_GT_LABELS_F64) != 0;
    const bool nhwc = h->out_vec_label || h->out_heat_label || h->out_vec_weights || h->out_heat_weights;
    const size_t esz = f64 ? 8 : 4;
    RMPE_REQUIRE(h->max_persons >= 0 && h->max_persons <= kMaxPersonsGt, "max_persons must be in [0,64]");
    if (!no_transform) RMPE_REQUIRE(h->src_mask && h->M && h->flip && h->src_height > 0 && h->src_width > 0, "missing source");
    if (!no_warp) RMPE_REQUIRE(h->src_img != nullptr, "src_img is null");
    RMPE_REQUIRE(h->n_persons != nullptr && (h->max_persons == 0 || h->joints != nullptr), "joints / n_persons");
    RMPE_REQUIRE(!no_transform || h->out_mask != nullptr, "with RMPE_GT_NO_TRANSFORM out_mask is the input mask");
    // n_persons lives on the host here: validate it instead of letting the kernels clamp it
    for (int i = 0; i < B; i++)
        RMPE_REQUIRE(h->n_persons[i] >= 0 && h->n_persons[i] <= h->max_persons, "n_persons[i] must be in [0, max_persons]");
    const size_t img_b = no_warp ? 0 : (size_t)B * h->src_height * h->src_width * 3;
    const size_t msk_b = no_transform ? 0 : (size_t)B * h->src_height * h->src_width;
    const size_t jnt_b = (size_t)B * h->max_persons * kParts * 3 * sizeof(double);
    const size_t oimg_b = no_warp ? 0 : (size_t)B * 3 * kOutW * kOutH;
    const size_t omsk_b = (size_t)B * kCells * esz;
    // planar labels are produced when asked for, and when nothing else would make the rasteriser run (it writes the joints)
    const bool planar = h->out_labels != nullptr || !nhwc || h->max_persons > 24;
    const size_t olab_b = planar ? (size_t)B * kLayers * kCells * esz : 0;
    const size_t o38_b = (size_t)B * kCells * 38 * esz, o19_b = (size_t)B * kCells * 19 * esz;
    const size_t ocnt_b = h->out_count ? (size_t)B * kLimbs * kCells * sizeof(int32_t) : 0;

    size_t need = Arena::need(img_b) + Arena::need(msk_b) + Arena::need(B * sizeof(RmpeSrcDesc)) + Arena::need(jnt_b) * 2 +
                  Arena::need(B * 4) * 2 + Arena::need(B * 48) + Arena::need(B) + Arena::need(oimg_b) +
                  Arena::need(omsk_b) + Arena::need(olab_b) + Arena::need(ocnt_b) +
                  (h->out_vec_label ? Arena::need(o38_b) : 0) + (h->out_vec_weights ? Arena::need(o38_b) : 0) +
                  (h->out_heat_label ? Arena::need(o19_b) : 0) + (h->out_heat_weights ? Arena::need(o19_b) : 0);
    rc = ctx->arena.reserve(need);
    if (rc != RMPE_OK) return rc;
    Arena &A = ctx->arena;
    A.reset();
    cudaStream_t st = ctx->stream;

    uint8_t *d_img = (uint8_t *)A.take(img_b);
    uint8_t *d_msk = (uint8_t *)A.take(msk_b);
    RmpeSrcDesc *d_desc = (RmpeSrcDesc *)A.take(B * sizeof(RmpeSrcDesc));
    double *d_j = (double *)A.take(jnt_b);
    double *d_jo = (double *)A.take(jnt_b);
    int32_t *d_np = (int32_t *)A.take(B * 4);
    int32_t *d_st = (int32_t *)A.take(B * 4);
    double *d_M = (double *)A.take(B * 48);
    uint8_t *d_flip = (uint8_t *)A.take(B);
    uint8_t *d_oimg = (uint8_t *)A.take(oimg_b);
    uint8_t *d_omsk = (uint8_t *)A.take(omsk_b);
    uint8_t *d_olab = olab_b ? (uint8_t *)A.take(olab_b) : nullptr;
    int32_t *d_ocnt = (int32_t *)A.take(ocnt_b);
    uint8_t *d_y1 = h->out_vec_label ? (uint8_t *)A.take(o38_b) : nullptr;
    uint8_t *d_x1 = h->out_vec_weights ? (uint8_t *)A.take(o38_b) : nullptr;
    uint8_t *d_y2 = h->out_heat_label ? (uint8_t *)A.take(o19_b) : nullptr;
    uint8_t *d_x2 = h->out_heat_weights ? (uint8_t *)A.take(o19_b) : nullptr;

    std::vector<RmpeSrcDesc> desc(B);
    for (int i = 0; i < B; i++) {
        desc[i].img_offset = (int64_t)i * h->src_height * h->src_width * 3;
        desc[i].mask_offset = (int64_t)i * h->src_height * h->src_width;
        desc[i].height = h->src_height; desc[i].width = h->src_width;
        desc[i].img_pitch = 3 * h->src_width; desc[i].mask_pitch = h->src_width;
    }
    // small per-sample tables first, on the main stream; the chunk streams wait for them
    RMPE_HOST_TRY(cudaMemcpyAsync(d_desc, desc.data(), B * sizeof(RmpeSrcDesc), cudaMemcpyHostToDevice, st));
    if (jnt_b) RMPE_HOST_TRY(cudaMemcpyAsync(d_j, h->joints, jnt_b, cudaMemcpyHostToDevice, st));
    RMPE_HOST_TRY(cudaMemcpyAsync(d_np, h->n_persons, B * 4, cudaMemcpyHostToDevice, st));
    if (!no_transform) {
        RMPE_HOST_TRY(cudaMemcpyAsync(d_M, h->M, B * 48, cudaMemcpyHostToDevice, st));
        RMPE_HOST_TRY(cudaMemcpyAsync(d_flip, h->flip, B, cudaMemcpyHostToDevice, st));
    }
    RMPE_HOST_TRY(cudaStreamSynchronize(st));

    // Chunk pipeline: sources in, kernels, results out, round-robin over three streams, so that the
    // host-to-device copy of chunk c+1, the kernels of chunk c and the device-to-host copy of chunk c-1
    // overlap (separate copy engines per direction).  Chunks are independent: samples never interact.
    static const int n_chunks = [] {
        const char *e = getenv("RMPE_HOST_CHUNKS");
        int v = e ? atoi(e) : 16;
        return (v >= 1 && v <= 64) ? v : 16;
    }();
    const int chunk = B <= 8 ? B : std::max(8, (B + n_chunks - 1) / n_chunks);
    const size_t src_img_b = (size_t)h->src_height * h->src_width * 3, src_msk_b = (size_t)h->src_height * h->src_width;
    const size_t lab1 = (size_t)kLayers * kCells * esz, msk1 = (size_t)kCells * esz;
    const size_t s38 = (size_t)kCells * 38 * esz, s19 = (size_t)kCells * 19 * esz;
    for (int c0 = 0, ci = 0; c0 < B; c0 += chunk, ci++) {
        const int n = std::min(chunk, B - c0);
        cudaStream_t cs = ctx->pipe[ci % 3];
        if (img_b) RMPE_HOST_TRY(cudaMemcpyAsync(d_img + c0 * src_img_b, h->src_img + c0 * src_img_b, n * src_img_b, cudaMemcpyHostToDevice, cs));
        if (msk_b) RMPE_HOST_TRY(cudaMemcpyAsync(d_msk + c0 * src_msk_b, h->src_mask + c0 * src_msk_b, n * src_msk_b, cudaMemcpyHostToDevice, cs));
        if (no_transform)
            RMPE_HOST_TRY(cudaMemcpyAsync(d_omsk + c0 * msk1, (const uint8_t *)h->out_mask + c0 * msk1, n * msk1, cudaMemcpyHostToDevice, cs));
        RmpeGtBatch d;
        memset(&d, 0, sizeof(d));
        d.batch = n; d.max_persons = h->max_persons; d.flags = h->flags;
        d.sigma = h->sigma; d.thre = h->thre;
        d.src_img = d_img; d.src_mask = d_msk; d.src_desc = d_desc + c0;     // descriptor offsets are batch-relative
        d.joints = d_j + (size_t)c0 * h->max_persons * kParts * 3; d.n_persons = d_np + c0;
        d.M = d_M + 6 * c0; d.flip = d_flip + c0;
        d.out_img = no_warp ? nullptr : d_oimg + (size_t)c0 * 3 * kOutW * kOutH;
        d.out_mask = d_omsk + c0 * msk1;
        d.out_labels = d_olab ? d_olab + c0 * lab1 : nullptr;
        d.out_joints = d_jo + (size_t)c0 * h->max_persons * kParts * 3;
        d.out_count = h->out_count ? d_ocnt + (size_t)c0 * kLimbs * kCells : nullptr;
        d.out_vec_label = d_y1 ? d_y1 + c0 * s38 : nullptr; d.out_vec_weights = d_x1 ? d_x1 + c0 * s38 : nullptr;
        d.out_heat_label = d_y2 ? d_y2 + c0 * s19 : nullptr; d.out_heat_weights = d_x2 ? d_x2 + c0 * s19 : nullptr;
        d.status = d_st + c0;
        rc = rmpe_gt_batch(&d, cs);
        if (rc != RMPE_OK) return host_fail(*ctx, rc);
        if (h->out_img && oimg_b)
            RMPE_HOST_TRY(cudaMemcpyAsync(h->out_img + (size_t)c0 * 3 * kOutW * kOutH, d.out_img, (size_t)n * 3 * kOutW * kOutH, cudaMemcpyDeviceToHost, cs));
        if (h->out_labels)
            RMPE_HOST_TRY(cudaMemcpyAsync((uint8_t *)h->out_labels + c0 * lab1, d.out_labels, n * lab1, cudaMemcpyDeviceToHost, cs));
        if (d_y1) RMPE_HOST_TRY(cudaMemcpyAsync((uint8_t *)h->out_vec_label + c0 * s38, d.out_vec_label, n * s38, cudaMemcpyDeviceToHost, cs));
        if (d_y2) RMPE_HOST_TRY(cudaMemcpyAsync((uint8_t *)h->out_heat_label + c0 * s19, d.out_heat_label, n * s19, cudaMemcpyDeviceToHost, cs));
        if (d_x1) RMPE_HOST_TRY(cudaMemcpyAsync((uint8_t *)h->out_vec_weights + c0 * s38, d.out_vec_weights, n * s38, cudaMemcpyDeviceToHost, cs));
        if (d_x2) RMPE_HOST_TRY(cudaMemcpyAsync((uint8_t *)h->out_heat_weights + c0 * s19, d.out_heat_weights, n * s19, cudaMemcpyDeviceToHost, cs));
        if (h->out_count)
            RMPE_HOST_TRY(cudaMemcpyAsync(h->out_count + (size_t)c0 * kLimbs * kCells, d.out_count, (size_t)n * kLimbs * kCells * 4, cudaMemcpyDeviceToHost, cs));
    }
    for (int i = 0; i < 3; i++) RMPE_HOST_TRY(cudaStreamSynchronize(ctx->pipe[i]));
    // the small outputs leave in one piece at the end: their host buffers are usually pageable, and a
    // device-to-host copy into pageable memory blocks the issuing thread -- inside the loop it would
    // serialise the chunk pipeline
    if (h->out_mask && !no_transform) RMPE_HOST_TRY(cudaMemcpyAsync(h->out_mask, d_omsk, omsk_b, cudaMemcpyDeviceToHost, st));
    if (h->out_joints && jnt_b) RMPE_HOST_TRY(cudaMemcpyAsync(h->out_joints, d_jo, jnt_b, cudaMemcpyDeviceToHost, st));
    if (h->status) RMPE_HOST_TRY(cudaMemcpyAsync(h->status, d_st, B * 4, cudaMemcpyDeviceToHost, st));
    RMPE_HOST_TRY(cudaStreamSynchronize(st));
    return RMPE_OK;
}

extern "C" int rmpe_keras_batch_host(const RmpeKerasBatch *h) {
    if (!g.init) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(h != nullptr && h->batch >= 0, "descriptor");
    if (h->batch == 0) return RMPE_OK;
    RMPE_REQUIRE(h->mask != nullptr, "mask is required");
    HostCtx *ctx = nullptr;
    int rc = host_ctx(&ctx);
    if (rc != RMPE_OK) return rc;
    Arena &A = ctx->arena;
    const size_t esz = (h->flags & RMPE_GT_LABELS_F64) ? 8 : 4;
    const size_t B = (size_t)h->batch;
    const size_t lab_b = h->labels ? B * kLayers * kCells * esz : 0, msk_b = B * kCells * esz;
    const size_t o38 = B * kCells * 38 * esz, o19 = B * kCells * 19 * esz;
    rc = A.reserve(Arena::need(lab_b) + Arena::need(msk_b) + 2 * Arena::need(o38) + 2 * Arena::need(o19));
    if (rc != RMPE_OK) return rc;
    A.reset();
    cudaStream_t st = ctx->stream;
    RmpeKerasBatch d = *h;
    void *d_lab = lab_b ? A.take(lab_b) : nullptr;
    void *d_msk = A.take(msk_b);
    if (lab_b) RMPE_HOST_TRY(cudaMemcpyAsync(d_lab, h->labels, lab_b, cudaMemcpyHostToDevice, st));
    RMPE_HOST_TRY(cudaMemcpyAsync(d_msk, h->mask, msk_b, cudaMemcpyHostToDevice, st));
    d.labels = d_lab; d.mask = d_msk;
    d.vec_weights = h->vec_weights ? A.take(o38) : nullptr;
    d.heat_weights = h->heat_weights ? A.take(o19) : nullptr;
    d.vec_label = h->vec_label ? A.take(o38) : nullptr;
    d.heat_label = h->heat_label ? A.take(o19) : nullptr;
    rc = rmpe_keras_batch(&d, st);
    if (rc != RMPE_OK) return host_fail(*ctx, rc);
    if (h->vec_weights) RMPE_HOST_TRY(cudaMemcpyAsync(h->vec_weights, d.vec_weights, o38, cudaMemcpyDeviceToHost, st));
    if (h->heat_weights) RMPE_HOST_TRY(cudaMemcpyAsync(h->heat_weights, d.heat_weights, o19, cudaMemcpyDeviceToHost, st));
    if (h->vec_label) RMPE_HOST_TRY(cudaMemcpyAsync(h->vec_label, d.vec_label, o38, cudaMemcpyDeviceToHost, st));
    if (h->heat_label) RMPE_HOST_TRY(cudaMemcpyAsync(h->heat_label, d.heat_label, o19, cudaMemcpyDeviceToHost, st));
    RMPE_HOST_TRY(cudaStreamSynchronize(st));
    return RMPE_OK;
}

// ------------------------------------------------------------------------------------------
// U1: util.padRightDownCorner (util.py:57-77)
// ------------------------------------------------------------------------------------------
namespace rmpe {
__global__ void k_pad_rd(const uint8_t *__restrict__ src, int H, int W, int C, int Hp, int Wp, int pad_value,
                         uint8_t *__restrict__ dst) {
    size_t n = (size_t)Hp * Wp * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t row = i / ((size_t)Wp * C);
        size_t rem = i - row * (size_t)Wp * C;
        int x = (int)(rem / C);
        dst[i] = (row < (size_t)H && x < W) ? src[(row * W + x) * C + (rem - (size_t)x * C)] : (uint8_t)pad_value;
    }
}
}  // namespace rmpe

extern "C" int rmpe_pad_right_down_corner(const uint8_t *src_dev, int height, int width, int channels, int stride,
                                          int pad_value, uint8_t *dst_dev, int *pad4_out, void *stream) {
    if (!g.init) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(src_dev && dst_dev && pad4_out && height > 0 && width > 0 && channels > 0 && stride > 0, "bad argument");
    int pd = (height % stride == 0) ? 0 : stride - height % stride;
    int pr = (width % stride == 0) ? 0 : stride - width % stride;
    pad4_out[0] = 0; pad4_out[1] = 0; pad4_out[2] = pd; pad4_out[3] = pr;
    size_t n = (size_t)(height + pd) * (width + pr) * channels;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_pad_rd<<<blocks, 256, 0, (cudaStream_t)stream>>>(src_dev, height, width, channels, height + pd, width + pr,
                                                        pad_value, dst_dev);
    count_launch();
    RMPE_CUDA_TRY(cudaGetLastError());
    return RMPE_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer decode wrapper
// ------------------------------------------------------------------------------------------
extern "C" int rmpe_decode_batch_host(const RmpeDecodeBatchHost *h) {
    if (!g.init) { set_error("rmpe_init not called"); return RMPE_E_NOTINIT; }
    RMPE_REQUIRE(h != nullptr, "descriptor is null");
    RMPE_REQUIRE(h->batch >= 0, "negative batch");
    if (h->batch == 0) return RMPE_OK;
    RMPE_REQUIRE(h->heat && h->paf && h->frames, "blobs / frames");
    RMPE_REQUIRE(h->candidate && h->n_peaks && h->subset && h->n_subset, "candidate / subset outputs");
    RMPE_REQUIRE(h->max_peaks > 0 && h->max_peaks <= 1024 && h->max_cand > 0 && h->max_cand <= 4096 &&
                     h->max_persons > 0 && h->max_persons <= 128, "capacities");
    RMPE_REQUIRE(h->stride > 0, "stride");
    const int B = h->batch, MP = h->max_peaks, MC = h->max_cand, MS = h->max_persons;
    // every blob a descriptor names must lie inside the caller's arrays (the kernels trust the offsets)
    for (int i = 0; i < B; i++) {
        const RmpeFrameDesc &f = h->frames[i];
        RMPE_REQUIRE(f.n_scales >= 1 && f.n_scales <= RMPE_MAX_SCALES, "frame descriptor: n_scales");
        for (int s = 0; s < f.n_scales; s++) {
            RMPE_REQUIRE(f.grid_h[s] > 0 && f.grid_w[s] > 0 && f.heat_offset[s] >= 0 && f.paf_offset[s] >= 0, "frame descriptor: grid / offsets");
            const size_t cells = (size_t)f.grid_h[s] * f.grid_w[s];
            RMPE_REQUIRE((size_t)f.heat_offset[s] + cells * 19 <= h->heat_elems, "frame descriptor: heat blob beyond heat_elems");
            RMPE_REQUIRE((size_t)f.paf_offset[s] + cells * 38 <= h->paf_elems, "frame descriptor: PAF blob beyond paf_elems");
        }
    }
    HostCtx *ctx = nullptr;
    int rc = host_ctx(&ctx);
    if (rc != RMPE_OK) return rc;
    Arena &A = ctx->arena;
    size_t ws_bytes = rmpe_decode_workspace_bytes(B, h->frames, MP, MC, h->stride);
    const size_t cand_b = (size_t)B * kParts * MP * 4 * 8, npk_b = (size_t)B * kParts * 4;
    const size_t conn_b = (size_t)B * kLimbs * MP * 5 * 8, nconn_b = (size_t)B * kLimbs * 4;
    const size_t lc_b = h->limb_cand ? (size_t)B * kLimbs * MC * 4 * 8 : 0;
    const size_t sub_b = (size_t)B * MS * 20 * 8;
    size_t need = Arena::need(h->heat_elems * 4) + Arena::need(h->paf_elems * 4) + Arena::need(B * sizeof(RmpeFrameDesc)) +
                  Arena::need(cand_b) + Arena::need(npk_b) + Arena::need(conn_b) + Arena::need(nconn_b) * 2 +
                  Arena::need(lc_b) + Arena::need(sub_b) + Arena::need(B * 4) * 2 + Arena::need(ws_bytes);
    rc = A.reserve(need);
    if (rc != RMPE_OK) return rc;
    A.reset();
    cudaStream_t st = ctx->stream;
    float *d_heat = (float *)A.take(h->heat_elems * 4);
    float *d_paf = (float *)A.take(h->paf_elems * 4);
    RmpeFrameDesc *d_fr = (RmpeFrameDesc *)A.take(B * sizeof(RmpeFrameDesc));
    double *d_cand = (double *)A.take(cand_b);
    int32_t *d_npk = (int32_t *)A.take(npk_b);
    double *d_conn = (double *)A.take(conn_b);
    int32_t *d_nconn = (int32_t *)A.take(nconn_b);
    int32_t *d_nlc = (int32_t *)A.take(nconn_b);
    double *d_lc = lc_b ? (double *)A.take(lc_b) : nullptr;
    double *d_sub = (double *)A.take(sub_b);
    int32_t *d_nsub = (int32_t *)A.take(B * 4);
    int32_t *d_st = (int32_t *)A.take(B * 4);
    void *d_ws = A.take(ws_bytes);
    RMPE_HOST_TRY(cudaMemcpyAsync(d_fr, h->frames, B * sizeof(RmpeFrameDesc), cudaMemcpyHostToDevice, st));
    RMPE_HOST_TRY(cudaMemsetAsync(d_npk, 0, npk_b, st));
    // Chunks of 16 or 64 frames (64 = rmpe_decode_batch's own chunk size): the blobs of chunk i + 1 cross the bus on the copy
    // stream while chunk i is decoded on the main stream -- a 1000-frame multi-scale list is 4.6 GB of blobs, 85 ms of
    // PCIe against 35 ms of kernels.  A chunk's blobs are the byte range its descriptors span (make_frames lays the
    // blobs out in frame order; overlapping ranges are merely copied twice).
    // small batches are cut finer so that there is something to overlap at all (64 single-scale frames: 4 x 16)
    const int kHostChunk = B < 256 ? 16 : 64;
    cudaStream_t cs = ctx->pipe[0];
    cudaEvent_t ev_start, ev_copied[2];
    RMPE_HOST_TRY(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) RMPE_HOST_TRY(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
    RMPE_HOST_TRY(cudaEventRecord(ev_start, st));
    RMPE_HOST_TRY(cudaStreamWaitEvent(cs, ev_start, 0));        // the arena may still be in use by this thread's previous call
    int rc_chunk = RMPE_OK;
    for (int f0 = 0, ci = 0; f0 < B; f0 += kHostChunk, ci++) {
        const int n = std::min(kHostChunk, B - f0);
        size_t h_lo = (size_t)-1, h_hi = 0, p_lo = (size_t)-1, p_hi = 0;
        for (int i = f0; i < f0 + n; i++) {
            const RmpeFrameDesc &f = h->frames[i];
            for (int s = 0; s < f.n_scales; s++) {
                const size_t cells = (size_t)f.grid_h[s] * f.grid_w[s];
                h_lo = std::min(h_lo, (size_t)f.heat_offset[s]); h_hi = std::max(h_hi, (size_t)f.heat_offset[s] + cells * 19);
                p_lo = std::min(p_lo, (size_t)f.paf_offset[s]); p_hi = std::max(p_hi, (size_t)f.paf_offset[s] + cells * 38);
            }
        }
        cudaError_t e = cudaMemcpyAsync(d_heat + h_lo, h->heat + h_lo, (h_hi - h_lo) * 4, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_paf + p_lo, h->paf + p_lo, (p_hi - p_lo) * 4, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess) e = cudaEventRecord(ev_copied[ci & 1], cs);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev_copied[ci & 1], 0);
        if (e != cudaSuccess) { set_error("decode host pipeline: %s", cudaGetErrorString(e)); rc_chunk = RMPE_E_CUDA; break; }
        RmpeDecodeBatch d;
        memset(&d, 0, sizeof(d));
        d.batch = n; d.max_peaks = MP; d.max_cand = MC; d.max_persons = MS; d.stride = h->stride; d.flags = h->flags;
        d.thre1 = h->thre1; d.thre2 = h->thre2;
        d.heat = d_heat; d.paf = d_paf; d.frames = d_fr + f0; d.frames_host = h->frames + f0;
        d.candidate = d_cand + (size_t)f0 * kParts * MP * 4; d.n_peaks = d_npk + (size_t)f0 * kParts;
        d.connections = d_conn + (size_t)f0 * kLimbs * MP * 5; d.n_conn = d_nconn + (size_t)f0 * kLimbs;
        d.limb_cand = d_lc ? d_lc + (size_t)f0 * kLimbs * MC * 4 : nullptr; d.n_limb_cand = d_nlc + (size_t)f0 * kLimbs;
        d.subset = d_sub + (size_t)f0 * MS * 20; d.n_subset = d_nsub + f0; d.status = d_st + f0;
        d.workspace = d_ws; d.workspace_bytes = ws_bytes;
        rc_chunk = rmpe_decode_batch(&d, st);
        if (rc_chunk != RMPE_OK) break;
    }
    cudaEventDestroy(ev_start);
    for (int i = 0; i < 2; i++) cudaEventDestroy(ev_copied[i]);
    if (rc_chunk != RMPE_OK) return host_fail(*ctx, rc_chunk);
    // Counts first; then only the filled prefix of every capacity-sized table comes back (a frame fills ~50 of its
    // 18 x max_peaks candidate rows): one strided copy per table, row = the largest prefix any frame / limb uses.
    std::vector<int32_t> npk((size_t)B * kParts), nconn((size_t)B * kLimbs), nlc((size_t)B * kLimbs), nsub(B);
    RMPE_HOST_TRY(cudaMemcpyAsync(npk.data(), d_npk, npk_b, cudaMemcpyDeviceToHost, st));
    RMPE_HOST_TRY(cudaMemcpyAsync(nconn.data(), d_nconn, nconn_b, cudaMemcpyDeviceToHost, st));
    if (h->limb_cand) RMPE_HOST_TRY(cudaMemcpyAsync(nlc.data(), d_nlc, nconn_b, cudaMemcpyDeviceToHost, st));
    RMPE_HOST_TRY(cudaMemcpyAsync(nsub.data(), d_nsub, B * 4, cudaMemcpyDeviceToHost, st));
    if (h->status) RMPE_HOST_TRY(cudaMemcpyAsync(h->status, d_st, B * 4, cudaMemcpyDeviceToHost, st));
    RMPE_HOST_TRY(cudaStreamSynchronize(st));
    int max_cand_rows = 0, max_conn = 0, max_lc = 0, max_sub = 0;
    for (int i = 0; i < B; i++) {
        int tot = 0;
        for (int p = 0; p < kParts; p++) tot += std::min(npk[(size_t)i * kParts + p], MP);
        max_cand_rows = std::max(max_cand_rows, tot);
        for (int k = 0; k < kLimbs; k++) {
            max_conn = std::max(max_conn, std::min(nconn[(size_t)i * kLimbs + k], MP));
            max_lc = std::max(max_lc, std::min(nlc[(size_t)i * kLimbs + k], MC));
        }
        max_sub = std::max(max_sub, std::min(nsub[i], MS));
    }
    memcpy(h->n_peaks, npk.data(), npk_b);
    if (h->n_conn) memcpy(h->n_conn, nconn.data(), nconn_b);
    if (h->n_limb_cand) {
        if (h->limb_cand) memcpy(h->n_limb_cand, nlc.data(), nconn_b);
        else memset(h->n_limb_cand, 0, nconn_b);
    }
    memcpy(h->n_subset, nsub.data(), B * 4);
    if (max_cand_rows)
        RMPE_HOST_TRY(cudaMemcpy2DAsync(h->candidate, (size_t)kParts * MP * 32, d_cand, (size_t)kParts * MP * 32, (size_t)max_cand_rows * 32,
                                        B, cudaMemcpyDeviceToHost, st));
    if (h->connections && max_conn)
        RMPE_HOST_TRY(cudaMemcpy2DAsync(h->connections, (size_t)MP * 40, d_conn, (size_t)MP * 40, (size_t)max_conn * 40, (size_t)B * kLimbs,
                                        cudaMemcpyDeviceToHost, st));
    if (h->limb_cand && max_lc)
        RMPE_HOST_TRY(cudaMemcpy2DAsync(h->limb_cand, (size_t)MC * 32, d_lc, (size_t)MC * 32, (size_t)max_lc * 32, (size_t)B * kLimbs,
                                        cudaMemcpyDeviceToHost, st));
    if (max_sub)
        RMPE_HOST_TRY(cudaMemcpy2DAsync(h->subset, (size_t)MS * 160, d_sub, (size_t)MS * 160, (size_t)max_sub * 160, B, cudaMemcpyDeviceToHost, st));
    RMPE_HOST_TRY(cudaStreamSynchronize(st));
    return RMPE_OK;
}
