/*
 * rmpe_b200.h -- C ABI of the B200-native OpenPose target-generation + decode hot path.
 *
 * The reference (GuruMulay/Adapting-RGB-Pose-Estimation-to-New-Domains) has no FFI of its own:
 * the path is reached through plain Python signatures.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference root); the Python
 * mirror of those signatures lives in adapting-rgb-pose-estimation-to-new-domains_b200/ and
 * binds this header with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every function returns RMPE_OK (0) or a negative RMPE_E_* code; rmpe_last_error() gives
 *     the message of the last failure on the calling thread.
 *   - "device" pointers are CUDA device pointers on the device passed to rmpe_init; the caller
 *     owns every buffer.  The library owns only a per-device constant area (interpolation
 *     tables) and, for the *_host entry points, a staging arena that grows on demand.
 *   - kernels are asynchronous on the given stream (a cudaStream_t / CUstream cast to void*);
 *     the *_host entry points synchronise before returning.
 *   - per-sample anomalies never become errors across the ABI: they are bits in status[].
 */
#ifndef RMPE_B200_H
#define RMPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMPE_ABI_VERSION 2   /* 2: sigma / thre and the NHWC (Keras) outputs in RmpeGtBatch[Host], stride in
                                  rmpe_decode_workspace_bytes, RMPE_ST_PERSONS_CLAMPED */

/* error codes */
#define RMPE_OK 0
#define RMPE_E_BADARG (-1)
#define RMPE_E_CUDA (-2)
#define RMPE_E_NOTINIT (-3)
#define RMPE_E_NOMEM (-4)

/* fixed geometry of the target generator: py_rmpe_server/py_rmpe_config.py:12-43 */
#define RMPE_OUT_W 368
#define RMPE_OUT_H 368
#define RMPE_STRIDE 8
#define RMPE_GRID 46
#define RMPE_NUM_PARTS 18
#define RMPE_NUM_LIMBS 19
#define RMPE_NUM_LAYERS 57 /* PAF 0..37, heat 38..55, background 56 */

/* status bits (per sample / per frame) */
#define RMPE_ST_ZERO_LIMB 0x1        /* a zero-length limb was skipped (py_rmpe_heatmapper.py:81-84) */
#define RMPE_ST_PEAK_OVERFLOW 0x2    /* more peaks of one part than max_peaks: extra peaks dropped */
#define RMPE_ST_CAND_OVERFLOW 0x4    /* more limb candidates than max_cand */
#define RMPE_ST_PERSON_OVERFLOW 0x8  /* more assembled persons than max_persons */
#define RMPE_ST_FOUND_GT2 0x10       /* the reference would raise IndexError (eval...:192-195) */
#define RMPE_ST_SINGULAR 0x20        /* affine matrix not invertible (cv2 would produce border) */
#define RMPE_ST_PERSONS_CLAMPED 0x40 /* n_persons[i] was outside [0, max_persons]: clamped, extra persons ignored */

/* ---------------------------------------------------------------------------------------- */
/* lifetime                                                                                   */
/* ---------------------------------------------------------------------------------------- */
int rmpe_init(int device);       /* idempotent; uploads the bicubic tables to `device` */
void rmpe_shutdown(void);
const char *rmpe_last_error(void);
int rmpe_abi_version(void);
int rmpe_device(void);           /* device given to rmpe_init, -1 before */

/* ---------------------------------------------------------------------------------------- */
/* T1  AugmentSelection.affine  (py_rmpe_server/py_rmpe_transformer.py:39-78)                 */
/* Host-side: n matrices from the augmentation parameters, bit-identical to the numpy chain   */
/* (cos/sin from the host libm).  M_out is n x 6 doubles, row-major 2x3.                      */
/* ---------------------------------------------------------------------------------------- */
int rmpe_aug_affine(int n, const uint8_t *flip, const double *degree, const int32_t *crop_xy,
                    const double *scale, const double *center_xy, const double *scale_self,
                    double *M_out);

/* AugmentSelection.random (py_rmpe_transformer.py:19-27): n draws from a CPython-compatible
 * Mersenne Twister seeded like random.seed(seed) (non-negative ints), same draw order:
 * flip, degree, scale condition, [scale], x_off, y_off.  Sample i uses seed seeds[i]. */
int rmpe_aug_random(int n, const uint64_t *seeds, uint8_t *flip, double *degree, int32_t *crop_xy,
                    double *scale);

/* ---------------------------------------------------------------------------------------- */
/* T2-T4 + H1-H5: Transformer.transform + Heatmapper.create_heatmaps                          */
/*   (py_rmpe_transformer.py:83-114, py_rmpe_heatmapper.py:32-138,                            */
/*    glue = RawDataIterator.transform_data, py_rmpe_data_iterator.py:68-74)                  */
/* ---------------------------------------------------------------------------------------- */
typedef struct RmpeSrcDesc {
    int64_t img_offset;   /* byte offset of this sample's BGR HWC u8 image from src_img */
    int64_t mask_offset;  /* byte offset of this sample's u8 miss-mask from src_mask */
    int32_t height;
    int32_t width;
    int32_t img_pitch;    /* bytes per image row (>= 3*width) */
    int32_t mask_pitch;   /* bytes per mask row (>= width) */
} RmpeSrcDesc;

#define RMPE_GT_IMG_CHW 0x1        /* out_img as (B,3,368,368) (RawDataIterator.gen :39) instead of HWC */
#define RMPE_GT_LABELS_F64 0x2     /* out_labels / out_mask as float64 (reference dtype) instead of float32 */
#define RMPE_GT_NO_TRANSFORM 0x4   /* joints are already in output coordinates, out_mask is an INPUT:
                                      Heatmapper.create_heatmaps(joints, mask) on its own */
#define RMPE_GT_NO_WARP 0x8        /* skip image warp (labels + mask + joints only) */
#define RMPE_GT_SIMPLE_KERNELS 0x10 /* debugging: straight-line kernels without smem staging */
#define RMPE_GT_WARP_ONLY 0x20      /* image warp only (per-kernel timing in bench.py) */
#define RMPE_GT_PAF_AVERAGE 0x40    /* NON-reference variant: PAF vectors averaged over the persons whose band covers a
                                       pixel (the lines commented out at py_rmpe_heatmapper.py:119-126) instead of the
                                       reference's last-person-wins overwrite */

typedef struct RmpeGtBatch {
    int32_t batch;
    int32_t max_persons;          /* person stride of joints / out_joints */
    int32_t flags;
    int32_t reserved;
    /* inputs (device) */
    const uint8_t *src_img;
    const uint8_t *src_mask;
    const RmpeSrcDesc *src_desc;  /* [batch] */
    const double *joints;         /* [batch][max_persons][18][3] (x, y, visibility) */
    const int32_t *n_persons;     /* [batch] */
    const double *M;              /* [batch][6] forward affine (rmpe_aug_affine) */
    const uint8_t *flip;          /* [batch] */
    /* outputs (device) */
    uint8_t *out_img;             /* [batch][368][368][3] u8 (or CHW) */
    void *out_mask;               /* [batch][46][46] f32 (f64 with RMPE_GT_LABELS_F64) = u8/255 */
    void *out_labels;             /* [batch][57][46][46] f32/f64 */
    double *out_joints;           /* [batch][max_persons][18][3] */
    int32_t *out_count;           /* optional [batch][19][46][46]: put_vector_maps' local `count` */
    int32_t *status;              /* [batch] */
    /* Heatmapper(sigma, thre) (py_rmpe_heatmapper.py:10-14); 0 selects the reference defaults 7.0 / 8.0 */
    double sigma;
    double thre;
    /* Keras-ready NHWC tensors of DataIteratorBase.gen (training/ds_generators.py:47-63), written by the rasteriser
     * itself; each may be NULL.  With any of them set, out_labels may be NULL.  Element type follows
     * RMPE_GT_LABELS_F64. */
    void *out_vec_label;          /* y1 [batch][46][46][38] = labels[0:38] */
    void *out_heat_label;         /* y2 [batch][46][46][19] = labels[38:57] */
    void *out_vec_weights;        /* x1 [batch][46][46][38] = mask repeated */
    void *out_heat_weights;       /* x2 [batch][46][46][19] */
} RmpeGtBatch;

int rmpe_gt_batch(const RmpeGtBatch *b, void *stream);

/* Same path with HOST buffers (the call the Python drop-in classes make).  Source images must
 * share one (height,width).  Copies are inside the call: cudaMemcpyAsync straight between the caller's
 * buffers and a device arena owned by the calling thread (no host staging copy -- pin the buffers to let
 * the chunk pipeline overlap copies and kernels; pageable buffers work but serialise).  Any output
 * pointer may be NULL to skip its read-back.  Calls from different host threads do not serialise: each
 * thread has its own arena and streams. */
typedef struct RmpeGtBatchHost {
    int32_t batch;
    int32_t max_persons;
    int32_t flags;
    int32_t src_height;
    int32_t src_width;
    int32_t reserved;
    const uint8_t *src_img;       /* [batch][H][W][3] */
    const uint8_t *src_mask;      /* [batch][H][W] */
    const double *joints;
    const int32_t *n_persons;
    const double *M;
    const uint8_t *flip;
    uint8_t *out_img;
    void *out_mask;
    void *out_labels;
    double *out_joints;
    int32_t *out_count;
    int32_t *status;
    double sigma;                 /* as in RmpeGtBatch */
    double thre;
    void *out_vec_label;
    void *out_heat_label;
    void *out_vec_weights;
    void *out_heat_weights;
} RmpeGtBatchHost;

int rmpe_gt_batch_host(const RmpeGtBatchHost *b);

/* ---------------------------------------------------------------------------------------- */
/* Batch assembly of DataIteratorBase.gen (training/ds_generators.py:31-106): the step right     */
/* after the GT path.  Turns the planar labels / mask of rmpe_gt_batch into the Keras-ready NHWC  */
/* tensors  x1 = mask repeated to 38 channels, x2 = mask repeated to 19, y1 = labels[0:38] as     */
/* (46,46,38), y2 = labels[38:57] as (46,46,19)  (reference :52-63).  Any output may be NULL.     */
/* ---------------------------------------------------------------------------------------- */
typedef struct RmpeKerasBatch {
    int32_t batch;
    int32_t flags;            /* RMPE_GT_LABELS_F64: all six arrays are float64 instead of float32 */
    const void *labels;       /* [batch][57][46][46] */
    const void *mask;         /* [batch][46][46] */
    void *vec_weights;        /* x1 [batch][46][46][38] */
    void *heat_weights;       /* x2 [batch][46][46][19] */
    void *vec_label;          /* y1 [batch][46][46][38] */
    void *heat_label;         /* y2 [batch][46][46][19] */
} RmpeKerasBatch;

int rmpe_keras_batch(const RmpeKerasBatch *b, void *stream);      /* device pointers, asynchronous */
int rmpe_keras_batch_host(const RmpeKerasBatch *b);               /* host pointers, synchronous */

/* ---------------------------------------------------------------------------------------- */
/* D1-D6: process_single_scale / process_multi_scale from the network blobs on               */
/*   (eval/eval_coco2014_multi_modes.py:263-415 and :58-231)                                  */
/* ---------------------------------------------------------------------------------------- */
#define RMPE_MAX_SCALES 4

typedef struct RmpeFrameDesc {
    int32_t height;                 /* oriImg.shape[0] */
    int32_t width;                  /* oriImg.shape[1] */
    int32_t n_scales;               /* 1 = process_single_scale; 2..4 = process_multi_scale */
    int32_t reserved;
    int32_t grid_h[RMPE_MAX_SCALES];   /* blob rows per scale */
    int32_t grid_w[RMPE_MAX_SCALES];
    int32_t pad_down[RMPE_MAX_SCALES]; /* padRightDownCorner pad[2], pad[3] (util.py:57-77); */
    int32_t pad_right[RMPE_MAX_SCALES];/* ignored when n_scales == 1 */
    int64_t heat_offset[RMPE_MAX_SCALES]; /* element offsets into `heat` / `paf` */
    int64_t paf_offset[RMPE_MAX_SCALES];
} RmpeFrameDesc;

/* RmpeDecodeBatch.flags */
#define RMPE_DECODE_REUSE_TABLES 0x1 /* the workspace still holds the up-sampling/smoothing operator tables of a
                                        previous call with the same frames, capacities and workspace: skip rebuilding
                                        them (they depend on frame geometry only).  Honoured when the batch is processed
                                        as one chunk (<= 64 frames that fit the workspace); larger batches reuse the
                                        table region per chunk and always rebuild */

typedef struct RmpeDecodeBatch {
    int32_t batch;
    int32_t max_peaks;     /* capacity per part (<= 1024) */
    int32_t max_cand;      /* capacity of limb candidates per limb (<= 4096) */
    int32_t max_persons;   /* capacity of subset rows (<= 128) */
    int32_t stride;        /* model_params['stride'] = 8 */
    int32_t flags;
    double thre1;          /* params['thre1'] */
    double thre2;          /* params['thre2'] */
    /* inputs (device) */
    const float *heat;     /* NHWC (h,w,19) blobs, one per frame and scale */
    const float *paf;      /* NHWC (h,w,38) */
    const RmpeFrameDesc *frames; /* [batch], DEVICE memory */
    const RmpeFrameDesc *frames_host; /* the same array in host memory (launch geometry) */
    /* outputs (device) */
    double *candidate;     /* [batch][18*max_peaks][4]  x, y, score, id  (compact, id order) */
    int32_t *n_peaks;      /* [batch][18] */
    double *connections;   /* [batch][19][max_peaks][5]  idA, idB, score, i, j */
    int32_t *n_conn;       /* [batch][19]  (-1 = special_k: a part with no peaks) */
    double *limb_cand;     /* optional [batch][19][max_cand][4]  i, j, score, score+sA+sB (unsorted) */
    int32_t *n_limb_cand;  /* [batch][19] */
    double *subset;        /* [batch][max_persons][20] */
    int32_t *n_subset;     /* [batch] */
    int32_t *status;       /* [batch] */
    /* caller-provided scratch (device), >= rmpe_decode_workspace_bytes() */
    void *workspace;
    size_t workspace_bytes;
} RmpeDecodeBatch;

size_t rmpe_decode_workspace_bytes(int batch, const RmpeFrameDesc *frames_host, int max_peaks,
                                   int max_cand, int stride);
int rmpe_decode_batch(const RmpeDecodeBatch *b, void *stream);

/* Host-buffer variant: blobs and results in host memory; frames is a host array. */
typedef struct RmpeDecodeBatchHost {
    int32_t batch;
    int32_t max_peaks;
    int32_t max_cand;
    int32_t max_persons;
    int32_t stride;
    int32_t flags;
    double thre1;
    double thre2;
    const float *heat;
    const float *paf;
    size_t heat_elems;     /* total floats behind `heat` / `paf` */
    size_t paf_elems;
    const RmpeFrameDesc *frames;
    double *candidate;
    int32_t *n_peaks;
    double *connections;
    int32_t *n_conn;
    double *limb_cand;
    int32_t *n_limb_cand;
    double *subset;
    int32_t *n_subset;
    int32_t *status;
} RmpeDecodeBatchHost;

int rmpe_decode_batch_host(const RmpeDecodeBatchHost *b);

/* debugging / stage-parity hooks (model stride 8 only): materialise D1 (upsampled or scale-averaged heat map, planar
 * [18][H][W], float32 for n_scales==1 else float64) and D2's smoothed map for one frame. */
int rmpe_debug_heat_maps(const RmpeFrameDesc *frame_host, const float *heat_dev, void *up_out_dev,
                         void *smooth_out_dev, void *stream);
/* evaluate the up-sampled PAF at integer points (n x (c,y,x) int32 triples) for one frame */
int rmpe_debug_paf_points(const RmpeFrameDesc *frame_host, const float *paf_dev, int n,
                          const int32_t *cyx_dev, double *out_dev, void *stream);

/* D5/D6 alone (eval...:364-415) on caller-made connection lists of one frame, all pointers DEVICE memory laid out as
 * in RmpeDecodeBatch with batch = 1.  Test hook for RMPE_ST_FOUND_GT2: connection lists made by the library itself are
 * one-to-one per limb and cannot produce the reference's IndexError. */
int rmpe_debug_assemble(int max_peaks, int max_persons, const double *candidate_dev, const double *connections_dev,
                        const int32_t *n_conn_dev, const int32_t *n_peaks_dev, double *subset_dev,
                        int32_t *n_subset_dev, int32_t *status_dev, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* U1  util.padRightDownCorner (util.py:57-77) on a device HWC u8 image                       */
/* ---------------------------------------------------------------------------------------- */
int rmpe_pad_right_down_corner(const uint8_t *src_dev, int height, int width, int channels,
                               int stride, int pad_value, uint8_t *dst_dev, int *pad4_out,
                               void *stream);

/* kernels launched by this library since rmpe_init (bench.py's gpu_launches) */
int64_t rmpe_launch_count(void);

/* Per-kernel device time (bench.py's roofline leg).  While enabled every kernel launch of this
 * library is bracketed by two CUDA events on its own stream; rmpe_profile_get() waits for the
 * recorded events and returns the accumulated milliseconds and launch count of kernel `index`
 * (0 <= index < rmpe_profile_count()).  Off by default; the reference's counterpart is the
 * time() prints of py_rmpe_server/rmpe_server.py:64-78. */
int rmpe_profile_enable(int on);
int rmpe_profile_reset(void);
int rmpe_profile_count(void);
int rmpe_profile_get(int index, char *name_out, int name_cap, double *total_ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* RMPE_B200_H */
