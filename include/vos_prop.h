/* vos_prop.h -- C ABI of the B200 label-propagation engine (libvosprop.so).
 *
 * Drop-in scope: the transductive label-propagation hot path of hynekdav/semi-supervised-VOS.
 * The reference has no FFI; its boundary is a set of Python callables.
 * Each entry point below replaces one of them (reference file:line in the comment).  The Python
 * mirror under semi-supervised-vos_b200/src/ binds these with ctypes.
 *
 * Conventions: plain C, raw device pointers, cudaStream_t passed as void*, every call is
 * stream-ordered and never synchronises the device; int return (0 = ok, <0 = error, message via
 * vosprop_last_error()); no exceptions cross the ABI; a handle is not thread-safe.
 * There is NO CPU fallback: on a machine without an sm_100 device vosprop_create() fails.
 */
#ifndef VOS_PROP_H_
#define VOS_PROP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOSPROP_ABI_VERSION 4
#define VOSPROP_MAX_REFS 32     /* reference frames per step (reference default ref_num = 9)   */
#define VOSPROP_MAX_CLASSES 24  /* d = objects + 1; 15..24 (validation: 22 annotation centroids, src/train.py:206) run with index
                                   labels only: no dense labels, no probability propagation, no top-k, W_d >= 32   */
#define VOSPROP_MAX_DENSE_CLASSES 14  /* DAVIS <= 11, YouTube-VOS <= 11: every label kind and kernel                 */
#define VOSPROP_FEAT_DIM 256    /* VOSNet embedding width, src/model/vos_net.py:22             */
#define VOSPROP_TILE 128        /* pixel tile of the affinity kernel                           */
#define VOSPROP_MAX_TOPK 64     /* largest k of the top-k extension                            */

enum vosprop_status {
    VOSPROP_OK = 0,
    VOSPROP_ERR_INVALID = -1,   /* bad argument                                                */
    VOSPROP_ERR_CUDA = -2,      /* CUDA runtime / driver error (string has the detail)         */
    VOSPROP_ERR_UNSUPPORTED = -3, /* no sm_100 device / shape beyond engine capacity           */
    VOSPROP_ERR_STATE = -4      /* call order violated (e.g. reference frame not in the ring)  */
};

enum vosprop_dtype { VOSPROP_F32 = 0, VOSPROP_F16 = 1, VOSPROP_BF16 = 2 };
/* How the reference memory stores embeddings, fixed per video by vosprop_reset():
 *   SPLIT3 : any input dtype; x = hi + lo (two bf16), S = Qhi.Rhi + Qlo.Rhi + Qhi.Rlo on the tensor cores
 *            (3 passes, fp32 accumulate) -- fp32-grade logits from fp32 embeddings.
 *   F16    : embeddings must arrive as fp16 -- what VOSNet emits under CUDA autocast, the reference's own
 *            GPU path (src/utils/inference_utils.py:35,52-53).  The product of two fp16 values is exact in
 *            the fp32 accumulator, so ONE tensor-core pass equals the fp32 contraction of those embeddings.
 *   BF16   : same for bf16 embeddings.
 * F16/BF16 refuse an append of any other dtype (it would round the embeddings). */
enum vosprop_precision { VOSPROP_PREC_SPLIT3 = 0, VOSPROP_PREC_F16 = 1, VOSPROP_PREC_BF16 = 2 };
/* memory order of a feature map handed to vosprop_append_features */
enum vosprop_layout { VOSPROP_NCHW = 0 /* (K, H_d*W_d), torch default */, VOSPROP_NHWC = 1 /* (H_d*W_d, K) */ };
enum vosprop_kernel {
    VOSPROP_KERNEL_TC = 0,   /* product path: TMA + tcgen05 bf16x3 fused affinity kernel; picks the
                                index-label variant (target tile in TMEM, analytic prior, class bytes)
                                when every reference carries index labels and W_d >= 32, else the
                                general variant                                                  */
    VOSPROP_KERNEL_SIMT = 1, /* on-device fp32 checker (CUDA cores), same data path; tests only */
    VOSPROP_KERNEL_TC_DENSE = 2 /* force the general tensor-core variant (dense label records)  */
};

typedef struct vosprop_engine vosprop_engine;

typedef struct vosprop_config {
    int32_t device;         /* CUDA ordinal                                                     */
    int32_t max_pixels;     /* largest H_d*W_d this engine will see (480p: 6420, 1080p: 32400)  */
    int32_t ring_slots;     /* reference-memory ring size; >= frame_range + 4 + 1 (44+1)        */
    int32_t max_fullres_pixels; /* largest H*W (for the full-resolution mask staging buffer)    */
} vosprop_config;

/* One propagation step == one call of the reference's predict() (src/model/predict.py:19-71)
 * plus the label / mask write-back of inference_single (src/utils/inference_utils.py:67-75). */
typedef struct vosprop_step {
    int32_t frame_idx;                       /* target frame; its features must be in the ring  */
    int32_t n_refs;                          /* number of reference frames                      */
    int32_t ref_frames[VOSPROP_MAX_REFS];    /* frame indices (duplicates allowed, as index_select) */
    float ref_sigma[VOSPROP_MAX_REFS];       /* Gaussian prior sigma per ref; <= 0: no prior    */
    float temperature;                       /* multiplies the logits (predict.py:52); >= 0     */
    int32_t probability_propagation;         /* 1: store raw prediction as the new label (inference_utils.py:67-68) */
    int32_t write_labels;                    /* 1: write the new label into the ring (normal); 0: pure predict()  */
    int32_t topk;                            /* 0: full softmax (the reference).  1..VOSPROP_MAX_TOPK: EXTENSION -- the softmax of
                                                predict.py:55 runs over the k largest logit*temperature per target pixel only
                                                (ties -> lowest reference index); prior and label gather unchanged            */
    int32_t kernel;                          /* enum vosprop_kernel                             */
    float* out_prediction;                   /* device (d, P) fp32 or NULL  -- predict()'s return value */
    uint8_t* out_mask_lowres;                /* device (P) uint8 or NULL    -- argmax over d at stride 8 */
    uint8_t* out_mask_fullres;               /* device (H, W) uint8 or NULL -- nearest up-sample + argmax (inference_utils.py:74-75) */
    int32_t* out_topk_idx;                   /* device (P, topk) int32 or NULL (top-k mode only): reference indices r*P + pixel
                                                (r = position in ref_frames), best first; -1 where fewer than k exist         */
    /* Several sequences in flight on one GPU (one engine + one stream each): the fused affinity kernel fills every SM,
     * so the affinity kernels of different sequences cannot overlap each other.  Chaining them with these two events
     * keeps them back to back on the device while the launch latencies, appends and merges of the other sequences
     * fill the gaps between them.  Both may be NULL. */
    void* wait_event;                        /* cudaEvent_t: the stream waits for it before the affinity kernel  */
    void* record_event;                      /* cudaEvent_t: recorded right after the affinity kernel            */
} vosprop_step;

const char* vosprop_last_error(void);
int vosprop_abi_version(void);

/* Engine lifetime (one per sequence in flight; owns ring buffer, TMA descriptors, scratch). */
int vosprop_create(const vosprop_config* cfg, vosprop_engine** out);
void vosprop_destroy(vosprop_engine* e);

/* New video: geometry + class count.  Replaces the state reset at a video boundary
 * (src/utils/inference_utils.py:28-48) and the prior set-up of prepare_first_frame
 * (src/model/predict.py:117-118: the (P,P) Gaussian matrices are never built; the kernel
 * evaluates the closed form).  H_d, W_d: feature-map size; H, W: full frame size; d: classes;
 * precision: storage / tensor-core mode of the reference memory for this video. */
int vosprop_reset(vosprop_engine* e, int32_t H_d, int32_t W_d, int32_t H, int32_t W, int32_t d,
                  int32_t precision /* enum vosprop_precision */, void* stream);

/* Append frame `frame_idx`'s embedding to the ring (slot = frame_idx % ring_slots).  Replaces
 * `feats_history = torch.cat(...)` (inference_utils.py:72, :36).  `features`: device pointer to
 * K x P (NCHW) or P x K (NHWC) elements of `dtype`. */
int vosprop_append_features(vosprop_engine* e, int32_t frame_idx, const void* features,
                            int32_t dtype, int32_t layout, void* stream);

/* Append `n_frames` consecutive frames first_frame_idx .. +n_frames-1 in ONE kernel launch: `features` holds n_frames
 * maps back to back (each K x P or P x K elements of `dtype`); frame first_frame_idx + i goes to ring slot
 * (first_frame_idx + i) % ring_slots.  When `class_idx` is not NULL it holds n_frames x P class bytes and every appended
 * frame also gets its index labels.  Same effect as n_frames x (vosprop_append_features [+ vosprop_set_labels_index]).
 * Two callers: a whole labelled clip installed at once -- the reference frames of a validation clip (src/train.py:181-207:
 * ref = features[:, 0:num_frames-1] with their annotations) -- and the inference loop, which appends the frames of a
 * backbone batch AHEAD of the frame being propagated (a ring of 45 + n slots keeps every frame sample_frames can still
 * pick, predict.py:74-89): one 3.3 MB copy per frame is launch-latency-bound (6-8 us at 0.8 TB/s), 19 frames in one
 * launch run at 2.9 TB/s and leave the per-frame chain affinity -> merge -> affinity. */
int vosprop_append_frames(vosprop_engine* e, int32_t first_frame_idx, int32_t n_frames, const void* features,
                          int32_t dtype, int32_t layout, const uint8_t* class_idx, void* stream);

/* Set frame `frame_idx`'s labels from a class-index map (device, P x uint8): the one-hot first
 * frame of get_labels (predict.py:92-96).  */
int vosprop_set_labels_index(vosprop_engine* e, int32_t frame_idx, const uint8_t* class_idx, void* stream);
/* Set frame `frame_idx`'s labels from a (d, P) fp32 device array (probability propagation or an
 * externally supplied history, label_history[:, t]). */
int vosprop_set_labels_dense(vosprop_engine* e, int32_t frame_idx, const float* labels, void* stream);

/* The hot path: fused affinity + softmax + prior + label gather, then merge / argmax /
 * ring-buffer label update / mask write-back.  3 kernel launches, no host sync. */
int vosprop_propagate(vosprop_engine* e, const vosprop_step* step, void* stream);

/* sample_frames (src/model/predict.py:74-89), host only.  Writes up to VOSPROP_MAX_REFS indices,
 * returns the count, or VOSPROP_ERR_INVALID where the reference raises (num_refs < 3 once
 * frame_idx > num_refs -> negative linspace count). */
int vosprop_sample_frames(int32_t frame_idx, int32_t take_range, int32_t num_refs, int32_t* out_idx);

/* Convenience: sample_frames + the sigma-per-ref rule of predict.py:60-66, filled into `step`
 * (frame_idx, n_refs, ref_frames, ref_sigma).  sigma <= 0 disables the prior (probability mode). */
int vosprop_plan_step(int32_t frame_idx, int32_t take_range, int32_t num_refs, float sigma_dense,
                      float sigma_sparse, int32_t probability_propagation, vosprop_step* step);

/* Introspection for tests / bench. */
int vosprop_ring_slots(const vosprop_engine* e);
int vosprop_num_sms(const vosprop_engine* e);
/* Input normalisation on the device (no engine state): `n` decoded frames, uint8, pixel-interleaved RGB (n, H, W, 3) as a
 * JPEG decoder leaves them -> (x / 255 - mean[c]) / std[c] in fp32, in the reference's order of operations
 * (torchvision ToTensor + Normalize, src/utils/datasets.py:128-131, :147), stored as VOSPROP_F32 or VOSPROP_F16 in the same
 * element order, i.e. an (n, 3, H, W) tensor in channels-last layout -- what the cuDNN backbone consumes.  The fp32
 * result is bit-identical to the reference's; the fp16 result is that value rounded to nearest, which is what autocast
 * feeds the first convolution.  Replaces ~7 ms of host arithmetic and a 4x larger host-to-device copy per 480p frame. */
int vosprop_normalize_u8(const uint8_t* rgb, int64_t n_pixels /* n * H * W */, const float* mean3, const float* std3,
                         void* out, int32_t out_dtype /* enum vosprop_dtype: F32 or F16 */, void* stream);

/* Work decomposition of the affinity kernel (host arithmetic shared with the device code):
 * fills grid size and per-CTA [begin,end) of the linearised (m_tile, n_tile) space. */
int vosprop_debug_decompose(int32_t n_pixels, int32_t n_refs, int32_t num_sms, int32_t* grid,
                            int64_t* cta_begin /* num_sms+1 entries or NULL */, int32_t* max_segments);
/* Block skipping in the fused index-label kernel: mode 0 never, 1 always, 2 auto (the default).  A 32 x 32 block of the affinity matrix whose logits all
 * lie more than 127 (log2 units) below a lower bound of the final maximum of their rows weighs less than 2^-127 of the
 * row's soft-max denominator per element -- at most ~1e-34 of a row's mass over all skipped blocks, 26 orders of magnitude
 * below fp32 resolution -- and is left out.  With peaked embeddings (|f|^2 ~ 256) about two thirds of the blocks go and
 * the launch gets 10-17 % shorter (reference tiles are then visited in a strided order that spreads the live tiles over
 * the CTAs).  It costs 6-9 % when nothing can be skipped, hence auto: a launch of the skipping kernel
 * reports how many blocks it left out into host-mapped memory, the first launches of an engine and 2 of every 256 later
 * ones run it as a probe, and the host switches it on at >= 5 % skipped blocks and off below 2 % -- reading the reports as
 * they arrive, never synchronising.  Results do not depend on the mode beyond fp32 rounding (2e-6, tile order). */
int vosprop_block_skip(vosprop_engine* e, int32_t mode);
/* 1 if the next launch of the fused index kernel would skip (mode 1, or auto mode after a report of >= 5 % dead blocks), else 0. */
int vosprop_block_skip_state(const vosprop_engine* e);
/* Development aid for profiling: a bit mask that switches off parts of the fused epilogue (bit 0: all per-step
 * arithmetic, 1: label gather, 2: prior, 3: running-max update, 4: prior and packed math).  Results are WRONG
 * with any bit set; production code never calls this (flags start at 0). */
int vosprop_debug_flags(vosprop_engine* e, int32_t flags);
/* Development aid: device buffer of num_sms*16 int64 that the fused kernel fills with cycle counters of its role
 * warps (time in each mbarrier wait); NULL (the default) switches the instrumentation off. */
int vosprop_debug_clocks(vosprop_engine* e, void* device_buffer);
/* kernel launches issued by this handle since creation (for bench.py's gpu_launches) */
int64_t vosprop_launch_count(const vosprop_engine* e);

/* Per-kernel device timing for bench.py's roofline: when enabled, every append / affinity / merge
 * launch is bracketed by CUDA events recorded on the launching stream (capacity = launches kept).
 * vosprop_timing_read() synchronises on the recorded events, returns the summed milliseconds and
 * launch counts per kernel class since the last read, and clears the log.
 * totals_ms[3] / counts[3] are indexed by enum vosprop_timed_kernel. */
enum vosprop_timed_kernel { VOSPROP_T_APPEND = 0, VOSPROP_T_AFFINITY = 1, VOSPROP_T_MERGE = 2 };
int vosprop_timing_enable(vosprop_engine* e, int32_t capacity);
/* Which kernel classes get event pairs: bit (1 << vosprop_timed_kernel); default all three.  Event records cost a few
 * microseconds of front-end time per frame, so bench.py times only the dominant kernel inside its timed region. */
int vosprop_timing_select(vosprop_engine* e, int32_t class_mask);
int vosprop_timing_read(vosprop_engine* e, double* totals_ms, int64_t* counts);

#ifdef __cplusplus
}
#endif
#endif /* VOS_PROP_H_ */
