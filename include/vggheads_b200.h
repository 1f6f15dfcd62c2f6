/* vggheads_b200 - C ABI of the B200-native VGGHeads inference hot path.
 *
 * The reference (KupynOrest/head_detector) has no FFI layer; its seam is three Python callables
 * (SURVEY.md 8b).  Each entry point below names the reference interface it replaces so a
 * maintainer can bind it from head_detector/detector.py (see INTEGRATION.md for the ctypes stub):
 *
 *   vgh_letterbox             <- `HeadDetector._transform_image` head_detector/detector.py:40-52
 *                                (cv2.resize INTER_LANCZOS4 + cv2.copyMakeBorder, bit-exact)
 *   vgh_detector_forward      <- `self.model(image)`            head_detector/detector.py:58-59
 *                                (TorchScript YoloHeads_L: yolo_head_training/yolo_head/
 *                                 yolo_head_ndfl_heads.py:117-175, yolo_head_dfl_head.py:141-186)
 *   vgh_select_nms            <- `utils.nms(...)`               head_detector/utils.py:159-194
 *                                (batched twin: yolo_heads_post_prediction_callback.py:55-97)
 *   vgh_flame_decode          <- `reproject_spatial_vertices`   head_detector/flame.py:179-208
 *                                + FLAMELayer.forward            head_detector/flame.py:122-169
 *                                + vertex un-letterboxing        head_detector/detector.py:66-69
 *   vgh_detector_postprocess  <- `HeadDetector._postprocess`    head_detector/detector.py:92-95
 *   vgh_detector_run_host     <- `HeadDetector.__call__` minus host glue, batched; HOST buffers
 *
 * Conventions: plain C, no exceptions across the boundary; every function returns 0 on success or a
 * non-zero status, and vgh_last_error() returns a thread-local message.  "dev" pointers are CUDA
 * device pointers owned by the caller unless stated; streams are cudaStream_t passed as void*.
 * Handles are immutable after creation except for their internal work buffers: use one handle per
 * stream.  There is NO CPU fallback: every entry point needs a CUDA device (sm_100a).
 */
#ifndef VGGHEADS_B200_H
#define VGGHEADS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGH_NUM_VERTS 5023
#define VGH_NUM_PARAMS 413 /* [shape300|expr100|jaw3|rot6d6|transl3|scale1], head_info.py:12-21,54-78 */

typedef struct vgh_flame vgh_flame;
typedef struct vgh_detector vgh_detector;

int vgh_version(void);
const char* vgh_last_error(void);

/* ---------------------------------------------------------------------------------- FLAME decode */
/* Constants exactly as FLAMELayer registers them (flame.py:43-95), fp32 HOST arrays:
 * v_template[5023*3], shapedirs[5023*3*400], posedirs[36*15069], j_regressor[5*5023],
 * lbs_weights[5023*5].  Uploaded (re-laid-out) once. */
int vgh_flame_create(const float* v_template, const float* shapedirs, const float* posedirs, const float* j_regressor,
                     const float* lbs_weights, vgh_flame** out);
void vgh_flame_destroy(vgh_flame* f);
/* params_dev [n,413]; n_shape_live / n_expr_live: leading shape / expression coefficients that may
 * be non-zero (300/100 = general; 128/64 for rows emitted by the network).  xform_dev: optional
 * [n,3] = (pad_x, pad_y, letterbox_scale) applied as detector.py:67-69 (NULL = (0,0,1)).
 * Outputs (device, fp32): verts_dev [n,5023,3] model-space vertices incl. +0.05 z (may be NULL),
 * rot_dev [n,9] row-major rotation (may be NULL), proj_dev [n,5023,3] = ((R v) * max(s,1e-8) + t
 * - pad) / scale.  n == 0 is a no-op (flame.py:186-189). */
int vgh_flame_decode(const vgh_flame* f, const float* params_dev, int n, int n_shape_live, int n_expr_live,
                     const float* xform_dev, float* verts_dev, float* rot_dev, float* proj_dev, void* stream);

/* ---------------------------------------------------------------------------------- select + NMS */
/* boxes_dev [B,A,4] xyxy, scores_dev [B,A] (any sign).  Per image: score >= conf_thr, best top_k
 * by score (top_k <= 1024: the candidate set lives in shared memory; the reference's only call site uses 1000,
 * utils.py:163-166), greedy NMS (suppress iff IoU > iou_thr), first keep_k survivors in descending-score order.  keep_idx_dev [B,keep_k] int32 ORIGINAL anchor ids (-1 padded),
 * keep_cnt_dev [B] int32; keep_boxes_dev [B,keep_k,4] / keep_scores_dev [B,keep_k] optional. */
int vgh_select_nms(const float* boxes_dev, const float* scores_dev, int B, int A, float conf_thr, float iou_thr,
                   int top_k, int keep_k, int32_t* keep_idx_dev, int32_t* keep_cnt_dev, float* keep_boxes_dev,
                   float* keep_scores_dev, void* stream);

/* ---------------------------------------------------------------------------------- letterbox */
/* HeadDetector._transform_image (detector.py:40-52) for n RGB uint8 images of different sizes in one
 * launch.  src_dev: the images packed back to back in DEVICE memory, image i starting at byte
 * offsets[i] as [heights[i], widths[i], 3] (no row padding).  offsets / heights / widths are HOST
 * arrays.  Per image: new size = longest side -> image_size with the reference's int() truncation
 * (detector.py:42-45), OpenCV's fixed-point 8-tap Lanczos-4 resize (bit-exact, border replicated),
 * centred in the square (pad_w//2 left, pad_h//2 top) on a border of (127,0,0) - what cv2.copyMakeBorder
 * makes of the reference's scalar `value=127`.  out_dev: uint8
 * [n,image_size,image_size,3], directly consumable by vgh_detector_forward.  xform_host (optional,
 * HOST) [n,3] = (pad_x, pad_y, scale) with scale = image_size / max(h, w) - the `cache` of
 * detector.py:52-56, in the layout vgh_detector_postprocess takes.  An image whose resized extent
 * would be empty (cv2.resize raises there) fails with a non-zero status. */
int vgh_letterbox(const uint8_t* src_dev, const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n,
                  int image_size, uint8_t* out_dev, float* xform_host, void* stream);

/* ---------------------------------------------------------------------------------- mesh consumers */
/* `PredictionResult.get_pncc()`: PNCCProcessor.__call__ (head_detector/pncc_processor.py:66-73) over the CPU z-buffer
 * rasteriser Sim3DR `_rasterize` (head_detector/Sim3DR/lib/rasterize_kernel.cpp:219-293), for all heads of an image in two
 * launches, bit-identical.  verts_dev [n,5023,3] image-space vertices as HeadMetadata.vertices_3d holds them (the kernel
 * uses depth = -z: the reference flips z in place before rasterising); tris_dev int32 [ntri,3]; colors_dev float
 * [5023,3] in [0,1]; image_dev uint8 [H,W,3] painted in place (zero it for the reference's output); keys_dev: H*W
 * uint64 workspace. */
int vgh_pncc_render(const float* verts_dev, int n, const int32_t* tris_dev, int ntri, const float* colors_dev, int H, int W,
                    uint8_t* image_dev, uint64_t* keys_dev, void* stream);
/* `refined_head_bbox` (head_detector/utils.py:26-35) for n heads: int-truncated min / max of x, y over the vertex subset
 * idx_dev int32 [n_idx] -> out_xywh_dev int32 [n,4] = (x, y, x1 - x, y1 - y). */
int vgh_head_bbox(const float* verts_dev, int n, const int32_t* idx_dev, int n_idx, int32_t* out_xywh_dev, void* stream);

/* ---------------------------------------------------------------------------------- conv network */
/* Execution plan of the deploy-form network, produced by head_detector_b200/arch.py from the
 * reference's arch yaml (yolo_heads_l_arch_params.yaml).  Activations are NHWC bf16 buffers;
 * every op reads/writes a channel slice so that no concat is ever materialised. */
typedef struct {
  int32_t H, W, C;   /* spatial size and channels per pixel */
  int32_t fp32;      /* 0 = bf16, 1 = fp32 (raw head outputs) */
  int32_t stack;     /* 0 = batched [B,H,W,C]; 1 = ONE stacked image [H,W,C] (survivor patches, sparse heads) */
} vgh_buf_desc;

/* VGH_OP_STEM: im2col of the uint8 image for the 3x3 stride-2 stem -> bf16 [B,S/2,S/2,32] (27 taps in
 * (ky,kx,c) order + 5 zeros); the stem itself is then a VGH_OP_CONV with cin = 32.
 * Sparse heads (n_dense_ops < n_ops): the FLAME branch of a head level is only read at the anchors that survive
 * NMS and its receptive field there is 7x7 pixels, so its ops run AFTER select/NMS on 8x8 windows gathered around
 * the survivors (VGH_OP_PATCH_GATHER), stacked as one tall image per level; VGH_OP_PATCH_MASK zeroes the window
 * pixels outside the feature map (the dense graph's conv padding).  Such ops carry level = 1 + head level.
 * VGH_OP_STEM_CONV: the whole stem in one kernel (uint8 image -> 3x3 stride-2 conv, K = 27 padded to 32, + bias + ReLU ->
 * bf16 with 64 stored channels); carries the weight fields of a conv op (n_pad rows of k_total = 32). */
enum { VGH_OP_STEM = 0, VGH_OP_CONV = 1, VGH_OP_SPP = 2, VGH_OP_PATCH_GATHER = 3, VGH_OP_PATCH_MASK = 4, VGH_OP_STEM_CONV = 5 };

typedef struct {
  int32_t kind;
  int32_t in_buf, in_coff, cin;      /* cin: channels per tap (multiple of 32) */
  int32_t out_buf, out_coff, cout;   /* cout: stored channels (multiple of 16) */
  int32_t ksize, stride;             /* 1 or 3; stride 1 or 2 */
  int32_t relu, up;                  /* up=1: 2x2 stride-2 transpose conv, cout = 4*up_cout */
  int32_t up_cout;
  int32_t res_buf, res_coff;         /* residual source (res_buf < 0: none) */
  float res_alpha;
  int32_t n_pad, k_total, block_n;   /* packed weight matrix [n_pad][k_total] bf16, UMMA N */
  int64_t w_off, b_off;              /* element offsets into the weight / bias blobs */
  int32_t lane, level;               /* independent graph branches run on separate lanes (streams), 0 = main; level: see above */
} vgh_op_desc;

typedef struct {
  int32_t batch, image_size;
  int32_t n_bufs, n_ops;
  const vgh_buf_desc* bufs;
  const vgh_op_desc* ops;
  const uint16_t* weights_host;  /* bf16 (or, with act_f16, fp16) bits, all packed conv weights */
  int64_t n_weights;
  const float* bias_host;
  int64_t n_bias;
  int32_t reg_buf[3], flame_buf[3]; /* raw head output buffers per level (fp32) */
  int32_t keep_k;                /* capacity of survivors per image (keep_top_k, utils.py:166) */
  int32_t n_dense_ops;           /* ops [0, n_dense_ops) run before select/NMS, the rest after it on survivor patches;
                                    0 or n_ops = single-phase (dense) plan */
  int32_t split;                 /* parity mode (fp32-class arithmetic for end-to-end checks against the reference's fp32 path):
                                    every bf16 activation y is stored as three bf16 terms h + m + l == y in six planes
                                    [h|m|h|m|h|l] per 32-channel granule (buffers carry 6x the channels) and weights are packed
                                    [w_h|w_h|w_m|w_m|w_l|w_h] along K, so that the bf16 MMAs accumulate the six leading partial
                                    products in fp32 (relative error ~2^-22 per product instead of 2^-9).  Dense plan only. */
  int32_t act_f16;               /* 1: the 16-bit activation buffers and `weights_host` are IEEE fp16 (11 significant bits: the
                                    precision class of the TF32 convs the reference runs on a GPU by default, and the format the
                                    released model was trained in - AMP, yolo_heads_l.yaml) instead of bf16 (8 bits); same kernels,
                                    same tensor-core rate (tcgen05 kind::f16 takes either).  0: bf16.  Must be 0 with `split`. */
} vgh_net_desc;

int vgh_detector_create(const vgh_net_desc* net, const vgh_flame* flame, vgh_detector** out);
void vgh_detector_destroy(vgh_detector* d);

/* images_dev: uint8 [B,S,S,3] RGB (letterboxed; the /255 of detector.py:51 is folded into the stem).
 * Fills the internal boxes [B,A,4] / scores [B,A] / raw-flame buffers. */
int vgh_detector_forward(vgh_detector* d, const uint8_t* images_dev, void* stream);
/* select+NMS, survivor FLAME rows, FLAME decode.  img_xform_dev optional [B,3] (pad_x,pad_y,scale). */
int vgh_detector_postprocess(vgh_detector* d, float conf_thr, float iou_thr, int top_k, const float* img_xform_dev,
                             void* stream);
/* The model-output boundary of the reference (boxes, scores, flame[B,A,413]) for parity tests:
 * expands the compact raw rows to the dense 413-wide tensor into flame_dev [B,A,413]. */
int vgh_detector_dense_flame(vgh_detector* d, float* flame_dev, void* stream);

/* Device views of the internal result buffers (valid until destroy). */
enum {
  VGH_OUT_BOXES = 0,       /* float [B,A,4] */
  VGH_OUT_SCORES = 1,      /* float [B,A] */
  VGH_OUT_KEEP_IDX = 2,    /* int32 [B,keep_k] */
  VGH_OUT_KEEP_CNT = 3,    /* int32 [B] */
  VGH_OUT_KEEP_BOXES = 4,  /* float [B,keep_k,4] */
  VGH_OUT_KEEP_SCORES = 5, /* float [B,keep_k] */
  VGH_OUT_HEAD_OFFSETS = 6,/* int32 [B+1] exclusive prefix of keep_cnt; [B] = total heads */
  VGH_OUT_HEAD_PARAMS = 7, /* float [total,413] packed image-major */
  VGH_OUT_HEAD_VERTS = 8,  /* float [total,5023,3] */
  VGH_OUT_HEAD_ROT = 9,    /* float [total,9] */
  VGH_OUT_INPUT = 10       /* uint8 [B,S,S,3] internal staging image buffer */
};
void* vgh_detector_output(vgh_detector* d, int which);
int vgh_detector_num_anchors(const vgh_detector* d);
/* Copy an activation buffer to the host (debug / layer-wise parity). */
int vgh_detector_read_buffer(vgh_detector* d, int buf, void* host_dst, size_t bytes);

/* Stage-wise parity aids (tests): overwrite an activation buffer from the host (same layout as read_buffer: NHWC,
 * bf16 bits or fp32), and run the dense plan ops [first_op, n_dense_ops) + box decode over the buffers as they are. */
int vgh_detector_write_buffer(vgh_detector* d, int buf, const void* host_src, size_t bytes);
int vgh_detector_forward_from(vgh_detector* d, int first_op, void* stream);

/* End to end with HOST buffers (pinned recommended): H2D of images, forward, postprocess, D2H of
 * counts / kept boxes / kept scores / packed params and vertices.  The whole device side replays
 * one CUDA graph.  verts_host capacity = max_heads*5023*3 floats; returns total heads in
 * *total_heads (heads beyond max_heads are not copied). */
int vgh_detector_run_host(vgh_detector* d, const uint8_t* images_host, const float* img_xform_host, float conf_thr,
                          float iou_thr, int top_k, int32_t* keep_cnt_host, float* keep_boxes_host,
                          float* keep_scores_host, float* params_host, float* verts_host, int max_heads,
                          int32_t* total_heads, void* stream);
/* Two-deep pipelined form of vgh_detector_run_host for streams of batches: submit() enqueues the H2D
 * upload (own copy stream), the graph replay and the staging of the results; collect() downloads the
 * results of the OLDEST outstanding submission (own copy stream) and blocks until they are on the
 * host.  At most two submissions may be outstanding.  Each step still does its own H2D and D2H; they
 * overlap the compute of the neighbouring steps.  Host buffers should be pinned. */
int vgh_detector_submit_host(vgh_detector* d, const uint8_t* images_host, float conf_thr, float iou_thr, int top_k);
int vgh_detector_collect_host(vgh_detector* d, int32_t* keep_cnt_host, float* keep_boxes_host, float* keep_scores_host,
                              float* params_host, float* verts_host, int max_heads, int32_t* total_heads);
/* ---------------------------------------------------------------------------------- multi-GPU gather
 * The reference is single-GPU; its batched twin (yolo_heads_post_prediction_callback.py:55-97) treats every image
 * independently, so the batch is sharded over ranks (one process per GPU) and the one exchange of the path is the
 * gather of every rank's predictions to rank 0 (BASELINE north_star).  Here that exchange is fused into
 * the per-step result snapshot: ONE kernel packs the step's results into a RECORD and writes it straight into its
 * destination - a local buffer, or rank 0's receive ring mapped over NVLink (vgh_peer_*), with device-side flags for
 * flow control; no size exchange and no host synchronisation.
 *
 * Record (4-byte words; n = heads of the step): [0] n, [1] B, [2] keep_k, [3] sequence, [4..15] reserved,
 * keep_cnt int32[B] (padded to 4 words), keep_boxes float[B,keep_k,4], keep_scores float[B,keep_k] (padded to 4) ->
 * `fixed_words`; then params float[n,413] (padded to 4), then vertices float[n,5023,3].
 * vgh_detector_record_layout: out4 = {fixed_words, 413, 15069, capacity_words (n = B*keep_k)}. */
int vgh_detector_record_layout(const vgh_detector* d, int64_t* out4);
/* Arms the NEXT vgh_detector_submit_device / vgh_detector_submit_host: after that step's graph replay its record is
 * packed into dst_record_dev (16-byte aligned, capacity_words words; local or peer-mapped).  wait_flag_dev (optional,
 * device memory): the pack first waits on the device until *wait_flag_dev >= wait_val (slot handed back by the
 * consumer; bounded by $VGGHEADS_B200_PUSH_TIMEOUT_MS, default 4000 -> sticky bit 1 in vgh_detector_push_status).
 * done_flag_dev (optional, may be peer memory): set to done_val once the whole record is visible system-wide. */
int vgh_detector_arm_push(vgh_detector* d, float* dst_record_dev, const uint64_t* wait_flag_dev, uint64_t wait_val,
                          uint64_t* done_flag_dev, uint64_t done_val);
int vgh_detector_push_status(vgh_detector* d, int32_t* status_host);
/* Device-resident step: graph replay over the internal staging input on `stream` (+ the armed record push). */
int vgh_detector_submit_device(vgh_detector* d, float conf_thr, float iou_thr, int top_k, void* stream);
/* Peer memory through CUDA IPC (one process per GPU): alloc returns a zeroed device allocation and its 64-byte handle;
 * another process opens the handle (peer access over NVLink is enabled lazily) and may hand the pointer to
 * vgh_detector_arm_push as destination / flag. */
int vgh_peer_alloc(size_t bytes, void** dev_ptr, uint8_t* handle64);
int vgh_peer_free(void* dev_ptr);
int vgh_peer_open(const uint8_t* handle64, void** dev_ptr);
int vgh_peer_close(void* dev_ptr);
/* Consumer (rank 0): a one-block kernel on `stream` waits until ready_dev[i*stride] >= value for all i < n (n <= 32;
 * bounded by timeout_ms, 0 = default -> bit 2 of *status_dev), stores the sum of the n records' head counts (record i
 * at records_dev + i*record_stride_words; NULL = skip) to *total_out_dev, then writes ack_val to every ack_ptrs_dev[i]
 * (device array of n device-accessible pointers, typically peer memory; NULL entries skipped). */
int vgh_gather_wait(const uint64_t* ready_dev, int n, int stride, uint64_t value, const float* records_dev,
                    int64_t record_stride_words, uint64_t* const* ack_ptrs_dev, uint64_t ack_val, int32_t* total_out_dev,
                    int32_t* status_dev, int timeout_ms, void* stream);
/* Same device work (graph replay) with inputs already resident in the internal staging buffer and
 * results left on the device - the kernel-only timing path. */
int vgh_detector_run_device(vgh_detector* d, float conf_thr, float iou_thr, int top_k, void* stream);
/* Synthetic-workload hook (bench/tests): after box decode, overwrite the internal boxes [B,A,4] and
 * scores [B,A] with caller data (random weights never yield detections).
 * The copy is part of the timed device work.  NULL, NULL disables it. */
int vgh_detector_set_override(vgh_detector* d, const float* boxes_dev, const float* scores_dev);
/* Measurement aid: eager (non-graph) execution over the internal staging image with one CUDA-event
 * pair around every plan op and every post-processing stage, averaged over `iters`.  ms_out
 * (capacity >= n_ops + 4) = [op_0 .. op_{n-1}, box_decode, select_nms, gather, flame_decode]. */
int vgh_detector_profile(vgh_detector* d, int iters, float conf_thr, float iou_thr, int top_k, float* ms_out,
                         int capacity, void* stream);
/* Device-side configuration search: times every conv op under a few (M-tiles per CTA, pipeline
 * depth) candidates and keeps the fastest.  Optional; call once after create. */
int vgh_detector_autotune(vgh_detector* d, int iters, void* stream);
/* Configuration of plan op `op`: out6 = {m_tiles_per_cta, stages, block_n, block_k, tile_w, tile_h}. */
int vgh_detector_op_config(const vgh_detector* d, int op, int32_t* out6);
/* Number of kernel launches one forward+postprocess issues (graph nodes), for reporting. */
int vgh_detector_launch_count(const vgh_detector* d);

#ifdef __cplusplus
}
#endif
#endif /* VGGHEADS_B200_H */
