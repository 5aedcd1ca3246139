/* boxpath.h — C ABI of libboxpath.so: the B200-native box-processing hot path.
 *
 * Drop-in boundary for the box-processing functions of irvingzhang0512/tf_eager_object_detection.
 * The reference has no FFI of its own: the boundary is the Python-object API of the
 * classes cited on each entry point below (paths relative to /root/reference/object_detection/).
 * Each entry point states which reference call it replaces; the Python mirror of those classes
 * lives in tf_eager_object_detection_b200/ and binds this header through ctypes.
 *
 * Conventions
 *  - plain C: pointers + sizes only.  All tensor pointers are DEVICE pointers (fp32 / int32,
 *    C-contiguous, 16-byte aligned where the element is a box) unless the name ends in `_host`.
 *  - every call returns 0 on success or a negative bx_status; bx_last_error() gives the message of the
 *    last failure on the calling thread.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *    the library never calls cudaDeviceSynchronize and never frees caller memory.  Only bx_destroy and
 *    bx_profile_read wait (on the handle's own stream / events).
 *  - every call runs on the HANDLE's device whatever the caller's current device is, and leaves the caller's current
 *    device unchanged (bx_create does not change it either).  A handle is used from one host thread and one stream at a
 *    time (its workspace is shared by consecutive calls and ordered by that stream); use one handle per stream for
 *    concurrent streams.
 *  - boxes are (x1, y1, x2, y2) in image pixels, as everywhere on the reference's path.
 *  - no CPU fallback: without a CUDA device bx_create fails and nothing else can be called.
 *  - BX_NVTX=1 in the environment (read once): every entry point that takes a handle runs inside an NVTX push/pop range
 *    named after it, so `ncu --nvtx --nvtx-include "bx_roi_pool/"` or a timeline tool attributes kernels to calls.
 */
#ifndef BOXPATH_H_
#define BOXPATH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BX_VERSION 100 /* 0.1.0 */

typedef enum {
  BX_OK = 0,
  BX_ERR_INVALID = -1,      /* bad argument (shape, range, NULL, alignment) — TF InvalidArgumentError analogue */
  BX_ERR_CUDA = -2,         /* a CUDA runtime/driver call failed */
  BX_ERR_UNSUPPORTED = -3,  /* valid request outside what the kernels are built for (documented limits) */
  BX_ERR_DLPACK = -4        /* DLTensor is not kDLCUDA / wrong dtype / not contiguous / wrong shape */
} bx_status;

typedef struct bx_handle bx_handle; /* per (device, host thread) context: workspace + TMA descriptors */

int bx_version(void);
const char* bx_last_error(void);
int bx_create(int device, bx_handle** out);
int bx_destroy(bx_handle* h);

/* Validate a DLPack tensor (struct DLTensor of dlpack.h, passed as void*) and return its data pointer
 * (data + byte_offset).  dtype_code/bits: 2/32 = float32, 0/32 = int32 (DLDataTypeCode).  `shape` may
 * contain -1 wildcards.  Rejects non-CUDA, wrong device, wrong dtype, non-contiguous, misaligned. */
int bx_dlpack_data(const void* dltensor, int device, int dtype_code, int bits, int ndim,
                   const int64_t* shape, int align_bytes, void** out_data);

/* ---- a1+a2: utils/bbox_transform.py:32-55 decode_bbox_with_mean_and_std, then
 *      utils/bbox_tf.py:59-78 bboxes_clip_filter(min_edge=None).  anchors [n,4] (anchor_batch_stride 0: shared by
 *      the batch) or [batch,n,4]; deltas [batch,n,4]; out [batch,n,4].  clip: x to [0,W-1], y to [0,H-1]; pass
 *      image_h = image_w = 0 to skip the clip (pure decode). */
int bx_decode_clip(bx_handle* h, const float* anchors, int anchors_batched, const float* deltas, int batch, int n,
                   const float means[4], const float stds[4], int image_h, int image_w, float* out_boxes,
                   void* stream);

/* ---- utils/bbox_transform.py:4-29 encode_bbox_with_mean_and_std: src [n,4], dst [n,4] -> out [n,4]. */
int bx_encode(bx_handle* h, const float* src, const float* dst, int n, const float means[4], const float stds[4],
              float* out, void* stream);

/* ---- utils/bbox_tf.py:80-84: min-edge filter of bboxes_clip_filter: keep rows with (x2-x1+1)>=min_edge and
 *      (y2-y1+1)>=min_edge; boxes are clipped first exactly as bx_decode_clip does.  out_boxes [n,4] compacted
 *      (ascending index), out_idx [n] int32, out_count [1]. */
int bx_clip_filter(bx_handle* h, const float* boxes, int n, float min_value, int image_h, int image_w,
                   float min_edge, float* out_boxes, int* out_idx, int* out_count, void* stream);

/* ---- utils/bbox_tf.py:87-101 bboxes_range_filter: out_mask [n] int32 (1 = inside), out_idx [n] compacted
 *      ascending, out_count [1]. */
int bx_range_filter(bx_handle* h, const float* anchors, int n, int image_h, int image_w, int* out_idx,
                    int* out_count, void* stream);

/* ---- tf.image.non_max_suppression as called at model/region_proposal.py:74-76 (TF r1.13 CPU kernel semantics:
 *      greedy by descending score, ties -> lower index; IoU without +1 on min/max-normalised corners;
 *      non-positive-area boxes neither suppress nor are suppressed; strict `>`).
 *      boxes [batch,n,4], scores [batch,n] -> out_idx [batch,max_out] int32 (selection order, -1 padded),
 *      out_count [batch].  Limits: max_out <= 2048, n < 2^22. */
int bx_nms(bx_handle* h, const float* boxes, const float* scores, int batch, int n, int max_out,
           float iou_threshold, int* out_idx, int* out_count, void* stream);

/* ---- a3: model/region_proposal.py:37-81 RegionProposal.call, batched over images.
 *      pre_nms_top_k = 0 and min_size <= 0 reproduce the reference (top-k commented out at :65-69, min_edge=None
 *      at :63); >0 enable the py-faster-rcnn steps in the order filter -> top-k -> NMS. */
typedef struct {
  float means[4];
  float stds[4];
  int image_h, image_w;
  int pre_nms_top_k;   /* 0 = all anchors (reference behaviour) */
  int post_nms;        /* max_output_size: 300 / 2000 (C4), 1000 / 2000 (FPN); <= 2048 */
  float iou_threshold; /* 0.7 */
  float min_size;      /* <= 0: no min-size filter (reference behaviour) */
} bx_proposal_params;

/* anchors [n,4] shared by the batch; deltas [batch,n,4]; scores [batch,n] ->
 * out_boxes [batch,post_nms,4] (selection order = descending score, zero padded), out_idx [batch,post_nms]
 * int32 anchor indices (-1 padded), out_count [batch]. */
int bx_proposals(bx_handle* h, const float* anchors, const float* deltas, const float* scores, int batch, int n,
                 const bx_proposal_params* p, float* out_boxes, int* out_idx, int* out_count, void* stream);

/* ---- tf.image.crop_and_resize(bilinear, extrapolation 0) as called at model/roi_pooling.py:37,79,86,134.
 *      image [b,ih,iw,c] NHWC; boxes [r,4] = (y1,x1,y2,x2) normalised; box_ind [r] int32 -> out [r,ch,cw,c]. */
int bx_crop_and_resize(bx_handle* h, const float* image, int b, int ih, int iw, int c, const float* boxes,
                       const int* box_ind, int r, int crop_h, int crop_w, float extrapolation_value, float* out,
                       void* stream);

/* ---- a4/a5/a8: the three RoI feature extractors of model/roi_pooling.py, batched.
 *      BX_ROI_STRIDE_NORM : RoiPoolingCropAndResize.call :53-90  (rois/stride, normalised by (fh-1),(fw-1))
 *      BX_ROI_IMAGE_NORM  : RoiPoolingCropAndResize2.call :15-42 (normalised by image H, W)
 *      BX_ROI_ALIGN_PAD   : RoiPoolingRoiAlign.call :158-176 -> roi_align :140-155 -> crop_and_resize(pad_border) :93-137
 *      pool: BX_POOL_NONE crop PxP directly (ResNet C4, :85-90); BX_POOL_MAX2 crop 2Px2P then MaxPooling2D 2x2 'same'
 *      (:75-84, :42); BX_POOL_AVG2 crop 2Px2P then avg_pool 2x2 (:154; only with BX_ROI_ALIGN_PAD).
 *      feat [b,fh,fw,c]; rois [r,4] image coordinates; box_ind [r] int32 or NULL (all zeros, as the reference);
 *      roi_counts NULL, or [b] int32 with rois laid out [b, r/b, 4]: rows >= count are zero-filled (padding of
 *      bx_proposals).  out [r,P,P,c]. */
typedef enum { BX_ROI_STRIDE_NORM = 0, BX_ROI_IMAGE_NORM = 1, BX_ROI_ALIGN_PAD = 2 } bx_roi_mode;
typedef enum { BX_POOL_NONE = 0, BX_POOL_MAX2 = 1, BX_POOL_AVG2 = 2 } bx_pool_op;
int bx_roi_pool(bx_handle* h, int mode, int pool, int pool_size, const float* feat, int b, int fh, int fw, int c,
                const float* rois, const int* box_ind, const int* roi_counts, int r, float stride, int image_h,
                int image_w, float* out, void* stream);

/* ---- f2 ("next" row): the inputs of the path, produced on the device.
 *      bx_generate_anchors: utils/anchor_generator.py:46-60 generate_by_anchor_base_tf (one level; offsets = the
 *      anchor base rows) and :137-162 make_anchors over P2..P6 as concatenated at fpn/base_fpn_model.py:163-186
 *      (offsets = (-0.5ws, -0.5hs, +0.5ws, +0.5hs) per anchor): anchor = (x*stride, y*stride, x*stride, y*stride) +
 *      offsets[level][k], cells row-major, anchors fastest.  fh/fw/stride [n_levels] and offsets [n_levels, A, 4] are
 *      HOST arrays (passed as kernel parameters); out_anchors [sum fh*fw*A, 4].  n_levels <= 5, A <= 32.
 *      bx_rpn_scores: raw RPN logits -> foreground probability (tf.nn.softmax over the (bg, fg) pair):
 *        BX_RPN_CAFFE  faster_rcnn/base_faster_rcnn_model.py:149-152, logits [batch, n/A cells, 2A] = [bg x A | fg x A];
 *        BX_RPN_PAIRS  fpn/base_fpn_model.py:223, logits [batch, n, 2] = (bg, fg) per anchor.
 *      bx_proposals_rpn: bx_proposals on the raw logits; for n <= 24576 without min_size the softmax runs inside the
 *      proposal kernel's key pass (no extra launch, no score tensor); out_scores [batch,n] is optional. */
typedef enum { BX_RPN_CAFFE = 0, BX_RPN_PAIRS = 1 } bx_rpn_layout;
int bx_generate_anchors(bx_handle* h, int n_levels, const int* fh, const int* fw, const float* stride,
                        int anchors_per_cell, const float* offsets, float* out_anchors, void* stream);
int bx_rpn_scores(bx_handle* h, const float* logits, int layout, int anchors_per_cell, int batch, int n,
                  float* out_scores, void* stream);
int bx_proposals_rpn(bx_handle* h, const float* anchors, const float* deltas, const float* logits, int layout,
                     int anchors_per_cell, int batch, int n, const bx_proposal_params* p, float* out_boxes,
                     int* out_idx, int* out_count, float* out_scores, void* stream);

/* ---- f3 ("next" row): gradient of bx_roi_pool w.r.t. the feature map, i.e. the backward pass TF runs through
 *      tf.image.crop_and_resize (+ the 2x2 pool) when scripts/train.py:99-103 differentiates the model (the boxes are
 *      under tf.stop_gradient, model/roi_pooling.py:37,79,86).  Same arguments as bx_roi_pool; grad_out [r,P,P,c];
 *      grad_feat [b,fh,fw,c] is written completely by the call.  feat is read only for BX_POOL_MAX2 (argmax).  c must
 *      be a multiple of 4.  Two kernels: a row-owned one without atomics — every pixel's sum has one
 *      fixed order, the result is bit-reproducible — which is the default for BX_POOL_NONE and BX_POOL_AVG2, and a
 *      scatter kernel with fp32 atomics (last bits vary from run to run), the default for BX_POOL_MAX2 where it is the
 *      faster one.  bx_set_deterministic(h, 1) selects the row-owned kernel for every mode. */
int bx_roi_pool_grad(bx_handle* h, int mode, int pool, int pool_size, const float* feat, int b, int fh, int fw, int c,
                     const float* rois, const int* box_ind, const int* roi_counts, int r, float stride, int image_h,
                     int image_w, const float* grad_out, float* grad_feat, void* stream);

/* ---- f3 ("next" row): model/losses.py:16-28 smooth_l1_loss as called at faster_rcnn/base_faster_rcnn_model.py:209-211
 *      (dim=[0,1], reduce_all=1: the total sum) and :220-222 (dim=[1], reduce_all=0: sum / n).  pred, target, in_w,
 *      out_w [n,d]; out_loss [1]; out_grad [n,d] or NULL = d loss / d pred (the `sign` mask is a constant,
 *      losses.py:21).  Deterministic reduction (fixed tree, fp64 partials). */
int bx_smooth_l1_loss(bx_handle* h, const float* pred, const float* target, const float* in_w, const float* out_w,
                      long long n, int d, float sigma, int reduce_all, float* out_loss, float* out_grad, void* stream);

/* ---- f3: model/losses.py:4-13 cls_loss = tf.losses.sparse_softmax_cross_entropy(logits, to_int32(labels), weight)
 *      with the caller's `labels >= 0` gather (base_faster_rcnn_model.py:204-207) folded in: rows whose label is
 *      negative are skipped.  logits [n,c]; labels [n] fp32 (what bx_anchor_target / bx_proposal_target emit);
 *      loss = weight * sum(logsumexp(x) - x[label]) / #selected (0 when none, or weight == 0:
 *      SUM_BY_NONZERO_WEIGHTS).  out_count (nullable) = #selected; out_grad [n,c] or NULL = d loss / d logits. */
int bx_cls_loss(bx_handle* h, const float* logits, const float* labels, int n, int c, float weight, float* out_loss,
                int* out_count, float* out_grad, void* stream);

/* ---- a6: model/fpn/base_fpn_model.py:303-324 BaseFPN._assign_levels.
 *      rois [r,4] -> out_level [r] int32 in [min_level,max_level]; out_order [r] int32 = concat over levels of the
 *      ascending indices of each level (the reference's `assign_level_idx`); out_counts [max_level-min_level+1]. */
int bx_fpn_assign_levels(bx_handle* h, const float* rois, int r, int min_level, int max_level, int* out_level,
                         int* out_order, int* out_counts, void* stream);

/* ---- a7: model/fpn/base_fpn_model.py:152-161 BaseFPN._get_roi_features (+ a6): RoiPoolingCropAndResize2 on the
 *      level each roi is assigned to, rows written in level-major order (out row j <- roi out_order[j]).
 *      feats[l] -> [b,fh[l],fw[l],c] for l = min_level..max_level (n_levels pointers, host arrays of pointers/dims);
 *      rois [r,4]; box_ind [r] or NULL; out [r,P,P,c]; out_order [r]; out_counts [n_levels]. */
int bx_fpn_roi_features(bx_handle* h, const float* const* feats, const int* fh, const int* fw, int n_levels,
                        int min_level, int b, int c, const float* rois, const int* box_ind, int r, int image_h,
                        int image_w, int pool_size, float* out, int* out_level, int* out_order, int* out_counts,
                        void* stream);

/* ---- a9: utils/bbox_tf.py:37-56 pairwise_iou ("+1" areas; 0 where the intersection is 0).
 *      a [n,4], b [m,4] -> out [n,m] row-major. */
int bx_pairwise_iou(bx_handle* h, const float* a, int n, const float* b, int m, float* out, void* stream);

/* ---- a10: model/anchor_target.py:29-107 AnchorTarget.call (+_unmap :110-125), batched over images.
 *      anchors [n,4] shared; gt [batch,max_gt,4] with gt_counts [batch] (NULL: all max_gt valid);
 *      perm [batch,n] int32: sampling priority per anchor replacing the reference's unseeded tf.random_shuffle
 *      (shuffle(idx) := idx sorted ascending by perm; DESIGN.md "Sampling").
 *      -> labels [batch,n] fp32 (-1/0/1), targets, in_w, out_w [batch,n,4]; counts [batch,2] = (#fg,#bg) sampled.
 *      An image without ground truth (max_gt == 0, gt may be NULL, or gt_counts[i] == 0) — where the reference's argmax
 *      over an empty axis raises — is treated as max overlap 0: inside anchors are background (bx_proposal_target: every
 *      roi is a background candidate when neg_iou_threshold <= 0). */
typedef struct {
  float pos_iou_threshold;  /* 0.7 */
  float neg_iou_threshold;  /* 0.3 */
  int total_num_samples;    /* 256 */
  int max_pos_samples;      /* 128 */
  float means[4];
  float stds[4];
  int image_h, image_w;
} bx_anchor_target_params;
int bx_anchor_target(bx_handle* h, const float* anchors, int n, const float* gt, const int* gt_counts, int batch,
                     int max_gt, const int* perm, const bx_anchor_target_params* p, float* out_labels,
                     float* out_targets, float* out_in_w, float* out_out_w, int* out_counts, void* stream);

/* ---- a11: model/proposal_target.py:32-124 ProposalTarget.call, batched over images.
 *      rois [batch,k,4] with roi_counts [batch] (NULL: k each); gt [batch,max_gt,4]; gt_labels [batch,max_gt] int32;
 *      perm [batch,k] int32 (as above; background padding cycles through the shuffled bg set).
 *      -> out_rois [batch,S,4], out_labels [batch,S] int32, out_targets/in_w/out_w [batch,S,4*num_classes],
 *      out_keep [batch,S] int32 roi indices, out_counts [batch,2] = (#fg kept, status: 0 ok, 1 = empty background
 *      set where the reference's np.random.choice raises ValueError, :77). Keeps the `labels[idx]` quirk (:99,117). */
typedef struct {
  int num_classes;          /* 21 */
  float pos_iou_threshold;  /* 0.5 */
  float neg_iou_threshold;  /* 0.0 in all reference configs (ctor default 0.5) */
  int total_num_samples;    /* S: 128 (C4) / 256 (FPN) */
  int max_pos_samples;      /* 32 / 64 */
  float means[4];
  float stds[4];
} bx_proposal_target_params;
int bx_proposal_target(bx_handle* h, const float* rois, const int* roi_counts, int k, const float* gt,
                       const int* gt_labels, const int* gt_counts, int batch, int max_gt, const int* perm,
                       const bx_proposal_target_params* p, float* out_rois, int* out_labels, float* out_targets,
                       float* out_in_w, float* out_out_w, int* out_keep, int* out_counts, void* stream);

/* ---- f1 ("next" row): model/prediction.py:103-163 post_ops_prediction, batched over images.  Per foreground class:
 *      score > score_threshold -> decode with the roi-head means/stds (utils/bbox_transform.py:32-55) -> clip to the image
 *      + min-edge filter (utils/bbox_tf.py:59-84, min_edge = extractor_stride) -> tf.image.non_max_suppression
 *      (max_per_class, nms_iou_threshold) -> concatenation over classes -> top max_per_image by score (descending; ties
 *      to the lower class, then the earlier NMS pick).
 *      scores [batch,r,C] softmax; deltas [batch,r,C,4]; rois [batch,r,4]; roi_counts [batch] or NULL ->
 *      out_det [batch,max_per_image,6] = (x1,y1,x2,y2,score,class) zero padded — the record layout that
 *      distributed.allgather_detections ships — and out_count [batch].  Limits: (C-1)*max_per_class <= 8192. */
typedef struct {
  float means[4];
  float stds[4];            /* roi head: (0.1, 0.1, 0.2, 0.2) */
  int image_h, image_w;
  int num_classes;          /* C, class 0 = background */
  int max_per_class;        /* 50 */
  int max_per_image;        /* 150 (ctor default; the configs use 50) */
  float nms_iou_threshold;  /* 0.3 */
  float score_threshold;    /* 0.05 (configs: 0.0) */
  float min_edge;           /* extractor_stride = 16; <= 0 disables the filter */
} bx_prediction_params;
int bx_post_ops_prediction(bx_handle* h, const float* scores, const float* deltas, const float* rois,
                           const int* roi_counts, int batch, int r, const bx_prediction_params* p, float* out_det,
                           int* out_count, void* stream);

/* ---- f1, the two copies of that logic inside the reference's EVALUATION loops, which differ from post_ops_prediction in
 *      three ways (evaluation/pascal_eval_files_utils.py:76-106, scripts/eval_coco.py:116-153):
 *        - the rois come from im_detect, which divides them by the image's resize factor first
 *          (faster_rcnn/base_faster_rcnn_model.py:304 `rois / img_scale`): img_scale [batch] fp32 or NULL;
 *        - boxes are clipped to the image's own RAW size: image_sizes [batch,2] = (raw_h, raw_w) fp32 (x to [0, raw_w-1],
 *          y to [0, raw_h-1]) or NULL (params->image_h/w for the whole batch); min_edge = the loops' min_size;
 *        - the per-image cut: BX_CUT_TOP_K = top max_per_image by score (eval_coco.py:148, prediction.py:158);
 *          BX_CUT_SCORE_GE = every detection with score >= the max_per_image-th largest score
 *          (pascal_eval_files_utils.py:99-106: ties with the cut are KEPT, so an image can return more than max_per_image;
 *          out_rows >= max_per_image is the room the caller provides, anything beyond it is dropped in score order).
 *      out_det [batch,out_rows,6] zero padded, out_count [batch]; everything else as bx_post_ops_prediction. */
typedef enum { BX_CUT_TOP_K = 0, BX_CUT_SCORE_GE = 1 } bx_cut_mode;
int bx_eval_detections(bx_handle* h, const float* scores, const float* deltas, const float* rois, const int* roi_counts,
                       const float* image_sizes, const float* img_scale, int batch, int r, const bx_prediction_params* p,
                       int cut_mode, int out_rows, float* out_det, int* out_count, void* stream);

/* ---- composite used by the benchmark and by BaseFasterRcnn.call eval (faster_rcnn/base_faster_rcnn_model.py:153,182):
 *      bx_proposals followed by bx_roi_pool(BX_ROI_STRIDE_NORM) on the kept boxes, device-resident tensors. */
int bx_c4_proposal_roi(bx_handle* h, const float* anchors, const float* deltas, const float* scores,
                       const float* feat, int batch, int n, int fh, int fw, int c, const bx_proposal_params* p,
                       float stride, int pool_size, int pool, float* out_rois, int* out_idx, int* out_count,
                       float* out_feat, void* stream);

/* ---- same, HOST buffers (pinned recommended): copies deltas/scores/feat host->device, runs, copies
 *      rois/idx/count/features device->host, all on `stream`; anchors stay device-resident (they depend on the image
 *      shape only).  Returns after enqueueing; synchronise `stream` before reading the host outputs. */
int bx_c4_proposal_roi_host(bx_handle* h, const float* anchors_dev, const float* deltas_host,
                            const float* scores_host, const float* feat_host, int batch, int n, int fh, int fw, int c,
                            const bx_proposal_params* p, float stride, int pool_size, int pool,
                            float* out_rois_host, int* out_idx_host, int* out_count_host, float* out_feat_host,
                            void* stream);

/* ---- e (multi-GPU): the path's only exchange — all-gather of the per-image detection records for evaluation, which
 *      replaces the single-process accumulation loops of evaluation/pascal_eval_files_utils.py:73-107 and
 *      scripts/eval_coco.py:116-164.  `nccl_comm` is the host framework's ncclComm_t (torch:
 *      ProcessGroupNCCL._comm_ptr()); NCCL is not linked: ncclAllGather / ncclGroupStart / ncclGroupEnd are resolved
 *      from the libnccl already loaded in the process (BX_ERR_UNSUPPORTED if there is none).  Every rank passes the same
 *      b_local (pad uneven shards first).  records [b_local,kmax,fields] fp32 + counts [b_local] int32 ->
 *      out_records [world*b_local,kmax,fields], out_counts [world*b_local] in rank order; one fused NCCL group on
 *      `stream`. */
int bx_allgather_detections(bx_handle* h, void* nccl_comm, const float* records, const int* counts, int b_local,
                            int kmax, int fields, int world, float* out_records, int* out_counts, void* stream);

/* on != 0: calls on this handle use bit-reproducible kernels where the default one is not (today: bx_roi_pool_grad with
 * BX_POOL_MAX2).  The Python mirror sets it from torch.are_deterministic_algorithms_enabled(). */
int bx_set_deterministic(bx_handle* h, int on);

/* number of kernels launched by this handle since creation (bench.py "gpu_launches") */
long long bx_launch_count(const bx_handle* h);

/* Counters and sizes of the handle: out[0] = kernels launched, out[1] = RoI launches served by the TMA band kernel,
 * out[2] = plain-crop RoI launches that FELL BACK to a gather kernel (shape outside the band kernel's limits: C % 32 != 0,
 * map too wide for a 3-row band, no cuTensorMapEncodeTiled; about 0.75x the speed — never silent: it is counted here),
 * out[3..5] = bytes held by the workspace / RoI plan area / host-entry staging area.  Writes min(n, 6) values. */
int bx_stats(const bx_handle* h, long long* out, int n);

/* Grow the handle's three device areas to at least the given sizes now (stream-ordered on `stream`).  Every entry point
 * grows them on demand — stream-ordered (cudaMallocAsync / cudaFreeAsync), without synchronising the device — the first
 * time a call needs more than any earlier call did; bx_reserve() (or one warm-up call of each op at its largest shape)
 * moves that growth out of the steady state, and is required before a CUDA-graph capture, during which growth is refused
 * with BX_ERR_UNSUPPORTED.  Read the sizes a workload reached with bx_stats(). */
int bx_reserve(bx_handle* h, size_t workspace_bytes, size_t plan_bytes, size_t stage_bytes, void* stream);

/* Measurement support (bench.py roofline leg): when enabled, every RoI-pooling launch of this handle is bracketed by a
 * cudaEvent pair recorded on the launch stream (up to `capacity` launches, then recording stops).
 * bx_profile_read synchronises the recorded events and returns their durations in milliseconds. */
int bx_profile_roi(bx_handle* h, int enable, int capacity);
int bx_profile_read(bx_handle* h, float* ms_out, int max_records, int* n_out);

#ifdef __cplusplus
}
#endif
#endif /* BOXPATH_H_ */
