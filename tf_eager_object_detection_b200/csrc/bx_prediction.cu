// Post-head detection filtering ("next" row f1): model/prediction.py:103-163 post_ops_prediction, batched over images
// and classes.  Every (image, foreground class) pair is one pseudo-image of the proposal-stage NMS kernel
// (bx_proposals.cu): a prepare kernel applies the score threshold, decodes with the roi-head stds, clips and applies the
// min-edge filter (excluded candidates get key 0), the NMS kernel keeps <= max_per_class per pair, and a per-image
// kernel picks the top max_per_image of the concatenation.
#include "bx_common.cuh"

namespace {

constexpr int kMaxCand = 8192;   // (C-1) * max_per_class candidates per image in the final top-k

struct PredArgs {
  const float* scores;     // [B,R,C]
  const float4* deltas;    // [B,R,C]
  const float4* rois;      // [B,R]
  const int* roi_counts;   // [B] or null
  const float* image_sizes;  // [B,2] = (raw_h, raw_w) per image or null (codec.max_x / max_y for the whole batch)
  const float* img_scale;    // [B] or null: rois are divided by it first (im_detect, base_faster_rcnn_model.py:304)
  int B, R, C;
  BoxCodec codec;
  float score_thr, min_edge;
  float4* ws_boxes;        // [B*(C-1), R]
  uint32_t* ws_keys;       // [B*(C-1), R]
};

__global__ void __launch_bounds__(256) prediction_prepare_kernel(const PredArgs a) {
  const int fg = a.C - 1;
  const long long total = static_cast<long long>(a.B) * a.R * fg;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int c = static_cast<int>(i % fg) + 1;            // class fastest: coalesced reads of scores / deltas
    const long long br = i / fg;
    const int r = static_cast<int>(br % a.R), b = static_cast<int>(br / a.R);
    const float s = a.scores[br * a.C + c];
    float4 roi = a.rois[br];
    if (a.img_scale) {
      const float sc = a.img_scale[b];
      roi = make_float4(roi.x / sc, roi.y / sc, roi.z / sc, roi.w / sc);
    }
    BoxCodec codec = a.codec;
    if (a.image_sizes) {           // clip box of THIS image: x to [0, raw_w - 1], y to [0, raw_h - 1] (utils/bbox_tf.py:71-74)
      codec.max_y = a.image_sizes[2 * b] - 1.0f;
      codec.max_x = a.image_sizes[2 * b + 1] - 1.0f;
    }
    const float4 box = bx_decode_clip_one(roi, a.deltas[br * a.C + c], codec);
    bool ok = s > a.score_thr;                                                               // prediction.py:135
    if (a.roi_counts) ok = ok && (r < a.roi_counts[b]);
    if (a.min_edge > 0.0f)                                                                   // utils/bbox_tf.py:80-83
      ok = ok && ((box.z - box.x + 1.0f) >= a.min_edge) && ((box.w - box.y + 1.0f) >= a.min_edge);
    const size_t o = (static_cast<size_t>(b) * fg + (c - 1)) * a.R + r;
    a.ws_boxes[o] = box;
    a.ws_keys[o] = ok ? bx_score_key(s + 0.0f) : 0u;
  }
}

struct TopkArgs {
  const float* scores;       // [B,R,C]
  const float4* kept_boxes;  // [B*(C-1), Kc]
  const int* kept_idx;       // [B*(C-1), Kc]
  const int* kept_count;     // [B*(C-1)]
  int R, C, Kc, max_per_image;
  int cut_mode;              // bx_cut_mode
  int out_rows;              // rows of out_det per image (>= max_per_image)
  float* out_det;            // [B, out_rows, 6]
  int* out_count;            // [B]
};

__global__ void __launch_bounds__(1024) prediction_topk_kernel(const TopkArgs a) {
  extern __shared__ unsigned long long comp[];
  __shared__ int s_total;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int fg = a.C - 1, n = fg * a.Kc;
  int pow2 = 64;
  while (pow2 < n) pow2 <<= 1;
  if (tid == 0) s_total = 0;
  __syncthreads();
  int mine = 0;
  for (int e = tid; e < pow2; e += 1024) {
    unsigned long long v = 0ull;
    if (e < n) {
      const int cls = e / a.Kc, k = e - cls * a.Kc;
      const int pair = b * fg + cls;
      if (k < a.kept_count[pair]) {
        const int roi = a.kept_idx[static_cast<size_t>(pair) * a.Kc + k];
        const float s = a.scores[(static_cast<size_t>(b) * a.R + roi) * a.C + cls + 1];
        v = (static_cast<unsigned long long>(bx_score_key(s + 0.0f)) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(e));
        ++mine;
      }
    }
    comp[e] = v;
  }
  mine = __reduce_add_sync(0xFFFFFFFFu, mine);
  if ((tid & 31) == 0 && mine) atomicAdd(&s_total, mine);
  __syncthreads();
  bx_bitonic_sort<true>(comp, pow2);
  int keep = min(s_total, a.max_per_image);
  if (a.cut_mode == BX_CUT_SCORE_GE && s_total > a.max_per_image) {
    // evaluation/pascal_eval_files_utils.py:99-106: image_thresh = the max_per_image-th largest score, keep score >= it —
    // every detection that TIES with the cut survives (sorted descending, so the ties follow the cut directly)
    __shared__ int s_extra;
    if (tid == 0) s_extra = 0;
    __syncthreads();
    const unsigned long long cut_key = comp[a.max_per_image - 1] >> 32;
    int extra = 0;
    for (int e = a.max_per_image + tid; e < s_total; e += 1024) extra += ((comp[e] >> 32) == cut_key) ? 1 : 0;
    extra = __reduce_add_sync(0xFFFFFFFFu, extra);
    if ((tid & 31) == 0 && extra) atomicAdd(&s_extra, extra);
    __syncthreads();
    keep = min(a.max_per_image + s_extra, a.out_rows);
  }
  float* det = a.out_det + static_cast<size_t>(b) * a.out_rows * 6;
  for (int t = tid; t < a.out_rows; t += 1024) {
    float rec[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (t < keep) {
      const int e = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(comp[t] & 0xFFFFFFFFull));
      const int cls = e / a.Kc, k = e - cls * a.Kc;
      const int pair = b * fg + cls;
      const float4 box = a.kept_boxes[static_cast<size_t>(pair) * a.Kc + k];
      const int roi = a.kept_idx[static_cast<size_t>(pair) * a.Kc + k];
      rec[0] = box.x; rec[1] = box.y; rec[2] = box.z; rec[3] = box.w;
      rec[4] = a.scores[(static_cast<size_t>(b) * a.R + roi) * a.C + cls + 1];
      rec[5] = static_cast<float>(cls + 1);
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) det[t * 6 + q] = rec[q];
  }
  if (tid == 0) a.out_count[b] = keep;
}

}  // namespace

static int prediction_impl(bx_handle* h, const float* scores, const float* deltas, const float* rois,
                           const int* roi_counts, const float* image_sizes, const float* img_scale, int batch, int r,
                           const bx_prediction_params* p, int cut_mode, int out_rows, float* out_det, int* out_count,
                           void* stream);

extern "C" int bx_post_ops_prediction(bx_handle* h, const float* scores, const float* deltas, const float* rois,
                                      const int* roi_counts, int batch, int r, const bx_prediction_params* p,
                                      float* out_det, int* out_count, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(p, BX_ERR_INVALID, "bx_post_ops_prediction: NULL argument");
  return prediction_impl(h, scores, deltas, rois, roi_counts, nullptr, nullptr, batch, r, p, BX_CUT_TOP_K, p->max_per_image,
                         out_det, out_count, stream);
}

extern "C" int bx_eval_detections(bx_handle* h, const float* scores, const float* deltas, const float* rois,
                                  const int* roi_counts, const float* image_sizes, const float* img_scale, int batch,
                                  int r, const bx_prediction_params* p, int cut_mode, int out_rows, float* out_det,
                                  int* out_count, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(p, BX_ERR_INVALID, "bx_eval_detections: NULL argument");
  BX_REQUIRE(cut_mode == BX_CUT_TOP_K || cut_mode == BX_CUT_SCORE_GE, BX_ERR_INVALID, "bx_eval_detections: bad cut_mode %d", cut_mode);
  BX_REQUIRE(out_rows >= p->max_per_image, BX_ERR_INVALID, "bx_eval_detections: out_rows %d < max_per_image %d", out_rows,
             p->max_per_image);
  return prediction_impl(h, scores, deltas, rois, roi_counts, image_sizes, img_scale, batch, r, p, cut_mode, out_rows, out_det,
                         out_count, stream);
}

static int prediction_impl(bx_handle* h, const float* scores, const float* deltas, const float* rois,
                           const int* roi_counts, const float* image_sizes, const float* img_scale, int batch, int r,
                           const bx_prediction_params* p, int cut_mode, int out_rows, float* out_det, int* out_count,
                           void* stream) {
  BX_REQUIRE(h && scores && deltas && rois && p && out_det && out_count, BX_ERR_INVALID,
             "bx_post_ops_prediction: NULL argument");
  BX_REQUIRE(batch >= 0 && r >= 0 && p->num_classes >= 2, BX_ERR_INVALID, "bx_post_ops_prediction: bad size");
  BX_REQUIRE(p->max_per_class > 0 && p->max_per_image > 0, BX_ERR_INVALID, "bx_post_ops_prediction: bad limits");
  BX_REQUIRE(p->nms_iou_threshold >= 0.0f && p->nms_iou_threshold <= 1.0f, BX_ERR_INVALID,
             "bx_post_ops_prediction: iou_threshold must be in [0, 1]");
  BX_REQUIRE(image_sizes || (p->image_h > 0 && p->image_w > 0), BX_ERR_INVALID,
             "bx_post_ops_prediction: image shape must be positive");
  const int fg = p->num_classes - 1;
  BX_REQUIRE(static_cast<long long>(fg) * p->max_per_class <= kMaxCand, BX_ERR_UNSUPPORTED,
             "bx_post_ops_prediction: (C-1) * max_per_class = %lld > %d", (long long)fg * p->max_per_class, kMaxCand);
  BX_REQUIRE(bx_aligned(deltas, 16) && bx_aligned(rois, 16), BX_ERR_INVALID,
             "bx_post_ops_prediction: box tensors must be 16-byte aligned");
  if (batch == 0) return BX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t pairs = static_cast<size_t>(batch) * fg;
  const int rr = r > 0 ? r : 1;
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  const size_t b_boxes = up(pairs * rr * sizeof(float4)), b_keys = up(pairs * rr * sizeof(uint32_t));
  const size_t b_kbox = up(pairs * p->max_per_class * sizeof(float4)), b_kidx = up(pairs * p->max_per_class * sizeof(int));
  const size_t b_kcnt = up(pairs * sizeof(int));
  int rc = bx_ws_reserve(h, b_boxes + b_keys + b_kbox + b_kidx + b_kcnt, st);
  if (rc) return rc;
  char* base = static_cast<char*>(h->ws);
  float4* ws_boxes = reinterpret_cast<float4*>(base); base += b_boxes;
  uint32_t* ws_keys = reinterpret_cast<uint32_t*>(base); base += b_keys;
  float4* kept_boxes = reinterpret_cast<float4*>(base); base += b_kbox;
  int* kept_idx = reinterpret_cast<int*>(base); base += b_kidx;
  int* kept_cnt = reinterpret_cast<int*>(base);

  if (r > 0) {
    PredArgs a = {};
    a.scores = scores;
    a.deltas = reinterpret_cast<const float4*>(deltas);
    a.rois = reinterpret_cast<const float4*>(rois);
    a.roi_counts = roi_counts;
    a.image_sizes = image_sizes;
    a.img_scale = img_scale;
    a.B = batch; a.R = r; a.C = p->num_classes;
    a.codec.m0 = p->means[0]; a.codec.m1 = p->means[1]; a.codec.m2 = p->means[2]; a.codec.m3 = p->means[3];
    a.codec.s0 = p->stds[0]; a.codec.s1 = p->stds[1]; a.codec.s2 = p->stds[2]; a.codec.s3 = p->stds[3];
    a.codec.clip = 1;
    a.codec.max_x = static_cast<float>(p->image_w - 1);
    a.codec.max_y = static_cast<float>(p->image_h - 1);
    a.score_thr = p->score_threshold;
    a.min_edge = p->min_edge;
    a.ws_boxes = ws_boxes;
    a.ws_keys = ws_keys;
    const long long total = static_cast<long long>(batch) * r * fg;
    prediction_prepare_kernel<<<static_cast<int>(bx_min_ll(bx_div_up(total, 256), 8ll * h->num_sms)), 256, 0, st>>>(a);
    BX_LAUNCH_CHECK(h);
    rc = bx_internal_nms_keys(h, reinterpret_cast<const float*>(ws_boxes), ws_keys, static_cast<int>(pairs), r,
                              p->max_per_class, p->nms_iou_threshold, reinterpret_cast<float*>(kept_boxes), kept_idx,
                              kept_cnt, st);
    if (rc) return rc;
  } else {
    BX_CUDA(cudaMemsetAsync(kept_cnt, 0, pairs * sizeof(int), st));
  }
  TopkArgs t = {};
  t.scores = scores;
  t.kept_boxes = kept_boxes;
  t.kept_idx = kept_idx;
  t.kept_count = kept_cnt;
  t.R = r; t.C = p->num_classes; t.Kc = p->max_per_class; t.max_per_image = p->max_per_image;
  t.cut_mode = cut_mode;
  t.out_rows = out_rows;
  t.out_det = out_det;
  t.out_count = out_count;
  int pow2 = 64;
  while (pow2 < fg * p->max_per_class) pow2 <<= 1;
  const size_t smem = static_cast<size_t>(pow2) * sizeof(unsigned long long);
  BX_CUDA(cudaFuncSetAttribute(prediction_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prediction_topk_kernel<<<batch, 1024, smem, st>>>(t);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}
