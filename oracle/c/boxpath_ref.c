/* CPU ORACLE, C twin (TEST INFRASTRUCTURE / CPU BASELINE ONLY — never linked into the product).
 *
 * Plain-C restatement of the heavy part of the reference's path, used (a) as a second checker beside the numpy
 * oracle and (b) as the timed "reference arm" / cpu_baseline of bench.py, since the reference's real TF CPU kernels
 * cannot run here (TensorFlow not installable).  PARITY UNPINNED at the TF-kernel boundary.
 * Follows (paths relative to /root/reference/object_detection/):
 *   orc_decode_clip      utils/bbox_transform.py:32-55 + utils/bbox_tf.py:71-74
 *   orc_nms              tf.image.non_max_suppression as called at model/region_proposal.py:74-76 (TF r1.13 CPU kernel,
 *                        SURVEY App. B.1): single-threaded greedy, like TF's kernel
 *   orc_crop_and_resize  tf.image.crop_and_resize at model/roi_pooling.py:37,79,86, sharded over
 *                        boxes across host threads like TF's Shard()
 *   orc_roi_pool_c4      model/roi_pooling.py:53-90
 *   orc_c4_proposal_roi  model/region_proposal.py:37-81 then model/roi_pooling.py:53-90, per image
 * Built with -ffp-contract=off so that no FMA contraction changes the fp32 results.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must still use all host cores */
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_decode_clip(const float* anchors, const float* deltas, int n, const float* means, const float* stds, int H,
                     int W, float* out) {
  const float mx = (float)(W - 1), my = (float)(H - 1);
  for (int i = 0; i < n; ++i) {
    const float* a = anchors + 4 * i;
    const float* t = deltas + 4 * i;
    const float dx = t[0] * stds[0] + means[0], dy = t[1] * stds[1] + means[1];
    const float dw = t[2] * stds[2] + means[2], dh = t[3] * stds[3] + means[3];
    float w = a[2] - a[0] + 1.0f, h = a[3] - a[1] + 1.0f;
    float cx = a[0] + 0.5f * w, cy = a[1] + 0.5f * h;
    cx = cx + dx * w;
    cy = cy + dy * h;
    w = w * expf(dw);
    h = h * expf(dh);
    float x1 = cx - 0.5f * w, y1 = cy - 0.5f * h;
    float x2 = x1 + w, y2 = y1 + h;
    if (H > 0 && W > 0) {
      x1 = fmaxf(fminf(x1, mx), 0.0f); y1 = fmaxf(fminf(y1, my), 0.0f);
      x2 = fmaxf(fminf(x2, mx), 0.0f); y2 = fmaxf(fminf(y2, my), 0.0f);
    }
    out[4 * i + 0] = x1; out[4 * i + 1] = y1; out[4 * i + 2] = x2; out[4 * i + 3] = y2;
  }
}

typedef struct { float s; int i; } cand_t;
static int cand_cmp(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);   /* ties: lower index first */
}

static int iou_gt(const float* b, int i, int j, float thr) {
  const float lo0i = fminf(b[4*i], b[4*i+2]), hi0i = fmaxf(b[4*i], b[4*i+2]);
  const float lo1i = fminf(b[4*i+1], b[4*i+3]), hi1i = fmaxf(b[4*i+1], b[4*i+3]);
  const float lo0j = fminf(b[4*j], b[4*j+2]), hi0j = fmaxf(b[4*j], b[4*j+2]);
  const float lo1j = fminf(b[4*j+1], b[4*j+3]), hi1j = fmaxf(b[4*j+1], b[4*j+3]);
  const float ai = (hi0i - lo0i) * (hi1i - lo1i), aj = (hi0j - lo0j) * (hi1j - lo1j);
  if (ai <= 0.0f || aj <= 0.0f) return 0;
  const float i0 = fmaxf(fminf(hi0i, hi0j) - fmaxf(lo0i, lo0j), 0.0f);
  const float i1 = fmaxf(fminf(hi1i, hi1j) - fmaxf(lo1i, lo1j), 0.0f);
  const float inter = i0 * i1;
  const float iou = inter / (ai + aj - inter);
  return iou > thr;
}

/* returns the number of kept indices written to out_idx (selection order); top_k > 0 limits the candidates */
int orc_nms(const float* boxes, const float* scores, int n, int top_k, int max_out, float thr, int* out_idx) {
  cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) { c[i].s = scores[i] + 0.0f; c[i].i = i; }
  qsort(c, (size_t)n, sizeof(cand_t), cand_cmp);
  const int lim = (top_k > 0 && top_k < n) ? top_k : n;
  int kept = 0;
  for (int q = 0; q < lim && kept < max_out; ++q) {
    const int i = c[q].i;
    int keep = 1;
    for (int k = kept - 1; k >= 0; --k)          /* most recently selected first, as TF does */
      if (iou_gt(boxes, i, out_idx[k], thr)) { keep = 0; break; }
    if (keep) out_idx[kept++] = i;
  }
  free(c);
  return kept;
}

void orc_crop_and_resize(const float* image, int b, int h, int w, int c, const float* boxes, const int* box_ind,
                         int r, int ch, int cw, float ext, float* out) {
  (void)b;
#pragma omp parallel for schedule(dynamic, 4)
  for (int k = 0; k < r; ++k) {
    const float y1 = boxes[4*k], x1 = boxes[4*k+1], y2 = boxes[4*k+2], x2 = boxes[4*k+3];
    const float* img = image + (size_t)box_ind[k] * h * w * c;
    const float hs = (ch > 1) ? (y2 - y1) * (float)(h - 1) / (float)(ch - 1) : 0.0f;
    const float ws = (cw > 1) ? (x2 - x1) * (float)(w - 1) / (float)(cw - 1) : 0.0f;
    for (int y = 0; y < ch; ++y) {
      float* orow = out + (((size_t)k * ch + y) * cw) * c;
      const float in_y = (ch > 1) ? y1 * (float)(h - 1) + (float)y * hs : 0.5f * (y1 + y2) * (float)(h - 1);
      if (in_y < 0 || in_y > (float)(h - 1)) {
        for (int e = 0; e < cw * c; ++e) orow[e] = ext;
        continue;
      }
      const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
      const float ly = in_y - (float)top;
      for (int x = 0; x < cw; ++x) {
        float* o = orow + (size_t)x * c;
        const float in_x = (cw > 1) ? x1 * (float)(w - 1) + (float)x * ws : 0.5f * (x1 + x2) * (float)(w - 1);
        if (in_x < 0 || in_x > (float)(w - 1)) {
          for (int d = 0; d < c; ++d) o[d] = ext;
          continue;
        }
        const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
        const float lx = in_x - (float)left;
        const float* tl = img + ((size_t)top * w + left) * c;
        const float* tr = img + ((size_t)top * w + right) * c;
        const float* bl = img + ((size_t)bot * w + left) * c;
        const float* br = img + ((size_t)bot * w + right) * c;
        for (int d = 0; d < c; ++d) {
          const float t = tl[d] + (tr[d] - tl[d]) * lx;
          const float bb = bl[d] + (br[d] - bl[d]) * lx;
          o[d] = t + (bb - t) * ly;
        }
      }
    }
  }
}

void orc_max_pool_2x2(const float* in, int n, int h, int w, int c, float* out) {
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < n; ++k)
    for (int i = 0; i < oh; ++i)
      for (int j = 0; j < ow; ++j) {
        float* o = out + (((size_t)k * oh + i) * ow + j) * c;
        for (int d = 0; d < c; ++d) {
          float m = -INFINITY;
          for (int di = 0; di < 2 && 2*i+di < h; ++di)
            for (int dj = 0; dj < 2 && 2*j+dj < w; ++dj) {
              const float v = in[(((size_t)k * h + 2*i+di) * w + 2*j+dj) * c + d];
              m = v > m ? v : m;
            }
          o[d] = m;
        }
      }
}

/* model/roi_pooling.py:53-90; box_ind NULL = all zeros.  Returns 0, or -1 on allocation failure. */
int orc_roi_pool_c4(const float* feat, int b, int h, int w, int c, const float* rois, const int* box_ind, int r,
                    float stride, int P, int max_flag, float* out) {
  float* nb = (float*)malloc(sizeof(float) * 4 * (size_t)(r > 0 ? r : 1));
  int* bi = (int*)calloc((size_t)(r > 0 ? r : 1), sizeof(int));
  if (!nb || !bi) return -1;
  for (int k = 0; k < r; ++k) {
    const float x1 = rois[4*k] / stride, y1 = rois[4*k+1] / stride, x2 = rois[4*k+2] / stride, y2 = rois[4*k+3] / stride;
    nb[4*k] = y1 / (float)(h - 1); nb[4*k+1] = x1 / (float)(w - 1);
    nb[4*k+2] = y2 / (float)(h - 1); nb[4*k+3] = x2 / (float)(w - 1);
    if (box_ind) bi[k] = box_ind[k];
  }
  int rc = 0;
  if (max_flag) {
    const int Q = 2 * P;
    float* tmp = (float*)malloc(sizeof(float) * (size_t)r * Q * Q * c);
    if (!tmp) rc = -1;
    else {
      orc_crop_and_resize(feat, b, h, w, c, nb, bi, r, Q, Q, 0.0f, tmp);
      orc_max_pool_2x2(tmp, r, Q, Q, c, out);
      free(tmp);
    }
  } else {
    orc_crop_and_resize(feat, b, h, w, c, nb, bi, r, P, P, 0.0f, out);
  }
  free(nb); free(bi);
  return rc;
}

/* The bench composite on the host: per image decode+clip -> (top-k) -> NMS (one thread per image, TF's NMS kernel is
 * single-threaded; images of the batch run on different threads), then crop_and_resize sharded over all boxes.
 * out_rois [batch,post,4] zero padded, out_idx [batch,post] -1 padded, out_count [batch], out_feat [batch*post,P,P,c]. */
int orc_c4_proposal_roi(const float* anchors, const float* deltas, const float* scores, const float* feat, int batch,
                        int n, int fh, int fw, int c, const float* means, const float* stds, int H, int W,
                        int pre_nms_top_k, int post_nms, float thr, float stride, int P, int max_flag,
                        float* out_rois, int* out_idx, int* out_count, float* out_feat) {
  int fail = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < batch; ++b) {
    float* dec = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    if (!dec) { fail = 1; continue; }
    orc_decode_clip(anchors, deltas + (size_t)b * n * 4, n, means, stds, H, W, dec);
    int* idx = out_idx + (size_t)b * post_nms;
    const int kept = orc_nms(dec, scores + (size_t)b * n, n, pre_nms_top_k, post_nms, thr, idx);
    out_count[b] = kept;
    float* ro = out_rois + (size_t)b * post_nms * 4;
    for (int k = 0; k < post_nms; ++k) {
      if (k < kept) memcpy(ro + 4*k, dec + 4*(size_t)idx[k], sizeof(float) * 4);
      else { memset(ro + 4*k, 0, sizeof(float) * 4); idx[k] = -1; }
    }
    free(dec);
  }
  if (fail) return -1;
  const int r = batch * post_nms;
  int* bi = (int*)malloc(sizeof(int) * (size_t)r);
  if (!bi) return -1;
  for (int k = 0; k < r; ++k) bi[k] = k / post_nms;
  const int rc = orc_roi_pool_c4(feat, batch, fh, fw, c, out_rois, bi, r, stride, P, max_flag, out_feat);
  /* padded rois (k >= count) pool the all-zero box; zero them like the GPU path does */
  for (int b = 0; b < batch; ++b)
    for (int k = out_count[b]; k < post_nms; ++k)
      memset(out_feat + ((size_t)b * post_nms + k) * P * P * c, 0, sizeof(float) * (size_t)P * P * c);
  free(bi);
  return rc;
}

/* FPN composite on the host (bench.py --workload cfg3|cfg5): per image decode+clip -> NMS over the concatenated P2..P6
 * anchors (model/region_proposal.py:37-81), then level assignment (fpn/base_fpn_model.py:303-324) and, per level,
 * crop_and_resize 2P x 2P on boxes normalised by the image size + 2x2 max pool (model/roi_pooling.py:15-42), written
 * level-major over the whole batch like bx_fpn_roi_features.  feats[l] = [batch, fh[l], fw[l], c], 4 levels. */
int orc_fpn_proposal_roi(const float* anchors, const float* deltas, const float* scores, const float* const* feats,
                         const int* fh, const int* fw, int batch, int n, int c, const float* means, const float* stds,
                         int H, int W, int pre_nms_top_k, int post_nms, float thr, int P, float* out_rois,
                         int* out_idx, int* out_count, float* out_feat, int* out_order) {
  int fail = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < batch; ++b) {
    float* dec = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    if (!dec) { fail = 1; continue; }
    orc_decode_clip(anchors, deltas + (size_t)b * n * 4, n, means, stds, H, W, dec);
    int* idx = out_idx + (size_t)b * post_nms;
    const int kept = orc_nms(dec, scores + (size_t)b * n, n, pre_nms_top_k, post_nms, thr, idx);
    out_count[b] = kept;
    float* ro = out_rois + (size_t)b * post_nms * 4;
    for (int k = 0; k < post_nms; ++k) {
      if (k < kept) memcpy(ro + 4*k, dec + 4*(size_t)idx[k], sizeof(float) * 4);
      else { memset(ro + 4*k, 0, sizeof(float) * 4); idx[k] = -1; }
    }
    free(dec);
  }
  if (fail) return -1;
  const int r = batch * post_nms, Q = 2 * P;
  int* lvl = (int*)malloc(sizeof(int) * (size_t)(r > 0 ? r : 1));
  float* nb = (float*)malloc(sizeof(float) * 4 * (size_t)(r > 0 ? r : 1));
  int* bi = (int*)malloc(sizeof(int) * (size_t)(r > 0 ? r : 1));
  if (!lvl || !nb || !bi) return -1;
  for (int k = 0; k < r; ++k) {
    const float* q = out_rois + 4 * (size_t)k;
    const float hh = fmaxf(0.0f, q[3] - q[1]), ww = fmaxf(0.0f, q[2] - q[0]);
    float lv = floorf(4.0f + logf(sqrtf(ww * hh + 1e-8f) / 224.0f) / logf(2.0f));
    lv = fminf(fmaxf(lv, 2.0f), 5.0f);
    lvl[k] = (int)lv - 2;
  }
  int pos = 0, rc = 0;
  for (int l = 0; l < 4 && !rc; ++l) {
    int m = 0;
    for (int k = 0; k < r; ++k)
      if (lvl[k] == l) {
        const float* q = out_rois + 4 * (size_t)k;
        nb[4*m] = q[1] / (float)H; nb[4*m+1] = q[0] / (float)W; nb[4*m+2] = q[3] / (float)H; nb[4*m+3] = q[2] / (float)W;
        bi[m] = k / post_nms;
        out_order[pos + m] = k;
        ++m;
      }
    if (m) {
      float* tmp = (float*)malloc(sizeof(float) * (size_t)m * Q * Q * c);
      if (!tmp) { rc = -1; break; }
      orc_crop_and_resize(feats[l], batch, fh[l], fw[l], c, nb, bi, m, Q, Q, 0.0f, tmp);
      orc_max_pool_2x2(tmp, m, Q, Q, c, out_feat + (size_t)pos * P * P * c);
      free(tmp);
    }
    pos += m;
  }
  free(lvl); free(nb); free(bi);
  return rc;
}
