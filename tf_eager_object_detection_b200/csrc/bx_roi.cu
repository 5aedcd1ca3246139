// RoI feature extraction: tf.image.crop_and_resize (bilinear, extrapolation 0) fused with the 2x2 max / avg pool and
// with FPN level routing.  Replaces model/roi_pooling.py:8-42 (RoiPoolingCropAndResize2), :45-90
// (RoiPoolingCropAndResize), :93-176 (crop_and_resize(pad_border)/roi_align/RoiPoolingRoiAlign) and
// model/fpn/base_fpn_model.py:152-161,303-324 (_get_roi_features, _assign_levels).
//
// Layout: features NHWC fp32, so the channel axis is the coalesced / float4 axis.  Sample coordinates follow the TF r1.13
// kernel's fp32 op order: in_y = y1*(h-1) + y*((y2-y1)*(h-1)/(Q-1)); a sample outside
// [0,h-1]x[0,w-1] is 0.  Three forward kernels share that arithmetic bit for bit:
//   * roi_band_kernel (bx_roi_band.cu)  plain crops: feature band stationary in shared memory (TMA), the headline kernel;
//   * roi_pool2_kernel (here)           pooled crops (2x2 max / avg): one CTA per roi, shared taps of the 2x2 sample block
//                                       loaded once, per-pixel records, packed fp32x2 lerps; FPN level routing reads
//                                       `order` / `level` so the output lands in the reference's concatenated order;
//   * roi_pool_kernel (here)            generic gather for every other shape (C % 4 != 0, wide maps, extrapolation != 0).
// roi_pool_grad_kernel is the backward w.r.t. the feature map ("next" row f3).
#include <stdlib.h>

#include "bx_roi.cuh"

using namespace bxroi;

namespace {

// kWhole = false: one CTA per (roi, pooled row).  kWhole = true: one CTA per roi, all pooled rows in order — adjacent
// rows share feature rows, which then hit in L1 instead of crossing L2 -> SM again (the FPN extractor, 16 taps per
// output, is bound by exactly that traffic).
template <int POOL, typename VecT, bool kWhole>
__global__ void __launch_bounds__(256) roi_pool_kernel(const RoiArgs a) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;  // crop samples per output pixel per axis
  constexpr int V = sizeof(VecT) / sizeof(float);
  __shared__ Axis ax_y[kWhole ? kMaxQ : 2];
  __shared__ Axis ax_x[kMaxQ];
  __shared__ int s_meta[4];  // image index, level, zero-fill flag

  const int P = a.P, Q = a.Q;
  const int j = kWhole ? blockIdx.x : blockIdx.x / P;   // output roi row
  const int py = kWhole ? 0 : blockIdx.x % P;
  const int tid = threadIdx.x;
  const int cv = a.c / V;
  VecT* out_row = reinterpret_cast<VecT*>(a.out + (static_cast<size_t>(j) * P + py) * P * a.c);

  if (tid < (kWhole ? 2 * Q : Q + 2)) {
    const int src = a.order ? a.order[j] : j;
    const int lvl = a.level ? a.level[src] - a.level_base : 0;
    int img = a.box_ind ? a.box_ind[src] : 0;
    int zero = 0;
    if (a.roi_counts) {
      img = src / a.rois_per_image;
      zero = (src % a.rois_per_image) >= a.roi_counts[img];
    }
    if (tid == 0) {
      s_meta[0] = img;
      s_meta[1] = lvl;
      s_meta[2] = zero || img < 0 || img >= a.b;
    }
    const float4 roi = a.rois[src];
    const int fh = a.lv[lvl].fh, fw = a.lv[lvl].fw;
    const NormBox nb = roi_norm_box(a, roi, fh, fw);
    const float y1n = nb.y1, x1n = nb.x1, y2n = nb.y2, x2n = nb.x2;
    const int dimy = nb.dimy, dimx = nb.dimx, pad = nb.pad;
    if (tid < Q) ax_x[tid] = sample_axis(x1n, x2n, tid, Q, dimx, pad);
    else if (kWhole) ax_y[tid - Q] = sample_axis(y1n, y2n, tid - Q, Q, dimy, pad);
    else if (tid - Q < S) ax_y[tid - Q] = sample_axis(y1n, y2n, py * S + (tid - Q), Q, dimy, pad);
  }
  __syncthreads();

  const int row_items = P * cv;
  const int items = kWhole ? P * row_items : row_items;
  if (s_meta[2]) {
    VecT z;
    float* zp = reinterpret_cast<float*>(&z);
#pragma unroll
    for (int v = 0; v < V; ++v) zp[v] = 0.0f;
    for (int it = tid; it < items; it += 256) out_row[it] = z;
    return;
  }
  const LevelFeat lf = a.lv[s_meta[1]];
  const VecT* feat = reinterpret_cast<const VecT*>(lf.feat + static_cast<size_t>(s_meta[0]) * lf.fh * lf.fw * a.c);
  const float ext = a.extrapolation;

  for (int it = tid; it < items; it += 256) {
    const int prow = kWhole ? it / row_items : 0;
    const int rit = it - prow * row_items;
    const int px = rit / cv, cg = rit % cv;
    float acc[V];
#pragma unroll
    for (int sy = 0; sy < S; ++sy) {
      const Axis ay = ax_y[prow * S + sy];
#pragma unroll
      for (int sx = 0; sx < S; ++sx) {
        const Axis axx = ax_x[px * S + sx];
        float val[V];
        if (ay.valid && axx.valid) {
          const VecT tl = __ldg(feat + (static_cast<size_t>(ay.lo) * lf.fw + axx.lo) * cv + cg);
          const VecT tr = __ldg(feat + (static_cast<size_t>(ay.lo) * lf.fw + axx.hi) * cv + cg);
          const VecT bl = __ldg(feat + (static_cast<size_t>(ay.hi) * lf.fw + axx.lo) * cv + cg);
          const VecT br = __ldg(feat + (static_cast<size_t>(ay.hi) * lf.fw + axx.hi) * cv + cg);
          const float* ptl = reinterpret_cast<const float*>(&tl);
          const float* ptr = reinterpret_cast<const float*>(&tr);
          const float* pbl = reinterpret_cast<const float*>(&bl);
          const float* pbr = reinterpret_cast<const float*>(&br);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float top = ptl[v] + (ptr[v] - ptl[v]) * axx.lerp;
            const float bot = pbl[v] + (pbr[v] - pbl[v]) * axx.lerp;
            val[v] = top + (bot - top) * ay.lerp;
          }
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) val[v] = ext;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          if (sy == 0 && sx == 0) acc[v] = val[v];
          else if (POOL == BX_POOL_MAX2) acc[v] = fmaxf(acc[v], val[v]);
          else acc[v] = acc[v] + val[v];  // AVG2: ((s00+s01)+s10)+s11
        }
      }
    }
    VecT o;
    float* po = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int v = 0; v < V; ++v) po[v] = (POOL == BX_POOL_AVG2) ? acc[v] / 4.0f : acc[v];
    out_row[it] = o;
  }
}


// ---- pooled extractors (2x2 max / avg over a 2P x 2P crop) on wide maps: the FPN extractor and every C4 shape the
// band kernel does not take.  One CTA per roi; thread = (channel group of 4, pixel lane) so that a warp works on one
// output pixel and every table lookup and branch below is warp-uniform.  Versus the generic kernel it
//   * x-interpolates each feature row once per pixel and reuses it: the lower row of sample row 2py is the upper row
//     of sample row 2py+1 whenever the two fall into adjacent pixel intervals (same loads, same operands, same
//     rounding), likewise the shared pixel column of the two x samples: 9-16 tap loads instead of 16;
//   * does the lerps as packed fp32x2 (FADD2 / FFMA2 with the -0.0 addend), 32-bit tap offsets from tables that are
//     pre-multiplied by the row / pixel pitch.
// Op order per sample is TF's (top = tl + (tr-tl)*lx; bot likewise; top + (bot-top)*ly), pooling order s00,s01,s10,s11.
#ifndef BX_POOL2_CTAS
#define BX_POOL2_CTAS 3
#endif
struct TapEnt {
  int lo, hi;    // offsets in float4 units (x: pixel*cv, y: row*fw*cv)
  float lerp;
  int valid;
};

struct F4x2 { ulonglong2 a, b; };   // x-lerped row values for the two x samples of a pixel

// everything one output pixel needs, precomputed once per roi so that the per-pixel dependency chain in front of the
// tap loads is one 64-byte shared-memory record (no division, no table indexing, no compares)
constexpr int kPool2MaxPix = 64;    // P * P
struct __align__(16) PixRec {
  int yo[4];                        // row offsets: y0.lo, y0.hi, y1.lo, y1.hi   (float4 units)
  int xo[4];                        // pixel offsets: x0.lo, x0.hi, x1.lo, x1.hi
  float wx0, wx1, wy0, wy1;
  int flags;                        // 1: all four samples valid, 2: x0.hi == x1.lo, 4: y0.hi == y1.lo, 16/32/64/128: y0/y1/x0/x1 valid
  int pad[3];
};

// L2 prefetch of the feature pixels a roi will sample (its bounding box on its level: up to ~30 rows of w_f * C * 4
// contiguous bytes in the NHWC map), one cp.async.bulk.prefetch.L2 per footprint row, issued by one warp.  The kernel is
// bound by the latency of the compulsory DRAM misses its tap loads take in-line (ncu: long-scoreboard stalls, every
// iteration waits for one DRAM round trip); prefetching the footprint of the roi that will run `dist` CTAs later turns
// them into L2 hits.  Pure hint: no effect on results.  (Prefetching FOR LATER CTAs measured slower — see the launch site.)
__device__ __forceinline__ void prefetch_footprint(const RoiArgs& a, int jp, int lane) {
  if (jp >= a.r) return;
  const int src = a.order ? a.order[jp] : jp;
  const int lvl = a.level ? a.level[src] - a.level_base : 0;
  int img = a.box_ind ? a.box_ind[src] : 0;
  if (a.roi_counts) {
    img = src / a.rois_per_image;
    if ((src % a.rois_per_image) >= a.roi_counts[img]) return;
  }
  if (img < 0 || img >= a.b) return;
  const int fh = a.lv[lvl].fh, fw = a.lv[lvl].fw;
  const NormBox nb = roi_norm_box(a, a.rois[src], fh, fw);
  const float dy = static_cast<float>(nb.dimy - 1), dx = static_cast<float>(nb.dimx - 1);
  // conservative bounding box of the sample coordinates (un-padded map indices), clamped to the map
  const float ya = fminf(nb.y1, nb.y2) * dy - static_cast<float>(nb.pad), yb = fmaxf(nb.y1, nb.y2) * dy - static_cast<float>(nb.pad);
  const float xa = fminf(nb.x1, nb.x2) * dx - static_cast<float>(nb.pad), xb = fmaxf(nb.x1, nb.x2) * dx - static_cast<float>(nb.pad);
  if (!(ya <= static_cast<float>(fh) && yb >= -1.0f && xa <= static_cast<float>(fw) && xb >= -1.0f)) return;   // also NaN
  const int y_lo = max(static_cast<int>(floorf(fmaxf(ya, 0.0f))), 0), y_hi = min(static_cast<int>(ceilf(fminf(yb, static_cast<float>(fh - 1)))), fh - 1);
  const int x_lo = max(static_cast<int>(floorf(fmaxf(xa, 0.0f))), 0), x_hi = min(static_cast<int>(ceilf(fminf(xb, static_cast<float>(fw - 1)))), fw - 1);
  if (y_hi < y_lo || x_hi < x_lo) return;
  const unsigned bytes = static_cast<unsigned>(x_hi - x_lo + 1) * static_cast<unsigned>(a.c) * 4u;
  const float* base = a.lv[lvl].feat + (static_cast<size_t>(img) * fh * fw + x_lo) * a.c;
  for (int y = y_lo + lane; y <= y_hi; y += 32) {
    const float* p = base + static_cast<size_t>(y) * fw * a.c;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
  }
}

template <int POOL>
__global__ void __launch_bounds__(256, BX_POOL2_CTAS) roi_pool2_kernel(const RoiArgs a, const float neg_zero, const int pf_dist) {
  __shared__ PixRec recs[kPool2MaxPix];
  __shared__ int s_meta[4];
  const int P = a.P, Q = a.Q;
  const int j = blockIdx.x, tid = threadIdx.x;
  const int cv = a.c >> 2;
  if (pf_dist >= 0 && tid >= 224) prefetch_footprint(a, j + pf_dist, tid - 224);   // last warp: not a record builder (P*P <= 64)
  if (tid < P * P) {                             // one thread per output pixel builds its record (single barrier)
    const int src = a.order ? a.order[j] : j;
    const int lvl = a.level ? a.level[src] - a.level_base : 0;
    int img = a.box_ind ? a.box_ind[src] : 0;
    int zero = 0;
    if (a.roi_counts) {
      img = src / a.rois_per_image;
      zero = (src % a.rois_per_image) >= a.roi_counts[img];
    }
    if (tid == 0) {
      s_meta[0] = img;
      s_meta[1] = lvl;
      s_meta[2] = zero || img < 0 || img >= a.b;
    }
    const int fh = a.lv[lvl].fh, fw = a.lv[lvl].fw;
    const NormBox nb = roi_norm_box(a, a.rois[src], fh, fw);
    const int prow = tid / P, px = tid - prow * P;
    const Axis y0 = sample_axis(nb.y1, nb.y2, 2 * prow, Q, nb.dimy, nb.pad);
    const Axis y1 = sample_axis(nb.y1, nb.y2, 2 * prow + 1, Q, nb.dimy, nb.pad);
    const Axis x0 = sample_axis(nb.x1, nb.x2, 2 * px, Q, nb.dimx, nb.pad);
    const Axis x1 = sample_axis(nb.x1, nb.x2, 2 * px + 1, Q, nb.dimx, nb.pad);
    const int pitch = fw * cv;
    PixRec r;
    r.yo[0] = y0.valid ? y0.lo * pitch : 0; r.yo[1] = y0.valid ? y0.hi * pitch : 0;
    r.yo[2] = y1.valid ? y1.lo * pitch : 0; r.yo[3] = y1.valid ? y1.hi * pitch : 0;
    r.xo[0] = x0.valid ? x0.lo * cv : 0; r.xo[1] = x0.valid ? x0.hi * cv : 0;
    r.xo[2] = x1.valid ? x1.lo * cv : 0; r.xo[3] = x1.valid ? x1.hi * cv : 0;
    r.wx0 = x0.lerp; r.wx1 = x1.lerp; r.wy0 = y0.lerp; r.wy1 = y1.lerp;
    r.flags = ((y0.valid && y1.valid && x0.valid && x1.valid) ? 1 : 0) | ((r.xo[1] == r.xo[2]) ? 2 : 0) |
              ((r.yo[1] == r.yo[2]) ? 4 : 0) | (y0.valid ? 16 : 0) | (y1.valid ? 32 : 0) | (x0.valid ? 64 : 0) |
              (x1.valid ? 128 : 0);
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
    recs[tid] = r;
  }
  __syncthreads();
  const int cg = tid % cv, grp = tid / cv;
  const int groups = 256 / cv;                   // pixels in flight per CTA (launch guarantees 256 % cv == 0, cv >= 32)
  float4* out = reinterpret_cast<float4*>(a.out) + static_cast<size_t>(j) * P * P * cv + cg;
  if (s_meta[2]) {
    for (int pix = grp; pix < P * P; pix += groups) out[static_cast<size_t>(pix) * cv] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const LevelFeat lf = a.lv[s_meta[1]];
  const ulonglong2* feat = reinterpret_cast<const ulonglong2*>(lf.feat) +
                           static_cast<size_t>(s_meta[0]) * lf.fh * lf.fw * cv + cg;
  const unsigned long long nz = f2_splat(neg_zero);
  const float ext = a.extrapolation;

  for (int pix = grp; pix < P * P; pix += groups) {
    const PixRec rec = recs[pix];
    const TapEnt y0 = {rec.yo[0], rec.yo[1], rec.wy0, rec.flags & 16}, y1 = {rec.yo[2], rec.yo[3], rec.wy1, rec.flags & 32};
    const TapEnt x0 = {rec.xo[0], rec.xo[1], rec.wx0, rec.flags & 64}, x1 = {rec.xo[2], rec.xo[3], rec.wx1, rec.flags & 128};
    ulonglong2 res;
    if (rec.flags & 1) {
      const unsigned long long w0 = f2_splat(x0.lerp), w1 = f2_splat(x1.lerp);
      const bool xsh = (rec.flags & 2) != 0;
      const bool ysh = (rec.flags & 4) != 0;
      // all taps of the pixel first (9-16 independent 16-byte loads in flight per thread), then the arithmetic
      const ulonglong2* ra = feat + y0.lo;
      const ulonglong2* rb = feat + y0.hi;
      const ulonglong2* rc = feat + y1.lo;
      const ulonglong2* rd = feat + y1.hi;
      const ulonglong2 a0 = __ldg(ra + x0.lo), a1 = __ldg(ra + x0.hi), a3 = __ldg(ra + x1.hi);
      const ulonglong2 b0 = __ldg(rb + x0.lo), b1 = __ldg(rb + x0.hi), b3 = __ldg(rb + x1.hi);
      const ulonglong2 d0 = __ldg(rd + x0.lo), d1 = __ldg(rd + x0.hi), d3 = __ldg(rd + x1.hi);
      ulonglong2 a2 = a1, b2 = b1, d2 = d1, c0 = b0, c1 = b1, c2 = b1, c3 = b3;
      if (!xsh) {
        a2 = __ldg(ra + x1.lo);
        b2 = __ldg(rb + x1.lo);
        d2 = __ldg(rd + x1.lo);
      }
      if (!ysh) {
        c0 = __ldg(rc + x0.lo);
        c1 = __ldg(rc + x0.hi);
        c3 = __ldg(rc + x1.hi);
        c2 = c1;
        if (!xsh) c2 = __ldg(rc + x1.lo);
      } else {
        c2 = b2;
      }
      // one feature row, x-interpolated at both x samples of the pixel
      auto row = [&](const ulonglong2 p0, const ulonglong2 p1, const ulonglong2 p2, const ulonglong2 p3) {
        F4x2 h;
        h.a.x = f2_add(p0.x, f2_mul(f2_sub(p1.x, p0.x), w0, nz));
        h.a.y = f2_add(p0.y, f2_mul(f2_sub(p1.y, p0.y), w0, nz));
        h.b.x = f2_add(p2.x, f2_mul(f2_sub(p3.x, p2.x), w1, nz));
        h.b.y = f2_add(p2.y, f2_mul(f2_sub(p3.y, p2.y), w1, nz));
        return h;
      };
      const F4x2 r0 = row(a0, a1, a2, a3), r1 = row(b0, b1, b2, b3);
      F4x2 r2 = r1;
      if (!ysh) r2 = row(c0, c1, c2, c3);
      const F4x2 r3 = row(d0, d1, d2, d3);
      const unsigned long long v0 = f2_splat(y0.lerp), v1 = f2_splat(y1.lerp);
      ulonglong2 s00, s01, s10, s11;
      s00.x = f2_add(r0.a.x, f2_mul(f2_sub(r1.a.x, r0.a.x), v0, nz));
      s00.y = f2_add(r0.a.y, f2_mul(f2_sub(r1.a.y, r0.a.y), v0, nz));
      s01.x = f2_add(r0.b.x, f2_mul(f2_sub(r1.b.x, r0.b.x), v0, nz));
      s01.y = f2_add(r0.b.y, f2_mul(f2_sub(r1.b.y, r0.b.y), v0, nz));
      s10.x = f2_add(r2.a.x, f2_mul(f2_sub(r3.a.x, r2.a.x), v1, nz));
      s10.y = f2_add(r2.a.y, f2_mul(f2_sub(r3.a.y, r2.a.y), v1, nz));
      s11.x = f2_add(r2.b.x, f2_mul(f2_sub(r3.b.x, r2.b.x), v1, nz));
      s11.y = f2_add(r2.b.y, f2_mul(f2_sub(r3.b.y, r2.b.y), v1, nz));
      res = pool2<POOL>(pool2<POOL>(pool2<POOL>(s00, s01), s10), s11);
    } else {
      // a sample outside the map (extrapolation value): plain per-sample form
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int sy = 0; sy < 2; ++sy) {
        const TapEnt ye = sy ? y1 : y0;
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
          const TapEnt xe = sx ? x1 : x0;
          float4 v = make_float4(ext, ext, ext, ext);
          if (ye.valid && xe.valid) {
            const ulonglong2 tl = __ldg(feat + ye.lo + xe.lo), tr = __ldg(feat + ye.lo + xe.hi);
            const ulonglong2 bl = __ldg(feat + ye.hi + xe.lo), br = __ldg(feat + ye.hi + xe.hi);
            const unsigned long long wx = f2_splat(xe.lerp), wy = f2_splat(ye.lerp);
            ulonglong2 t, b, o;
            t.x = f2_add(tl.x, f2_mul(f2_sub(tr.x, tl.x), wx, nz));
            t.y = f2_add(tl.y, f2_mul(f2_sub(tr.y, tl.y), wx, nz));
            b.x = f2_add(bl.x, f2_mul(f2_sub(br.x, bl.x), wx, nz));
            b.y = f2_add(bl.y, f2_mul(f2_sub(br.y, bl.y), wx, nz));
            o.x = f2_add(t.x, f2_mul(f2_sub(b.x, t.x), wy, nz));
            o.y = f2_add(t.y, f2_mul(f2_sub(b.y, t.y), wy, nz));
            v = *reinterpret_cast<const float4*>(&o);
          }
          if (sy == 0 && sx == 0) acc = v;
          else if (POOL == BX_POOL_MAX2) acc = make_float4(fmaxf(acc.x, v.x), fmaxf(acc.y, v.y), fmaxf(acc.z, v.z), fmaxf(acc.w, v.w));
          else acc = make_float4(acc.x + v.x, acc.y + v.y, acc.z + v.z, acc.w + v.w);
        }
      }
      res = *reinterpret_cast<const ulonglong2*>(&acc);
    }
    float4 o = *reinterpret_cast<const float4*>(&res);
    if (POOL == BX_POOL_AVG2) o = make_float4(o.x / 4.0f, o.y / 4.0f, o.z / 4.0f, o.w / 4.0f);
    out[static_cast<size_t>(pix) * cv] = o;
  }
}

// ---- backward of the RoI extractors w.r.t. the feature map ("next" row f3): gradient of tf.image.crop_and_resize
//      (TF CropAndResizeGradImage: each sample scatters (1-ly)(1-lx), (1-ly)lx, ly(1-lx), ly lx times its gradient to its 4
//      taps) composed with the gradient of the 2x2 pool (max: all of it to the first maximal sample of the window in
//      row-major order, as TF's MaxPoolGrad; avg: a quarter to each).  The boxes get no gradient (tf.stop_gradient,
//      roi_pooling.py:37,79,86).  One CTA per (roi, py); fp32 atomics (red.global.add.v4.f32).  This scatter form is the
//      fallback (crops wider than 32 samples) and the A/B partner (BX_ROI_GRAD_ATOMIC=1) of the row-owned, atomic-free
//      kernel in bx_roi_grad.cu, which bx_roi_pool_grad uses by default.
template <int POOL>
__global__ void __launch_bounds__(256) roi_pool_grad_kernel(const RoiGradArgs g) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;
  const RoiArgs& a = g.a;
  __shared__ Axis ax_y[2];
  __shared__ Axis ax_x[kMaxQ];
  __shared__ int s_meta[4];
  const int P = a.P, Q = a.Q;
  const int j = blockIdx.x / P, py = blockIdx.x % P;
  const int tid = threadIdx.x;
  const int cv = a.c / 4;
  if (tid < Q + 2) {
    int img = a.box_ind ? a.box_ind[j] : 0;
    int zero = 0;
    if (a.roi_counts) {
      img = j / a.rois_per_image;
      zero = (j % a.rois_per_image) >= a.roi_counts[img];
    }
    if (tid == 0) {
      s_meta[0] = img;
      s_meta[2] = zero || img < 0 || img >= a.b;
    }
    const NormBox nb = roi_norm_box(a, a.rois[j], a.lv[0].fh, a.lv[0].fw);
    if (tid < Q) ax_x[tid] = sample_axis(nb.x1, nb.x2, tid, Q, nb.dimx, nb.pad);
    else if (tid - Q < S) ax_y[tid - Q] = sample_axis(nb.y1, nb.y2, py * S + (tid - Q), Q, nb.dimy, nb.pad);
  }
  __syncthreads();
  if (s_meta[2]) return;   // padded roi: its output was a constant, no gradient
  const LevelFeat lf = a.lv[0];
  const size_t img_off = static_cast<size_t>(s_meta[0]) * lf.fh * lf.fw * a.c;
  const float4* feat = reinterpret_cast<const float4*>(lf.feat + img_off);
  float4* gfeat = reinterpret_cast<float4*>(g.grad_feat + img_off);
  const float4* gout = reinterpret_cast<const float4*>(g.grad_out + (static_cast<size_t>(j) * P + py) * P * a.c);
  const int items = P * cv;
  for (int it = tid; it < items; it += 256) {
    const int px = it / cv, cg = it % cv;
    const float4 go = gout[it];
    float gs[S * S][4];   // gradient routed to each of the S*S samples of this output pixel, per channel
    if (POOL == BX_POOL_NONE) {
      gs[0][0] = go.x; gs[0][1] = go.y; gs[0][2] = go.z; gs[0][3] = go.w;
    } else if (POOL == BX_POOL_AVG2) {
#pragma unroll
      for (int s = 0; s < S * S; ++s) { gs[s][0] = go.x / 4.0f; gs[s][1] = go.y / 4.0f; gs[s][2] = go.z / 4.0f; gs[s][3] = go.w / 4.0f; }
    } else {
      // recompute the four sample values and give the gradient to the first maximum per channel
      float best[4];
      int arg[4];
#pragma unroll
      for (int sy = 0; sy < S; ++sy)
#pragma unroll
        for (int sx = 0; sx < S; ++sx) {
          const Axis ay = ax_y[sy], axx = ax_x[px * S + sx];
          float val[4] = {a.extrapolation, a.extrapolation, a.extrapolation, a.extrapolation};
          if (ay.valid && axx.valid) {
            const float4 tl = __ldg(feat + (static_cast<size_t>(ay.lo) * lf.fw + axx.lo) * cv + cg);
            const float4 tr = __ldg(feat + (static_cast<size_t>(ay.lo) * lf.fw + axx.hi) * cv + cg);
            const float4 bl = __ldg(feat + (static_cast<size_t>(ay.hi) * lf.fw + axx.lo) * cv + cg);
            const float4 br = __ldg(feat + (static_cast<size_t>(ay.hi) * lf.fw + axx.hi) * cv + cg);
            const float* ptl = reinterpret_cast<const float*>(&tl); const float* ptr = reinterpret_cast<const float*>(&tr);
            const float* pbl = reinterpret_cast<const float*>(&bl); const float* pbr = reinterpret_cast<const float*>(&br);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const float top = ptl[v] + (ptr[v] - ptl[v]) * axx.lerp;
              const float bot = pbl[v] + (pbr[v] - pbl[v]) * axx.lerp;
              val[v] = top + (bot - top) * ay.lerp;
            }
          }
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if ((sy == 0 && sx == 0) || val[v] > best[v]) { best[v] = val[v]; arg[v] = sy * S + sx; }
        }
      const float gov[4] = {go.x, go.y, go.z, go.w};
#pragma unroll
      for (int s = 0; s < S * S; ++s)
#pragma unroll
        for (int v = 0; v < 4; ++v) gs[s][v] = (arg[v] == s) ? gov[v] : 0.0f;
    }
#pragma unroll
    for (int sy = 0; sy < S; ++sy)
#pragma unroll
      for (int sx = 0; sx < S; ++sx) {
        const Axis ay = ax_y[sy], axx = ax_x[px * S + sx];
        if (!(ay.valid && axx.valid)) continue;   // extrapolated sample: constant, no gradient
        const float* q = gs[sy * S + sx];
        if (q[0] == 0.0f && q[1] == 0.0f && q[2] == 0.0f && q[3] == 0.0f) continue;
        const float wy1 = ay.lerp, wy0 = 1.0f - ay.lerp, wx1 = axx.lerp, wx0 = 1.0f - axx.lerp;
        float4 t;
        const float dt0 = wy0 * q[0], dt1 = wy0 * q[1], dt2 = wy0 * q[2], dt3 = wy0 * q[3];   // dtop = (1 - ly) * g
        const float db0 = wy1 * q[0], db1 = wy1 * q[1], db2 = wy1 * q[2], db3 = wy1 * q[3];   // dbottom = ly * g
        t = make_float4(wx0 * dt0, wx0 * dt1, wx0 * dt2, wx0 * dt3);
        atomicAdd(gfeat + (static_cast<size_t>(ay.lo) * lf.fw + axx.lo) * cv + cg, t);
        t = make_float4(wx1 * dt0, wx1 * dt1, wx1 * dt2, wx1 * dt3);
        atomicAdd(gfeat + (static_cast<size_t>(ay.lo) * lf.fw + axx.hi) * cv + cg, t);
        t = make_float4(wx0 * db0, wx0 * db1, wx0 * db2, wx0 * db3);
        atomicAdd(gfeat + (static_cast<size_t>(ay.hi) * lf.fw + axx.lo) * cv + cg, t);
        t = make_float4(wx1 * db0, wx1 * db1, wx1 * db2, wx1 * db3);
        atomicAdd(gfeat + (static_cast<size_t>(ay.hi) * lf.fw + axx.hi) * cv + cg, t);
      }
  }
}

template <typename VecT>
int launch_roi_v(bx_handle* h, const RoiArgs& a, int pool, cudaStream_t st) {
  static const int whole_env = getenv("BX_ROI_WHOLE") ? atoi(getenv("BX_ROI_WHOLE")) : -1;   // A/B switch
  // whole-roi CTAs when there are enough rois to fill the device and 2x2 pooling makes adjacent rows share taps
  const bool whole = whole_env >= 0 ? whole_env != 0 : (pool != BX_POOL_NONE && a.r >= 4 * h->num_sms);
  if (whole) {
    if (pool == BX_POOL_NONE) roi_pool_kernel<BX_POOL_NONE, VecT, true><<<a.r, 256, 0, st>>>(a);
    else if (pool == BX_POOL_MAX2) roi_pool_kernel<BX_POOL_MAX2, VecT, true><<<a.r, 256, 0, st>>>(a);
    else roi_pool_kernel<BX_POOL_AVG2, VecT, true><<<a.r, 256, 0, st>>>(a);
    BX_LAUNCH_CHECK(h);
    return BX_OK;
  }
  const int grid = a.r * a.P;
  if (pool == BX_POOL_NONE) roi_pool_kernel<BX_POOL_NONE, VecT, false><<<grid, 256, 0, st>>>(a);
  else if (pool == BX_POOL_MAX2) roi_pool_kernel<BX_POOL_MAX2, VecT, false><<<grid, 256, 0, st>>>(a);
  else roi_pool_kernel<BX_POOL_AVG2, VecT, false><<<grid, 256, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

// pooled fast kernel: float4-aligned maps whose channel-group count divides the CTA into whole-warp pixel lanes
bool roi_pool2_ok(const RoiArgs& a, int pool) {
  const bool off = getenv("BX_ROI_NO_POOL2") != nullptr;                      // A/B / test switch (read per call)
  if (off || pool == BX_POOL_NONE || (a.c & 3)) return false;
  const int cv = a.c >> 2;
  if (cv < 32 || cv > 256 || (256 % cv) != 0 || !bx_aligned(a.out, 16) || a.P * a.P > kPool2MaxPix) return false;
  for (int l = 0; l < a.n_levels; ++l) {
    if (!bx_aligned(a.lv[l].feat, 16)) return false;
    if (static_cast<long long>(a.lv[l].fh) * a.lv[l].fw * cv >= (1ll << 31)) return false;   // 32-bit tap offsets
  }
  return true;
}

int launch_roi(bx_handle* h, const RoiArgs& a, int pool, cudaStream_t st) {
  if (a.r == 0) return BX_OK;
  const bool prof = h->prof_on && h->prof_n < h->prof_cap;
  if (prof) BX_CUDA(cudaEventRecord(h->prof_ev[2 * h->prof_n], st));
  int used = 0, rc = 0;
  // pooled crops (2x2 max / avg): the roi-stationary kernel with tap reuse wins on every measured shape (VGG16
  // 14x14+max C=512: 36 vs 76 us at B=1, 168 vs 225 us at B=8; profiles/micro/vgg_ab.py), plain crops go to the band kernel
  const bool band_pooled = getenv("BX_ROI_BAND_POOLED") != nullptr;             // A/B / test switch (read per call)
  if (pool == BX_POOL_NONE || band_pooled || !roi_pool2_ok(a, pool)) {
    rc = roi_band_launch(h, a, pool, st, &used);     // TMA band-stationary kernel when the shape allows it
    if (rc) return rc;
    // a plain crop the band kernel could not take (C % 32, map too wide, no TMA entry point ...) runs on a gather kernel
    // at roughly 0.75x the speed: not an error, but visible — bx_stats() reports the count next to bx_launch_count()
    if (!used && pool == BX_POOL_NONE && !band_pooled) h->band_fallbacks++;
  }
  if (!used && pool != BX_POOL_NONE && !getenv("BX_ROI_NO_POOL2")) {
    rc = roi_stage_launch(h, a, pool, st, &used);    // opt-in: roi footprint staged in shared memory by TMA (bx_roi_stage.cu)
    if (rc) return rc;
  }
  if (!used && roi_pool2_ok(a, pool)) {
    // L2 prefetch distance in CTAs.  Measured (cfg3 B=16 / cfg5 B=32, us): off 597 / 700, own roi (0) 578 / 653, one wave
    // ahead (444) 625 / 830, two waves 650 / 843: only the CTA's own footprint pays.  BX_POOL2_PF overrides (-1 = off)
    static const int pf_env = getenv("BX_POOL2_PF") ? atoi(getenv("BX_POOL2_PF")) : -2;
    const int pf = pf_env != -2 ? pf_env : 0;
    // (a two-channel-per-thread variant — half the tap registers, 4 CTAs / SM — measured SLOWER, 733 vs 563 us at cfg3
    //  B = 16: the doubled load instructions cost more than the occupancy gains; profiles/README.md)
    if (pool == BX_POOL_MAX2) roi_pool2_kernel<BX_POOL_MAX2><<<a.r, 256, 0, st>>>(a, -0.0f, pf);
    else roi_pool2_kernel<BX_POOL_AVG2><<<a.r, 256, 0, st>>>(a, -0.0f, pf);
    BX_LAUNCH_CHECK(h);
    used = 1;
  }
  if (!used) {
    bool vec = (a.c % 4 == 0) && bx_aligned(a.out, 16);
    for (int l = 0; l < a.n_levels; ++l) vec = vec && bx_aligned(a.lv[l].feat, 16);
    rc = vec ? launch_roi_v<float4>(h, a, pool, st) : launch_roi_v<float>(h, a, pool, st);
    if (rc) return rc;
  }
  if (prof) {
    BX_CUDA(cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], st));
    h->prof_n++;
  }
  return BX_OK;
}

// ---- FPN level assignment (base_fpn_model.py:303-324): single CTA, stable level-major order
__device__ __forceinline__ int roi_level(const float4 roi, int min_level, int max_level) {
  const float hh = fmaxf(0.0f, roi.w - roi.y);
  const float ww = fmaxf(0.0f, roi.z - roi.x);
  float lv = floorf(4.0f + logf(sqrtf(ww * hh + 1e-8f) / 224.0f) / logf(2.0f));
  lv = fmaxf(lv, static_cast<float>(min_level));
  lv = fminf(lv, static_cast<float>(max_level));
  return static_cast<int>(lv) - min_level;
}

// One CTA; thread t owns the contiguous chunk [t * chunk, (t + 1) * chunk) of the rois, so "ascending index within a
// level" is "thread order, then position inside the thread's chunk": per level one block-wide exclusive scan of the
// per-thread counts gives every thread the first output slot of its chunk.  Two passes over the rois (count, place), the
// second one hitting L1; 4 scans + 3 barriers in total (the previous form walked the rois 1024 at a time with a ballot
// round and two barriers per level and step: 45 us for 16 000 rois, now 9 us).
__global__ void __launch_bounds__(1024) assign_levels_kernel(const float4* __restrict__ rois, int r, int min_level,
                                                             int max_level, int* __restrict__ out_level,
                                                             int* __restrict__ out_order, int* __restrict__ out_counts) {
  __shared__ int warp_tot[kMaxLevels][32];
  __shared__ int level_base[kMaxLevels + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nl = max_level - min_level + 1;
  const int chunk = (r + 1023) / 1024;
  const int lo = min(r, tid * chunk), hi = min(r, lo + chunk);
  int cnt[kMaxLevels];
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) cnt[l] = 0;
  for (int i = lo; i < hi; ++i) {
    const int lv = roi_level(rois[i], min_level, max_level);
    if (out_level) out_level[i] = lv + min_level;
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) cnt[l] += (lv == l) ? 1 : 0;
  }
  // exclusive scan of cnt[l] over the threads of the block, per level
  int pre[kMaxLevels];
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) {
    pre[l] = 0;
    if (l < nl) {
      int v = cnt[l];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += o;
      }
      if (lane == 31) warp_tot[l][warp] = v;
      pre[l] = v - cnt[l];                     // exclusive prefix inside the warp
    }
  }
  __syncthreads();
  if (warp == 0) {
    for (int l = 0; l < nl; ++l) {
      const int t = warp_tot[l][lane];
      int v = t;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += o;
      }
      warp_tot[l][lane] = v - t;               // exclusive prefix of the warp totals
      if (lane == 31) level_base[l + 1] = v;   // level total (prefix-summed over the levels below)
    }
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int l = 0; l < nl; ++l) {
      const int tot = level_base[l + 1];
      if (out_counts) out_counts[l] = tot;
      level_base[l] = run;
      run += tot;
    }
  }
  __syncthreads();
  int pos[kMaxLevels];
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) pos[l] = (l < nl) ? level_base[l] + warp_tot[l][warp] + pre[l] : 0;
  for (int i = lo; i < hi; ++i) {
    const int lv = roi_level(rois[i], min_level, max_level);
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) {
      if (lv == l) out_order[pos[l]++] = i;
    }
  }
}

int check_roi_common(const char* fn, bx_handle* h, int pool_size, int c, int r, const void* rois, const void* out) {
  BX_REQUIRE(h && out && (rois || r == 0), BX_ERR_INVALID, "%s: NULL argument", fn);
  BX_REQUIRE(r >= 0 && c > 0, BX_ERR_INVALID, "%s: bad size", fn);
  BX_REQUIRE(pool_size > 0 && 2 * pool_size <= kMaxQ, BX_ERR_UNSUPPORTED, "%s: pool_size must be in [1, %d]", fn, kMaxQ / 2);
  BX_REQUIRE(bx_aligned(rois, 16), BX_ERR_INVALID, "%s: rois must be 16-byte aligned", fn);
  return BX_OK;
}

}  // namespace

extern "C" int bx_crop_and_resize(bx_handle* h, const float* image, int b, int ih, int iw, int c, const float* boxes,
                                  const int* box_ind, int r, int crop_h, int crop_w, float extrapolation_value,
                                  float* out, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && image && out && (boxes || r == 0), BX_ERR_INVALID, "bx_crop_and_resize: NULL argument");
  BX_REQUIRE(crop_h > 0 && crop_w > 0, BX_ERR_INVALID, "bx_crop_and_resize: crop_size must be 2 positive ints");
  BX_REQUIRE(crop_h == crop_w && crop_h <= kMaxQ, BX_ERR_UNSUPPORTED,
             "bx_crop_and_resize: only square crops up to %d are built (the reference uses 7 and 14)", kMaxQ);
  BX_REQUIRE(b > 0 && ih > 0 && iw > 0 && c > 0 && r >= 0, BX_ERR_INVALID, "bx_crop_and_resize: bad size");
  BX_REQUIRE(bx_aligned(boxes, 16), BX_ERR_INVALID, "bx_crop_and_resize: boxes must be 16-byte aligned");
  RoiArgs a = {};
  a.lv[0] = {image, ih, iw};
  a.n_levels = 1;
  a.rois = reinterpret_cast<const float4*>(boxes);
  a.box_ind = box_ind;
  a.r = r; a.b = b; a.c = c;
  a.mode = 3;
  a.P = crop_h; a.Q = crop_h;
  a.extrapolation = extrapolation_value;
  a.out = out;
  return launch_roi(h, a, BX_POOL_NONE, static_cast<cudaStream_t>(stream));
}

extern "C" int bx_roi_pool(bx_handle* h, int mode, int pool, int pool_size, const float* feat, int b, int fh, int fw,
                           int c, const float* rois, const int* box_ind, const int* roi_counts, int r, float stride,
                           int image_h, int image_w, float* out, void* stream) {
  BxEnter guard(h, stream);
  int rc = check_roi_common("bx_roi_pool", h, pool_size, c, r, rois, out);
  if (rc) return rc;
  BX_REQUIRE(feat && b > 0 && fh > 1 && fw > 1, BX_ERR_INVALID, "bx_roi_pool: bad feature map");
  BX_REQUIRE(mode >= 0 && mode <= 2 && pool >= 0 && pool <= 2, BX_ERR_INVALID, "bx_roi_pool: bad mode/pool enum");
  BX_REQUIRE(mode == BX_ROI_IMAGE_NORM ? (image_h > 0 && image_w > 0) : (stride > 0.0f), BX_ERR_INVALID,
             "bx_roi_pool: stride (or image shape) must be positive");
  BX_REQUIRE(!roi_counts || (r % b == 0), BX_ERR_INVALID, "bx_roi_pool: with roi_counts, r must be batch * rois_per_image");
  RoiArgs a = {};
  a.lv[0] = {feat, fh, fw};
  a.n_levels = 1;
  a.rois = reinterpret_cast<const float4*>(rois);
  a.box_ind = box_ind;
  a.roi_counts = roi_counts;
  a.rois_per_image = roi_counts ? r / b : 0;
  a.r = r; a.b = b; a.c = c;
  a.mode = mode;
  a.P = pool_size;
  a.Q = (pool == BX_POOL_NONE) ? pool_size : 2 * pool_size;
  a.stride = stride;
  a.image_h = static_cast<float>(image_h);
  a.image_w = static_cast<float>(image_w);
  a.extrapolation = 0.0f;
  a.out = out;
  return launch_roi(h, a, pool, static_cast<cudaStream_t>(stream));
}

extern "C" int bx_roi_pool_grad(bx_handle* h, int mode, int pool, int pool_size, const float* feat, int b, int fh,
                                int fw, int c, const float* rois, const int* box_ind, const int* roi_counts, int r,
                                float stride, int image_h, int image_w, const float* grad_out, float* grad_feat,
                                void* stream) {
  BxEnter guard(h, stream);
  int rc = check_roi_common("bx_roi_pool_grad", h, pool_size, c, r, rois, grad_feat);
  if (rc) return rc;
  BX_REQUIRE(grad_out || r == 0, BX_ERR_INVALID, "bx_roi_pool_grad: NULL grad_out");
  BX_REQUIRE(b > 0 && fh > 1 && fw > 1, BX_ERR_INVALID, "bx_roi_pool_grad: bad feature map");
  BX_REQUIRE(mode >= 0 && mode <= 2 && pool >= 0 && pool <= 2, BX_ERR_INVALID, "bx_roi_pool_grad: bad mode/pool enum");
  BX_REQUIRE(pool != BX_POOL_MAX2 || feat, BX_ERR_INVALID, "bx_roi_pool_grad: the max-pool gradient needs the features");
  BX_REQUIRE(mode == BX_ROI_IMAGE_NORM ? (image_h > 0 && image_w > 0) : (stride > 0.0f), BX_ERR_INVALID,
             "bx_roi_pool_grad: stride (or image shape) must be positive");
  BX_REQUIRE(!roi_counts || (r % b == 0), BX_ERR_INVALID, "bx_roi_pool_grad: with roi_counts, r must be batch * rois_per_image");
  BX_REQUIRE(c % 4 == 0 && bx_aligned(grad_feat, 16) && bx_aligned(grad_out, 16) && bx_aligned(feat, 16),
             BX_ERR_UNSUPPORTED, "bx_roi_pool_grad: channels must be a multiple of 4 and tensors 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t feat_bytes = sizeof(float) * static_cast<size_t>(b) * fh * fw * c;
  if (r == 0) {
    BX_CUDA(cudaMemsetAsync(grad_feat, 0, feat_bytes, st));
    return BX_OK;
  }
  RoiGradArgs g = {};
  g.a.lv[0] = {feat, fh, fw};
  g.a.n_levels = 1;
  g.a.rois = reinterpret_cast<const float4*>(rois);
  g.a.box_ind = box_ind;
  g.a.roi_counts = roi_counts;
  g.a.rois_per_image = roi_counts ? r / b : 0;
  g.a.r = r; g.a.b = b; g.a.c = c;
  g.a.mode = mode;
  g.a.P = pool_size;
  g.a.Q = (pool == BX_POOL_NONE) ? pool_size : 2 * pool_size;
  g.a.stride = stride;
  g.a.image_h = static_cast<float>(image_h);
  g.a.image_w = static_cast<float>(image_w);
  g.a.extrapolation = 0.0f;
  g.grad_out = grad_out;
  g.grad_feat = grad_feat;
  int used = 0;
  if (int rc2 = roi_grad_rows_launch(h, g, pool, st, &used)) return rc2;   // row-owned kernel: writes every element itself
  if (used) return BX_OK;
  BX_CUDA(cudaMemsetAsync(grad_feat, 0, feat_bytes, st));
  const int grid = r * pool_size;
  if (pool == BX_POOL_NONE) roi_pool_grad_kernel<BX_POOL_NONE><<<grid, 256, 0, st>>>(g);
  else if (pool == BX_POOL_MAX2) roi_pool_grad_kernel<BX_POOL_MAX2><<<grid, 256, 0, st>>>(g);
  else roi_pool_grad_kernel<BX_POOL_AVG2><<<grid, 256, 0, st>>>(g);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

extern "C" int bx_fpn_assign_levels(bx_handle* h, const float* rois, int r, int min_level, int max_level,
                                    int* out_level, int* out_order, int* out_counts, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && out_order && (rois || r == 0), BX_ERR_INVALID, "bx_fpn_assign_levels: NULL argument");
  BX_REQUIRE(r >= 0 && max_level >= min_level && max_level - min_level < kMaxLevels, BX_ERR_INVALID,
             "bx_fpn_assign_levels: bad level range");
  BX_REQUIRE(bx_aligned(rois, 16), BX_ERR_INVALID, "bx_fpn_assign_levels: rois must be 16-byte aligned");
  assign_levels_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(rois), r,
                                                                          min_level, max_level, out_level, out_order,
                                                                          out_counts);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

extern "C" int bx_fpn_roi_features(bx_handle* h, const float* const* feats, const int* fh, const int* fw,
                                   int n_levels, int min_level, int b, int c, const float* rois, const int* box_ind,
                                   int r, int image_h, int image_w, int pool_size, float* out, int* out_level,
                                   int* out_order, int* out_counts, void* stream) {
  BxEnter guard(h, stream);
  int rc = check_roi_common("bx_fpn_roi_features", h, pool_size, c, r, rois, out);
  if (rc) return rc;
  BX_REQUIRE(feats && fh && fw && out_level && out_order, BX_ERR_INVALID, "bx_fpn_roi_features: NULL argument");
  BX_REQUIRE(n_levels > 0 && n_levels <= kMaxLevels && b > 0, BX_ERR_INVALID, "bx_fpn_roi_features: bad level count");
  BX_REQUIRE(image_h > 0 && image_w > 0, BX_ERR_INVALID, "bx_fpn_roi_features: image shape must be positive");
  rc = bx_fpn_assign_levels(h, rois, r, min_level, min_level + n_levels - 1, out_level, out_order, out_counts, stream);
  if (rc) return rc;
  RoiArgs a = {};
  for (int l = 0; l < n_levels; ++l) {
    BX_REQUIRE(feats[l] && fh[l] > 1 && fw[l] > 1, BX_ERR_INVALID, "bx_fpn_roi_features: bad feature map %d", l);
    a.lv[l] = {feats[l], fh[l], fw[l]};
  }
  a.n_levels = n_levels;
  a.rois = reinterpret_cast<const float4*>(rois);
  a.box_ind = box_ind;
  a.order = out_order;
  a.level = out_level;
  a.r = r; a.b = b; a.c = c;
  a.mode = BX_ROI_IMAGE_NORM;
  a.P = pool_size;
  a.Q = 2 * pool_size;
  a.image_h = static_cast<float>(image_h);
  a.image_w = static_cast<float>(image_w);
  a.out = out;
  a.level_base = min_level;
  return launch_roi(h, a, BX_POOL_MAX2, static_cast<cudaStream_t>(stream));
}
