// Shared declarations of the RoI-pooling kernels (bx_roi.cu: direct-gather kernel; bx_roi_band.cu: TMA band-stationary kernel).
#pragma once
#include "bx_common.cuh"

namespace bxroi {


constexpr int kMaxLevels = 8;
constexpr int kMaxQ = 64;  // max crop size (2*pool_size)

struct LevelFeat {
  const float* feat;  // [b,fh,fw,c]
  int fh, fw;
};

struct RoiArgs {
  LevelFeat lv[kMaxLevels];
  int n_levels;
  const float4* rois;     // [r] image coordinates (or normalised y1,x1,y2,x2 for mode RAW)
  const int* box_ind;     // [r] or null
  const int* roi_counts;  // [b] or null
  const int* order;       // [r] output row j reads roi order[j] (FPN level-major) or null (identity)
  const int* level;       // [r] absolute level per roi or null (level 0)
  int level_base;         // min_level: level[src] - level_base indexes lv[]
  int r, b, c;
  int rois_per_image;     // r / b when roi_counts is given
  int mode;               // bx_roi_mode or 3 = RAW normalised boxes
  int P;                  // output size
  int Q;                  // crop size (P or 2P)
  float stride;
  float image_h, image_w;
  float extrapolation;
  float* out;             // [r,P,P,c]
};

struct RoiGradArgs {
  RoiArgs a;              // forward description; a.out is unused
  const float* grad_out;  // [r,P,P,c]
  float* grad_feat;       // [b,fh,fw,c]
};

struct Axis {
  int lo, hi;   // tap indices (already mapped to the un-padded map)
  float lerp;
  int valid;
};

// coordinate of crop sample `s` along one axis, TF op order.  n1,n2: normalised box ends; dim: (padded) map size.
__device__ __forceinline__ Axis sample_axis(float n1, float n2, int s, int Q, int dim, int pad) {
  Axis a;
  const float dm1 = static_cast<float>(dim - 1);
  float in;
  if (Q > 1) {
    const float scale = (n2 - n1) * dm1 / static_cast<float>(Q - 1);
    in = n1 * dm1 + static_cast<float>(s) * scale;
  } else {
    in = 0.5f * (n1 + n2) * dm1;
  }
  a.valid = !(in < 0.0f || in > dm1);
  const float lo = floorf(in), hi = ceilf(in);
  a.lerp = in - lo;
  int ilo = static_cast<int>(lo), ihi = static_cast<int>(hi);
  if (pad) {  // SYMMETRIC pad by 1 (roi_pooling.py:100): padded index p -> original clamp(p-1, 0, dim-3)
    ilo = min(max(ilo - 1, 0), dim - 3);
    ihi = min(max(ihi - 1, 0), dim - 3);
  }
  a.lo = ilo;
  a.hi = ihi;
  return a;
}

// normalised crop box (y1,x1,y2,x2) of one roi for the given mode + the (padded) map size the crop op sees.
struct NormBox {
  float y1, x1, y2, x2;
  int dimy, dimx, pad;
};

__device__ __forceinline__ NormBox roi_norm_box(const RoiArgs& a, const float4 roi, int fh, int fw) {
  NormBox n;
  n.dimy = fh; n.dimx = fw; n.pad = 0;
  const int Q = a.Q;
  if (a.mode == BX_ROI_STRIDE_NORM) {           // roi_pooling.py:64-74
    const float fy = static_cast<float>(fh - 1), fx = static_cast<float>(fw - 1);
    n.y1 = (roi.y / a.stride) / fy;
    n.x1 = (roi.x / a.stride) / fx;
    n.y2 = (roi.w / a.stride) / fy;
    n.x2 = (roi.z / a.stride) / fx;
  } else if (a.mode == BX_ROI_IMAGE_NORM) {     // roi_pooling.py:26-35
    n.y1 = roi.y / a.image_h;
    n.x1 = roi.x / a.image_w;
    n.y2 = roi.w / a.image_h;
    n.x2 = roi.z / a.image_w;
  } else if (a.mode == BX_ROI_ALIGN_PAD) {      // roi_pooling.py:175,101,103-130
    n.pad = 1;
    n.dimy = fh + 2;
    n.dimx = fw + 2;
    const float x0 = roi.x / a.stride + 1.0f, y0 = roi.y / a.stride + 1.0f;
    const float x1 = roi.z / a.stride + 1.0f, y1 = roi.w / a.stride + 1.0f;
    const float qf = static_cast<float>(Q);
    const float sw = (x1 - x0) / qf, sh = (y1 - y0) / qf;
    const float ih = static_cast<float>(n.dimy - 1), iw = static_cast<float>(n.dimx - 1);
    n.x1 = (x0 + sw / 2.0f - 0.5f) / iw;
    n.y1 = (y0 + sh / 2.0f - 0.5f) / ih;
    const float nw = sw * static_cast<float>(Q - 1) / iw;
    const float nh = sh * static_cast<float>(Q - 1) / ih;
    n.x2 = n.x1 + nw;
    n.y2 = n.y1 + nh;
  } else {                                      // RAW: boxes are (y1,x1,y2,x2) normalised
    n.y1 = roi.x; n.x1 = roi.y; n.y2 = roi.z; n.x2 = roi.w;
  }
  return n;
}

// Packed fp32x2 arithmetic (sm_100 FADD2 / FFMA2): two IEEE fp32 operations per issued instruction, each lane rounded
// exactly like the scalar op.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (single rounding) even with
// -fmad=false, so the multiply is written as fma(a, b, -0.0) with the -0.0 arriving as a kernel argument: a * b + (-0.0)
// is the correctly rounded product (sign of zero included) and cannot be fused with the following add.
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b, unsigned long long nz) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
  return r;
}
__device__ __forceinline__ unsigned long long f2_splat(float v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
  return r;
}
// running 2x2 pool over packed values: max (Keras MaxPooling2D) or sum (tf.nn.avg_pool, divided by 4 afterwards)
template <int POOL>
__device__ __forceinline__ ulonglong2 pool2(const ulonglong2 a, const ulonglong2 b) {
  ulonglong2 r;
  if (POOL == BX_POOL_MAX2) {
    const float4 x = *reinterpret_cast<const float4*>(&a), y = *reinterpret_cast<const float4*>(&b);
    const float4 m = make_float4(fmaxf(x.x, y.x), fmaxf(x.y, y.y), fmaxf(x.z, y.z), fmaxf(x.w, y.w));
    r = *reinterpret_cast<const ulonglong2*>(&m);
  } else {
    r.x = f2_add(a.x, b.x);
    r.y = f2_add(a.y, b.y);
  }
  return r;
}

// implemented in bx_roi_band.cu: returns BX_OK and sets *used = 1 when the band kernel handled the launch
int roi_band_launch(bx_handle* h, const RoiArgs& a, int pool, cudaStream_t st, int* used);
// implemented in bx_roi_stage.cu: pooled crops with the roi footprint staged in shared memory by TMA
int roi_stage_launch(bx_handle* h, const RoiArgs& a, int pool, cudaStream_t st, int* used);

// implemented in bx_roi_grad.cu: row-owned, atomic-free backward; *used = 0 -> the caller runs the scatter kernel
int roi_grad_rows_launch(bx_handle* h, const RoiGradArgs& g, int pool, cudaStream_t st, int* used);

}  // namespace bxroi
