// Band-stationary RoI pooling (sm_100a): the feature map is the stationary operand.
//
// One CTA owns (image, 32-channel slice, band of feature rows).  It stages the band [rows r0..r1] x all columns x 32
// channels (128 B per pixel) in shared memory with TMA tile loads (cp.async.bulk.tensor.2d over the [B*fh*fw, C] view of
// the NHWC map), then produces every output row (roi, py) whose first sample row lands in the band: all bilinear taps
// come from shared memory, so L2->SM traffic is the band once (instead of 4 taps per sample), and every output float4
// is written once, 128 contiguous bytes per pixel-slice.
//
// Same arithmetic as bx_roi.cu (TF r1.13 crop_and_resize op order, no FMA); replaces the same reference code:
// model/roi_pooling.py:8-42, :45-90, :93-176.
#include <cuda.h>
#include <stdlib.h>

#include "bx_roi.cuh"

namespace bxroi {
namespace {

constexpr int kThreads = 1024;
constexpr int kSlice = 32;          // channels per CTA (128 B per pixel)
constexpr int kBoxRows = 256;       // pixels per TMA box
constexpr int kBoxBytes = kBoxRows * kSlice * 4;

struct BandArgs {
  RoiArgs r;
  int rows_per_band, n_bands, n_slices;
  int chunk;      // rois per plan block
  int n_chunks;   // plan blocks per (image, band)
  int nbox;       // TMA boxes per band
  int scan_all;   // 1: rois of any image may be anywhere (box_ind given) -> every block scans every roi
  unsigned char* plan;   // [b, n_bands, n_chunks] plan blocks (global workspace)
  float neg_zero;        // -0.0f at run time: addend that turns fma.rn.f32x2 into an exact, non-contractible multiply
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Plan: everything about the crop geometry that does not depend on the channel slice is computed once per
// (image, band, roi chunk) by roi_plan_kernel and stored in global memory as one contiguous "plan block"; each band CTA
// pulls its block into shared memory with one bulk copy (cp.async.bulk) next to the TMA tile loads of the band.
//
// Plan block layout (byte offsets, chunk = rois per block, Q = crop size):
//   [0]                      int   n_runs  (+ padding to 16 B)
//   [16]                     u64   xval [chunk]     validity bit per sample column
//   [.. ]                    uint2 xtab [chunk*Q]   (lo_off | hi_off << 16: byte offsets of the tap columns in a band row, lerp_x)
//   [.. ]                    uint2 ytab [chunk*Q]   (a | in_band << 15 | b << 16 | valid << 31, lerp_y): in_band: a, b = offsets
//                                                   (16 B units) of the top / bottom tap rows in the staged band (the zero row
//                                                   when the sample is invalid); else absolute feature rows (pooled pairs only)
//   [.. ]                    uint2 runs [4*chunk]   (roi_local | zero << 12 | all_x_valid << 13 | py_begin << 16 | py_end << 24,
//                                                    offset of output pixel (roi, py_begin, 0) in float4 units from the first roi
//                                                    of the block)
struct PlanLayout {
  uint32_t off_xval, off_xtab, off_ytab, off_runs, bytes;
};

__host__ __device__ inline PlanLayout plan_layout(int chunk, int Q) {
  PlanLayout L;
  L.off_xval = 16;
  L.off_xtab = L.off_xval + 8u * chunk;
  L.off_ytab = L.off_xtab + 8u * chunk * Q;
  L.off_runs = L.off_ytab + 8u * chunk * Q;
  L.bytes = (L.off_runs + 32u * chunk + 127u) & ~127u;
  return L;
}

// one axis sample from (base, scale): same arithmetic as sample_axis() in bx_roi.cuh
__device__ __forceinline__ void axis_from_par(float base, float scale, int s, int dim, int pad, int& lo, int& hi,
                                              float& lerp, bool& valid) {
  const float dm1 = static_cast<float>(dim - 1);
  const float in = base + static_cast<float>(s) * scale;
  valid = !(in < 0.0f || in > dm1);
  const float flo = floorf(in), fhi = ceilf(in);
  lerp = in - flo;
  lo = static_cast<int>(flo);
  hi = static_cast<int>(fhi);
  if (pad) {
    lo = min(max(lo - 1, 0), dim - 3);
    hi = min(max(hi - 1, 0), dim - 3);
  }
}

constexpr int kPlanThreads = 512;
constexpr int kPlanMaxChunk = 1024;

template <int S>
__global__ void __launch_bounds__(kPlanThreads) roi_plan_kernel(const BandArgs a) {
  __shared__ float4 rpar[kPlanMaxChunk];
  __shared__ uint32_t pym[kPlanMaxChunk];
  __shared__ uint32_t xv_lo[kPlanMaxChunk], xv_hi[kPlanMaxChunk];
  __shared__ unsigned char rflag[kPlanMaxChunk];
  __shared__ int n_runs;
  const RoiArgs& r = a.r;
  const int tid = threadIdx.x;
  int u = blockIdx.x;
  const int ch = u % a.n_chunks; u /= a.n_chunks;
  const int band_i = u % a.n_bands;
  const int img = u / a.n_bands;
  const int P = r.P, Q = r.Q;
  const int fh = r.lv[0].fh, fw = r.lv[0].fw;
  const int r0 = band_i * a.rows_per_band;
  const int r1 = min(fh, r0 + a.rows_per_band);
  const int rows_loaded = min(r1 + 1, fh) - r0;
  const int pad = (r.mode == BX_ROI_ALIGN_PAD) ? 1 : 0;
  const int dimy = fh + 2 * pad, dimx = fw + 2 * pad;
  const uint32_t row_bytes = static_cast<uint32_t>(fw) * (kSlice * 4);
  const uint32_t zrow_off = static_cast<uint32_t>(a.nbox) * kBoxBytes;
  const PlanLayout L = plan_layout(a.chunk, Q);
  unsigned char* blk = a.plan + static_cast<size_t>(blockIdx.x) * L.bytes;
  unsigned long long* g_xval = reinterpret_cast<unsigned long long*>(blk + L.off_xval);
  uint2* g_xtab = reinterpret_cast<uint2*>(blk + L.off_xtab);
  uint2* g_ytab = reinterpret_cast<uint2*>(blk + L.off_ytab);
  uint2* g_runs = reinterpret_cast<uint2*>(blk + L.off_runs);

  int g_begin = 0, g_end = r.r;
  if (!a.scan_all && r.roi_counts) {
    g_begin = img * r.rois_per_image;
    g_end = g_begin + r.rois_per_image;
  }
  const int c0 = g_begin + ch * a.chunk;
  const int nroi = max(0, min(a.chunk, g_end - c0));

  // ---- per-roi crop parameters
  for (int l = tid; l < nroi; l += kPlanThreads) {
    const int g = c0 + l;
    int rimg = 0;
    uint32_t flag = 0;
    if (r.roi_counts) {
      rimg = g / r.rois_per_image;
      if ((g % r.rois_per_image) >= r.roi_counts[rimg]) flag |= 2u;
    } else if (r.box_ind) {
      rimg = r.box_ind[g];
    }
    if (rimg == img) flag |= 1u;
    else if (r.box_ind && (rimg < 0 || rimg >= r.b) && img == 0) flag |= 3u;   // bad box_ind: zero-filled by image 0
    const NormBox nb = roi_norm_box(r, r.rois[g], fh, fw);
    const float dmx = static_cast<float>(dimx - 1), dmy = static_cast<float>(dimy - 1);
    float4 par;
    if (Q > 1) {
      par.x = nb.x1 * dmx;
      par.y = (nb.x2 - nb.x1) * dmx / static_cast<float>(Q - 1);
      par.z = nb.y1 * dmy;
      par.w = (nb.y2 - nb.y1) * dmy / static_cast<float>(Q - 1);
    } else {
      par.x = 0.5f * (nb.x1 + nb.x2) * dmx;
      par.y = 0.0f;
      par.z = 0.5f * (nb.y1 + nb.y2) * dmy;
      par.w = 0.0f;
    }
    rpar[l] = par;
    pym[l] = 0u;
    xv_lo[l] = 0u;
    xv_hi[l] = 0u;
    rflag[l] = static_cast<unsigned char>(flag);
  }
  if (tid == 0) n_runs = 0;
  __syncthreads();
  // ---- x / y sample tables, ownership of output rows
  for (int idx = tid; idx < nroi * Q; idx += kPlanThreads) {
    const int l = idx / Q, s = idx - l * Q;
    const uint32_t flag = rflag[l];
    if (!(flag & 1u)) continue;
    const bool zero = flag & 2u;
    const float4 par = rpar[l];
    int lo, hi;
    float lerp;
    bool valid;
    axis_from_par(par.x, par.y, s, dimx, pad, lo, hi, lerp, valid);
    valid = valid && !zero;
    g_xtab[idx] = make_uint2(valid ? (static_cast<uint32_t>(lo * (kSlice * 4)) | (static_cast<uint32_t>(hi * (kSlice * 4)) << 16)) : 0u,
                             __float_as_uint(lerp));
    if (valid) atomicOr(s < 32 ? &xv_lo[l] : &xv_hi[l], 1u << (s & 31));
    axis_from_par(par.z, par.w, s, dimy, pad, lo, hi, lerp, valid);
    valid = valid && !zero;
    const bool inb = valid && (lo >= r0 && hi < r0 + rows_loaded);
    uint32_t ya, yb;
    if (inb) {
      ya = (static_cast<uint32_t>(lo - r0) * row_bytes) >> 4;
      yb = (static_cast<uint32_t>(hi - r0) * row_bytes) >> 4;
    } else if (!valid) {
      ya = yb = zrow_off >> 4;
    } else {
      ya = static_cast<uint32_t>(lo);
      yb = static_cast<uint32_t>(hi);
    }
    g_ytab[idx] = make_uint2(ya | ((inb || !valid) ? (1u << 15) : 0u) | (yb << 16) | (valid ? (1u << 31) : 0u),
                             __float_as_uint(valid ? lerp : 0.0f));
    // the band that owns output row py = s / S is the one holding the top tap row of its first valid sample row
    if (S == 1 || (s & 1) == 0) {
      int owner_row = valid ? lo : -1;
      if (S == 2 && !valid) {
        int lo2, hi2; float l2; bool v2;
        axis_from_par(par.z, par.w, s + 1, dimy, pad, lo2, hi2, l2, v2);
        if (v2 && !zero) owner_row = lo2;
      }
      const int owner = owner_row < 0 ? 0 : min(owner_row / a.rows_per_band, a.n_bands - 1);
      if (owner == band_i) atomicOr(&pym[l], 1u << (s / S));
    }
  }
  __syncthreads();
  // ---- one work item per run of consecutive owned rows of a roi
  for (int l = tid; l < nroi; l += kPlanThreads) {
    const unsigned long long xv = static_cast<unsigned long long>(xv_lo[l]) | (static_cast<unsigned long long>(xv_hi[l]) << 32);
    g_xval[l] = xv;
    const unsigned long long full = (Q >= 64) ? ~0ull : ((1ull << Q) - 1ull);
    uint32_t m = pym[l];
    const uint32_t z = ((static_cast<uint32_t>(rflag[l]) & 2u) << 11) | ((xv == full) ? (1u << 13) : 0u);
    while (m) {
      const int b0 = __ffs(m) - 1;
      const int len = __ffs(~(m >> b0)) - 1;               // first zero above b0 ends the run (len <= P < 32)
      g_runs[atomicAdd(&n_runs, 1)] = make_uint2(static_cast<uint32_t>(l) | z | (static_cast<uint32_t>(b0) << 16) |
                                                     (static_cast<uint32_t>(b0 + len) << 24),
                                                 static_cast<uint32_t>((l * P + b0) * P) * static_cast<uint32_t>(r.c / 4));
      m &= ~(((1u << len) - 1u) << b0);
    }
  }
  __syncthreads();
  if (tid == 0) *reinterpret_cast<int*>(blk) = n_runs;
}

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float4 lerp2(const float4 tl, const float4 tr, const float4 bl, const float4 br, float wx,
                                        float wy) {
  const float t0 = tl.x + (tr.x - tl.x) * wx, b0 = bl.x + (br.x - bl.x) * wx;
  const float t1 = tl.y + (tr.y - tl.y) * wx, b1 = bl.y + (br.y - bl.y) * wx;
  const float t2 = tl.z + (tr.z - tl.z) * wx, b2 = bl.z + (br.z - bl.z) * wx;
  const float t3 = tl.w + (tr.w - tl.w) * wx, b3 = bl.w + (br.w - bl.w) * wx;
  return make_float4(t0 + (b0 - t0) * wy, t1 + (b1 - t1) * wy, t2 + (b2 - t2) * wy, t3 + (b3 - t3) * wy);
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
// t = tl + (tr - tl) * wx ; b = bl + (br - bl) * wx ; out = t + (b - t) * wy     (op order of TF's crop_and_resize)
__device__ __forceinline__ ulonglong2 lerp2_packed(const ulonglong2 tl, const ulonglong2 tr, const ulonglong2 bl,
                                                   const ulonglong2 br, unsigned long long wx, unsigned long long wy,
                                                   unsigned long long nz) {
  const unsigned long long t01 = f2_add(tl.x, f2_mul(f2_sub(tr.x, tl.x), wx, nz));
  const unsigned long long t23 = f2_add(tl.y, f2_mul(f2_sub(tr.y, tl.y), wx, nz));
  const unsigned long long b01 = f2_add(bl.x, f2_mul(f2_sub(br.x, bl.x), wx, nz));
  const unsigned long long b23 = f2_add(bl.y, f2_mul(f2_sub(br.y, bl.y), wx, nz));
  ulonglong2 o;
  o.x = f2_add(t01, f2_mul(f2_sub(b01, t01), wy, nz));
  o.y = f2_add(t23, f2_mul(f2_sub(b23, t23), wy, nz));
  return o;
}

template <int POOL>
__device__ __forceinline__ void pool_acc(float4& acc, const float4 v, bool first) {
  if (first) acc = v;
  else if (POOL == BX_POOL_MAX2) acc = make_float4(fmaxf(acc.x, v.x), fmaxf(acc.y, v.y), fmaxf(acc.z, v.z), fmaxf(acc.w, v.w));
  else acc = make_float4(acc.x + v.x, acc.y + v.y, acc.z + v.z, acc.w + v.w);
}

template <int POOL>
__global__ void __launch_bounds__(kThreads, 1) roi_band_kernel(const __grid_constant__ CUtensorMap tmap, const BandArgs a) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;
  extern __shared__ __align__(128) unsigned char smem[];
  const RoiArgs& r = a.r;
  const int fh = r.lv[0].fh, fw = r.lv[0].fw;
  const uint32_t row_bytes = static_cast<uint32_t>(fw) * (kSlice * 4);
  const uint32_t zrow_off = static_cast<uint32_t>(a.nbox) * kBoxBytes;      // one all-zero band row after the TMA boxes
  const PlanLayout L = plan_layout(a.chunk, r.Q);
  unsigned char* tab = smem + zrow_off + ((row_bytes + 127u) & ~127u);
  const unsigned long long* xval = reinterpret_cast<const unsigned long long*>(tab + L.off_xval);
  const uint2* xtab = reinterpret_cast<const uint2*>(tab + L.off_xtab);
  const uint2* ytab = reinterpret_cast<const uint2*>(tab + L.off_ytab);
  const uint2* runs = reinterpret_cast<const uint2*>(tab + L.off_runs);
  __shared__ uint64_t mbar;
  __shared__ int next_item;

  const int tid = threadIdx.x, lane = tid & 31;
  const int P = r.P, Q = r.Q, C = r.c;
  int u = blockIdx.x;
  const int band_i = u % a.n_bands; u /= a.n_bands;
  const int slice = u % a.n_slices;
  const int img = u / a.n_slices;
  const int r0 = band_i * a.rows_per_band;
  const float* feat_img = r.lv[0].feat + static_cast<size_t>(img) * fh * fw * C;
  const unsigned char* plan0 = a.plan + static_cast<size_t>((img * a.n_bands + band_i) * a.n_chunks) * L.bytes;

  if (tid == 0) {
    mbar_init(&mbar, 1);
    mbar_expect_tx(&mbar, static_cast<uint32_t>(a.nbox) * kBoxBytes + L.bytes);
    bulk_load(tab, plan0, L.bytes, &mbar);
    const int row0 = (img * fh + r0) * fw;
#if defined(BX_EXP) && BX_EXP == 5
    for (int bx = 0; bx < a.nbox; ++bx)   // experiment: band boxes all read the same rows (L2-hot, no DRAM)
      tma_load_2d(smem + static_cast<size_t>(bx) * kBoxBytes, &tmap, 0, 0, &mbar);
#else
    for (int bx = 0; bx < a.nbox; ++bx)
      tma_load_2d(smem + static_cast<size_t>(bx) * kBoxBytes, &tmap, slice * kSlice, row0 + bx * kBoxRows, &mbar);
#endif
    next_item = 0;
  }
  for (uint32_t i = tid * 16u; i < row_bytes; i += kThreads * 16u)
    *reinterpret_cast<float4*>(smem + zrow_off + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  int g_begin = 0;   // first roi of plan block 0
  if (!a.scan_all && r.roi_counts) g_begin = img * r.rois_per_image;
  const int q = lane & 7, sub = lane >> 3;
  const unsigned char* band_q = smem + q * 16;
  const bool ext_zero = (r.extrapolation == 0.0f);
  const unsigned long long nz2 = f2_splat(a.neg_zero);

  for (int ch = 0; ch < a.n_chunks; ++ch) {
    if (ch > 0) {
      __syncthreads();                       // everyone is done with the previous plan block
      if (tid == 0) {
        next_item = 0;
        mbar_expect_tx(&mbar, L.bytes);
        bulk_load(tab, plan0 + static_cast<size_t>(ch) * L.bytes, L.bytes, &mbar);
      }
      __syncthreads();
    }
    mbar_wait(&mbar, static_cast<uint32_t>(ch & 1));
#if defined(BX_EXP) && BX_EXP == 4
    if (a.neg_zero == 0.0f) return;   // experiment: load + wait only (-0.0f == 0.0f is true)
#endif
    const int c0 = g_begin + ch * a.chunk;
    const int n_items = *reinterpret_cast<const int*>(tab);
    // output pixel (roi c0, py 0, px 0), this CTA's channel slice, this lane's 4 channels
    float4* out_c0 = reinterpret_cast<float4*>(r.out + static_cast<size_t>(c0) * P * P * C + slice * kSlice + q * 4);
    const uint32_t c4 = static_cast<uint32_t>(C) >> 2;
    // ---- one warp per run of output rows of a roi; 8 lanes (32 channels) per pixel, 4 pixels per pass.  Runs are handed
    //      out dynamically; the next ticket is drawn before the current run is processed so its latency is hidden.
    int it = 0;
    if (lane == 0) it = atomicAdd(&next_item, 1);
    it = __shfl_sync(0xFFFFFFFFu, it, 0);
    while (it < n_items) {
      int nxt = 0;
      if (lane == 0) nxt = atomicAdd(&next_item, 1);
      const uint2 we = runs[it];
      const int l = we.x & 0xFFF;
      const bool zero = (we.x >> 12) & 1u;
      const int py_begin = (we.x >> 16) & 0xFF, py_end = we.x >> 24;
      const float ev = zero ? 0.0f : r.extrapolation;
      const unsigned long long xv = xval[l];
      const uint2* xrow = xtab + l * Q;
      const uint2* yrow = ytab + l * Q;
      const bool fast = (S == 1) && (ext_zero || zero) && ((we.x >> 13) & 1u);
      for (int px0 = 0; px0 < P; px0 += 4) {
        const int px = px0 + sub;
        const bool act = px < P;
        const int pxc = act ? px : 0;
        float4* out_px = out_c0 + (we.y + static_cast<uint32_t>(px) * c4);
        if (fast) {
          // every sample column valid and extrapolation 0: invalid sample rows read the zero row, no selects needed
          const uint2 xp = xrow[pxc];
          const unsigned char* band_lo = band_q + (xp.x & 0xFFFFu);
          const unsigned char* band_hi = band_q + (xp.x >> 16);
          const unsigned long long wx2 = f2_splat(__uint_as_float(xp.y));
          for (int py = py_begin; py < py_end; ++py) {
            const uint2 ye = yrow[py];
            const uint32_t ta = (ye.x & 0x3FFFu) << 4, tb = ((ye.x >> 16) & 0x3FFFu) << 4;
#if defined(BX_EXP) && (BX_EXP == 2 || BX_EXP == 3)
            // experiment: no shared-memory tap loads
            ulonglong2 fake; fake.x = wx2 + ta; fake.y = wx2 + tb;
            const ulonglong2 o = lerp2_packed(fake, fake, fake, fake, wx2, f2_splat(__uint_as_float(ye.y)), nz2);
#else
            const ulonglong2 o = lerp2_packed(*reinterpret_cast<const ulonglong2*>(band_lo + ta),
                                              *reinterpret_cast<const ulonglong2*>(band_hi + ta),
                                              *reinterpret_cast<const ulonglong2*>(band_lo + tb),
                                              *reinterpret_cast<const ulonglong2*>(band_hi + tb), wx2,
                                              f2_splat(__uint_as_float(ye.y)), nz2);
#endif
#if defined(BX_EXP) && (BX_EXP == 1 || BX_EXP == 3)
            if (act && o.x == 0x123456789ull) *reinterpret_cast<ulonglong2*>(out_px) = o;   // experiment: (almost) no stores
#else
            if (act) *reinterpret_cast<ulonglong2*>(out_px) = o;
#endif
            out_px += static_cast<uint32_t>(P) * c4;
          }
        } else {
          uint32_t xlo[S], xhi[S];
          float lx[S];
          bool xok[S];
#pragma unroll
          for (int sx = 0; sx < S; ++sx) {
            const uint2 xp = xrow[pxc * S + sx];
            xlo[sx] = xp.x & 0xFFFFu;
            xhi[sx] = xp.x >> 16;
            lx[sx] = __uint_as_float(xp.y);
            xok[sx] = (xv >> (pxc * S + sx)) & 1ull;
          }
          for (int py = py_begin; py < py_end; ++py) {
            float4 acc;
#pragma unroll
            for (int sy = 0; sy < S; ++sy) {
              const uint2 ye = yrow[py * S + sy];
              const bool yok = ye.x >> 31;
              const bool inb = (ye.x >> 15) & 1u;
              const uint32_t ya = ye.x & 0x3FFFu, yb = (ye.x >> 16) & 0x3FFFu;
#pragma unroll
              for (int sx = 0; sx < S; ++sx) {
                float4 tl, tr, bl, br;
                if (inb) {
                  const unsigned char* bt = band_q + (ya << 4);
                  const unsigned char* bb = band_q + (yb << 4);
                  tl = *reinterpret_cast<const float4*>(bt + xlo[sx]);
                  tr = *reinterpret_cast<const float4*>(bt + xhi[sx]);
                  bl = *reinterpret_cast<const float4*>(bb + xlo[sx]);
                  br = *reinterpret_cast<const float4*>(bb + xhi[sx]);
                } else {  // sample row of a pooled pair reaching past the staged band: read it from L2
                  const float* gt = feat_img + (static_cast<size_t>(ya) * fw) * C + slice * kSlice + q * 4;
                  const float* gb = feat_img + (static_cast<size_t>(yb) * fw) * C + slice * kSlice + q * 4;
                  const size_t lo = xlo[sx] / (kSlice * 4), hi = xhi[sx] / (kSlice * 4);
                  tl = __ldg(reinterpret_cast<const float4*>(gt + lo * C));
                  tr = __ldg(reinterpret_cast<const float4*>(gt + hi * C));
                  bl = __ldg(reinterpret_cast<const float4*>(gb + lo * C));
                  br = __ldg(reinterpret_cast<const float4*>(gb + hi * C));
                }
                float4 v = lerp2(tl, tr, bl, br, lx[sx], __uint_as_float(ye.y));
                if (!(yok && xok[sx])) v = make_float4(ev, ev, ev, ev);
                pool_acc<POOL>(acc, v, sy == 0 && sx == 0);
              }
            }
            if (POOL == BX_POOL_AVG2) acc = make_float4(acc.x / 4.0f, acc.y / 4.0f, acc.z / 4.0f, acc.w / 4.0f);
            if (act) *out_px = acc;
            out_px += static_cast<uint32_t>(P) * c4;
          }
        }
      }
      it = __shfl_sync(0xFFFFFFFFu, nxt, 0);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int POOL>
int launch_band(bx_handle* h, const BandArgs& a, const CUtensorMap& tmap, size_t smem, cudaStream_t st) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;
  roi_plan_kernel<S><<<a.r.b * a.n_bands * a.n_chunks, kPlanThreads, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  BX_CUDA(cudaFuncSetAttribute(roi_band_kernel<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  roi_band_kernel<POOL><<<a.r.b * a.n_slices * a.n_bands, kThreads, smem, st>>>(tmap, a);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

}  // namespace

int roi_band_launch(bx_handle* h, const RoiArgs& ra, int pool, cudaStream_t st, int* used) {
  *used = 0;
  if (getenv("BX_ROI_DIRECT")) return BX_OK;                      // measurement switch: force the direct-gather kernel
  if (ra.n_levels != 1 || ra.order || ra.level) return BX_OK;     // FPN routing stays on the direct kernel
  if (ra.c % kSlice != 0 || !bx_aligned(ra.out, 16) || !bx_aligned(ra.lv[0].feat, 16)) return BX_OK;
  const int fh = ra.lv[0].fh, fw = ra.lv[0].fw;
  if (fh >= 4096 || fw >= 4096 || ra.r > 4096 * 1024) return BX_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return BX_OK;
  if (fw * kSlice * 4 > 65535 || ra.P > 31 || ra.Q > 64) return BX_OK;   // 16-bit tap offsets, 32-bit row masks
  const size_t budget = h->smem_optin - 1024;
  // rois per plan block: everything of one image when that fits in ~1/5 of shared memory
  const int rois_range = ra.roi_counts ? ra.rois_per_image : ra.r;
  int chunk = rois_range < kPlanMaxChunk ? rois_range : kPlanMaxChunk;
  while (chunk > 32 && plan_layout(chunk, ra.Q).bytes > budget / 5) chunk = (chunk + 1) / 2;
  if (chunk < 1) return BX_OK;
  if (static_cast<unsigned long long>(chunk) * ra.P * ra.P * (ra.c / 4) >= (1ull << 31)) return BX_OK;
  const int n_chunks = (rois_range + chunk - 1) / chunk;
  const size_t tab = plan_layout(chunk, ra.Q).bytes;
  // band height: as many rows (+1 halo) as fit, then balanced over the bands
  const size_t row_bytes = static_cast<size_t>(fw) * kSlice * 4;
  const size_t zrow = (row_bytes + 127) & ~static_cast<size_t>(127);       // the all-zero row behind the TMA boxes
  if (budget < tab + zrow + kBoxBytes) return BX_OK;
  int max_rows_loaded = static_cast<int>(((budget - tab - zrow) / kBoxBytes) * kBoxBytes / row_bytes);
  if (max_rows_loaded < 3) return BX_OK;                          // map too wide for a useful band: direct kernel
  int rows_per_band = max_rows_loaded - 1;
  int n_bands = (fh + rows_per_band - 1) / rows_per_band;
  rows_per_band = (fh + n_bands - 1) / n_bands;
  const int rows_loaded = (rows_per_band + 1 < fh) ? rows_per_band + 1 : fh;
  const int nbox = static_cast<int>((static_cast<size_t>(rows_loaded) * fw + kBoxRows - 1) / kBoxRows);
  const size_t smem = static_cast<size_t>(nbox) * kBoxBytes + zrow + tab;
  if (static_cast<size_t>(nbox) * kBoxBytes + zrow > (1u << 18)) return BX_OK;    // 14-bit row offsets in 16 B units
  if (smem > budget) return BX_OK;
  if (static_cast<size_t>(nbox) * kBoxBytes + tab >= (1u << 20)) return BX_OK;    // mbarrier tx-count limit

  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ra.c), static_cast<cuuint64_t>(ra.b) * fh * fw};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ra.c) * sizeof(float)};
  const cuuint32_t box[2] = {kSlice, kBoxRows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ra.lv[0].feat), gdim, gstride,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BX_REQUIRE(cr == CUDA_SUCCESS, BX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);

  const size_t plan_bytes = static_cast<size_t>(ra.b) * n_bands * n_chunks * tab;
  int rc = bx_plan_reserve(h, plan_bytes);
  if (rc) return rc;

  BandArgs a;
  a.r = ra;
  a.rows_per_band = rows_per_band;
  a.n_bands = n_bands;
  a.n_slices = ra.c / kSlice;
  a.chunk = chunk;
  a.n_chunks = n_chunks;
  a.nbox = nbox;
  a.scan_all = (ra.box_ind != nullptr && !ra.roi_counts) ? 1 : 0;
  a.plan = static_cast<unsigned char*>(h->plan);
  a.neg_zero = -0.0f;
  if (pool == BX_POOL_NONE) rc = launch_band<BX_POOL_NONE>(h, a, tmap, smem, st);
  else if (pool == BX_POOL_MAX2) rc = launch_band<BX_POOL_MAX2>(h, a, tmap, smem, st);
  else rc = launch_band<BX_POOL_AVG2>(h, a, tmap, smem, st);
  if (rc == BX_OK) *used = 1;
  return rc;
}

}  // namespace bxroi
