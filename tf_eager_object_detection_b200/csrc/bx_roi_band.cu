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

constexpr int kSlice = 32;          // channels per CTA (128 B per pixel)
constexpr int kBoxRows = 256;       // max pixels per TMA box

struct BandArgs {
  RoiArgs r;
  int rows_per_band, n_bands, n_slices;
  int cap;          // rois per plan block (shared-memory table capacity of the band kernel)
  int n_blocks_max; // plan blocks reserved per (image, band)
  int nbox;         // TMA boxes per band
  int box_rows;     // pixels per TMA box (<= 256)
  int band_bytes;   // nbox * box_rows * 128: the all-zero row starts here
  int scan_all;     // 1: rois of any image may be anywhere (box_ind given) -> every plan CTA scans every roi
  unsigned char* plan;   // [b, n_bands, n_blocks_max] plan blocks (global workspace)
  float neg_zero;        // -0.0f at run time: addend that turns fma.rn.f32x2 into an exact, non-contractible multiply
  unsigned long long* dbg;   // optional [grid,4] timestamps (globaltimer ns: start, band ready, done; smid) — BX_BAND_DEBUG
  // persistent kernel: units beyond the first one of a CTA are drawn from *unit_counter (zeroed by the plan kernel)
  int* unit_counter;
  int n_units;               // b * n_slices * n_bands
  int band_stride;           // bytes between the two band buffers (band + zero row, 128-byte aligned); 0: one buffer
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
// (image, band) by roi_plan_kernel and stored in global memory as contiguous "plan blocks" of `cap` rois; only the rois
// that own output rows in the band are listed (compacted).  Each band CTA pulls a block into shared memory with one
// bulk copy (cp.async.bulk) next to the TMA tile loads of the band.
//
// Plan block layout (byte offsets; Q = crop size):
//   [0]   int n_runs ; [4] int n_blocks_used (block 0 only) ; padding to 16 B
//   [16]  u64   xval [cap]      validity bit per sample column
//   [..]  uint2 xtab [cap*Q]    (lo_off | hi_off << 16: byte offsets of the tap columns in a band row, lerp_x)
//   [..]  uint2 ytab [cap*Q]    (a | in_band << 15 | b << 16 | valid << 31, lerp_y): in_band: a, b = offsets (16 B units) of
//                               the top / bottom tap rows in the staged band (the zero row when the sample is invalid);
//                               else absolute feature rows (pooled pairs only)
//   [..]  uint2 runs [4*cap]    (slot | zero << 12 | all_x_valid << 13 | all_rows_in_band << 14 | py_begin << 16 | py_end << 24,
//                                offset of output pixel (roi, py_begin, 0) in float4 units from the image's first roi)
struct PlanLayout {
  uint32_t off_xval, off_xtab, off_ytab, off_runs, bytes;
};

__host__ __device__ inline PlanLayout plan_layout(int cap, int Q) {
  PlanLayout L;
  L.off_xval = 16;
  L.off_xtab = L.off_xval + 8u * cap;
  L.off_ytab = L.off_xtab + 8u * cap * Q;
  L.off_runs = L.off_ytab + 8u * cap * Q;
  L.bytes = (L.off_runs + 32u * cap + 127u) & ~127u;
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
constexpr int kPlanPass = 1024;      // rois examined per pass of the plan kernel
constexpr int kPlanMaxBlocks = 64;

template <int S>
__global__ void __launch_bounds__(kPlanThreads) roi_plan_kernel(const BandArgs a) {
  __shared__ float4 rpar[kPlanPass];
  __shared__ uint32_t pym[kPlanPass];
  __shared__ uint32_t xv_lo[kPlanPass], xv_hi[kPlanPass];
  __shared__ uint32_t oobm[kPlanPass];   // output rows with a valid sample row outside the staged band (read from L2)
  __shared__ uint32_t cidx[kPlanPass];
  __shared__ unsigned char rflag[kPlanPass];
  __shared__ int blk_runs[kPlanMaxBlocks];
  __shared__ int total_j;
  // programmatic dependent launch: let the band kernel start (and issue its TMA band loads) while the plan is computed;
  // it waits on griddepcontrol.wait before touching the plan
  asm volatile("griddepcontrol.launch_dependents;");
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.unit_counter) *a.unit_counter = 0;   // read behind griddepcontrol.wait only
  const RoiArgs& r = a.r;
  const int tid = threadIdx.x, lane = tid & 31;
  const int band_i = blockIdx.x % a.n_bands;
  const int img = blockIdx.x / a.n_bands;
  const int P = r.P, Q = r.Q;
  const int fh = r.lv[0].fh, fw = r.lv[0].fw;
  const int r0 = band_i * a.rows_per_band;
  const int r1 = min(fh, r0 + a.rows_per_band);
  const int rows_loaded = min(r1 + 1, fh) - r0;
  const int pad = (r.mode == BX_ROI_ALIGN_PAD) ? 1 : 0;
  const int dimy = fh + 2 * pad, dimx = fw + 2 * pad;
  const uint32_t row_bytes = static_cast<uint32_t>(fw) * (kSlice * 4);
  const uint32_t zrow_off = static_cast<uint32_t>(a.band_bytes);
  const PlanLayout L = plan_layout(a.cap, Q);
  unsigned char* blk0 = a.plan + static_cast<size_t>(blockIdx.x) * a.n_blocks_max * L.bytes;

  int g_begin = 0, g_end = r.r;
  if (!a.scan_all && r.roi_counts) {
    g_begin = img * r.rois_per_image;
    g_end = g_begin + r.rois_per_image;
  }
  if (tid < kPlanMaxBlocks) blk_runs[tid] = 0;
  if (tid == 0) total_j = 0;
  __syncthreads();

  for (int c0 = g_begin; c0 < g_end; c0 += kPlanPass) {
    const int nroi = min(kPlanPass, g_end - c0);
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
      oobm[l] = 0u;
      rflag[l] = static_cast<unsigned char>(flag);
    }
    __syncthreads();
    // ---- ownership: the band that owns output row py is the one holding the top tap row of its first valid sample row
    for (int idx = tid; idx < nroi * P; idx += kPlanThreads) {
      const int l = idx / P, py = idx - l * P;
      const uint32_t flag = rflag[l];
      if (!(flag & 1u)) continue;
      const bool zero = flag & 2u;
      const float4 par = rpar[l];
      int owner_row = -1;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        int lo, hi; float lerp; bool valid;
        axis_from_par(par.z, par.w, py * S + s, dimy, pad, lo, hi, lerp, valid);
        if (valid && !zero && owner_row < 0) owner_row = lo;
      }
      const int owner = owner_row < 0 ? 0 : min(owner_row / a.rows_per_band, a.n_bands - 1);
      if (owner == band_i) atomicOr(&pym[l], 1u << py);
    }
    __syncthreads();
    // ---- compact the rois that own rows here (order within the plan is irrelevant: outputs are disjoint)
    for (int l0 = 0; l0 < nroi; l0 += kPlanThreads) {
      const int l = l0 + tid;
      const bool has = (l < nroi) && pym[l] != 0u;
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
      if (bal) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&total_j, __popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (has) cidx[l] = static_cast<uint32_t>(base + __popc(bal & ((1u << lane) - 1u)));
      }
    }
    __syncthreads();
    // ---- x / y sample tables of the listed rois
    for (int idx = tid; idx < nroi * Q; idx += kPlanThreads) {
      const int l = idx / Q, s = idx - l * Q;
      if (pym[l] == 0u) continue;
      const int j = cidx[l], bi = j / a.cap, slot = j - bi * a.cap;
      if (bi >= a.n_blocks_max) continue;
      unsigned char* blk = blk0 + static_cast<size_t>(bi) * L.bytes;
      const bool zero = rflag[l] & 2u;
      const float4 par = rpar[l];
      int lo, hi;
      float lerp;
      bool valid;
      axis_from_par(par.x, par.y, s, dimx, pad, lo, hi, lerp, valid);
      valid = valid && !zero;
      reinterpret_cast<uint2*>(blk + L.off_xtab)[slot * Q + s] =
          make_uint2(valid ? (static_cast<uint32_t>(lo * (kSlice * 4)) | (static_cast<uint32_t>(hi * (kSlice * 4)) << 16)) : 0u,
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
        atomicOr(&oobm[l], 1u << (s / S));
      }
      reinterpret_cast<uint2*>(blk + L.off_ytab)[slot * Q + s] =
          make_uint2(ya | ((inb || !valid) ? (1u << 15) : 0u) | (yb << 16) | (valid ? (1u << 31) : 0u),
                     __float_as_uint(valid ? lerp : 0.0f));
    }
    __syncthreads();
    // ---- validity masks
    for (int l = tid; l < nroi; l += kPlanThreads) {
      if (pym[l] == 0u) continue;
      const int j = cidx[l], bi = j / a.cap, slot = j - bi * a.cap;
      if (bi >= a.n_blocks_max) continue;
      unsigned char* blk = blk0 + static_cast<size_t>(bi) * L.bytes;
      reinterpret_cast<unsigned long long*>(blk + L.off_xval)[slot] =
          static_cast<unsigned long long>(xv_lo[l]) | (static_cast<unsigned long long>(xv_hi[l]) << 32);
    }
    // ---- one work item per run of consecutive owned rows, longest runs first: the band kernel deals runs to its warps
    //      round-robin, so descending length keeps the warps balanced without any run-time scheduling
    for (int want = P; want >= 1; --want) {
      for (int l = tid; l < nroi; l += kPlanThreads) {
        uint32_t m = pym[l];
        if (m == 0u) continue;
        const int j = cidx[l], bi = j / a.cap, slot = j - bi * a.cap;
        if (bi >= a.n_blocks_max) continue;
        unsigned char* blk = blk0 + static_cast<size_t>(bi) * L.bytes;
        const unsigned long long xv = static_cast<unsigned long long>(xv_lo[l]) | (static_cast<unsigned long long>(xv_hi[l]) << 32);
        const unsigned long long full = (Q >= 64) ? ~0ull : ((1ull << Q) - 1ull);
        const uint32_t z = ((static_cast<uint32_t>(rflag[l]) & 2u) << 11) | ((xv == full) ? (1u << 13) : 0u);
        uint2* g_runs = reinterpret_cast<uint2*>(blk + L.off_runs);
        const uint32_t roi_in_img = static_cast<uint32_t>(c0 - g_begin + l);
        while (m) {
          const int b0 = __ffs(m) - 1;
          const int len = __ffs(~(m >> b0)) - 1;               // first zero above b0 ends the run (len <= P < 32)
          if (len == want) {
            const int pos = atomicAdd(&blk_runs[bi], 1);
            const uint32_t inb_all = ((oobm[l] >> b0) & ((1u << len) - 1u)) == 0u ? (1u << 14) : 0u;
            if (pos < 4 * a.cap)
              g_runs[pos] = make_uint2(static_cast<uint32_t>(slot) | z | inb_all | (static_cast<uint32_t>(b0) << 16) |
                                           (static_cast<uint32_t>(b0 + len) << 24),
                                       ((roi_in_img * P + b0) * P) * static_cast<uint32_t>(r.c / 4));
          }
          m &= ~(((1u << len) - 1u) << b0);
        }
      }
      __syncthreads();
    }
  }
  if (tid < a.n_blocks_max) {
    int* hdr = reinterpret_cast<int*>(blk0 + static_cast<size_t>(tid) * L.bytes);
    hdr[0] = min(blk_runs[tid], 4 * a.cap);
    hdr[1] = min((total_j + a.cap - 1) / a.cap, a.n_blocks_max);
  }
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

// One plan block of one unit (image, channel slice, band): every run of output rows listed in `tab`, taps from the band staged
// at `band` (zero row at band + a.band_bytes).  Shared by the per-unit kernel and the persistent kernel.
template <int POOL, int THREADS>
__device__ __forceinline__ void band_process_block(const BandArgs& a, const unsigned char* __restrict__ tab,
                                                   const unsigned char* __restrict__ band, const float* __restrict__ feat_img,
                                                   int slice, int g_begin, int tid) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;
  const RoiArgs& r = a.r;
  const int fw = r.lv[0].fw;
  const int P = r.P, Q = r.Q, C = r.c;
  const PlanLayout L = plan_layout(a.cap, Q);
  const unsigned long long* xval = reinterpret_cast<const unsigned long long*>(tab + L.off_xval);
  const uint2* xtab = reinterpret_cast<const uint2*>(tab + L.off_xtab);
  const uint2* ytab = reinterpret_cast<const uint2*>(tab + L.off_ytab);
  const uint2* runs = reinterpret_cast<const uint2*>(tab + L.off_runs);
  const int lane = tid & 31;
  const int q = lane & 7, sub = lane >> 3;
  const unsigned char* band_q = band + q * 16;
  const bool ext_zero = (r.extrapolation == 0.0f);
  const unsigned long long nz2 = f2_splat(a.neg_zero);
  // output pixel (first roi of the image, py 0, px 0), this unit's channel slice, this lane's 4 channels
  float4* out_c0 = reinterpret_cast<float4*>(r.out + static_cast<size_t>(g_begin) * P * P * C + slice * kSlice + q * 4);
  const uint32_t c4 = static_cast<uint32_t>(C) >> 2;
  const int n_items = *reinterpret_cast<const int*>(tab);
  // ---- one warp per run of output rows of a roi; 8 lanes (32 channels) per pixel.  Runs are sorted by descending
  //      length in the plan and dealt round-robin to the warps.
  for (int it = tid >> 5; it < n_items; it += THREADS / 32) {
    const uint2 we = runs[it];
    const int l = we.x & 0xFFF;
    const bool zero = (we.x >> 12) & 1u;
    const int py_begin = (we.x >> 16) & 0xFF, py_end = we.x >> 24;
    const float ev = zero ? 0.0f : r.extrapolation;
    const unsigned long long xv = xval[l];
    const uint2* xrow = xtab + l * Q;
    const uint2* yrow = ytab + l * Q;
    // fast paths: every sample column valid, extrapolation 0 and (pooled crops) every valid sample row inside the
    // staged band: invalid sample rows read the zero row, so no selects and no validity tests are needed
    // (sample columns outside the map — rois clipped at the right / bottom image border end there in the C4 geometry —
    //  read pixel 0 and are zeroed by one select per output, so they stay on the fast paths)
    const bool fast = (ext_zero || zero) && (S == 1 || ((we.x >> 14) & 1u));
    if (fast && S == 1) {
      // two pixels per lane (px and px + 4) share each row's y parameters: 8 independent tap loads in flight per row
      for (int px0 = 0; px0 < P; px0 += 8) {
        const int pxA = px0 + sub, pxB = pxA + 4;
        const bool actA = pxA < P, actB = pxB < P;
        const uint2 xa = xrow[actA ? pxA : 0], xb = xrow[actB ? pxB : 0];
        const unsigned char* a_lo = band_q + (xa.x & 0xFFFFu);
        const unsigned char* a_hi = band_q + (xa.x >> 16);
        const unsigned char* b_lo = band_q + (xb.x & 0xFFFFu);
        const unsigned char* b_hi = band_q + (xb.x >> 16);
        const unsigned long long wa2 = f2_splat(__uint_as_float(xa.y)), wb2 = f2_splat(__uint_as_float(xb.y));
        const bool okA = (xv >> (actA ? pxA : 0)) & 1ull, okB = (xv >> (actB ? pxB : 0)) & 1ull;
        float4* out_a = out_c0 + (we.y + static_cast<uint32_t>(pxA) * c4);
        const bool two = px0 + 4 < P;                        // warp-uniform: is there a second group of pixels?
        for (int py = py_begin; py < py_end; ++py) {
          const uint2 ye = yrow[py];
          const uint32_t ta = (ye.x & 0x3FFFu) << 4, tb = ((ye.x >> 16) & 0x3FFFu) << 4;
          const unsigned long long wy2 = f2_splat(__uint_as_float(ye.y));
          ulonglong2 oa = lerp2_packed(*reinterpret_cast<const ulonglong2*>(a_lo + ta),
                                       *reinterpret_cast<const ulonglong2*>(a_hi + ta),
                                       *reinterpret_cast<const ulonglong2*>(a_lo + tb),
                                       *reinterpret_cast<const ulonglong2*>(a_hi + tb), wa2, wy2, nz2);
          if (!okA) oa = make_ulonglong2(0ull, 0ull);
          if (two) {
            // lanes whose second pixel does not exist (px >= P) issue no loads: their 128 B would be a wasted
            // shared-memory wavefront per tap, and this kernel is bound by the LSU data pipe
            ulonglong2 b0 = make_ulonglong2(0ull, 0ull), b1 = b0, b2 = b0, b3 = b0;
            if (actB) {
              b0 = *reinterpret_cast<const ulonglong2*>(b_lo + ta);
              b1 = *reinterpret_cast<const ulonglong2*>(b_hi + ta);
              b2 = *reinterpret_cast<const ulonglong2*>(b_lo + tb);
              b3 = *reinterpret_cast<const ulonglong2*>(b_hi + tb);
            }
            ulonglong2 ob = lerp2_packed(b0, b1, b2, b3, wb2, wy2, nz2);
            if (!okB) ob = make_ulonglong2(0ull, 0ull);
            if (actB) *reinterpret_cast<ulonglong2*>(out_a + 4 * c4) = ob;
          }
          if (actA) *reinterpret_cast<ulonglong2*>(out_a) = oa;
          out_a += static_cast<uint32_t>(P) * c4;
        }
      }
    } else if (fast) {
      // pooled crop (2x2 samples per output pixel), packed math, max / mean in the slow path's order (sy major)
      for (int px0 = 0; px0 < P; px0 += 4) {
        const int px = px0 + sub;
        const bool act = px < P;
        const int pxc = act ? px : 0;
        const uint2 x0 = xrow[pxc * 2], x1 = xrow[pxc * 2 + 1];
        const unsigned char* lo0 = band_q + (x0.x & 0xFFFFu);
        const unsigned char* hi0 = band_q + (x0.x >> 16);
        const unsigned char* lo1 = band_q + (x1.x & 0xFFFFu);
        const unsigned char* hi1 = band_q + (x1.x >> 16);
        const unsigned long long w0 = f2_splat(__uint_as_float(x0.y)), w1 = f2_splat(__uint_as_float(x1.y));
        const bool ok0 = (xv >> (pxc * 2)) & 1ull, ok1 = (xv >> (pxc * 2 + 1)) & 1ull;
        float4* out_px = out_c0 + (we.y + static_cast<uint32_t>(px) * c4);
        for (int py = py_begin; py < py_end; ++py) {
          ulonglong2 acc;
#pragma unroll
          for (int sy = 0; sy < 2; ++sy) {
            const uint2 ye = yrow[py * 2 + sy];
            const uint32_t ta = (ye.x & 0x3FFFu) << 4, tb = ((ye.x >> 16) & 0x3FFFu) << 4;
            const unsigned long long wy2 = f2_splat(__uint_as_float(ye.y));
            ulonglong2 v0 = lerp2_packed(*reinterpret_cast<const ulonglong2*>(lo0 + ta),
                                         *reinterpret_cast<const ulonglong2*>(hi0 + ta),
                                         *reinterpret_cast<const ulonglong2*>(lo0 + tb),
                                         *reinterpret_cast<const ulonglong2*>(hi0 + tb), w0, wy2, nz2);
            ulonglong2 v1 = lerp2_packed(*reinterpret_cast<const ulonglong2*>(lo1 + ta),
                                         *reinterpret_cast<const ulonglong2*>(hi1 + ta),
                                         *reinterpret_cast<const ulonglong2*>(lo1 + tb),
                                         *reinterpret_cast<const ulonglong2*>(hi1 + tb), w1, wy2, nz2);
            if (!ok0) v0 = make_ulonglong2(0ull, 0ull);
            if (!ok1) v1 = make_ulonglong2(0ull, 0ull);
            if (sy == 0) acc = v0; else acc = pool2<POOL>(acc, v0);
            acc = pool2<POOL>(acc, v1);
          }
          if (POOL == BX_POOL_AVG2) {
            const unsigned long long q4 = f2_splat(0.25f);   // x / 4 == x * 0.25 exactly (power of two)
            acc.x = f2_mul(acc.x, q4, nz2);
            acc.y = f2_mul(acc.y, q4, nz2);
          }
          if (act) *reinterpret_cast<ulonglong2*>(out_px) = acc;
          out_px += static_cast<uint32_t>(P) * c4;
        }
      }
    } else {
      for (int px0 = 0; px0 < P; px0 += 4) {
        const int px = px0 + sub;
        const bool act = px < P;
        const int pxc = act ? px : 0;
        float4* out_px = out_c0 + (we.y + static_cast<uint32_t>(px) * c4);
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
  }
}

template <int POOL, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 512) ? 2 : 1)
roi_band_kernel(const __grid_constant__ CUtensorMap tmap, const BandArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const RoiArgs& r = a.r;
  const int fh = r.lv[0].fh, fw = r.lv[0].fw;
  const uint32_t row_bytes = static_cast<uint32_t>(fw) * (kSlice * 4);
  const uint32_t zrow_off = static_cast<uint32_t>(a.band_bytes);            // one all-zero band row after the TMA boxes
  const PlanLayout L = plan_layout(a.cap, r.Q);
  // two plan-block buffers: block ch + 1 is copied in while block ch is processed (a (image, band) unit lists
  // ~150 rois at cfg2, more than one block holds; waiting for the second block used to cost 6 % of the kernel)
  unsigned char* tab0 = smem + zrow_off + ((row_bytes + 127u) & ~127u);
  const uint32_t tab_step = a.n_blocks_max > 1 ? L.bytes : 0u;      // offset of the second buffer
  __shared__ uint64_t mbar[2];

  const int tid = threadIdx.x;
  const int C = r.c;
  if (a.dbg && tid == 0) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    a.dbg[blockIdx.x * 4 + 0] = t; a.dbg[blockIdx.x * 4 + 3] = sm;
  }
  // band-major CTA order: band 0 (which also owns every roi's invalid rows) is the heaviest, the last band the
  // lightest, so the hardware dispatcher hands out the long units first and the tail of the launch is short
  int u = blockIdx.x;
  const int per_band = a.n_slices * r.b;
  const int band_i = u / per_band; u -= band_i * per_band;
  const int slice = u % a.n_slices;
  const int img = u / a.n_slices;
  const int r0 = band_i * a.rows_per_band;
  const float* feat_img = r.lv[0].feat + static_cast<size_t>(img) * fh * fw * C;
  const unsigned char* plan0 = a.plan + static_cast<size_t>(img * a.n_bands + band_i) * a.n_blocks_max * L.bytes;

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_expect_tx(&mbar[0], static_cast<uint32_t>(a.band_bytes) + L.bytes);
    const int row0 = (img * fh + r0) * fw;
    for (int bx = 0; bx < a.nbox; ++bx)
      tma_load_2d(smem + static_cast<size_t>(bx) * a.box_rows * (kSlice * 4), &tmap, slice * kSlice, row0 + bx * a.box_rows, &mbar[0]);
    asm volatile("griddepcontrol.wait;" ::: "memory");    // the plan kernel (previous launch in the stream) is complete
    bulk_load(tab0, plan0, L.bytes, &mbar[0]);
  }
  for (uint32_t i = tid * 16u; i < row_bytes; i += THREADS * 16u)
    *reinterpret_cast<float4*>(smem + zrow_off + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  int g_begin = 0;   // first roi of the image's range
  if (!a.scan_all && r.roi_counts) g_begin = img * r.rois_per_image;

  int n_blocks = 1;
  for (int ch = 0; ch < n_blocks; ++ch) {
    const unsigned char* tab = tab0 + (ch & 1) * tab_step;
    mbar_wait(&mbar[ch & 1], static_cast<uint32_t>((ch >> 1) & 1));
    if (ch == 0) n_blocks = reinterpret_cast<const int*>(tab)[1];
    if (ch + 1 < n_blocks) {                 // prefetch the next plan block into the other buffer
      if (ch >= 1) __syncthreads();          // ... which block ch - 1 used: everyone is done with it
      if (tid == 0) {
        mbar_expect_tx(&mbar[(ch + 1) & 1], L.bytes);
        bulk_load(tab0 + ((ch + 1) & 1) * tab_step, plan0 + static_cast<size_t>(ch + 1) * L.bytes, L.bytes, &mbar[(ch + 1) & 1]);
      }
    }
    if (a.dbg && tid == 0 && ch == 0) {
      unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.dbg[blockIdx.x * 4 + 1] = t;
    }
    band_process_block<POOL, THREADS>(a, tab, smem, feat_img, slice, g_begin, tid);
  }
  if (a.dbg) {
    __syncthreads();
    if (tid == 0) {
      unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.dbg[blockIdx.x * 4 + 2] = t;
    }
  }
}

// ---- persistent form: one 1024-thread CTA per SM loops over units (image, channel slice, band), heaviest bands first.
// The per-unit kernel keeps two 512-thread CTAs on an SM so that one CTA's TMA wait overlaps the other's arithmetic, but
// the %globaltimer timeline shows a CTA waiting 5.8 us for its band against 20.8 us of work, and the SM's two slots 90 %
// occupied.  Here the band (and the first plan block) of the NEXT unit is copied into a second buffer while all 32 warps
// work on the current one, so after the first unit nothing waits for TMA, and there is no wave quantisation: units are
// drawn from a global counter (the first one is the CTA's index).  Arithmetic and results are unchanged.
// MEASURED SLOWER than the per-unit kernel (see roi_band_launch): opt-in, kept for comparison.
template <int POOL>
__global__ void __launch_bounds__(1024, 1) roi_band_persist_kernel(const __grid_constant__ CUtensorMap tmap, const BandArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t band_bar[2], plan_bar[2];
  __shared__ int s_next[2];
  const RoiArgs& r = a.r;
  const int fh = r.lv[0].fh, fw = r.lv[0].fw, C = r.c;
  const uint32_t row_bytes = static_cast<uint32_t>(fw) * (kSlice * 4);
  const uint32_t zrow_bytes = (row_bytes + 127u) & ~127u;
  const PlanLayout L = plan_layout(a.cap, r.Q);
  const int n_band_buf = a.band_stride ? 2 : 1;
  unsigned char* tab0 = smem + (a.band_stride ? 2u * static_cast<uint32_t>(a.band_stride) : static_cast<uint32_t>(a.band_bytes) + zrow_bytes);
  const int tid = threadIdx.x;
  const int per_band = a.n_slices * r.b;
  auto unit_of = [&](int u, int& band_i, int& slice, int& img) {
    band_i = u / per_band;
    const int v = u - band_i * per_band;
    slice = v % a.n_slices;
    img = v / a.n_slices;
  };
  auto issue_band = [&](int u, int buf) {       // thread 0: TMA boxes of unit u's band into band buffer `buf`
    int band_i, slice, img;
    unit_of(u, band_i, slice, img);
    unsigned char* dst = smem + static_cast<size_t>(buf) * a.band_stride;
    mbar_expect_tx(&band_bar[buf], static_cast<uint32_t>(a.band_bytes));
    const int row0 = (img * fh + band_i * a.rows_per_band) * fw;
    for (int bx = 0; bx < a.nbox; ++bx)
      tma_load_2d(dst + static_cast<size_t>(bx) * a.box_rows * (kSlice * 4), &tmap, slice * kSlice, row0 + bx * a.box_rows, &band_bar[buf]);
  };
  auto plan_of = [&](int u) {
    int band_i, slice, img;
    unit_of(u, band_i, slice, img);
    return a.plan + static_cast<size_t>(img * a.n_bands + band_i) * a.n_blocks_max * L.bytes;
  };
  auto issue_plan = [&](const unsigned char* src, uint32_t seq) {   // thread 0: one plan block into buffer seq & 1
    mbar_expect_tx(&plan_bar[seq & 1u], L.bytes);
    bulk_load(tab0 + (seq & 1u) * L.bytes, src, L.bytes, &plan_bar[seq & 1u]);
  };

  int u = blockIdx.x;                           // gridDim.x <= n_units
  if (tid == 0) {
    mbar_init(&band_bar[0], 1); mbar_init(&band_bar[1], 1);
    mbar_init(&plan_bar[0], 1); mbar_init(&plan_bar[1], 1);
    issue_band(u, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");    // the plan kernel is complete: plan blocks and the unit counter are valid
    issue_plan(plan_of(u), 0u);
  }
  // the all-zero row behind every band buffer
  for (int bfi = 0; bfi < n_band_buf; ++bfi)
    for (uint32_t i = tid * 16u; i < row_bytes; i += 1024u * 16u)
      *reinterpret_cast<float4*>(smem + static_cast<size_t>(bfi) * a.band_stride + a.band_bytes + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  uint32_t seq = 0;                             // plan blocks processed so far by this CTA (buffer = seq & 1)
  for (int k = 0; u < a.n_units; ++k) {         // k-th unit of this CTA: band buffer k & 1 (0 when single-buffered)
    const int bbuf = (n_band_buf == 2) ? (k & 1) : 0;
    int band_i, slice, img;
    unit_of(u, band_i, slice, img);
    const float* feat_img = r.lv[0].feat + static_cast<size_t>(img) * fh * fw * C;
    const unsigned char* plan0 = plan_of(u);
    int g_begin = 0;
    if (!a.scan_all && r.roi_counts) g_begin = img * r.rois_per_image;
    // The next unit is drawn now and its band requested at once: the other band buffer is free, everybody is past the
    // previous unit (the barrier that closes every unit, or the start-up barrier).  s_next is double-buffered by k: the
    // slot written here is read behind the closing barrier of THIS unit and rewritten two units later.
    if (tid == 0) {
      const int nu = atomicAdd(a.unit_counter, 1) + static_cast<int>(gridDim.x);
      s_next[k & 1] = nu;
      if (n_band_buf == 2 && nu < a.n_units) issue_band(nu, bbuf ^ 1);
    }
    mbar_wait(&band_bar[bbuf], static_cast<uint32_t>(((n_band_buf == 2) ? (k >> 1) : k) & 1));
    const unsigned char* band = smem + static_cast<size_t>(bbuf) * a.band_stride;
    int n_blocks = 1;
    for (int ch = 0; ch < n_blocks; ++ch, ++seq) {
      const unsigned char* tab = tab0 + (seq & 1u) * L.bytes;
      mbar_wait(&plan_bar[seq & 1u], (seq >> 1) & 1u);
      if (ch == 0) n_blocks = reinterpret_cast<const int*>(tab)[1];
      // prefetch the plan block that comes next in this CTA's sequence — the unit's next block, or block 0 of the next
      // unit — into the other buffer, which block seq - 1 used: everyone is done with it (barrier below for ch >= 1, the
      // closing barrier of the previous unit for ch == 0)
      if (ch >= 1) __syncthreads();
      if (tid == 0) {
        const int nu = s_next[k & 1];           // thread 0's own write
        if (ch + 1 < n_blocks) issue_plan(plan0 + static_cast<size_t>(ch + 1) * L.bytes, seq + 1u);
        else if (nu < a.n_units) issue_plan(plan_of(nu), seq + 1u);
      }
      band_process_block<POOL, 1024>(a, tab, band, feat_img, slice, g_begin, tid);
    }
    __syncthreads();                            // closes the unit: its band buffer and last plan buffer are free, s_next is visible
    const int nu = s_next[k & 1];
    if (n_band_buf == 1 && tid == 0 && nu < a.n_units) issue_band(nu, 0);   // single band buffer: only now
    u = nu;
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

template <int POOL, int THREADS>
int launch_band_t(bx_handle* h, const BandArgs& a, const CUtensorMap& tmap, size_t smem, cudaStream_t st) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;
  roi_plan_kernel<S><<<a.r.b * a.n_bands, kPlanThreads, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  BX_CUDA(cudaFuncSetAttribute(roi_band_kernel<POOL, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BX_CUDA(cudaFuncSetAttribute(roi_band_kernel<POOL, THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.r.b * a.n_slices * a.n_bands);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BX_CUDA(cudaLaunchKernelEx(&cfg, roi_band_kernel<POOL, THREADS>, tmap, a));
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

template <int POOL>
int launch_band(bx_handle* h, const BandArgs& a, const CUtensorMap& tmap, size_t smem, int threads, cudaStream_t st) {
  return threads == 512 ? launch_band_t<POOL, 512>(h, a, tmap, smem, st) : launch_band_t<POOL, 1024>(h, a, tmap, smem, st);
}

// shared-memory plan of one band configuration; returns false when it does not fit `budget`
struct BandCfg {
  int rows_per_band, n_bands, nbox, box_rows, band_bytes, cap, n_blocks_max;
  size_t smem;
};

bool make_band_cfg(int fh, int fw, int Q, int rois_range, size_t budget, BandCfg* c) {
  const size_t row_bytes = static_cast<size_t>(fw) * kSlice * 4;
  const size_t zrow = (row_bytes + 127) & ~static_cast<size_t>(127);       // the all-zero row behind the TMA boxes
  // table capacity: the rois of one image that own rows in one band — about (band height + roi height) / map height
  // of them; sized generously, overflow spills into further plan blocks handled sequentially by the same CTA
  // (two buffers of `cap` rois when a unit can need more than one block: the next block is prefetched)
  int cap = rois_range < 112 ? rois_range : 112;
  if (cap < 1) cap = 1;
  while (cap > 32 && 2 * plan_layout(cap, Q).bytes > budget / 3) cap = (cap + 1) / 2;
  const size_t tab = plan_layout(cap, Q).bytes * (((rois_range + cap - 1) / cap > 1) ? 2 : 1);
  if (budget < tab + zrow + 3 * row_bytes + 256) return false;
  int max_rows_loaded = static_cast<int>((budget - tab - zrow - 256) / row_bytes);
  if (max_rows_loaded < 3) return false;                                    // map too wide for a useful band
  if (max_rows_loaded > fh) max_rows_loaded = fh;
  if (const char* e = getenv("BX_ROI_BAND_ROWS")) {   // measurement switch: cap on the rows a band owns
    const int cap_rows = atoi(e);
    if (cap_rows >= 2 && cap_rows + 1 < max_rows_loaded) max_rows_loaded = cap_rows + 1;
  }
  int rows_per_band = (max_rows_loaded >= fh) ? fh : max_rows_loaded - 1;
  const int n_bands = (fh + rows_per_band - 1) / rows_per_band;
  rows_per_band = (fh + n_bands - 1) / n_bands;
  const int rows_loaded = (rows_per_band + 1 < fh) ? rows_per_band + 1 : fh;
  const int px = rows_loaded * fw;
  const int nbox = (px + kBoxRows - 1) / kBoxRows;
  const int box_rows = (px + nbox - 1) / nbox;
  c->rows_per_band = rows_per_band;
  c->n_bands = n_bands;
  c->nbox = nbox;
  c->box_rows = box_rows;
  c->band_bytes = nbox * box_rows * kSlice * 4;
  c->cap = cap;
  c->n_blocks_max = (rois_range + cap - 1) / cap;
  c->smem = static_cast<size_t>(c->band_bytes) + zrow + tab;
  if (c->n_blocks_max > kPlanMaxBlocks) return false;
  if (static_cast<size_t>(c->band_bytes) + zrow > (1u << 18)) return false;      // 14-bit row offsets in 16 B units
  if (static_cast<size_t>(c->band_bytes) + tab >= (1u << 20)) return false;       // mbarrier tx-count limit
  return c->smem <= budget;
}

// configuration of the persistent kernel: two plan buffers always (block 0 of the next unit is prefetched), two band
// buffers when they fit (else one: wide maps), bands as tall as the room allows
bool make_persist_cfg(int fh, int fw, int Q, int rois_range, size_t budget, BandCfg* c, int* band_stride) {
  const size_t row_bytes = static_cast<size_t>(fw) * kSlice * 4;
  const size_t zrow = (row_bytes + 127) & ~static_cast<size_t>(127);
  int cap = rois_range < 112 ? rois_range : 112;
  if (cap < 1) cap = 1;
  while (cap > 32 && 2 * plan_layout(cap, Q).bytes > budget / 5) cap = (cap + 1) / 2;
  const size_t tab = 2 * static_cast<size_t>(plan_layout(cap, Q).bytes);
  if (budget < tab + zrow + 3 * row_bytes + 256) return false;
  int bufs = 2;
  long long room = (static_cast<long long>(budget) - static_cast<long long>(tab)) / 2 - static_cast<long long>(zrow) - 128;
  if (room < static_cast<long long>(3 * row_bytes)) {
    bufs = 1;
    room = static_cast<long long>(budget) - static_cast<long long>(tab) - static_cast<long long>(zrow) - 128;
    if (room < static_cast<long long>(3 * row_bytes)) return false;
  }
  int max_rows_loaded = static_cast<int>(room / static_cast<long long>(row_bytes));
  if (max_rows_loaded > fh) max_rows_loaded = fh;
  int rows_per_band = (max_rows_loaded >= fh) ? fh : max_rows_loaded - 1;
  const int n_bands = (fh + rows_per_band - 1) / rows_per_band;
  rows_per_band = (fh + n_bands - 1) / n_bands;
  const int rows_loaded = (rows_per_band + 1 < fh) ? rows_per_band + 1 : fh;
  const int px = rows_loaded * fw;
  const int nbox = (px + kBoxRows - 1) / kBoxRows;
  const int box_rows = (px + nbox - 1) / nbox;
  c->rows_per_band = rows_per_band;
  c->n_bands = n_bands;
  c->nbox = nbox;
  c->box_rows = box_rows;
  c->band_bytes = nbox * box_rows * kSlice * 4;
  c->cap = cap;
  c->n_blocks_max = (rois_range + cap - 1) / cap;
  if (c->n_blocks_max > kPlanMaxBlocks) return false;
  if (static_cast<size_t>(c->band_bytes) + zrow > (1u << 18)) return false;        // 14-bit row offsets in 16 B units
  if (static_cast<size_t>(c->band_bytes) >= (1u << 20)) return false;               // mbarrier tx-count limit
  const size_t stride = (static_cast<size_t>(c->band_bytes) + zrow + 127) & ~static_cast<size_t>(127);
  *band_stride = bufs == 2 ? static_cast<int>(stride) : 0;
  c->smem = (bufs == 2 ? 2 * stride : static_cast<size_t>(c->band_bytes) + zrow) + tab;
  return c->smem <= budget;
}

template <int POOL>
int launch_band_persist(bx_handle* h, const BandArgs& a, const CUtensorMap& tmap, size_t smem, int grid, cudaStream_t st) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;
  roi_plan_kernel<S><<<a.r.b * a.n_bands, kPlanThreads, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  BX_CUDA(cudaFuncSetAttribute(roi_band_persist_kernel<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BX_CUDA(cudaLaunchKernelEx(&cfg, roi_band_persist_kernel<POOL>, tmap, a));
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
  const int rois_range = ra.roi_counts ? ra.rois_per_image : ra.r;
  if (static_cast<unsigned long long>(rois_range) * ra.P * ra.P * (ra.c / 4) >= (1ull << 31)) return BX_OK;
  // two 512-thread CTAs per SM (TMA waits of one overlap the other's compute) when the band fits in half an SM's
  // shared memory, else one 1024-thread CTA per SM
  BandCfg cfg;
  int threads = 512;
  // persistent kernel: opt-in (BX_ROI_BAND_PERSIST=1, read per call).  Measured SLOWER than the per-unit kernel at cfg2 —
  // 0.145 vs 0.133 ms alone, 0.1246 vs 0.1062 ms pipelined: with all 32 warps of the SM on one unit every plan-block and
  // unit barrier waits for the slowest warp, which two independent 16-warp CTAs hide from each other; reserving 4 / 8 SMs for
  // the proposal kernels of other steps gives 1 % back (profiles/README.md).  Kept as a second implementation the tests compare.
  const char* persist_s = getenv("BX_ROI_BAND_PERSIST");
  int band_stride = 0;
  bool persist = (persist_s && atoi(persist_s) != 0) && !getenv("BX_BAND_DEBUG") &&
                 make_persist_cfg(fh, fw, ra.Q, rois_range, h->smem_optin - 1024, &cfg, &band_stride);
  const char* force1 = getenv("BX_ROI_ONE_CTA");
  if (!persist && (force1 || !make_band_cfg(fh, fw, ra.Q, rois_range, (h->smem_sm - 2 * 1024) / 2 - 64, &cfg))) {
    threads = 1024;
    if (!make_band_cfg(fh, fw, ra.Q, rois_range, h->smem_optin - 1024, &cfg)) return BX_OK;
  }

  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ra.c), static_cast<cuuint64_t>(ra.b) * fh * fw};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ra.c) * sizeof(float)};
  const cuuint32_t box[2] = {kSlice, static_cast<cuuint32_t>(cfg.box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ra.lv[0].feat), gdim, gstride,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BX_REQUIRE(cr == CUDA_SUCCESS, BX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);

  const size_t plan_bytes = static_cast<size_t>(ra.b) * cfg.n_bands * cfg.n_blocks_max * plan_layout(cfg.cap, ra.Q).bytes;
  const size_t counter_off = (plan_bytes + 255) & ~static_cast<size_t>(255);
  int rc = bx_plan_reserve(h, counter_off + 256, st);
  if (rc) return rc;

  BandArgs a;
  a.r = ra;
  a.rows_per_band = cfg.rows_per_band;
  a.n_bands = cfg.n_bands;
  a.n_slices = ra.c / kSlice;
  a.cap = cfg.cap;
  a.n_blocks_max = cfg.n_blocks_max;
  a.nbox = cfg.nbox;
  a.box_rows = cfg.box_rows;
  a.band_bytes = cfg.band_bytes;
  a.scan_all = (ra.box_ind != nullptr && !ra.roi_counts) ? 1 : 0;
  a.plan = static_cast<unsigned char*>(h->plan);
  a.neg_zero = -0.0f;
  a.dbg = nullptr;
  a.unit_counter = persist ? reinterpret_cast<int*>(a.plan + counter_off) : nullptr;
  a.n_units = ra.b * a.n_slices * a.n_bands;
  a.band_stride = band_stride;
  if (persist) {
    // grid: one CTA per SM; BX_ROI_BAND_RESERVE leaves SMs to the small kernels of other steps (proposal / plan kernels)
    const char* res_s = getenv("BX_ROI_BAND_RESERVE");
    int grid = h->num_sms - (res_s ? atoi(res_s) : 0);
    if (grid < 1) grid = 1;
    if (grid > a.n_units) grid = a.n_units;
    if (pool == BX_POOL_NONE) rc = launch_band_persist<BX_POOL_NONE>(h, a, tmap, cfg.smem, grid, st);
    else if (pool == BX_POOL_MAX2) rc = launch_band_persist<BX_POOL_MAX2>(h, a, tmap, cfg.smem, grid, st);
    else rc = launch_band_persist<BX_POOL_AVG2>(h, a, tmap, cfg.smem, grid, st);
    if (rc == BX_OK) {
      *used = 1;
      h->band_launches++;
    }
    return rc;
  }
  if (getenv("BX_BAND_DEBUG")) {   // measurement aid: per-CTA timestamps appended to the plan buffer, dumped by the caller
    const size_t grid = static_cast<size_t>(ra.b) * a.n_slices * a.n_bands;
    rc = bx_plan_reserve(h, counter_off + 512 + grid * 32, st);
    if (rc) return rc;
    a.plan = static_cast<unsigned char*>(h->plan);
    a.dbg = reinterpret_cast<unsigned long long*>(a.plan + counter_off + 256);
    h->dbg_ptr = a.dbg;
    h->dbg_count = static_cast<long long>(grid);
    h->dbg_info[0] = a.n_bands; h->dbg_info[1] = a.n_slices; h->dbg_info[2] = threads; h->dbg_info[3] = (int)cfg.smem;
  }
  if (pool == BX_POOL_NONE) rc = launch_band<BX_POOL_NONE>(h, a, tmap, cfg.smem, threads, st);
  else if (pool == BX_POOL_MAX2) rc = launch_band<BX_POOL_MAX2>(h, a, tmap, cfg.smem, threads, st);
  else rc = launch_band<BX_POOL_AVG2>(h, a, tmap, cfg.smem, threads, st);
  if (rc == BX_OK) {
    *used = 1;
    h->band_launches++;
  }
  return rc;
}

}  // namespace bxroi
