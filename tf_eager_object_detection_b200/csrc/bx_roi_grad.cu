// Backward of the RoI extractors w.r.t. the feature map, row-owned form ("next" row f3; the gradient TF sends through
// tf.image.crop_and_resize (+ the 2x2 pool) at scripts/train.py:99-103 for model/roi_pooling.py:36-42,79-90,175).
//
// The scatter form (roi_pool_grad_kernel in bx_roi.cu) sends 4 float4 reductions to L2 per pooled gradient element: 1.9 GB
// of RED traffic for 481 MB of input at the cfg2 shape, and a result whose last bits depend on the order the atomics land
// in.  Here every feature-map row has ONE owner: a CTA owns (image, row, 64-pixel segment, group of channel slices), each
// of its warps owns one slice of 32*VEC channels and keeps that row segment of the gradient in its own shared memory.
// The CTA scans the image's rois (in roi order), keeps those with a crop sample whose upper or lower tap row is this row,
// and every warp adds the samples' shares to its private row with plain shared-memory read-modify-writes: lane = channel,
// so no two lanes touch the same word, and a warp executes in program order, so no atomics and no barriers inside the
// accumulation.  Each row is then written once with plain stores: no memset of grad_feat, no atomics, and the sum order
// of every pixel is fixed (roi, sample row, upper/lower tap, sample column, left/right tap): the result is bit-reproducible
// run to run.  A sample row whose two tap rows differ is read by two CTAs (the second read hits in L2: the two CTAs are
// neighbours in launch order), so HBM traffic stays at grad_out once + grad_feat once.
//
// 2x2 max pool: the arg-max of every pooling window (first maximum in row-major order, TF MaxPoolGrad) is computed once
// per call by roi_argmax_kernel into a byte per (roi, py, px, channel) in the handle's workspace, and the row kernel
// routes the pooled gradient to the sample the byte names.  2x2 avg pool: a quarter to each sample.
#include <stdio.h>
#include <stdlib.h>

#include "bx_roi.cuh"

namespace bxroi {

namespace {

constexpr int kGSeg = 64;    // pixels of a row one CTA owns
constexpr int kGHits = 48;   // rois per table fill (fewer when a roi can have more than 8 entries, see hits_for)
constexpr int kGEntries = 384;   // entries (roi, sample row, chunk of 8 output pixels) per table fill
constexpr int kGMaxQ = 32;   // sample rows are tracked in 32-bit masks

struct GradRowsArgs {
  RoiGradArgs g;
  const unsigned char* code;   // [r,P,P,c] arg-max sample of each pooling window (MAX2), else null
  int n_xseg, n_sg, xw_max;
  int hits;   // rois per table fill: hits * Q * ceil(P / 8) <= kGEntries
};

template <int VEC> struct VecOf;
template <> struct VecOf<1> { using T = float; using C = unsigned char; };
template <> struct VecOf<2> { using T = float2; using C = unsigned short; };

// ---- shared memory through 32-bit shared-window addresses.  All of them `asm volatile`: the accumulation is a sequence
//      of read-modify-writes whose order is the program order, and the compiler keeps volatile asm statements in order.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned lds32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds64(unsigned a) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(unsigned a, unsigned v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts64(unsigned a, unsigned x, unsigned y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y));
}

// VEC fp32 values of one lane, packed in pairs for the fp32x2 pipe (FMUL2 / FFMA2) when VEC > 1
template <int VEC> struct Pk { unsigned long long v[VEC / 2]; };
template <> struct Pk<1> { float v[1]; };

template <int VEC>
__device__ __forceinline__ Pk<VEC> pk_load(unsigned a) {
  Pk<VEC> r;
  if constexpr (VEC == 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r.v[0]) : "r"(a));
  else if constexpr (VEC == 2) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r.v[0]) : "r"(a));
  else asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(r.v[0]), "=l"(r.v[1]) : "r"(a));
  return r;
}
template <int VEC>
__device__ __forceinline__ void pk_store(unsigned a, const Pk<VEC>& r) {
  if constexpr (VEC == 1) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(r.v[0]));
  else if constexpr (VEC == 2) asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(r.v[0]));
  else asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a), "l"(r.v[0]), "l"(r.v[1]));
}
// w * q
template <int VEC>
__device__ __forceinline__ Pk<VEC> pk_scale(float w, const Pk<VEC>& q) {
  Pk<VEC> r;
  if constexpr (VEC == 1) r.v[0] = w * q.v[0];
  else {
    const unsigned long long ww = f2_splat(w);
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v[i]) : "l"(ww), "l"(q.v[i]));
  }
  return r;
}
// cur + w * d with one rounding (the scatter kernel rounds the product first: both are within an ulp of the exact share;
// the tolerance of this op is 1e-5 of max|grad|)
template <int VEC>
__device__ __forceinline__ void pk_fma(float w, const Pk<VEC>& d, Pk<VEC>& cur) {
  if constexpr (VEC == 1) cur.v[0] = __fmaf_rn(w, d.v[0], cur.v[0]);
  else {
    const unsigned long long ww = f2_splat(w);
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cur.v[i]) : "l"(ww), "l"(d.v[i]));
  }
}

// gradient of one crop sample from the pooled gradient `go`: identity, a quarter, or all of it if the sample is the
// window's arg-max (`code` holds one byte per channel, `want` = 2*(sy&1) + (sx&1))
template <int POOL, int VEC>
__device__ __forceinline__ Pk<VEC> route(const typename VecOf<VEC>::T& gov, unsigned code, int want) {
  const float* go = reinterpret_cast<const float*>(&gov);
  float q[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    if (POOL == BX_POOL_NONE) q[v] = go[v];
    else if (POOL == BX_POOL_AVG2) q[v] = go[v] / 4.0f;
    else q[v] = (static_cast<int>((code >> (8 * v)) & 0xffu) == want) ? go[v] : 0.0f;
  }
  Pk<VEC> r;
  if constexpr (VEC == 1) r.v[0] = q[0];
  else {
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(r.v[i]) : "f"(q[2 * i]), "f"(q[2 * i + 1]));
  }
  return r;
}

// One tap (left: kHi = 0, right: kHi = 1) of up to 8 samples of a sample row, added to the warp's row.  kBatch: the 8 taps
// are known to be different pixels (the roi's sample spacing is at least one pixel), so the 8 loads are issued before the
// 8 stores; otherwise read-modify-write one at a time, in the same order.  Taps outside the segment and extrapolated
// samples point at the spare pixel behind the row.
template <int VEC, bool kBatch, int kHi, int N>
__device__ __forceinline__ void add_taps(unsigned acc, const Pk<VEC> (&d)[8], const uint2 (&xe)[8], int n) {
  if (kBatch) {
    Pk<VEC> cur[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < N && k < n) cur[k] = pk_load<VEC>(acc + (kHi ? xe[k].x >> 16 : xe[k].x & 0xffffu));
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < N && k < n) pk_fma<VEC>(kHi ? __uint_as_float(xe[k].y) : 1.0f - __uint_as_float(xe[k].y), d[k], cur[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < N && k < n) pk_store<VEC>(acc + (kHi ? xe[k].x >> 16 : xe[k].x & 0xffffu), cur[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < N && k < n) {
        const unsigned p = acc + (kHi ? xe[k].x >> 16 : xe[k].x & 0xffffu);
        Pk<VEC> cur = pk_load<VEC>(p);
        pk_fma<VEC>(kHi ? __uint_as_float(xe[k].y) : 1.0f - __uint_as_float(xe[k].y), d[k], cur);
        pk_store<VEC>(p, cur);
      }
  }
}

// KP: the pooled size when it is at most 8 (one chunk per sample row, loop bounds known to the compiler), else 0.
// SPLIT = false: the W warps of a CTA own W channel slices.  SPLIT = true (small problems: too few rows to fill the device,
// and the length of a warp's entry chain is the run time): the W warps share ONE slice, take every W-th entry into private
// copies of the row, and the copies are summed in warp order at the end — still one fixed summation order.
template <int POOL, int VEC, int W, int KP, bool SPLIT>
__global__ void __launch_bounds__(W * 32, KP ? (VEC == 2 ? 3 : 2) : 1) roi_grad_rows_kernel(const GradRowsArgs ga) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;
  constexpr int kThreads = W * 32;
  constexpr int kPix = 32 * VEC * 4;   // bytes per pixel of a warp's row
  constexpr int N = KP ? KP : 8;       // samples per chunk
  using V = typename VecOf<VEC>::T;
  using CV = typename VecOf<VEC>::C;
  extern __shared__ __align__(16) unsigned char g_smem[];
  __shared__ int s_wcount[2][W];
  __shared__ int s_nent;
  const RoiArgs& a = ga.g.a;
  const int P = KP ? KP : a.P, Q = P * S;
  const int fh = a.lv[0].fh, fw = a.lv[0].fw;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int bid = blockIdx.x;
  const int sg = bid % ga.n_sg; bid /= ga.n_sg;
  const int xs = bid % ga.n_xseg; bid /= ga.n_xseg;
  const int row = bid % fh, img = bid / fh;
  const int x0 = xs * kGSeg, xw = min(kGSeg, fw - x0);
  const int row_pix = ga.xw_max + 1;   // + the spare pixel
  const unsigned spare = static_cast<unsigned>(ga.xw_max) * kPix;

  const unsigned smem0 = smem_addr(g_smem);
  const unsigned acc = smem0 + warp * row_pix * kPix + lane * VEC * 4;
  const int hits = ga.hits;
  const unsigned t_hdr = smem0 + W * row_pix * kPix;   // [hits] roi, masks of the sample rows hitting this row, batch flag
  const unsigned t_e = t_hdr + hits * 16;   // [kGEntries] slot | sy << 8 | p0 << 16 | lo << 24 | hi << 25 | batch << 26, row index
  const unsigned t_x = t_e + kGEntries * 8;           // [hits][Q] byte offsets of the two taps (16 + 16 bits), lerp
  const unsigned t_ly = t_x + hits * Q * 8;           // [hits][Q]

  for (int x = 0; x < row_pix; ++x) {
    Pk<VEC> z = {};
    pk_store<VEC>(acc + x * kPix, z);
  }
  if (SPLIT) __syncthreads();   // the final sum reads the other warps' copies even if no roi hits this row
  else __syncwarp();

  int cbeg = 0, cend = a.r;
  if (a.roi_counts) {
    cbeg = img * a.rois_per_image;
    cend = cbeg + min(max(a.roi_counts[img], 0), a.rois_per_image);
  }
  const int ch = (SPLIT ? sg : sg * W + warp) * 32 * VEC + lane * VEC;
  const bool ch_ok = ch < a.c;
  const float* go_lane = ga.g.grad_out + (ch_ok ? ch : 0);
  const unsigned char* code_lane = ga.code + (ch_ok ? ch : 0);
  const size_t cstride = a.c;

  int round = 0;
  for (int base = cbeg; base < cend; base += kThreads, ++round) {
    const int j = base + tid;
    unsigned mlo = 0, mhi = 0;
    NormBox nb = {};
    if (j < cend && (a.roi_counts || (a.box_ind ? a.box_ind[j] : 0) == img)) {
      nb = roi_norm_box(a, a.rois[j], fh, fw);
      // tap rows are monotonic in the sample index: the end samples bound them
      const Axis fy = sample_axis(nb.y1, nb.y2, 0, Q, nb.dimy, nb.pad), ly = sample_axis(nb.y1, nb.y2, Q - 1, Q, nb.dimy, nb.pad);
      const bool near = row >= min(min(fy.lo, fy.hi), min(ly.lo, ly.hi)) && row <= max(max(fy.lo, fy.hi), max(ly.lo, ly.hi));
      for (int sy = 0; near && sy < Q; ++sy) {
        const Axis ay = sample_axis(nb.y1, nb.y2, sy, Q, nb.dimy, nb.pad);
        if (ay.valid) {
          mlo |= static_cast<unsigned>(ay.lo == row) << sy;
          mhi |= static_cast<unsigned>(ay.hi == row) << sy;
        }
      }
      if ((mlo | mhi) && ga.n_xseg > 1) {   // taps are monotonic in the sample index: the end samples bound the columns
        const Axis f = sample_axis(nb.x1, nb.x2, 0, Q, nb.dimx, nb.pad), l = sample_axis(nb.x1, nb.x2, Q - 1, Q, nb.dimx, nb.pad);
        const int xmin = min(min(f.lo, f.hi), min(l.lo, l.hi)), xmax = max(max(f.lo, f.hi), max(l.lo, l.hi));
        if (xmax < x0 || xmin >= x0 + xw) mlo = mhi = 0;
      }
    }
    const bool hit = (mlo | mhi) != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_wcount[round & 1][warp] = __popc(bal);
    __syncthreads();
    int rank = __popc(bal & ((1u << lane) - 1)), total = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int n = s_wcount[round & 1][w];
      rank += (w < warp) ? n : 0;
      total += n;
    }
    for (int sub = 0; sub * hits < total; ++sub) {
      if (hit && rank / hits == sub) {
        const int slot = rank % hits;
        // batch = the taps of the samples one add_taps call handles (every S-th sample) are pairwise different pixels:
        // taps are monotonic in the sample index, so it is enough that neighbours differ
        unsigned batch = 1;
        int plo[2] = {-1, -1}, phi[2] = {-1, -1};
        for (int s = 0; s < Q; ++s) {
          sts32(t_ly + (slot * Q + s) * 4, __float_as_uint(sample_axis(nb.y1, nb.y2, s, Q, nb.dimy, nb.pad).lerp));
          const Axis ax = sample_axis(nb.x1, nb.x2, s, Q, nb.dimx, nb.pad);
          unsigned ol = spare, oh = spare;
          float lerp = 0.0f;
          if (ax.valid) {
            lerp = ax.lerp;
            const int xl = ax.lo - x0, xh = ax.hi - x0;
            if (static_cast<unsigned>(xl) < static_cast<unsigned>(xw)) ol = static_cast<unsigned>(xl) * kPix;
            if (static_cast<unsigned>(xh) < static_cast<unsigned>(xw)) oh = static_cast<unsigned>(xh) * kPix;
            if (ax.lo == plo[s % S] || ax.hi == phi[s % S]) batch = 0;
            plo[s % S] = ax.lo;
            phi[s % S] = ax.hi;
          }
          sts64(t_x + (slot * Q + s) * 8, ol | (oh << 16), __float_as_uint(lerp));
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_hdr + slot * 16), "r"(j), "r"(mlo), "r"(mhi), "r"(batch));
      }
      __syncthreads();
      const int n_slots = min(hits, total - sub * hits);
      // flat entry list of this fill, in (slot, sample row, chunk) order: warp 0, one or two slots per lane
      if (warp == 0) {
        int run = 0;
        for (int s0 = 0; s0 < n_slots; s0 += 32) {
          const int slot = s0 + lane;
          uint4 hd = make_uint4(0, 0, 0, 0);
          if (slot < n_slots)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hd.x), "=r"(hd.y), "=r"(hd.z), "=r"(hd.w) : "r"(t_hdr + slot * 16));
          const int chunks = (P + 7) >> 3;
          const int cnt = __popc(hd.y | hd.z) * chunks;
          int inc = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
          }
          int pos = run + inc - cnt;
          for (unsigned m = hd.y | hd.z; m; m &= m - 1) {
            const int sy = __ffs(m) - 1;
            for (int p0 = 0; p0 < P; p0 += 8, ++pos)
              sts64(t_e + pos * 8,
                    slot | (sy << 8) | (p0 << 16) | (((hd.y >> sy) & 1u) << 24) | (((hd.z >> sy) & 1u) << 25) | (hd.w << 26),
                    (hd.x * P + sy / S) * P + p0);
          }
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_nent = run;
      }
      __syncthreads();
      const int n_ent = s_nent;
      // the gradient rows of the next three entries are in flight while one entry is added: a warp walks its entries one
      // after the other (they may touch the same pixels), and few warps fit beside their rows in shared memory, so the
      // DRAM round trip of a row has to overlap the work on the rows before it
      auto issue = [&](int e, V (&g)[8], CV (&c)[8], uint2& er) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          V z = {};
          g[k] = z;
          c[k] = 0;
        }
        if (e < n_ent) er = lds64(t_e + e * 8);   // kept in registers until the entry is processed
        if (e < n_ent && ch_ok) {
          const int n = KP ? KP : min(8, P - static_cast<int>((er.x >> 16) & 0xffu));
          const float* gp = go_lane + static_cast<size_t>(er.y) * cstride;
          const unsigned char* cp = code_lane + static_cast<size_t>(er.y) * cstride;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k < N && k < n) {
              g[k] = __ldg(reinterpret_cast<const V*>(gp));
              if (POOL == BX_POOL_MAX2) c[k] = __ldg(reinterpret_cast<const CV*>(cp));
            }
            gp += cstride;
            cp += cstride;
          }
        }
      };
      // one entry: route the loaded gradients to this sample row's samples (which frees the load buffer: the loads of the
      // entry three ahead are issued into it right there, in front of the shared-memory work they overlap), then add the
      // upper and / or lower tap row's shares
      auto process = [&](uint2& er_buf, V (&gv)[8], CV (&cd)[8], int e_next) {
        const uint2 er = er_buf;
        const int hs = er.x & 0xffu, sy = (er.x >> 8) & 0xffu, p0 = (er.x >> 16) & 0xffu;
        const float ly = __uint_as_float(lds32(t_ly + (hs * Q + sy) * 4));
        const unsigned txa = t_x + (hs * Q + p0 * S) * 8;
        const int n = KP ? KP : min(8, P - p0);
        Pk<VEC> q[S][8];
        uint2 xe[S][8];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < N && k < n) {
              xe[s][k] = lds64(txa + (k * S + s) * 8);
              q[s][k] = route<POOL, VEC>(gv[k], static_cast<unsigned>(cd[k]), 2 * (sy & 1) + s);
            }
        issue(e_next, gv, cd, er_buf);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (!((er.x >> (24 + half)) & 1u)) continue;
          const float wy = half ? ly : 1.0f - ly;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            Pk<VEC> d[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < N && k < n) d[k] = pk_scale<VEC>(wy, q[s][k]);
            if ((er.x >> 26) & 1u) {
              add_taps<VEC, true, 0, N>(acc, d, xe[s], n);
              add_taps<VEC, true, 1, N>(acc, d, xe[s], n);
            } else {
              add_taps<VEC, false, 0, N>(acc, d, xe[s], n);
              add_taps<VEC, false, 1, N>(acc, d, xe[s], n);
            }
          }
        }
      };
      {
        V ga_[8], gb_[8], gc_[8];
        CV ca_[8], cb_[8], cc_[8];
        constexpr int kStep = SPLIT ? W : 1;
        const int e0 = SPLIT ? warp : 0;
        uint2 ea_ = make_uint2(0, 0), eb_ = ea_, ec_ = ea_;
        issue(e0, ga_, ca_, ea_);
        issue(e0 + kStep, gb_, cb_, eb_);
        issue(e0 + 2 * kStep, gc_, cc_, ec_);
        for (int e = e0; e < n_ent; e += 3 * kStep) {
          process(ea_, ga_, ca_, e + 3 * kStep);
          if (e + kStep < n_ent) process(eb_, gb_, cb_, e + 4 * kStep);
          if (e + 2 * kStep < n_ent) process(ec_, gc_, cc_, e + 5 * kStep);
        }
      }
      __syncthreads();
    }
  }
  if (ch_ok) {
    float* dst = ga.g.grad_feat + ((static_cast<size_t>(img) * fh + row) * fw + x0) * cstride + ch;
    if (SPLIT) {   // the loop above ended with a barrier (or never ran): all copies are complete
      const unsigned col = smem0 + lane * VEC * 4;
      for (int x = warp; x < xw; x += W) {
        Pk<VEC> r = pk_load<VEC>(col + x * kPix);
#pragma unroll
        for (int w = 1; w < W; ++w) {
          const Pk<VEC> o = pk_load<VEC>(col + (w * row_pix + x) * kPix);
          if constexpr (VEC == 1) r.v[0] = r.v[0] + o.v[0];
          else {
#pragma unroll
            for (int i = 0; i < VEC / 2; ++i) r.v[i] = f2_add(r.v[i], o.v[i]);
          }
        }
        *reinterpret_cast<V*>(dst + static_cast<size_t>(x) * cstride) = *reinterpret_cast<const V*>(&r);
      }
    } else {
      for (int x = 0; x < xw; ++x) {
        const Pk<VEC> r = pk_load<VEC>(acc + x * kPix);
        *reinterpret_cast<V*>(dst + static_cast<size_t>(x) * cstride) = *reinterpret_cast<const V*>(&r);
      }
    }
  }
}

// arg-max sample of every 2x2 pooling window, per channel: one byte 2*sy + sx, first maximum in row-major order
__global__ void __launch_bounds__(256) roi_argmax_kernel(const RoiArgs a, unsigned char* code) {
  __shared__ Axis ax_y[2];
  __shared__ Axis ax_x[kMaxQ];
  __shared__ int s_meta[2];
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
      s_meta[1] = zero || img < 0 || img >= a.b;
    }
    const NormBox nb = roi_norm_box(a, a.rois[j], a.lv[0].fh, a.lv[0].fw);
    if (tid < Q) ax_x[tid] = sample_axis(nb.x1, nb.x2, tid, Q, nb.dimx, nb.pad);
    else ax_y[tid - Q] = sample_axis(nb.y1, nb.y2, py * 2 + (tid - Q), Q, nb.dimy, nb.pad);
  }
  __syncthreads();
  if (s_meta[1]) return;   // never read: the row kernel skips rois outside the image's count
  const LevelFeat lf = a.lv[0];
  const float4* feat = reinterpret_cast<const float4*>(lf.feat + static_cast<size_t>(s_meta[0]) * lf.fh * lf.fw * a.c);
  uchar4* dst = reinterpret_cast<uchar4*>(code + (static_cast<size_t>(j) * P + py) * P * a.c);
  for (int it = tid; it < P * cv; it += 256) {
    const int px = it / cv, cg = it % cv;
    float best[4];
    unsigned char arg[4];
#pragma unroll
    for (int sy = 0; sy < 2; ++sy)
#pragma unroll
      for (int sx = 0; sx < 2; ++sx) {
        const Axis ay = ax_y[sy], axx = ax_x[px * 2 + sx];
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
          if ((sy == 0 && sx == 0) || val[v] > best[v]) { best[v] = val[v]; arg[v] = static_cast<unsigned char>(sy * 2 + sx); }
      }
    dst[it] = make_uchar4(arg[0], arg[1], arg[2], arg[3]);
  }
}

template <int POOL, int VEC, int W, int KP, bool SPLIT>
int launch_rows(bx_handle* h, const GradRowsArgs& ga, int grid, size_t smem, cudaStream_t st) {
  auto k = roi_grad_rows_kernel<POOL, VEC, W, KP, SPLIT>;
  BX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  k<<<grid, W * 32, smem, st>>>(ga);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

// warps per CTA for VEC channels per lane: the rows of a CTA take W * 65 pixels * 128 * VEC bytes of shared memory
constexpr int warps_for(int vec) { return vec == 1 ? 8 : 4; }

template <int POOL, int VEC>
int launch_rows_p(bx_handle* h, const GradRowsArgs& ga, bool split, int grid, size_t smem, cudaStream_t st) {
  constexpr int W = warps_for(VEC);
  if (ga.g.a.P == 7) return split ? launch_rows<POOL, VEC, W, 7, true>(h, ga, grid, smem, st) : launch_rows<POOL, VEC, W, 7, false>(h, ga, grid, smem, st);
  return split ? launch_rows<POOL, VEC, W, 0, true>(h, ga, grid, smem, st) : launch_rows<POOL, VEC, W, 0, false>(h, ga, grid, smem, st);
}

template <int POOL>
int launch_rows_cfg(bx_handle* h, const GradRowsArgs& ga, int vec, bool split, int grid, size_t smem, cudaStream_t st) {
  if (vec == 1) return launch_rows_p<POOL, 1>(h, ga, split, grid, smem, st);
  return launch_rows_p<POOL, 2>(h, ga, split, grid, smem, st);
}

}  // namespace

int roi_grad_rows_launch(bx_handle* h, const RoiGradArgs& g, int pool, cudaStream_t st, int* used) {
  *used = 0;
  const RoiArgs& a = g.a;
  const int fh = a.lv[0].fh, fw = a.lv[0].fw;
  // Which kernel (all switches are read per call):
  //   BX_ROI_GRAD_ATOMIC=1          the scatter kernel, always
  //   BX_ROI_GRAD_DETERMINISTIC=1   (or bx_set_deterministic) this kernel, always: bit-reproducible gradients
  //   default                       the faster of the two as measured (profiles/README.md): this kernel for plain crops and
  //                                 the 2x2 mean when the rows fill the device; the scatter kernel for the 2x2 max (three
  //                                 of four samples of a window carry no gradient, which the scatter form skips and the
  //                                 row form adds as zeros) and for small problems (a row's entries are a serial chain)
  const char* atomic_s = getenv("BX_ROI_GRAD_ATOMIC");
  const char* det_s = getenv("BX_ROI_GRAD_DETERMINISTIC");
  const bool deterministic = h->deterministic || (det_s && atoi(det_s) != 0);
  if (atomic_s && atoi(atomic_s) != 0) return BX_OK;
  if (a.Q > kGMaxQ || fw > 0x7fff || fh > 0x7fff || a.r > (1 << 24)) return BX_OK;
  // VEC channels per lane (a CTA has warps_for(VEC) warps); BX_ROI_GRAD_VEC overrides
  int vec = 2;
  if (const char* s = getenv("BX_ROI_GRAD_VEC")) {
    const int v = atoi(s);
    if (v == 1 || v == 2) vec = v;
  }
  while (vec > 1 && a.c < 32 * vec) vec >>= 1;   // narrow tensors: do not idle whole lanes
  const int warps = warps_for(vec);
  GradRowsArgs ga = {};
  ga.g = g;
  ga.xw_max = fw < kGSeg ? fw : kGSeg;
  ga.n_xseg = (fw + kGSeg - 1) / kGSeg;
  // few rows: the warps of a CTA share one channel slice and split its entries (BX_ROI_GRAD_SPLIT=0/1 overrides)
  const int slices = (a.c + 32 * vec - 1) / (32 * vec);
  const long long rows = static_cast<long long>(a.b) * fh * ga.n_xseg;
  bool split = rows * ((slices + warps - 1) / warps) < 3LL * h->num_sms;   // less than one CTA per resident slot (3 per SM)
  // default = the faster kernel as measured: the scatter kernel for the 2x2 max and for problems this small
  if (!deterministic && (pool == BX_POOL_MAX2 || split)) return BX_OK;
  if (const char* s = getenv("BX_ROI_GRAD_SPLIT")) split = atoi(s) != 0;
  ga.n_sg = split ? slices : (slices + warps - 1) / warps;
  const int per_roi = a.Q * ((a.P + 7) / 8);   // entries one roi can contribute to a row
  ga.hits = kGEntries / per_roi < kGHits ? kGEntries / per_roi : kGHits;
  if (static_cast<size_t>(ga.xw_max + 1) * 32 * vec * sizeof(float) > 0xffffu) return BX_OK;   // tap offsets are 16-bit
  const size_t smem = static_cast<size_t>(warps) * (ga.xw_max + 1) * 32 * vec * sizeof(float) +
                      static_cast<size_t>(ga.hits) * (a.Q * 12 + 16) + static_cast<size_t>(kGEntries) * 8 + 16;
  if (smem > h->smem_optin) return BX_OK;
  const long long grid = rows * ga.n_sg;
  if (grid > 0x7fffffffLL) return BX_OK;
  if (pool == BX_POOL_MAX2) {
    const size_t bytes = static_cast<size_t>(a.r) * a.P * a.P * a.c;
    if (int rc = bx_ws_reserve(h, bytes, st)) return rc;
    roi_argmax_kernel<<<a.r * a.P, 256, 0, st>>>(a, static_cast<unsigned char*>(h->ws));
    BX_LAUNCH_CHECK(h);
    ga.code = static_cast<const unsigned char*>(h->ws);
  }
  int rc;
  if (pool == BX_POOL_NONE) rc = launch_rows_cfg<BX_POOL_NONE>(h, ga, vec, split, static_cast<int>(grid), smem, st);
  else if (pool == BX_POOL_MAX2) rc = launch_rows_cfg<BX_POOL_MAX2>(h, ga, vec, split, static_cast<int>(grid), smem, st);
  else rc = launch_rows_cfg<BX_POOL_AVG2>(h, ga, vec, split, static_cast<int>(grid), smem, st);
  if (rc) return rc;
  *used = 1;
  return BX_OK;
}

}  // namespace bxroi
