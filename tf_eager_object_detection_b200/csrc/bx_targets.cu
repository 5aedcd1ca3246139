// Training-target stage: pairwise IoU ("+1" convention) and the fused AnchorTarget / ProposalTarget built on it.
// Replaces utils/bbox_tf.py:7-56 (pairwise_iou), model/anchor_target.py:29-125 (AnchorTarget.call, _unmap) and
// model/proposal_target.py:32-124 (ProposalTarget.call).  The [N,M] matrix is materialised only by bx_pairwise_iou
// (its contract); the fused targets recompute rows from gt boxes staged in shared memory.
//
// Sampling: the reference's unseeded tf.random_shuffle / np.random.choice are replaced by an injected priority array
// `perm`: shuffle(idx) := idx sorted ascending by (perm[idx], idx)  (DESIGN.md "Sampling").
#include "bx_common.cuh"

namespace {

constexpr int kMaxGt = 1024;     // gt boxes staged in shared memory
constexpr int kSelBins = 1024;
constexpr int kMaxRois = 2048;   // ProposalTarget: rois per image (== bx_proposals post_nms limit)

// ------------------------------------------------------------------------------------------- pairwise IoU
__global__ void __launch_bounds__(256) pairwise_iou_kernel(const float4* __restrict__ a, int n,
                                                           const float4* __restrict__ b, int m,
                                                           float* __restrict__ out) {
  const long long total = static_cast<long long>(n) * m;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += 256ll * gridDim.x) {
    const int i = static_cast<int>(e / m), j = static_cast<int>(e % m);
    const float4 x = __ldg(a + i), y = __ldg(b + j);
    out[e] = bx_iou_plus1(x, bx_area_plus1(x), y, bx_area_plus1(y));
  }
}

// ------------------------------------------------------------------------------------------- exact k-smallest select
// Block-wide: among elements i in [0,n) with flag(i) true, find the threshold composite T such that exactly k elements
// have comp(i) <= T, comp(i) = (prio[i] << 32) | i  (all distinct).  Adaptive 1024-bin radix refinement.
// All threads of the block must call; `hist` has kSelBins entries, `sc` 8 uint64 scratch words.  Requires 1 <= k <= #flagged.
// `val(i, v)` returns whether element i takes part and, if so, its composite in v.
template <typename ValFn>
__device__ uint64_t block_select_k_smallest_fn(int n, ValFn val, int k, uint32_t* hist, uint64_t* sc) {
  const int tid = threadIdx.x, nt = blockDim.x;
  uint64_t lo = 0ull, hi = 0xFFFFFFFFFFFFFFFFull;
  // tighten the range to [min,max] of the flagged composites
  {
    uint64_t mn = 0xFFFFFFFFFFFFFFFFull, mx = 0ull;
    for (int i = tid; i < n; i += nt) {
      uint64_t v;
      if (val(i, v)) {
        mn = min(mn, v);
        mx = max(mx, v);
      }
    }
    if (tid == 0) { sc[0] = 0xFFFFFFFFFFFFFFFFull; sc[1] = 0ull; }
    __syncthreads();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {                    // warp reduce first: 64-bit shared atomics are CAS loops
      mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
      mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
    }
    if ((tid & 31) == 0) {
      atomicMin(reinterpret_cast<unsigned long long*>(&sc[0]), static_cast<unsigned long long>(mn));
      atomicMax(reinterpret_cast<unsigned long long*>(&sc[1]), static_cast<unsigned long long>(mx));
    }
    __syncthreads();
    lo = sc[0];
    hi = sc[1];
    __syncthreads();
  }
  int need = k;
  for (;;) {
    const uint64_t span = hi - lo;
    int shift = 0;
    while ((span >> shift) >= static_cast<uint64_t>(kSelBins)) ++shift;
    for (int i = tid; i < kSelBins; i += nt) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      uint64_t v;
      if (val(i, v) && v >= lo && v <= hi) atomicAdd(&hist[static_cast<uint32_t>((v - lo) >> shift)], 1u);
    }
    __syncthreads();
    {
      // first bin b whose count does not fit entirely (run + c > need): block-wide exclusive scan, one bin per thread
      // (blockDim.x == kSelBins)
      __shared__ uint32_t wtot[32];
      const int lane = tid & 31, warp = tid >> 5;
      const uint32_t c = hist[tid];
      uint32_t pre = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, pre, d);
        if (lane >= d) pre += o;
      }
      if (lane == 31) wtot[warp] = pre;
      __syncthreads();
      if (warp == 0) {
        const uint32_t t = wtot[lane];
        uint32_t p2 = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, p2, d);
          if (lane >= d) p2 += o;
        }
        wtot[lane] = p2 - t;                                        // exclusive warp offsets
      }
      __syncthreads();
      const uint32_t before = wtot[warp] + pre - c;                 // count in bins < tid
      const bool crossing = (before <= static_cast<uint32_t>(need)) && (before + c > static_cast<uint32_t>(need));
      const bool none = (tid == kSelBins - 1) && (before + c <= static_cast<uint32_t>(need));
      if (crossing || none) {
        const int b = crossing ? tid : kSelBins;
        const int run = static_cast<int>(crossing ? before : before + c);
        // b: first bin that does not fit entirely (b == kSelBins: everything fits, then run == need by the precondition)
        if (run == need) {
          sc[2] = 1ull;                                              // done
          sc[3] = (b == 0) ? (lo - 1ull) : (lo + (static_cast<uint64_t>(b) << shift) - 1ull);  // T = last value below bin b
          if (b == kSelBins) sc[3] = hi;
        } else {
          sc[2] = 0ull;
          sc[4] = lo + (static_cast<uint64_t>(b) << shift);
          uint64_t nhi = sc[4] + ((1ull << shift) - 1ull);
          sc[5] = nhi > hi ? hi : nhi;
          sc[6] = static_cast<uint64_t>(need - run);
        }
      }
    }
    __syncthreads();
    const bool done = sc[2] != 0ull;
    const uint64_t T = sc[3];
    if (!done) {
      lo = sc[4];
      hi = sc[5];
      need = static_cast<int>(sc[6]);
    }
    __syncthreads();
    if (done) return T;
  }
}

template <typename FlagFn>
__device__ uint64_t block_select_k_smallest(int n, const int* __restrict__ prio, FlagFn flag, int k, uint32_t* hist,
                                            uint64_t* sc) {
  return block_select_k_smallest_fn(n, [&](int i, uint64_t& v) {
    if (!flag(i)) return false;
    v = (static_cast<uint64_t>(static_cast<uint32_t>(prio[i])) << 32) | static_cast<uint32_t>(i);
    return true;
  }, k, hist, sc);
}

// ------------------------------------------------------------------------------------------- anchor target
struct ATArgs {
  const float4* anchors;  // [n]
  const float4* gt;       // [batch,max_gt]
  const int* gt_counts;   // [batch] or null
  const int* perm;        // [batch,n]
  int n, max_gt;
  bx_anchor_target_params p;
  BoxCodec codec;
  // workspace
  float* row_max;         // [batch,n]
  int* row_arg;           // [batch,n]
  int* col_max;           // [batch,max_gt] (fp32 bits, iou >= 0 so int order == float order)
  int* warp_max;          // [batch, ceil(n/256)*8, max_gt] per-warp maxima of pass A (fp32 bits)
  int* inside_idx;        // [n] anchors inside the image (any order), shared by the batch
  int* n_inside;          // [1]
  int* label;             // [batch,n] pre-sampling then final label (-1/0/1), -2 = outside image
  int* counts;            // [batch,4] = #fg, #bg before sampling; after sampling: fg_final, bg_final
  // outputs
  float* out_labels;
  float4* out_targets;
  float4* out_in_w;
  float4* out_out_w;
  int* out_counts;        // [batch,2]
};

__device__ __forceinline__ bool anchor_inside(const float4 a, float max_x, float max_y) {
  return (a.x >= 0.0f) && (a.y >= 0.0f) && (a.z <= max_x) && (a.w <= max_y);  // utils/bbox_tf.py:95-100
}

// pass 0: labels <- -2 ("outside the image", anchor_target.py:48-49 keeps only the inside anchors) and the list of inside
// anchors, which is all passes A and B ever touch (8151 of 21 546 at cfg2/cfg4)
__global__ void __launch_bounds__(256) at_prepare_kernel(const ATArgs a) {
  const int img = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int i = blockIdx.x * 256 + tid;
  if (i < a.n) a.label[static_cast<size_t>(img) * a.n + i] = -2;
  if (img != 0) return;
  const bool inside = (i < a.n) && anchor_inside(a.anchors[i], static_cast<float>(a.p.image_w - 1), static_cast<float>(a.p.image_h - 1));
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, inside);
  int base = 0;
  if (lane == 0 && m) base = atomicAdd(a.n_inside, __popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  if (inside) a.inside_idx[base + __popc(m & ((1u << lane) - 1u))] = i;
}

// IoU of one (anchor, gt) pair exactly as bx_iou_plus1, with the division skipped for disjoint pairs (most of them)
__device__ __forceinline__ float at_iou(const float4 a, const float area_a, const float4 b, const float area_b) {
  const float ih = fmaxf(0.0f, fminf(a.w, b.w) - fmaxf(a.y, b.y) + 1.0f);
  const float iw = fmaxf(0.0f, fminf(a.z, b.z) - fmaxf(a.x, b.x) + 1.0f);
  const float inter = ih * iw;
  float v = 0.0f;
  if (inter != 0.0f) v = inter / (area_a + area_b - inter);
  return v;
}

// pass A: row max / argmax per inside anchor; per (warp, gt) maxima kept for pass B; column max per gt.
// dynamic smem: gt [m] float4 | area [m] float | warp maxima [8][m] int
__global__ void __launch_bounds__(256) at_rowstats_kernel(const ATArgs a) {
  extern __shared__ __align__(16) unsigned char at_smem[];
  const int img = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.gt_counts ? a.gt_counts[img] : a.max_gt;
  float4* s_gt = reinterpret_cast<float4*>(at_smem);
  float* s_area = reinterpret_cast<float*>(s_gt + a.max_gt);
  int* s_w = reinterpret_cast<int*>(s_area + a.max_gt);          // [8][max_gt]
  for (int j = tid; j < m; j += 256) {
    const float4 g = a.gt[static_cast<size_t>(img) * a.max_gt + j];
    s_gt[j] = g;
    s_area[j] = bx_area_plus1(g);
  }
  const int ni = *a.n_inside;
  if (static_cast<int>(blockIdx.x) * 256 >= ni) return;          // the grid covers n; only the inside anchors do work
  __syncthreads();
  const int ci = blockIdx.x * 256 + tid;
  const bool live = ci < ni;
  const int i = live ? a.inside_idx[ci] : 0;
  const float4 anc = live ? a.anchors[i] : make_float4(0, 0, 0, 0);
  const bool inside = live;
  const float area = bx_area_plus1(anc);
  float best = -1.0f;
  int arg = 0;
  const bool any_inside = __any_sync(0xFFFFFFFFu, inside);
  for (int j = 0; j < m; ++j) {
    int wmax = 0;
    if (any_inside) {
      const float v = inside ? at_iou(anc, area, s_gt[j], s_area[j]) : 0.0f;
      if (inside && v > best) { best = v; arg = j; }        // first max, like tf.argmax
      wmax = __reduce_max_sync(0xFFFFFFFFu, __float_as_int(v));  // int order == float order for v >= 0
    }
    if (lane == 0) s_w[warp * a.max_gt + j] = wmax;
  }
  if (live) {
    const size_t o = static_cast<size_t>(img) * a.n + i;
    a.row_max[o] = inside ? best : 0.0f;
    a.row_arg[o] = arg;
  }
  __syncthreads();
  int* g_w = a.warp_max + (static_cast<size_t>(img) * gridDim.x + blockIdx.x) * 8 * a.max_gt;
  for (int e = tid; e < 8 * m; e += 256) {
    const int w = e / m, j = e - w * m;
    g_w[w * a.max_gt + j] = s_w[w * a.max_gt + j];
  }
  for (int j = tid; j < m; j += 256) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) c = max(c, s_w[w * a.max_gt + j]);
    if (c) atomicMax(&a.col_max[static_cast<size_t>(img) * a.max_gt + j], c);
  }
}

// pass B: labels before subsampling (anchor_target.py:59-69) + fg/bg counts.  "anchor i attains the column maximum of
// gt j" (:64) is only possible in a warp whose own maximum for j equals the column maximum, so only those (warp, gt)
// pairs — about one per gt — recompute their IoUs.
__global__ void __launch_bounds__(256) at_label_kernel(const ATArgs a) {
  const int img = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.gt_counts ? a.gt_counts[img] : a.max_gt;
  const int ni = *a.n_inside;
  if (static_cast<int>(blockIdx.x) * 256 >= ni) return;
  const int ci = blockIdx.x * 256 + tid;
  const bool live = ci < ni;
  const int i = live ? a.inside_idx[ci] : 0;
  const float4 anc = live ? a.anchors[i] : make_float4(0, 0, 0, 0);
  const bool inside = live;
  const float area = bx_area_plus1(anc);
  const int* g_w = a.warp_max + ((static_cast<size_t>(img) * gridDim.x + blockIdx.x) * 8 + warp) * a.max_gt;
  const int* colmax = a.col_max + static_cast<size_t>(img) * a.max_gt;
  const float4* gt = a.gt + static_cast<size_t>(img) * a.max_gt;
  bool is_gt_arg = false;
  if (__any_sync(0xFFFFFFFFu, inside)) {
    for (int j0 = 0; j0 < m; j0 += 32) {
      const int j = j0 + lane;
      uint32_t hit = __ballot_sync(0xFFFFFFFFu, j < m && g_w[j] == colmax[j]);
      while (hit) {
        const int jj = j0 + __ffs(hit) - 1;
        hit &= hit - 1u;
        const float4 g = gt[jj];
        const float v = inside ? at_iou(anc, area, g, bx_area_plus1(g)) : -1.0f;
        is_gt_arg |= (__float_as_int(v) == colmax[jj]);                                       // :64
      }
    }
  }
  int lab = -2;
  if (live) {
    if (inside) {
      const float mx = a.row_max[static_cast<size_t>(img) * a.n + i];
      lab = -1;
      if (m > 0) {
        if (mx < a.p.neg_iou_threshold) lab = 0;      // :67
        if (is_gt_arg) lab = 1;                       // :68
        if (mx >= a.p.pos_iou_threshold) lab = 1;     // :69
      } else if (0.0f < a.p.neg_iou_threshold) {
        lab = 0;   // image without ground truth (the reference's argmax over an empty axis raises): max overlap 0 -> background
      }
    }
    a.label[static_cast<size_t>(img) * a.n + i] = lab;
  }
  const uint32_t fg = __ballot_sync(0xFFFFFFFFu, lab == 1), bg = __ballot_sync(0xFFFFFFFFu, lab == 0);
  if (lane == 0) {
    if (fg) atomicAdd(&a.counts[img * 4 + 0], __popc(fg));
    if (bg) atomicAdd(&a.counts[img * 4 + 1], __popc(bg));
  }
}

// pass C: subsample fg then bg by priority (anchor_target.py:72-84); one CTA per image.  kCompact: the (priority, index)
// composites of the fg and bg anchors are compacted into shared memory in one pass over the labels and every selection
// pass runs over those short lists ((nfg + nbg) * 8 bytes must fit); otherwise the passes stream labels from global.
template <bool kCompact>
__global__ void __launch_bounds__(1024) at_sample_kernel(const ATArgs a, const int cap) {
  extern __shared__ __align__(16) unsigned char at_smem[];
  __shared__ uint32_t hist[kSelBins];
  __shared__ uint64_t sc[8];
  __shared__ int s_cnt[2];
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  int* label = a.label + static_cast<size_t>(img) * a.n;
  const int* perm = a.perm + static_cast<size_t>(img) * a.n;
  const int nfg = a.counts[img * 4 + 0], nbg = a.counts[img * 4 + 1];
  const bool cut_fg = nfg > a.p.max_pos_samples;
  const int fg_final = cut_fg ? a.p.max_pos_samples : nfg;
  const int num_bg = a.p.total_num_samples - fg_final;   // :78
  const bool cut_bg = nbg > num_bg;
  const int bg_final = cut_bg ? max(num_bg, 0) : nbg;
  if (tid == 0) {
    a.counts[img * 4 + 2] = fg_final;
    a.counts[img * 4 + 3] = bg_final;
    if (a.out_counts) {
      a.out_counts[img * 2 + 0] = fg_final;
      a.out_counts[img * 2 + 1] = bg_final;
    }
  }
  if (!cut_fg && !cut_bg) return;
  if (kCompact && nfg + nbg <= cap) {
    uint64_t* fgc = reinterpret_cast<uint64_t*>(at_smem);
    uint64_t* bgc = fgc + nfg;
    if (tid < 2) s_cnt[tid] = 0;
    __syncthreads();
    constexpr int kU = 8;                                // independent label / priority loads in flight per thread
    for (int i0 = 0; i0 < a.n; i0 += 1024 * kU) {
      int labs[kU], prs[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * 1024 + tid;
        labs[u] = (i < a.n) ? label[i] : -2;
        prs[u] = (i < a.n) ? perm[i] : 0;
      }
      uint32_t mf[kU], mb[kU];
      int tf = 0, tb = 0;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        mf[u] = __ballot_sync(0xFFFFFFFFu, labs[u] == 1);
        mb[u] = __ballot_sync(0xFFFFFFFFu, labs[u] == 0);
        tf += __popc(mf[u]);
        tb += __popc(mb[u]);
      }
      int pf = 0, pb = 0;                                 // one reservation per warp and batch
      if (lane == 0) {
        if (tf) pf = atomicAdd(&s_cnt[0], tf);
        if (tb) pb = atomicAdd(&s_cnt[1], tb);
      }
      pf = __shfl_sync(0xFFFFFFFFu, pf, 0);
      pb = __shfl_sync(0xFFFFFFFFu, pb, 0);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * 1024 + tid;
        const uint64_t v = (static_cast<uint64_t>(static_cast<uint32_t>(prs[u])) << 32) | static_cast<uint32_t>(i);
        if (labs[u] == 1) fgc[pf + __popc(mf[u] & ((1u << lane) - 1u))] = v;
        if (labs[u] == 0) bgc[pb + __popc(mb[u] & ((1u << lane) - 1u))] = v;
        pf += __popc(mf[u]);
        pb += __popc(mb[u]);
      }
    }
    __syncthreads();
    if (cut_fg) {
      uint64_t T = 0ull;
      if (fg_final > 0) T = block_select_k_smallest_fn(nfg, [&](int i, uint64_t& v) { v = fgc[i]; return true; }, fg_final, hist, sc);
      for (int i = tid; i < nfg; i += 1024) {
        const uint64_t v = fgc[i];
        if (fg_final == 0 || v > T) label[static_cast<uint32_t>(v)] = -1;
      }
    }
    if (cut_bg) {
      uint64_t T = 0ull;
      if (bg_final > 0) T = block_select_k_smallest_fn(nbg, [&](int i, uint64_t& v) { v = bgc[i]; return true; }, bg_final, hist, sc);
      for (int i = tid; i < nbg; i += 1024) {
        const uint64_t v = bgc[i];
        if (bg_final == 0 || v > T) label[static_cast<uint32_t>(v)] = -1;
      }
    }
    return;
  }
  if (cut_fg) {
    uint64_t T = 0ull;
    if (fg_final > 0) T = block_select_k_smallest(a.n, perm, [&](int i) { return label[i] == 1; }, fg_final, hist, sc);
    for (int i = tid; i < a.n; i += 1024)
      if (label[i] == 1) {
        const uint64_t v = (static_cast<uint64_t>(static_cast<uint32_t>(perm[i])) << 32) | static_cast<uint32_t>(i);
        if (fg_final == 0 || v > T) label[i] = -1;
      }
    __syncthreads();
  }
  if (cut_bg) {
    uint64_t T = 0ull;
    if (bg_final > 0) T = block_select_k_smallest(a.n, perm, [&](int i) { return label[i] == 0; }, bg_final, hist, sc);
    for (int i = tid; i < a.n; i += 1024)
      if (label[i] == 0) {
        const uint64_t v = (static_cast<uint64_t>(static_cast<uint32_t>(perm[i])) << 32) | static_cast<uint32_t>(i);
        if (bg_final == 0 || v > T) label[i] = -1;
      }
    __syncthreads();
  }
}

// pass D: targets / weights / unmap (anchor_target.py:88-125)
__global__ void __launch_bounds__(256) at_finalize_kernel(const ATArgs a) {
  const int img = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n) return;
  const size_t o = static_cast<size_t>(img) * a.n + i;
  const int lab = a.label[o];
  float4 tg = make_float4(0, 0, 0, 0), iw = tg, ow = tg;
  float fl = -1.0f;
  if (lab != -2) {
    fl = static_cast<float>(lab);
    const int m = a.gt_counts ? a.gt_counts[img] : a.max_gt;
    if (m > 0) tg = bx_encode_one(a.anchors[i], a.gt[static_cast<size_t>(img) * a.max_gt + a.row_arg[o]], a.codec);
    if (lab == 1) iw = make_float4(1.f, 1.f, 1.f, 1.f);
    if (lab >= 0) {
      const float num_examples = static_cast<float>(a.counts[img * 4 + 2] + a.counts[img * 4 + 3]);  // :99
      const float w = 1.0f / num_examples;
      ow = make_float4(w, w, w, w);
    }
  }
  a.out_labels[o] = fl;
  a.out_targets[o] = tg;
  a.out_in_w[o] = iw;
  a.out_out_w[o] = ow;
}

// ------------------------------------------------------------------------------------------- proposal target
struct PTArgs {
  const float4* rois;      // [batch,k]
  const int* roi_counts;   // [batch] or null
  const float4* gt;        // [batch,max_gt]
  const int* gt_labels;    // [batch,max_gt]
  const int* gt_counts;    // [batch] or null
  const int* perm;         // [batch,k]
  int k, max_gt;
  bx_proposal_target_params p;
  BoxCodec codec;
  float4* out_rois;        // [batch,S]
  int* out_labels;         // [batch,S]
  float* out_targets;      // [batch,S,4C]
  float* out_in_w;
  float* out_out_w;
  int* out_keep;           // [batch,S]
  int* out_counts;         // [batch,2]
  float* ws_best;          // [batch,k] row max of the roi x gt IoU matrix (pt_rowstats_kernel)
  int* ws_arg;             // [batch,k] its first argmax
};

// roi x gt IoU rows over the whole device (proposal_target.py:56-58): row max and first argmax per roi
__global__ void __launch_bounds__(256) pt_rowstats_kernel(const PTArgs a) {
  __shared__ float4 s_gt[kMaxGt];
  __shared__ float s_area[kMaxGt];
  const int img = blockIdx.y, tid = threadIdx.x;
  const int m = a.gt_counts ? a.gt_counts[img] : a.max_gt;
  const int k = a.roi_counts ? min(a.roi_counts[img], a.k) : a.k;
  for (int j = tid; j < m; j += 256) {
    const float4 g = a.gt[static_cast<size_t>(img) * a.max_gt + j];
    s_gt[j] = g;
    s_area[j] = bx_area_plus1(g);
  }
  __syncthreads();
  const int i = blockIdx.x * 256 + tid;
  if (i >= k) return;
  const float4 r = a.rois[static_cast<size_t>(img) * a.k + i];
  const float ar = bx_area_plus1(r);
  float best = -1.0f;
  int arg = 0;
  for (int j = 0; j < m; ++j) {
    const float v = at_iou(r, ar, s_gt[j], s_area[j]);
    if (v > best) { best = v; arg = j; }
  }
  a.ws_best[static_cast<size_t>(img) * a.k + i] = best;
  a.ws_arg[static_cast<size_t>(img) * a.k + i] = arg;
}

__global__ void __launch_bounds__(1024) proposal_target_kernel(const PTArgs a) {
  extern __shared__ __align__(16) unsigned char pt_smem[];
  float4* s_gt = reinterpret_cast<float4*>(pt_smem);                       // kMaxGt
  uint64_t* s_fg = reinterpret_cast<uint64_t*>(s_gt + kMaxGt);             // kMaxRois
  uint64_t* s_bg = s_fg + kMaxRois;                                        // kMaxRois
  float* s_area = reinterpret_cast<float*>(s_bg + kMaxRois);               // kMaxGt
  unsigned short* s_assign = reinterpret_cast<unsigned short*>(s_area + kMaxGt);  // kMaxRois
  __shared__ int s_cnt[2];
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int m = a.gt_counts ? a.gt_counts[img] : a.max_gt;
  const int k = a.roi_counts ? min(a.roi_counts[img], a.k) : a.k;
  const float4* rois = a.rois + static_cast<size_t>(img) * a.k;
  const float4* gt = a.gt + static_cast<size_t>(img) * a.max_gt;
  const int* gt_labels = a.gt_labels + static_cast<size_t>(img) * a.max_gt;
  const int* perm = a.perm + static_cast<size_t>(img) * a.k;
  const int S = a.p.total_num_samples, C = a.p.num_classes;

  for (int j = tid; j < m; j += 1024) {
    s_gt[j] = gt[j];
    s_area[j] = bx_area_plus1(gt[j]);
  }
  if (tid < 2) s_cnt[tid] = 0;
  __syncthreads();

  // (1) max / argmax over gt, fg / bg membership (proposal_target.py:56-64)
  int pow2 = 64;
  while (pow2 < k) pow2 <<= 1;
  for (int i0 = 0; i0 < pow2; i0 += 1024) {
    const int i = i0 + tid;
    bool fg = false, bg = false;
    if (i < k) {
      const float best = a.ws_best[static_cast<size_t>(img) * a.k + i];
      s_assign[i] = static_cast<unsigned short>(a.ws_arg[static_cast<size_t>(img) * a.k + i]);
      if (m > 0) {
        fg = best >= a.p.pos_iou_threshold;
        bg = (best < a.p.pos_iou_threshold) && (best >= a.p.neg_iou_threshold);
      } else {
        bg = (0.0f < a.p.pos_iou_threshold) && (0.0f >= a.p.neg_iou_threshold);   // no ground truth: max overlap 0
      }
    }
    if (i < pow2) {
      // provisional keys by index; re-keyed by priority below when a set has to be shuffled
      s_fg[i] = fg ? static_cast<uint64_t>(i) : 0xFFFFFFFFFFFFFFFFull;
      s_bg[i] = bg ? static_cast<uint64_t>(i) : 0xFFFFFFFFFFFFFFFFull;
    }
    const uint32_t mf = __ballot_sync(0xFFFFFFFFu, fg), mb = __ballot_sync(0xFFFFFFFFu, bg);
    if (lane == 0) {
      if (mf) atomicAdd(&s_cnt[0], __popc(mf));
      if (mb) atomicAdd(&s_cnt[1], __popc(mb));
    }
  }
  __syncthreads();
  const int nfg_all = s_cnt[0], nbg_all = s_cnt[1];
  const bool shuffle_fg = nfg_all > a.p.max_pos_samples;                      // :67
  const int nfg = shuffle_fg ? a.p.max_pos_samples : nfg_all;
  const int want = S - nfg;
  const bool shuffle_bg = (nbg_all != want);                                  // :69-77 (> : subsample, < : padded cycle)
  // (2) order the two sets: ascending index (tf.where order) or ascending (perm, index) when shuffled
  for (int i = tid; i < pow2; i += 1024) {
    if (shuffle_fg && s_fg[i] != 0xFFFFFFFFFFFFFFFFull)
      s_fg[i] = (static_cast<uint64_t>(static_cast<uint32_t>(perm[i])) << 32) | static_cast<uint32_t>(i);
    if (shuffle_bg && s_bg[i] != 0xFFFFFFFFFFFFFFFFull)
      s_bg[i] = (static_cast<uint64_t>(static_cast<uint32_t>(perm[i])) << 32) | static_cast<uint32_t>(i);
  }
  __syncthreads();
  bx_bitonic_sort<false>(reinterpret_cast<unsigned long long*>(s_fg), pow2);
  bx_bitonic_sort<false>(reinterpret_cast<unsigned long long*>(s_bg), pow2);

  const int status = (want > 0 && nbg_all == 0) ? 1 : 0;   // np.random.choice on an empty set raises (:77)
  if (tid == 0) {
    a.out_counts[img * 2 + 0] = nfg;
    a.out_counts[img * 2 + 1] = status;
  }
  // (3) outputs
  float4* out_rois = a.out_rois + static_cast<size_t>(img) * S;
  int* out_labels = a.out_labels + static_cast<size_t>(img) * S;
  int* out_keep = a.out_keep + static_cast<size_t>(img) * S;
  const size_t row = static_cast<size_t>(4) * C;
  float* out_t = a.out_targets + static_cast<size_t>(img) * S * row;
  float* out_i = a.out_in_w + static_cast<size_t>(img) * S * row;
  float* out_o = a.out_out_w + static_cast<size_t>(img) * S * row;
  for (size_t e = tid; e < static_cast<size_t>(S) * row; e += 1024) {
    out_t[e] = 0.0f;
    out_i[e] = 0.0f;
    out_o[e] = 1.0f;                                                          // :122
  }
  __syncthreads();
  for (int s = tid; s < S; s += 1024) {
    int src = -1;
    if (s < nfg) src = static_cast<int>(s_fg[s] & 0xFFFFFFFFull);
    else if (nbg_all > 0) src = static_cast<int>(s_bg[(s - nfg) % nbg_all] & 0xFFFFFFFFull);
    out_keep[s] = src;
    if (src < 0) {
      out_rois[s] = make_float4(0, 0, 0, 0);
      out_labels[s] = 0;
      continue;
    }
    const float4 r = rois[src];
    out_rois[s] = r;                                                          // :82
    out_labels[s] = (s < nfg) ? gt_labels[s_assign[src]] : 0;                 // :83-86
    if (s < nfg) {
      const int cls = gt_labels[s_assign[s]];                                 // :99,117 `labels[idx]` quirk: roi #s, not roi fg[s]
      if (cls >= 0 && cls < C) {
        const float4 t = bx_encode_one(r, s_gt[s_assign[src]], a.codec);      // :105-108
        float* pt = out_t + static_cast<size_t>(s) * row + 4 * cls;
        float* pi = out_i + static_cast<size_t>(s) * row + 4 * cls;
        pt[0] = t.x; pt[1] = t.y; pt[2] = t.z; pt[3] = t.w;
        pi[0] = 1.f; pi[1] = 1.f; pi[2] = 1.f; pi[3] = 1.f;
      }
    }
  }
}

BoxCodec codec_of(const float means[4], const float stds[4]) {
  BoxCodec k = {};
  k.m0 = means[0]; k.m1 = means[1]; k.m2 = means[2]; k.m3 = means[3];
  k.s0 = stds[0]; k.s1 = stds[1]; k.s2 = stds[2]; k.s3 = stds[3];
  return k;
}

}  // namespace

extern "C" int bx_pairwise_iou(bx_handle* h, const float* a, int n, const float* b, int m, float* out, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && out && (a || n == 0) && (b || m == 0), BX_ERR_INVALID, "bx_pairwise_iou: NULL argument");
  BX_REQUIRE(n >= 0 && m >= 0, BX_ERR_INVALID, "bx_pairwise_iou: negative size");
  BX_REQUIRE(bx_aligned(a, 16) && bx_aligned(b, 16), BX_ERR_INVALID, "bx_pairwise_iou: boxes must be 16-byte aligned");
  const long long total = static_cast<long long>(n) * m;
  if (total == 0) return BX_OK;
  const int grid = static_cast<int>(bx_min_ll(bx_div_up(total, 256), 16ll * h->num_sms));
  pairwise_iou_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), n, reinterpret_cast<const float4*>(b), m, out);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

extern "C" int bx_anchor_target(bx_handle* h, const float* anchors, int n, const float* gt, const int* gt_counts,
                                int batch, int max_gt, const int* perm, const bx_anchor_target_params* p,
                                float* out_labels, float* out_targets, float* out_in_w, float* out_out_w,
                                int* out_counts, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && anchors && (gt || max_gt == 0) && perm && p && out_labels && out_targets && out_in_w && out_out_w,
             BX_ERR_INVALID, "bx_anchor_target: NULL argument");   // an image set without ground truth (max_gt == 0) is valid
  BX_REQUIRE(n > 0 && batch >= 0 && max_gt >= 0, BX_ERR_INVALID, "bx_anchor_target: bad size");
  BX_REQUIRE(max_gt <= kMaxGt, BX_ERR_UNSUPPORTED, "bx_anchor_target: max_gt %d > %d", max_gt, kMaxGt);
  BX_REQUIRE(p->max_pos_samples >= 0 && p->total_num_samples >= p->max_pos_samples, BX_ERR_INVALID,
             "bx_anchor_target: need 0 <= max_pos_samples <= total_num_samples");
  BX_REQUIRE(bx_aligned(anchors, 16) && bx_aligned(gt, 16) && bx_aligned(out_targets, 16) && bx_aligned(out_in_w, 16) &&
                 bx_aligned(out_out_w, 16), BX_ERR_INVALID, "bx_anchor_target: box tensors must be 16-byte aligned");
  if (batch == 0) return BX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bn = static_cast<size_t>(batch) * n;
  const int nblk = static_cast<int>(bx_div_up(n, 256));
  const size_t wm = static_cast<size_t>(batch) * nblk * 8 * (max_gt > 0 ? max_gt : 1);
  const size_t ws = bn * (sizeof(float) + 2 * sizeof(int)) + static_cast<size_t>(batch) * (max_gt + 4) * sizeof(int) +
                    4 * sizeof(int) + wm * sizeof(int) + static_cast<size_t>(n) * sizeof(int);
  int rc = bx_ws_reserve(h, ws, st);
  if (rc) return rc;
  ATArgs a = {};
  a.anchors = reinterpret_cast<const float4*>(anchors);
  a.gt = reinterpret_cast<const float4*>(gt);
  a.gt_counts = gt_counts;
  a.perm = perm;
  a.n = n;
  a.max_gt = max_gt;
  a.p = *p;
  a.codec = codec_of(p->means, p->stds);
  a.row_max = reinterpret_cast<float*>(h->ws);
  a.row_arg = reinterpret_cast<int*>(a.row_max + bn);
  a.label = a.row_arg + bn;
  a.col_max = a.label + bn;
  a.counts = a.col_max + static_cast<size_t>(batch) * max_gt;
  a.n_inside = a.counts + static_cast<size_t>(batch) * 4;
  a.warp_max = a.n_inside + 4;
  a.inside_idx = a.warp_max + wm;
  a.out_labels = out_labels;
  a.out_targets = reinterpret_cast<float4*>(out_targets);
  a.out_in_w = reinterpret_cast<float4*>(out_in_w);
  a.out_out_w = reinterpret_cast<float4*>(out_out_w);
  a.out_counts = out_counts;
  BX_CUDA(cudaMemsetAsync(a.col_max, 0, (static_cast<size_t>(batch) * (max_gt + 4) + 4) * sizeof(int), st));
  const dim3 grid(nblk, batch);
  at_prepare_kernel<<<grid, 256, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  const size_t smem_a = static_cast<size_t>(max_gt) * (sizeof(float4) + sizeof(float) + 8 * sizeof(int));
  if (smem_a > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(at_rowstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  at_rowstats_kernel<<<grid, 256, smem_a, st>>>(a);
  BX_LAUNCH_CHECK(h);
  at_label_kernel<<<grid, 256, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  // compaction lists of pass C: as many (priority, index) composites as shared memory holds, at most one per anchor
  const size_t room = h->smem_optin > 16 * 1024 ? h->smem_optin - 16 * 1024 : 0;
  int cap = static_cast<int>(room / sizeof(uint64_t));
  if (cap > n) cap = n;
  const size_t smem_c = static_cast<size_t>(cap) * sizeof(uint64_t);
  BX_CUDA(cudaFuncSetAttribute(at_sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  at_sample_kernel<true><<<batch, 1024, smem_c, st>>>(a, cap);
  BX_LAUNCH_CHECK(h);
  at_finalize_kernel<<<grid, 256, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

extern "C" int bx_proposal_target(bx_handle* h, const float* rois, const int* roi_counts, int k, const float* gt,
                                  const int* gt_labels, const int* gt_counts, int batch, int max_gt, const int* perm,
                                  const bx_proposal_target_params* p, float* out_rois, int* out_labels,
                                  float* out_targets, float* out_in_w, float* out_out_w, int* out_keep,
                                  int* out_counts, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && (rois || k == 0) && (gt || max_gt == 0) && (gt_labels || max_gt == 0) && (perm || k == 0) && p &&
                 out_rois && out_labels && out_targets && out_in_w && out_out_w && out_keep && out_counts,
             BX_ERR_INVALID, "bx_proposal_target: NULL argument");
  BX_REQUIRE(k >= 0 && batch >= 0 && max_gt >= 0, BX_ERR_INVALID, "bx_proposal_target: negative size");
  BX_REQUIRE(k <= kMaxRois, BX_ERR_UNSUPPORTED, "bx_proposal_target: %d rois per image > %d", k, kMaxRois);
  BX_REQUIRE(max_gt <= kMaxGt, BX_ERR_UNSUPPORTED, "bx_proposal_target: max_gt %d > %d", max_gt, kMaxGt);
  BX_REQUIRE(p->num_classes > 0 && p->total_num_samples > 0 && p->max_pos_samples >= 0 &&
                 p->max_pos_samples <= p->total_num_samples, BX_ERR_INVALID, "bx_proposal_target: bad sampling parameters");
  BX_REQUIRE(bx_aligned(rois, 16) && bx_aligned(gt, 16) && bx_aligned(out_rois, 16), BX_ERR_INVALID,
             "bx_proposal_target: box tensors must be 16-byte aligned");
  if (batch == 0) return BX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PTArgs a = {};
  a.rois = reinterpret_cast<const float4*>(rois);
  a.roi_counts = roi_counts;
  a.gt = reinterpret_cast<const float4*>(gt);
  a.gt_labels = gt_labels;
  a.gt_counts = gt_counts;
  a.perm = perm;
  a.k = k;
  a.max_gt = max_gt;
  a.p = *p;
  a.codec = codec_of(p->means, p->stds);
  a.out_rois = reinterpret_cast<float4*>(out_rois);
  a.out_labels = out_labels;
  a.out_targets = out_targets;
  a.out_in_w = out_in_w;
  a.out_out_w = out_out_w;
  a.out_keep = out_keep;
  a.out_counts = out_counts;
  const size_t smem = sizeof(float4) * kMaxGt + 2 * sizeof(uint64_t) * kMaxRois + sizeof(float) * kMaxGt +
                      sizeof(unsigned short) * kMaxRois;
  if (int rc = bx_ws_reserve(h, static_cast<size_t>(batch) * (k > 0 ? k : 1) * (sizeof(float) + sizeof(int)), st)) return rc;
  a.ws_best = reinterpret_cast<float*>(h->ws);
  a.ws_arg = reinterpret_cast<int*>(a.ws_best + static_cast<size_t>(batch) * k);
  if (k > 0) {
    pt_rowstats_kernel<<<dim3(bx_div_up(k, 256), batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    BX_LAUNCH_CHECK(h);
  }
  BX_CUDA(cudaFuncSetAttribute(proposal_target_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  proposal_target_kernel<<<batch, 1024, smem, static_cast<cudaStream_t>(stream)>>>(a);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}
