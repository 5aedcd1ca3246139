// Proposal stage: anchor decode + clip, candidate selection (adaptive radix select over 64-bit
// (score,index) composites staged in shared memory), bitonic sort of the selected chunk, greedy IoU NMS with
// 64-bit tile bitmasks and early exit at post_nms.  One CTA (1024 threads) per image; a batch is one launch.
//
// Replaces model/region_proposal.py:37-81 (RegionProposal.call): decode (utils/bbox_transform.py:32-55), clip
// (utils/bbox_tf.py:59-78) and tf.image.non_max_suppression (TF r1.13 CPU kernel semantics).
//
// Why "lazy" chunks: greedy NMS visits candidates in descending score order and stops after post_nms keeps, so only a
// prefix of the sorted order is ever needed.  Each round selects the next <= CHUNK best candidates (exact set, any
// order), sorts just those, and sweeps them; with pre_nms_top_k = 0 (reference behaviour) rounds continue until the
// quota is filled or every anchor was visited, so the result is identical to sorting all N.
//
// Two extensions keep that kernel fed at the larger configurations:
//   * n > 24 576 (FPN, 150 k - 267 k anchors): a multi-CTA radix select + stable compaction over the whole device
//     ("top set") hands the kernel the largest <= 24 576 keys, which it caches in shared memory; a per-image flag and an
//     early-out relaunch cover the rare case that this list runs dry;
//   * post_nms >= 512: the kernel runs as a thread-block cluster per image; helper CTAs hold a share of the kept list
//     and test every tile against it (tile boxes pushed over DSMEM, mbarrier signalling, partial 64-bit masks).
// The same file holds the inputs of the stage (anchors on the device, RPN score layouts, "next" row f2).
#include <stdlib.h>

#include "bx_common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kChunk = 2048;        // candidates sorted + swept per round
constexpr int kMaxClusterRanks = 16; // CTAs of one image's cluster in the cluster sweep (16 = non-portable size, BX_NMS_CLUSTER=16)
constexpr int kBins = 2048;         // histogram bins per select level
static_assert(kBins == 2 * kThreads, "the select scan gives every thread two bins");
constexpr int kMaxPost = 2048;      // kept-box list capacity (post_nms limit)
constexpr int kTile = 128;          // largest NMS tile: candidates resolved per round (two 64-bit mask words); the kernel
                                    // is instantiated for 64 (small quotas: fewer pair tests) and 128 (cluster sweep:
                                    // half the signalling rounds)
constexpr int tile_pairs(int t) { return t * (t - 1) / 2; }   // unordered candidate pairs of a tile
constexpr int kUnroll = 8;          // independent key loads in flight per thread in the streaming passes
constexpr int kKeyCacheMax = 24576; // keys cached in smem when n <= this (C4 600x1000: 21 546)

// IoU threshold with the two screening constants of iou_gt() below
struct IouThr {
  float thr, hi, lo;
};
__host__ inline IouThr make_iou_thr(float thr) {
  IouThr t;
  t.thr = thr;
  const bool screen = thr >= 1e-6f;
  t.hi = screen ? thr * 1.000002f : 0.0f;
  t.lo = screen ? thr * 0.999998f : 0.0f;
  return t;
}
struct ProposalArgs {
  const float4* anchors;  // [n] (shared) — decode mode
  const float4* deltas;   // [batch,n] — decode mode
  const float4* boxes;    // [batch,n] precomputed boxes (bx_nms / min-size path); overrides decode when non-null
  const float* scores;    // [batch,n]
  const uint32_t* keys;   // [batch,n] precomputed keys (0 = excluded) or null -> derived from scores
  int n;
  BoxCodec codec;
  int pre_nms_top_k;
  int post_nms;
  IouThr thr;
  float4* out_boxes;      // [batch,post_nms] or null
  int* out_idx;           // [batch,post_nms]
  int* out_count;         // [batch]
  int cache_keys;         // 1: keys live in smem
  // f2: raw RPN logits instead of scores (cached-key kernel only; larger n goes through rpn_scores_kernel first)
  const float* logits;    // [batch, n*2] in `logit_layout`, or null
  int logit_layout;       // bx_rpn_layout
  int anchors_per_cell;   // A (BX_RPN_CAFFE)
  float* out_scores;      // [batch,n] or null: the foreground probabilities the order was taken from
  // top-set prefilter (n > kKeyCacheMax): `keys` is a compacted [batch, n] list of the largest keys in ascending anchor
  // order, `src_idx` maps its positions back to anchors, boxes / deltas keep the full stride `src_stride`
  const int* src_idx;     // [batch,n] or null
  int src_stride;         // anchors per image in deltas / boxes when src_idx is set
  int first_chunk;        // candidates of the first select / sort / sweep round; 0: smallest power of two >= the quota
  const int* topset_info; // [batch,4] (threshold lo, m, n_valid_full, fallback flag) or null
  int* flag_out;          // &info[0][3]: set to 1 when the compacted list ran dry before the quota was met
  const int* run_flag;    // fallback launch: images whose flag is 0 return immediately
};

// Foreground probability of anchor i from the raw RPN logits: tf.nn.softmax over (bg, fg) — exp(x - max) / sum, fp32.
//   BX_RPN_CAFFE  (faster_rcnn/base_faster_rcnn_model.py:149-152): rows of 2A per cell, [bg x A | fg x A]
//   BX_RPN_PAIRS  (fpn/base_fpn_model.py:223):                     (bg, fg) per anchor
__device__ __forceinline__ float rpn_fg_prob(const float* __restrict__ logits, int layout, int A, int i) {
  float bg, fg;
  if (layout == BX_RPN_CAFFE) {
    const int cell = i / A, k = i - cell * A;
    const float* row = logits + static_cast<size_t>(cell) * 2 * A;
    bg = row[k];
    fg = row[A + k];
  } else {
    const float2 v = *reinterpret_cast<const float2*>(logits + 2 * static_cast<size_t>(i));
    bg = v.x;
    fg = v.y;
  }
  const float m = fmaxf(bg, fg);
  const float e0 = expf(bg - m), e1 = expf(fg - m);
  return e1 / (e0 + e1);
}

__device__ __forceinline__ uint64_t composite(uint32_t key, uint32_t idx) {
  return (static_cast<uint64_t>(key) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}

// TF NonMaxSuppressionV3 IoU test on min/max-normalised corners (x=lo0,y=lo1,z=hi0,w=hi1).  TF decides
// `inter / (area_a + area_b - inter) > thr` with one correctly rounded fp32 division; that exact test is the fallback
// below.  In front of it sits a division-free screen that is decision-EQUIVALENT (not an approximation):
//   thr_hi = fl(thr * 1.000002f), thr_lo = fl(thr * 0.999998f)  (host; both 0 when thr < 1e-6, which disables the screen)
//   p_hi = fl(thr_hi * uni) >= thr * uni * (1 + 2e-6)(1 - 2^-24)^2 > thr * uni * (1 + 1.8e-6)
//   inter > p_hi  =>  inter / uni > thr (1 + 1.8e-6)  =>  fl(inter / uni) >= thr (1 + 1.8e-6)(1 - 2^-24) > thr
//   inter < p_lo  =>  inter / uni < thr (1 - 1.8e-6)  =>  fl(inter / uni) <= thr (1 - 1.8e-6)(1 + 2^-24) < thr
// The bounds need normal-range products (relative rounding error 2^-24); `p_lo >= 1e-30f` guarantees that for p_lo and
// p_hi, so tiny / denormal unions and thr ~ 0 always take the exact division.  Pairs within 2e-6 (relative) of the
// threshold take it too.  tests/test_gpu_parity.py::test_nms_iou_threshold_guard_band pins the decisions at, one ulp
// above and one ulp below the quotient.  (Measured: extra early-out branches — per-axis disjointness, area ratio — make
// the sweep slower, the straight-line form below is the fastest.)
__device__ __forceinline__ bool iou_gt(const float4 a, const float area_a, const float4 b, const IouThr t) {
  const float area_b = (b.z - b.x) * (b.w - b.y);
  if (area_a <= 0.0f || area_b <= 0.0f) return false;
  const float i0 = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.0f);
  const float i1 = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.0f);
  const float inter = i0 * i1;
  if (inter <= 0.0f) return false;  // iou == 0, never > thr for thr in [0,1]
  const float uni = area_a + area_b - inter;
  const float p_lo = t.lo * uni;
  if (p_lo >= 1e-30f) {
    if (inter > t.hi * uni) return true;
    if (inter < p_lo) return false;
  }
  return inter / uni > t.thr;       // TF's exact test (fp32 division, strict >)
}

__device__ __forceinline__ float4 normalise(const float4 b) {
  return make_float4(fminf(b.x, b.z), fminf(b.y, b.w), fmaxf(b.x, b.z), fmaxf(b.y, b.w));
}

// what a helper needs to know about a tile (48 bytes, pushed with three 16-byte remote stores)
struct ClusterCmd {
  int cmd;                // 1 = tile, 2 = done
  int t0, tn;             // tile start / size in the leader's cand_box
  int kept;               // kept count before this tile
  int p_valid, p_kept;    // previous tile: valid flag, kept count before it
  int pad0, pad1;
  uint64_t p_keepmask[2]; // previous tile: keep mask (128 bits)
};
static_assert(sizeof(ClusterCmd) == 48, "ClusterCmd is pushed as three 16-byte remote stores");

// bits [0, c) of a 128-bit mask held as two words
__device__ __forceinline__ uint64_t below_w0(int c) { return c >= 64 ? ~0ull : ((1ull << c) - 1ull); }
__device__ __forceinline__ uint64_t below_w1(int c) { return c <= 64 ? 0ull : (c >= 128 ? ~0ull : ((1ull << (c - 64)) - 1ull)); }

struct Shared {
  uint64_t lo, hi;        // current select range (inclusive)
  uint64_t prev;          // exclusive upper bound: composites already consumed are >= prev
  uint64_t thresh;        // selected threshold of this round
  uint64_t sup[2];        // tile: candidates suppressed by the kept list
  uint64_t keepmask[2];   // tile: candidates kept
  uint32_t kmin, kmax;    // key range of the image
  int n_valid;            // keys > 0
  int cand_count;
  int kept;
  int state;
  // thread-block-cluster sweep: the leader (cluster rank 0) pushes this block into every helper's copy of Shared
  alignas(16) ClusterCmd cc;
  int p_valid, p_kept;    // leader: result of the previous tile (copied into cc for the next one)
  uint64_t p_keepmask[2];
  uint64_t sup_part[kMaxClusterRanks][2];  // partial suppression masks pushed by the helpers
  uint64_t mb_tile;       // helpers: "tile pushed" (1 arrival per tile, from the leader)
  uint64_t mb_part;       // leader: "partial masks delivered" (cs - 1 arrivals per tile)
};

// ---- cluster primitives (no-ops / rank 0 of 1 when the kernel is launched without a cluster dimension)
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_smem_ptr, uint32_t rank) {
  const uint32_t la = static_cast<uint32_t>(__cvta_generic_to_shared(local_smem_ptr));
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  return ra;
}
__device__ __forceinline__ void dsmem_st_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// mbarrier signalling across the cluster: one thread arrives (release at cluster scope) on a barrier that lives in
// another CTA's shared memory; the waiters acquire at cluster scope, so the remote stores issued before the arrive
// (ordered behind a __syncthreads when other threads made them) are visible after the wait.
__device__ __forceinline__ void cmbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void cmbar_arrive_remote(uint64_t* local_bar, uint32_t rank) {
  const uint32_t ra = dsmem_addr(local_bar, rank);
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(ra) : "memory");
}
__device__ __forceinline__ void cmbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CM_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra CM_DONE;\n"
      "bra CM_WAIT;\n"
      "CM_DONE:\n"
      "}\n" :: "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void dsmem_st_u64(uint32_t addr, uint64_t v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" :: "r"(addr), "l"(static_cast<unsigned long long>(v)) : "memory");
}

template <bool kCache>
__device__ __forceinline__ uint32_t load_key(const ProposalArgs& a, const uint32_t* skeys, const float* scores,
                                             const uint32_t* gkeys, int i) {
  if (kCache) return skeys[i];
  if (gkeys) return gkeys[i];
  return bx_score_key(scores[i] + 0.0f);
}

// Helper CTA of a cluster (rank > 0): owns the kept boxes g with g % cs == rank and, for every tile the leader
// announces, tests the tile's 64 candidates against them.  Two mbarrier signals per tile instead of whole-cluster
// barriers: A = "tile pushed" (the leader has written the command block and the tile's boxes into this CTA's shared
// memory and arrives on this CTA's mb_tile), B = "partial mask delivered" (this CTA arrives on the leader's mb_part).
template <int TILE>
__device__ void nms_cluster_helper(const ProposalArgs& a, float4* tilebuf /* [2][TILE] */, float4* kept_box,
                                   Shared* sh, uint32_t rank, uint32_t cs) {
  __shared__ uint64_t h_sup[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lsup = dsmem_addr(&sh->sup_part[rank][0], 0);
  int cur = 0;
  for (;;) {
    cmbar_wait(&sh->mb_tile, static_cast<uint32_t>(cur));                  // A: the leader pushed a tile (or "done")
    const ClusterCmd cc = sh->cc;
    if (tid < 2) h_sup[tid] = 0ull;
    if (cc.cmd != 1) return;
    const int tn = cc.tn, kept = cc.kept;
    if (cc.p_valid && tid < TILE && ((cc.p_keepmask[tid >> 6] >> (tid & 63)) & 1ull)) {   // adopt this CTA's share of the last keeps
      const uint32_t g = static_cast<uint32_t>(cc.p_kept) + __popcll(cc.p_keepmask[0] & below_w0(tid)) +
                         __popcll(cc.p_keepmask[1] & below_w1(tid));
      if (g % cs == rank) kept_box[g / cs] = normalise(tilebuf[(cur ^ 1) * TILE + tid]);
    }
    __syncthreads();
    const int kl = (kept > static_cast<int>(rank)) ? (kept - static_cast<int>(rank) + static_cast<int>(cs) - 1) / static_cast<int>(cs) : 0;
    const int c = tid & (TILE - 1);
    bool sflag = false;
    if (c < tn) {
      const float4 cb = normalise(tilebuf[cur * TILE + c]);
      const float ca = (cb.z - cb.x) * (cb.w - cb.y);
      for (int j = tid / TILE; j < kl && !sflag; j += kThreads / TILE) sflag = iou_gt(cb, ca, kept_box[j], a.thr);
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, sflag);
    if (lane == 0 && m) atomicOr(reinterpret_cast<unsigned long long*>(&h_sup[((warp * 32) & (TILE - 1)) >> 6]),
                                 static_cast<unsigned long long>(m) << ((warp & 1) * 32));
    __syncthreads();
    if (tid == 0) {
      dsmem_st_u64(lsup, h_sup[0]);
      dsmem_st_u64(lsup + 8u, h_sup[1]);
      cmbar_arrive_remote(&sh->mb_part, 0);                                // B: this helper's partial mask is in
    }
    cur ^= 1;
  }
}

template <bool kCache, int TILE>
__global__ void __launch_bounds__(kThreads, 1) proposals_kernel(const ProposalArgs a) {
  constexpr int kPairs = TILE * (TILE - 1) / 2;   // unordered candidate pairs of a tile
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // carve-up (all 16B aligned)
  float4* cand_box = reinterpret_cast<float4*>(smem_raw);                           // kChunk
  float4* kept_box = cand_box + kChunk;                                             // kMaxPost
  uint64_t* cand_key = reinterpret_cast<uint64_t*>(kept_box + kMaxPost);            // kChunk
  uint64_t* rowmask = cand_key + kChunk;                                            // kTile rows x 2 words
  uint32_t* hist = reinterpret_cast<uint32_t*>(rowmask + 2 * kTile);                // kBins
  uint32_t* red = hist + kBins;                                                     // 64 (block reductions)
  Shared* sh = reinterpret_cast<Shared*>(red + 64);
  uint16_t* kept_ci = reinterpret_cast<uint16_t*>(sh + 1);                          // kMaxPost: chunk position of each keep
  float4* tile_nb = reinterpret_cast<float4*>(kept_ci + kMaxPost);                  // kTile: normalised tile boxes
  float* tile_area = reinterpret_cast<float*>(tile_nb + kTile);                     // kTile
  uint16_t* pair_tab = reinterpret_cast<uint16_t*>(tile_area + kTile);              // kTile*(kTile-1)/2: (i | j << 8), i < j
  uint32_t* skeys = reinterpret_cast<uint32_t*>(pair_tab + kPairs);                 // n (when cached)

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t cs = cluster_size(), crank = cluster_rank();   // 1, 0 unless launched with a cluster dimension
  const int img = blockIdx.x / cs;
  if (a.run_flag && a.run_flag[img * 4] == 0) return;           // fallback launch: nothing to redo for this image
  if (cs > 1) {
    if (tid == 0) {
      cmbar_init(&sh->mb_tile, 1);
      cmbar_init(&sh->mb_part, cs - 1);
      sh->p_valid = 0;
    }
    __syncthreads();
    cluster_sync_all();                      // every CTA's barriers exist before anybody signals
  }
  if (crank != 0) {
    nms_cluster_helper<TILE>(a, cand_box, kept_box, sh, crank, cs);
    return;
  }
  if (tid < TILE - 1) {                          // row tid of the upper triangle: pairs (tid, tid+1 .. TILE-1)
    int o = tid * (2 * TILE - 1 - tid) / 2;
    for (int j = tid + 1; j < TILE; ++j) pair_tab[o++] = static_cast<uint16_t>(tid | (j << 8));
  }
  const int n = a.topset_info ? min(a.topset_info[img * 4 + 1], a.n) : a.n;
  const size_t full = a.src_idx ? static_cast<size_t>(a.src_stride) : static_cast<size_t>(a.n);
  const float* scores = a.scores ? a.scores + static_cast<size_t>(img) * a.n : nullptr;
  const float* logits = a.logits ? a.logits + static_cast<size_t>(img) * a.n * 2 : nullptr;
  float* out_scores = a.out_scores ? a.out_scores + static_cast<size_t>(img) * a.n : nullptr;
  const uint32_t* gkeys = a.keys ? a.keys + static_cast<size_t>(img) * a.n : nullptr;
  const int* src_idx = a.src_idx ? a.src_idx + static_cast<size_t>(img) * a.n : nullptr;
  const float4* deltas = a.deltas ? a.deltas + static_cast<size_t>(img) * full : nullptr;
  const float4* boxes = a.boxes ? a.boxes + static_cast<size_t>(img) * full : nullptr;
  float4* out_boxes = a.out_boxes ? a.out_boxes + static_cast<size_t>(img) * a.post_nms : nullptr;
  int* out_idx = a.out_idx + static_cast<size_t>(img) * a.post_nms;

  // ---- pass 0: keys -> smem (if cached), min / max / count of valid keys
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  int nvalid = 0;
  for (int base = tid; base < n; base += kThreads * kUnroll) {
    uint32_t kk[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {   // all loads of the batch in flight before any is consumed
      const int i = base + u * kThreads;
      if (kCache && logits) {           // f2: softmax fused into the key pass
        const float sc = (i < n) ? rpn_fg_prob(logits, a.logit_layout, a.anchors_per_cell, i) : 0.0f;
        if (out_scores && i < n) out_scores[i] = sc;
        kk[u] = (i < n) ? bx_score_key(sc + 0.0f) : 0u;
      } else {
        kk[u] = (i < n) ? (gkeys ? gkeys[i] : bx_score_key(scores[i] + 0.0f)) : 0u;
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kThreads;
      const uint32_t k = kk[u];
      if (kCache && i < n) skeys[i] = k;
      if (k) {
        kmin = min(kmin, k);
        kmax = max(kmax, k);
        ++nvalid;
      }
    }
  }
  kmin = __reduce_min_sync(0xFFFFFFFFu, kmin);
  kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
  nvalid = __reduce_add_sync(0xFFFFFFFFu, nvalid);
  if (lane == 0) {
    red[warp] = kmin;
    red[32 + warp] = kmax;
    hist[warp] = nvalid;
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t mn = __reduce_min_sync(0xFFFFFFFFu, red[lane]);
    uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, red[32 + lane]);
    int nv = __reduce_add_sync(0xFFFFFFFFu, static_cast<int>(hist[lane]));
    if (lane == 0) {
      sh->kmin = mn;
      sh->kmax = mx;
      sh->n_valid = nv;
      sh->prev = 0xFFFFFFFFFFFFFFFFull;
      sh->kept = 0;
    }
  }
  __syncthreads();

  const int n_valid = sh->n_valid;
  const int limit = (a.pre_nms_top_k > 0) ? min(a.pre_nms_top_k, n_valid) : n_valid;
  const uint64_t global_lo = static_cast<uint64_t>(sh->kmin) << 32;
  int consumed = 0;
  int tile_seq = 0;                        // tiles announced to the helpers so far (double-buffer parity)

  // first round: the smallest power of two (>= 512) that can fill the quota; later rounds take full chunks.  Measured
  // (profiles/README.md): at quota 1000 a first chunk of 1024 beats 2048 by 15 - 18 % (sorting and decoding twice the quota
  // costs more than the second round it sometimes saves); at quota 300, 512 beats 256 and 1024.
  int round_cap = 512;
  while (round_cap < kChunk && round_cap < a.post_nms) round_cap <<= 1;
  if (a.first_chunk > 0) round_cap = min(a.first_chunk, kChunk);
  while (consumed < limit && sh->kept < a.post_nms) {
    const int want = min(round_cap, limit - consumed);
    round_cap = kChunk;
    const uint64_t prev = sh->prev;

    // ---- select: threshold T with  #{v : T <= v < prev} in [1, want]  (as large as the bins allow)
    if (tid == 0) {
      sh->lo = global_lo;
      sh->hi = prev - 1;   // prev > global_lo while consumed < limit
      if (prev == 0xFFFFFFFFFFFFFFFFull) sh->hi = (static_cast<uint64_t>(sh->kmax) << 32) | 0xFFFFFFFFull;
    }
    __syncthreads();
    for (;;) {
      const uint64_t lo = sh->lo, hi = sh->hi;
      const uint64_t span = hi - lo;  // inclusive span - 1
      int shift = 0;
      while ((span >> shift) >= static_cast<uint64_t>(kBins)) ++shift;
      for (int i = tid; i < kBins; i += kThreads) hist[i] = 0;
      __syncthreads();
      for (int base = tid; base < n; base += kThreads * kUnroll) {
        uint32_t kk[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int i = base + u * kThreads;
          kk[u] = (i < n) ? load_key<kCache>(a, skeys, scores, gkeys, i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (!kk[u]) continue;
          const uint64_t v = composite(kk[u], static_cast<uint32_t>(base + u * kThreads));
          if (v >= lo && v <= hi) atomicAdd(&hist[static_cast<uint32_t>((v - lo) >> shift)], 1u);
        }
      }
      __syncthreads();
      // suffix scan from the top bin: largest suffix whose count <= want.  Block-wide: thread t owns bins 2t, 2t+1
      // (kBins == 2 * kThreads); warp suffix by shuffles, warp totals through `red`.
      {
        const uint32_t c0 = hist[2 * tid], c1 = hist[2 * tid + 1];
        const uint32_t mine = c0 + c1;
        uint32_t suf = mine;                                  // inclusive suffix over the lanes of the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_down_sync(0xFFFFFFFFu, suf, d);
          if (lane + d < 32) suf += o;
        }
        if (lane == 0) red[warp] = suf;                       // warp totals
        __syncthreads();
        uint32_t above_warp = 0;
        for (int w = warp + 1; w < kThreads / 32; ++w) above_warp += red[w];
        const uint32_t higher = above_warp + suf - mine;      // count in the bins above this thread's two
        const uint32_t w32 = static_cast<uint32_t>(want);
        if (higher <= w32 && higher + mine > w32) {           // the crossing is here (exactly one thread, if any)
          int ob;
          uint32_t cnt_above;
          if (higher + c1 > w32) { ob = 2 * tid + 1; cnt_above = higher; }
          else { ob = 2 * tid; cnt_above = higher + c1; }
          if (cnt_above > 0) {
            // a non-empty suffix fits: T = lower edge of bin ob+1
            sh->thresh = lo + (static_cast<uint64_t>(ob + 1) << shift);
            sh->state = 1;
          } else {
            // the top non-empty bin alone holds more than `want`: refine inside it
            const uint64_t nlo = lo + (static_cast<uint64_t>(ob) << shift);
            uint64_t nhi = nlo + ((1ull << shift) - 1ull);
            if (nhi > hi) nhi = hi;
            sh->lo = nlo;
            sh->hi = nhi;
            sh->state = 0;
          }
        } else if (tid == 0 && higher + mine <= w32) {        // everything in [lo,hi] fits
          sh->thresh = lo;
          sh->state = 1;
        }
      }
      __syncthreads();
      if (sh->state == 1) break;
      // state 0 with shift == 0 cannot happen: a bin of width 1 holds exactly one composite (<= want since want >= 1)
    }
    const uint64_t T = sh->thresh;

    // ---- compaction of {T <= v < prev} into cand_key (any order), then pad to a power of two
    if (tid == 0) sh->cand_count = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kThreads * kUnroll) {
      uint32_t kk[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int i = base + u * kThreads + tid;
        kk[u] = (i < n) ? load_key<kCache>(a, skeys, scores, gkeys, i) : 0u;
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        uint64_t v = 0;
        bool take = false;
        if (kk[u]) {
          v = composite(kk[u], static_cast<uint32_t>(base + u * kThreads + tid));
          take = (v >= T) && (v < prev);
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, take);
        if (m) {
          int pos = 0;
          if (lane == 0) pos = atomicAdd(&sh->cand_count, __popc(m));
          pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
          if (take) cand_key[pos + __popc(m & ((1u << lane) - 1u))] = v;
        }
      }
    }
    __syncthreads();
    const int cnt = sh->cand_count;  // 1..want
    int pow2 = 64;
    while (pow2 < cnt) pow2 <<= 1;
    for (int i = cnt + tid; i < pow2; i += kThreads) cand_key[i] = 0ull;
    __syncthreads();

    // ---- bitonic sort, descending
    bx_bitonic_sort<true>(reinterpret_cast<unsigned long long*>(cand_key), pow2);

    // ---- decode + clip (or gather) the candidates' boxes, normalised corners for the IoU test
    for (int i = tid; i < cnt; i += kThreads) {
      uint32_t idx = 0xFFFFFFFFu - static_cast<uint32_t>(cand_key[i] & 0xFFFFFFFFull);
      if (src_idx) idx = static_cast<uint32_t>(src_idx[idx]);
      float4 b;
      if (boxes) b = boxes[idx];
      else b = bx_decode_clip_one(a.anchors[idx], deltas[idx], a.codec);
      cand_box[i] = b;  // original orientation (written to out_boxes); normalised on load below
    }
    __syncthreads();

    // ---- greedy sweep in tiles of TILE candidates
    const int chunk_kept0 = sh->kept;
    for (int t0 = 0; t0 < cnt && sh->kept < a.post_nms; t0 += TILE) {
      const int tn = min(TILE, cnt - t0);
      const int kept = sh->kept;
      if (tid == 0) {
        sh->sup[0] = 0ull;
        sh->sup[1] = 0ull;
        if (cs > 1) {                        // the tile's command block (the previous tile's keeps ride along)
          sh->cc.cmd = 1;
          sh->cc.t0 = t0;
          sh->cc.tn = tn;
          sh->cc.kept = kept;
          sh->cc.p_valid = sh->p_valid;
          sh->cc.p_kept = sh->p_kept;
          sh->cc.p_keepmask[0] = sh->p_keepmask[0];
          sh->cc.p_keepmask[1] = sh->p_keepmask[1];
        }
      }
      if (cs > 1 && tid < 2 * kMaxClusterRanks) sh->sup_part[tid >> 1][tid & 1] = 0ull;
      if (tid >= 128 && tid < 128 + TILE) { // normalised corners + area of the tile's candidates, masks cleared
        const int c = tid - 128;
        const float4 nb = normalise(cand_box[t0 + min(c, tn - 1)]);
        tile_nb[c] = nb;
        tile_area[c] = (nb.z - nb.x) * (nb.w - nb.y);
        rowmask[2 * c] = 0ull;
        rowmask[2 * c + 1] = 0ull;
      }
      __syncthreads();
      if (cs > 1) {
        // push the tile's boxes and the command block into every helper's shared memory, then signal A
        const int n_push = TILE * (static_cast<int>(cs) - 1);                   // 896 at 8 CTAs, 1920 at 16
        for (int i = tid; i < n_push; i += kThreads) {
          const uint32_t helper = 1u + static_cast<uint32_t>(i) / TILE;
          const int c = i % TILE;
          if (c < tn) {
            const float4 b = cand_box[t0 + c];
            dsmem_st_v4(dsmem_addr(&cand_box[(tile_seq & 1) * TILE + c], helper),
                        make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)));
          }
        }
        if (tid >= kThreads - 64 && tid < kThreads - 64 + 3 * (static_cast<int>(cs) - 1)) {   // the last two warps: 3 x 16 B per helper
          const int e = tid - (kThreads - 64), part = e % 3;
          const uint32_t helper = 1u + static_cast<uint32_t>(e / 3);
          const uint4* src = reinterpret_cast<const uint4*>(&sh->cc) + part;
          dsmem_st_v4(dsmem_addr(src, helper), *src);
        }
        __syncthreads();                     // every push is issued before the signal
        if (tid < static_cast<int>(cs) - 1) cmbar_arrive_remote(&sh->mb_tile, 1u + static_cast<uint32_t>(tid));   // A
        ++tile_seq;
      }
      // kept boxes are dealt round-robin over the cluster: this CTA holds g = 0, cs, 2cs, ... at kept_box[g / cs]
      const int kl = (cs > 1) ? (kept + static_cast<int>(cs) - 1) / static_cast<int>(cs) : kept;
      {
        // (1) tile candidates vs kept list: candidate c = tid mod TILE, kept subset j = tid / TILE (mod 1024 / TILE)
        const int c = tid & (TILE - 1);
        bool s = false;
        if (c < tn) {
          const float4 cb = tile_nb[c];
          const float ca = tile_area[c];
          for (int j = tid / TILE; j < kl && !s; j += kThreads / TILE) s = iou_gt(cb, ca, kept_box[j], a.thr);
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, s);
        if (lane == 0 && m) atomicOr(reinterpret_cast<unsigned long long*>(&sh->sup[((warp * 32) & (TILE - 1)) >> 6]),
                                     static_cast<unsigned long long>(m) << ((warp & 1) * 32));
        // (2) intra-tile masks: every unordered pair (i < j) once — the test is symmetric bit for bit, so a hit sets
        //     bit j of row i and bit i of row j (the bits below i are "the earlier candidates that suppress i")
        for (int p = tid; p < kPairs; p += kThreads) {
          const int e = pair_tab[p], i = e & 0xFF, j = e >> 8;
          if (j < tn && iou_gt(tile_nb[i], tile_area[i], tile_nb[j], a.thr)) {
            atomicOr(reinterpret_cast<unsigned long long*>(&rowmask[2 * i + (j >> 6)]), 1ull << (j & 63));
            atomicOr(reinterpret_cast<unsigned long long*>(&rowmask[2 * j + (i >> 6)]), 1ull << (i & 63));
          }
        }
      }
      __syncthreads();
      if (cs > 1) cmbar_wait(&sh->mb_part, static_cast<uint32_t>((tile_seq - 1) & 1));   // B: all partial masks are in
      if (warp == 0) {
        // greedy resolve of the tile by warp 0 as a fixed point: candidate i is kept iff it is alive and no KEPT earlier
        // candidate suppresses it.  Iterating K <- {i alive : full[i] & below(i) & K == 0} from K = alive fixes
        // candidates in index order (after t rounds every i < t is final), the greedy set is its only fixed point,
        // and suppression chains inside a tile are short: a few ballot rounds instead of a 128-step dependent chain.
        // Lane l owns candidates l, l+32, l+64, l+96.
        uint64_t f[4][2], bl[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = lane + 32 * k;
          f[k][0] = rowmask[2 * c];
          f[k][1] = rowmask[2 * c + 1];
          bl[k][0] = below_w0(c);
          bl[k][1] = below_w1(c);
        }
        uint64_t rem0 = sh->sup[0], rem1 = sh->sup[1];
        if (cs > 1) {
#pragma unroll
          for (int r = 1; r < kMaxClusterRanks; ++r) { rem0 |= sh->sup_part[r][0]; rem1 |= sh->sup_part[r][1]; }
        }
        rem0 |= ~below_w0(tn);
        rem1 |= ~below_w1(tn);
        const uint64_t al0 = ~rem0, al1 = ~rem1;
        const bool a0 = (al0 >> lane) & 1ull, a1 = (al0 >> (lane + 32)) & 1ull;
        const bool a2 = (al1 >> lane) & 1ull, a3 = (al1 >> (lane + 32)) & 1ull;
        uint64_t K0 = al0, K1 = al1;
        for (;;) {
          const bool k0 = a0 && (((f[0][0] & bl[0][0] & K0) | (f[0][1] & bl[0][1] & K1)) == 0ull);
          const bool k1 = a1 && (((f[1][0] & bl[1][0] & K0) | (f[1][1] & bl[1][1] & K1)) == 0ull);
          const bool k2 = a2 && (((f[2][0] & bl[2][0] & K0) | (f[2][1] & bl[2][1] & K1)) == 0ull);
          const bool k3 = a3 && (((f[3][0] & bl[3][0] & K0) | (f[3][1] & bl[3][1] & K1)) == 0ull);
          const uint64_t n0 = static_cast<uint64_t>(__ballot_sync(0xFFFFFFFFu, k0)) |
                              (static_cast<uint64_t>(__ballot_sync(0xFFFFFFFFu, k1)) << 32);
          const uint64_t n1 = static_cast<uint64_t>(__ballot_sync(0xFFFFFFFFu, k2)) |
                              (static_cast<uint64_t>(__ballot_sync(0xFFFFFFFFu, k3)) << 32);
          if (n0 == K0 && n1 == K1) break;
          K0 = n0;
          K1 = n1;
        }
        const int room = a.post_nms - kept;                      // quota: only the first `room` keeps count
        if (__popcll(K0) + __popcll(K1) > room) {
          bool kk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = lane + 32 * k;
            const bool set = (((k < 2 ? K0 : K1) >> (c & 63)) & 1ull) != 0ull;
            kk[k] = set && (__popcll(K0 & bl[k][0]) + __popcll(K1 & bl[k][1]) < room);
          }
          K0 = static_cast<uint64_t>(__ballot_sync(0xFFFFFFFFu, kk[0])) | (static_cast<uint64_t>(__ballot_sync(0xFFFFFFFFu, kk[1])) << 32);
          K1 = static_cast<uint64_t>(__ballot_sync(0xFFFFFFFFu, kk[2])) | (static_cast<uint64_t>(__ballot_sync(0xFFFFFFFFu, kk[3])) << 32);
        }
        if (lane == 0) {
          sh->keepmask[0] = K0;
          sh->keepmask[1] = K1;
          sh->kept = kept + __popcll(K0) + __popcll(K1);
          sh->p_valid = 1;
          sh->p_kept = kept;
          sh->p_keepmask[0] = K0;
          sh->p_keepmask[1] = K1;
        }
      }
      __syncthreads();
      if (tid < tn) {
        const uint64_t K0 = sh->keepmask[0], K1 = sh->keepmask[1];
        if (((tid < 64 ? K0 : K1) >> (tid & 63)) & 1ull) {
          const int pos = kept + __popcll(K0 & below_w0(tid)) + __popcll(K1 & below_w1(tid));
          const float4 b = cand_box[t0 + tid];
          if (cs == 1) kept_box[pos] = normalise(b);
          else if (pos % cs == 0) kept_box[pos / cs] = normalise(b);
          kept_ci[pos] = static_cast<uint16_t>(t0 + tid);        // outputs are written once per chunk, below
        }
      }
      __syncthreads();
    }
    // ---- this chunk's keeps -> global (kept off the per-tile critical path: src_idx is a dependent global load)
    for (int pos = chunk_kept0 + tid; pos < sh->kept; pos += kThreads) {
      const int ci = kept_ci[pos];
      const int p = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(cand_key[ci] & 0xFFFFFFFFull));
      out_idx[pos] = src_idx ? src_idx[p] : p;
      if (out_boxes) out_boxes[pos] = cand_box[ci];
    }
    __syncthreads();

    consumed += cnt;
    if (tid == 0) sh->prev = T;
    __syncthreads();
  }

  if (cs > 1) {                              // release the helpers; stay until they have read the command
    if (tid == 0) sh->cc.cmd = 2;
    __syncthreads();
    if (tid >= kThreads - 64 && tid < kThreads - 64 + 3 * (static_cast<int>(cs) - 1)) {
      const int e = tid - (kThreads - 64), part = e % 3;
      const uint4* src = reinterpret_cast<const uint4*>(&sh->cc) + part;
      dsmem_st_v4(dsmem_addr(src, 1u + static_cast<uint32_t>(e / 3)), *src);
    }
    __syncthreads();
    if (tid < static_cast<int>(cs) - 1) cmbar_arrive_remote(&sh->mb_tile, 1u + static_cast<uint32_t>(tid));
  }
  // ---- pad the tail, publish the count
  const int kept = sh->kept;
  for (int i = kept + tid; i < a.post_nms; i += kThreads) {
    out_idx[i] = -1;
    if (out_boxes) out_boxes[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid == 0) {
    a.out_count[img] = kept;
    if (a.flag_out) {
      // the compacted list is a prefix of the full descending order: the result stands unless it ran dry early
      const int n_full = a.topset_info[img * 4 + 2];
      const int limit_full = (a.pre_nms_top_k > 0) ? min(a.pre_nms_top_k, n_full) : n_full;
      a.flag_out[img * 4] = (kept < a.post_nms && consumed < limit_full) ? 1 : 0;
    }
  }
}


// ------------------------------------------------------------------ top-set prefilter (n > kKeyCacheMax)
// One CTA per image cannot stream a quarter of a million keys several times per chunk at any speed; the whole device
// can.  A multi-CTA radix select (11 + 11 + 10 bits, later levels only for images whose crossing bin overflows the
// budget) finds a threshold T with  m_lo <= #{key >= T} <= cap  (fewer only under massive exact ties), a two-pass
// stable compaction writes those keys in ascending anchor order, and the single-CTA kernel then runs on the compacted
// list with its keys cached in shared memory.  Its result is the full result unless the list runs dry before the quota
// is met; that (rare) case raises a per-image flag and the full-length kernel, launched right behind with an
// early-out on the flag, redoes just those images.
constexpr int kTopBins = 2048;
constexpr int kTopThreads = 512;
constexpr int kTopMaxSlices = 64;

struct TopsetResult;
struct TopsetArgs {
  const float* scores;     // [batch,n] or null
  const uint32_t* keys;    // [batch,n] or null (precomputed, 0 = excluded)
  int n, slices, slice_len;
  int m_lo, cap;
  uint32_t* hist;          // [batch][3][kTopBins], zeroed
  int* slice_count;        // [batch][kTopMaxSlices]
  int* info;               // [batch][4]: threshold (low 32 bits), m, n_valid_full, fallback flag
  int* t_hi;               // [batch]: 1 when the threshold is 2^32 (nothing taken)
  int* ticket;             // [batch][3], zeroed: CTAs of an image that finished histogram level L
  struct TopsetResult* state;  // [batch]: threshold search state after the last finished level
  uint32_t* out_keys;      // [batch][cap]
  int* out_src;            // [batch][cap]
};

struct TopsetResult {
  unsigned long long T;    // take key >= T (T may be 2^32: take nothing)
  uint32_t prefix;         // bins fixed so far when !done
  int done, count, n_valid;
  int b, above, cb;        // scratch of suffix_find
};

__device__ __forceinline__ uint32_t topset_key(const TopsetArgs& a, size_t base, int i) {
  return a.keys ? a.keys[base + i] : bx_score_key(a.scores[base + i] + 0.0f);
}

// Highest bin b of bins[0..nb) whose inclusive suffix count reaches `need` (need >= 1); r->b = -1 when the total is
// smaller.  r->above = count strictly above b, r->cb = bins[b], r->n_valid = total.  All kTopThreads threads call it.
__device__ void suffix_find(const uint32_t* __restrict__ bins, int nb, int need, TopsetResult* r, uint32_t* s_warp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t c[4], mine = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int bi = tid * 4 + j;
    c[j] = (bi < nb) ? __ldcg(bins + bi) : 0u;               // written by other CTAs' atomics: read at L2
    mine += c[j];
  }
  uint32_t suf = mine;                                     // inclusive suffix over lanes (lane 31 first)
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_down_sync(0xFFFFFFFFu, suf, d);
    if (lane + d < 32) suf += o;
  }
  if (lane == 0) s_warp[warp] = suf;                       // warp totals
  if (tid == 0) r->b = -1;
  __syncthreads();
  uint32_t higher = 0, total = 0;
  for (int w = 0; w < kTopThreads / 32; ++w) {
    const uint32_t t = s_warp[w];
    total += t;
    if (w > warp) higher += t;
  }
  const uint32_t incl = higher + suf, excl = incl - mine;  // counts at / above this thread's 4 bins
  if (excl < static_cast<uint32_t>(need) && incl >= static_cast<uint32_t>(need)) {   // exactly one thread
    uint32_t run = excl;
    for (int j = 3; j >= 0; --j) {
      if (run + c[j] >= static_cast<uint32_t>(need)) {
        r->b = tid * 4 + j;
        r->above = static_cast<int>(run);
        r->cb = static_cast<int>(c[j]);
        break;
      }
      run += c[j];
    }
  }
  if (tid == 0) r->n_valid = static_cast<int>(total);
  __syncthreads();
}

// Resolves the threshold from the first `levels` histogram levels.  r->done == 0 means level `levels` is still needed
// (r->prefix = the bins fixed so far).
__device__ void topset_resolve(const uint32_t* __restrict__ hist, int levels, int m_lo, int cap, TopsetResult* r,
                               uint32_t* s_warp) {
  suffix_find(hist, kTopBins, m_lo, r, s_warp);
  const int n_valid = r->n_valid;
  int b = r->b, above = r->above, cb = r->cb;
  __syncthreads();
  int need = m_lo, budget = cap, acc = 0;
  unsigned long long T = 1ull;
  int done = 0, count = n_valid;
  uint32_t prefix = 0;
  if (b < 0) {
    done = 1;                                              // fewer than m_lo valid keys: take them all
  } else if (above + cb <= budget) {
    T = static_cast<unsigned long long>(b) << 21;
    if (T == 0ull) T = 1ull;
    count = above + cb;
    done = 1;
  } else {
    prefix = static_cast<uint32_t>(b);
    need -= above; budget -= above; acc += above;
    if (levels >= 1) {
      suffix_find(hist + kTopBins, kTopBins, need, r, s_warp);
      b = r->b; above = r->above; cb = r->cb;
      __syncthreads();
      if (above + cb <= budget) {
        T = (static_cast<unsigned long long>(prefix) << 21) | (static_cast<unsigned long long>(b) << 10);
        count = acc + above + cb;
        done = 1;
      } else {
        prefix = (prefix << 11) | static_cast<uint32_t>(b);
        need -= above; budget -= above; acc += above;
        if (levels >= 2) {
          suffix_find(hist + 2 * kTopBins, 1024, need, r, s_warp);
          b = r->b; above = r->above; cb = r->cb;
          __syncthreads();
          const unsigned long long key = (static_cast<unsigned long long>(prefix) << 10) | static_cast<unsigned long long>(b);
          if (above + cb <= budget) {
            T = key;
            count = acc + above + cb;
          } else {                                         // more exact ties than the budget holds: stop above them
            T = key + 1ull;
            count = acc + above;
          }
          done = 1;
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    r->T = T; r->done = done; r->count = count; r->prefix = prefix; r->n_valid = n_valid;
  }
  __syncthreads();
}

template <int LEVEL>
__global__ void __launch_bounds__(kTopThreads) topset_hist_kernel(const TopsetArgs a) {
  __shared__ uint32_t s_hist[kTopBins];
  __shared__ uint32_t s_warp[kTopThreads / 32];
  __shared__ TopsetResult s_r;
  const int img = blockIdx.y, tid = threadIdx.x;
  uint32_t* hist = a.hist + static_cast<size_t>(img) * 3 * kTopBins;
  __shared__ int s_last;
  uint32_t prefix = 0;
  if (LEVEL > 0) {                                   // state left by the last CTA of the previous level
    if (a.state[img].done) return;
    prefix = a.state[img].prefix;
  }
  for (int i = tid; i < kTopBins; i += kTopThreads) s_hist[i] = 0u;
  __syncthreads();
  const size_t base = static_cast<size_t>(img) * a.n;
  const int lo = blockIdx.x * a.slice_len, hi = min(a.n, lo + a.slice_len);
  for (int i0 = lo + tid; i0 < hi; i0 += kTopThreads * 4) {
    uint32_t k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kTopThreads;
      k[u] = (i < hi) ? topset_key(a, base, i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!k[u]) continue;
      if (LEVEL == 0) atomicAdd(&s_hist[k[u] >> 21], 1u);
      else if (LEVEL == 1) { if ((k[u] >> 21) == prefix) atomicAdd(&s_hist[(k[u] >> 10) & 2047u], 1u); }
      else { if ((k[u] >> 10) == prefix) atomicAdd(&s_hist[k[u] & 1023u], 1u); }
    }
  }
  __syncthreads();
  for (int i = tid; i < kTopBins; i += kTopThreads)
    if (s_hist[i]) atomicAdd(&hist[LEVEL * kTopBins + i], s_hist[i]);
  // the image's last CTA to get here resolves the threshold search once for everybody downstream
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&a.ticket[img * 3 + LEVEL], 1) == static_cast<int>(gridDim.x) - 1) ? 1 : 0;
  __syncthreads();
  if (s_last) {
    __threadfence();
    topset_resolve(hist, LEVEL, a.m_lo, a.cap, &s_r, s_warp);
    if (tid == 0) a.state[img] = s_r;
  }
}

__global__ void __launch_bounds__(kTopThreads) topset_count_kernel(const TopsetArgs a) {
  __shared__ TopsetResult s_r;
  __shared__ int s_cnt;
  const int img = blockIdx.y, tid = threadIdx.x;
  if (tid == 0) { s_r = a.state[img]; s_cnt = 0; }     // final: level 2 always terminates the search
  __syncthreads();
  const unsigned long long T = s_r.T;
  __syncthreads();
  const size_t base = static_cast<size_t>(img) * a.n;
  const int lo = blockIdx.x * a.slice_len, hi = min(a.n, lo + a.slice_len);
  int cnt = 0;
  for (int i = lo + tid; i < hi; i += kTopThreads) {
    const uint32_t k = topset_key(a, base, i);
    cnt += (k != 0u && static_cast<unsigned long long>(k) >= T) ? 1 : 0;
  }
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (tid == 0) {
    a.slice_count[img * kTopMaxSlices + blockIdx.x] = s_cnt;
    if (blockIdx.x == 0) {
      a.info[img * 4 + 0] = static_cast<int>(static_cast<uint32_t>(T & 0xFFFFFFFFull));
      a.info[img * 4 + 1] = s_r.count;
      a.info[img * 4 + 2] = s_r.n_valid;
      a.info[img * 4 + 3] = 0;
      a.t_hi[img] = (T >> 32) ? 1 : 0;
    }
  }
}

__global__ void __launch_bounds__(kTopThreads) topset_write_kernel(const TopsetArgs a) {
  // stable compaction of the slice: every warp owns a contiguous sub-chunk, counts its takes, and after one block-wide
  // scan of the 16 warp counts writes them in order — no block barrier inside the element loops
  __shared__ int s_wcnt[kTopThreads / 32];
  const int img = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long T = (static_cast<unsigned long long>(a.t_hi[img]) << 32) |
                               static_cast<unsigned long long>(static_cast<uint32_t>(a.info[img * 4 + 0]));
  int pos = 0;                                             // first output slot of this slice
  for (int s = 0; s < static_cast<int>(blockIdx.x); ++s) pos += a.slice_count[img * kTopMaxSlices + s];
  const size_t base = static_cast<size_t>(img) * a.n;
  uint32_t* out_keys = a.out_keys + static_cast<size_t>(img) * a.cap;
  int* out_src = a.out_src + static_cast<size_t>(img) * a.cap;
  const int lo = blockIdx.x * a.slice_len, hi = min(a.n, lo + a.slice_len);
  constexpr int kWarps = kTopThreads / 32;
  const int wlen = ((hi - lo + kWarps - 1) / kWarps + 31) & ~31;      // multiple of 32: whole-warp steps
  const int wlo = min(hi, lo + warp * wlen), whi = min(hi, wlo + wlen);
  int mine = 0;
  for (int i0 = wlo; i0 < whi; i0 += 32 * 4) {
    uint32_t k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 32 + lane;
      k[u] = (i < whi) ? topset_key(a, base, i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) mine += (k[u] != 0u && static_cast<unsigned long long>(k[u]) >= T) ? 1 : 0;
  }
  mine = __reduce_add_sync(0xFFFFFFFFu, mine);
  if (lane == 0) s_wcnt[warp] = mine;
  __syncthreads();
  for (int w = 0; w < warp; ++w) pos += s_wcnt[w];
  for (int i0 = wlo; i0 < whi; i0 += 32 * 4) {
    uint32_t k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 32 + lane;
      k[u] = (i < whi) ? topset_key(a, base, i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool take = (k[u] != 0u) && (static_cast<unsigned long long>(k[u]) >= T);
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, take);
      if (take) {
        const int o = pos + __popc(m & ((1u << lane) - 1u));
        if (o < a.cap) {                                   // always true: count <= cap by construction
          out_keys[o] = k[u];
          out_src[o] = i0 + u * 32 + lane;
        }
      }
      pos += __popc(m);
    }
  }
}

size_t proposals_smem_bytes(int n, bool cache, int tile) {
  size_t b = sizeof(float4) * (kChunk + kMaxPost) + sizeof(uint64_t) * (kChunk + 2 * kTile) +
             sizeof(uint32_t) * (kBins + 64) + sizeof(Shared) + sizeof(uint16_t) * kMaxPost +
             sizeof(float4) * kTile + sizeof(float) * kTile + sizeof(uint16_t) * tile_pairs(tile);
  if (cache) b += sizeof(uint32_t) * static_cast<size_t>(n);
  return (b + 15) & ~static_cast<size_t>(15);
}

static int launch_proposals_kernel(bx_handle* h, ProposalArgs& a, int batch, cudaStream_t st) {
  const bool cache = a.n <= kKeyCacheMax;
  a.cache_keys = cache ? 1 : 0;
  if (const char* fc = getenv("BX_PROP_FIRST_CHUNK")) a.first_chunk = atoi(fc);   // A/B switch, read per call
  const size_t smem64 = proposals_smem_bytes(a.n, cache, 64);
  BX_REQUIRE(smem64 <= h->smem_optin, BX_ERR_UNSUPPORTED, "proposals: %zu B shared memory > device limit %zu", smem64,
             h->smem_optin);
  if (!cache) {
    BX_CUDA(cudaFuncSetAttribute(proposals_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
    proposals_kernel<false, 64><<<batch, kThreads, smem64, st>>>(a);
    BX_LAUNCH_CHECK(h);
    return BX_OK;
  }
  // large quotas: the sweep against the kept list dominates -> spread it over a thread-block cluster per image, with
  // 128-candidate tiles (half the leader <-> helper signalling rounds); small quotas keep 64-candidate tiles on one CTA
  static const int cs_env = getenv("BX_NMS_CLUSTER") ? atoi(getenv("BX_NMS_CLUSTER")) : 0;   // A/B switch
  const size_t smem128 = proposals_smem_bytes(a.n, cache, 128);
  int cs = (a.post_nms >= 512 && smem128 <= h->smem_optin) ? 8 : 1;
  if (cs_env > 0) cs = cs_env > kMaxClusterRanks ? kMaxClusterRanks : cs_env;
  if (cs > 1 && smem128 > h->smem_optin) cs = 1;
  if (cs > 1) BX_CUDA(cudaFuncSetAttribute(proposals_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem128));
  if (cs > 8) BX_CUDA(cudaFuncSetAttribute(proposals_kernel<true, 128>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (; cs > 1; cs >>= 1) {               // largest cluster size whose clusters are all co-resident (one wave)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(batch * cs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem128;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, proposals_kernel<true, 128>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    if (max_clusters < batch && cs_env <= 0) continue;
    if (max_clusters < 1) continue;
    BX_CUDA(cudaLaunchKernelEx(&cfg, proposals_kernel<true, 128>, a));
    break;
  }
  if (cs == 1) {
    BX_CUDA(cudaFuncSetAttribute(proposals_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
    proposals_kernel<true, 64><<<batch, kThreads, smem64, st>>>(a);
  }
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

size_t topset_ws_bytes(int batch, int cap) {
  return static_cast<size_t>(batch) * (3 * kTopBins * sizeof(uint32_t) + 4 * sizeof(int) + kTopMaxSlices * sizeof(int) +
                                       4 * sizeof(int) + sizeof(int) + sizeof(TopsetResult) +
                                       static_cast<size_t>(cap) * (sizeof(uint32_t) + sizeof(int))) + 512;
}

// `ws` = workspace region of topset_ws_bytes(batch, cap) bytes reserved by the caller (16-byte aligned)
int launch_proposals(bx_handle* h, ProposalArgs& a, int batch, cudaStream_t st, void* ws = nullptr) {
  static const bool no_topset = getenv("BX_NO_TOPSET") != nullptr;          // A/B switch for profiles/micro
  if (a.n <= kKeyCacheMax || !ws || no_topset) return launch_proposals_kernel(h, a, batch, st);
  // ---- top-set prefilter over the whole device, then the cached-key kernel on the compacted list
  const int cap = kKeyCacheMax;
  int m_lo = 8 * a.post_nms;
  if (m_lo < 4096) m_lo = 4096;
  if (a.pre_nms_top_k > 0 && a.pre_nms_top_k < m_lo) m_lo = a.pre_nms_top_k;   // never need more than the pre-NMS cut
  if (a.pre_nms_top_k > m_lo && a.pre_nms_top_k <= cap) m_lo = a.pre_nms_top_k; // ... and take all of it when it fits
  if (m_lo > cap) m_lo = cap;
  TopsetArgs t = {};
  t.scores = a.scores;
  t.keys = a.keys;
  t.n = a.n;
  int slices = (2 * h->num_sms + batch - 1) / batch;
  if (slices > kTopMaxSlices) slices = kTopMaxSlices;
  if (slices < 1) slices = 1;
  t.slice_len = (a.n + slices - 1) / slices;
  t.slices = (a.n + t.slice_len - 1) / t.slice_len;
  t.m_lo = m_lo;
  t.cap = cap;
  char* p = static_cast<char*>(ws);
  t.hist = reinterpret_cast<uint32_t*>(p);          p += static_cast<size_t>(batch) * 3 * kTopBins * sizeof(uint32_t);
  t.ticket = reinterpret_cast<int*>(p);             p += static_cast<size_t>(batch) * 4 * sizeof(int);   // zeroed with hist
  t.slice_count = reinterpret_cast<int*>(p);        p += static_cast<size_t>(batch) * kTopMaxSlices * sizeof(int);
  t.info = reinterpret_cast<int*>(p);               p += static_cast<size_t>(batch) * 4 * sizeof(int);
  t.t_hi = reinterpret_cast<int*>(p);               p += ((static_cast<size_t>(batch) * sizeof(int) + 15) & ~size_t(15));
  t.state = reinterpret_cast<TopsetResult*>(p);     p += ((static_cast<size_t>(batch) * sizeof(TopsetResult) + 15) & ~size_t(15));
  t.out_keys = reinterpret_cast<uint32_t*>(p);      p += static_cast<size_t>(batch) * cap * sizeof(uint32_t);
  t.out_src = reinterpret_cast<int*>(p);
  BX_CUDA(cudaMemsetAsync(t.hist, 0, static_cast<size_t>(batch) * (3 * kTopBins * sizeof(uint32_t) + 4 * sizeof(int)), st));
  const dim3 grid(t.slices, batch);
  topset_hist_kernel<0><<<grid, kTopThreads, 0, st>>>(t);
  BX_LAUNCH_CHECK(h);
  topset_hist_kernel<1><<<grid, kTopThreads, 0, st>>>(t);
  BX_LAUNCH_CHECK(h);
  topset_hist_kernel<2><<<grid, kTopThreads, 0, st>>>(t);
  BX_LAUNCH_CHECK(h);
  topset_count_kernel<<<grid, kTopThreads, 0, st>>>(t);
  BX_LAUNCH_CHECK(h);
  topset_write_kernel<<<grid, kTopThreads, 0, st>>>(t);
  BX_LAUNCH_CHECK(h);
  ProposalArgs c = a;                     // compacted list: keys cached in smem, boxes / deltas through src_idx
  c.scores = nullptr;
  c.logits = nullptr;
  c.out_scores = nullptr;
  c.keys = t.out_keys;
  c.src_idx = t.out_src;
  c.src_stride = a.n;
  c.n = cap;
  c.topset_info = t.info;
  c.flag_out = t.info + 3;
  if (int rc = launch_proposals_kernel(h, c, batch, st)) return rc;
  a.run_flag = t.info + 3;                // full-length redo of the images whose list ran dry (normally none)
  return launch_proposals_kernel(h, a, batch, st);
}

// ------------------------------------------------------------------ elementwise helpers
__global__ void __launch_bounds__(256) decode_clip_kernel(const float4* __restrict__ anchors, int anchors_batched,
                                                          const float4* __restrict__ deltas, int batch, int n,
                                                          BoxCodec codec, float4* __restrict__ out) {
  const long long total = static_cast<long long>(batch) * n;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int j = static_cast<int>(i % n);
    const float4 a = anchors_batched ? anchors[i] : anchors[j];
    out[i] = bx_decode_clip_one(a, deltas[i], codec);
  }
}

// decode + clip + min-size key masking for the min_size > 0 path (boxes kept for the later gather)
__global__ void __launch_bounds__(256) decode_filter_keys_kernel(const float4* __restrict__ anchors,
                                                                 const float4* __restrict__ deltas,
                                                                 const float* __restrict__ scores, int batch, int n,
                                                                 BoxCodec codec, float min_size,
                                                                 float4* __restrict__ boxes,
                                                                 uint32_t* __restrict__ keys) {
  const long long total = static_cast<long long>(batch) * n;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const float4 b = bx_decode_clip_one(anchors[i % n], deltas[i], codec);
    boxes[i] = b;
    const bool ok = ((b.z - b.x + 1.0f) >= min_size) && ((b.w - b.y + 1.0f) >= min_size);  // utils/bbox_tf.py:80-83
    keys[i] = ok ? bx_score_key(scores[i] + 0.0f) : 0u;
  }
}

__global__ void __launch_bounds__(256) encode_kernel(const float4* __restrict__ src, const float4* __restrict__ dst,
                                                     int n, BoxCodec codec, float4* __restrict__ out) {
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += 256 * gridDim.x) out[i] = bx_encode_one(src[i], dst[i], codec);
}

// single-CTA ordered compaction (clip + min-edge filter, or inside-image filter); n is small on these call sites
__global__ void __launch_bounds__(1024) filter_compact_kernel(const float4* __restrict__ in, int n, int mode,
                                                              float lo, float max_x, float max_y, float min_edge,
                                                              float4* __restrict__ out_boxes, int* __restrict__ out_idx,
                                                              int* __restrict__ out_count) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n; b0 += 1024) {
    const int i = b0 + tid;
    bool keep = false;
    float4 v = make_float4(0, 0, 0, 0);
    if (i < n) {
      v = in[i];
      if (mode == 0) {  // bboxes_clip_filter with min_edge
        v.x = fmaxf(fminf(v.x, max_x), lo);
        v.y = fmaxf(fminf(v.y, max_y), lo);
        v.z = fmaxf(fminf(v.z, max_x), lo);
        v.w = fmaxf(fminf(v.w, max_y), lo);
        keep = ((v.z - v.x + 1.0f) >= min_edge) && ((v.w - v.y + 1.0f) >= min_edge);
      } else {          // bboxes_range_filter
        keep = (v.x >= 0.0f) && (v.y >= 0.0f) && (v.z <= max_x) && (v.w <= max_y);
      }
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
      const int c = warp_tot[w];
      if (w < warp) off += c;
      tot += c;
    }
    const int base = base_s;
    if (keep) {
      const int pos = base + off + __popc(m & ((1u << lane) - 1u));
      if (out_boxes) out_boxes[pos] = v;
      out_idx[pos] = i;
    }
    __syncthreads();
    if (tid == 0) base_s = base + tot;
    __syncthreads();
  }
  if (tid == 0) *out_count = base_s;
}

BoxCodec make_codec(const float means[4], const float stds[4], int image_h, int image_w) {
  BoxCodec k;
  k.m0 = means[0]; k.m1 = means[1]; k.m2 = means[2]; k.m3 = means[3];
  k.s0 = stds[0]; k.s1 = stds[1]; k.s2 = stds[2]; k.s3 = stds[3];
  k.clip = (image_h > 0 && image_w > 0) ? 1 : 0;
  k.max_x = static_cast<float>(image_w - 1);
  k.max_y = static_cast<float>(image_h - 1);
  return k;
}

}  // namespace

extern "C" int bx_decode_clip(bx_handle* h, const float* anchors, int anchors_batched, const float* deltas,
                              int batch, int n, const float means[4], const float stds[4], int image_h, int image_w,
                              float* out_boxes, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && anchors && deltas && out_boxes && means && stds, BX_ERR_INVALID, "bx_decode_clip: NULL argument");
  BX_REQUIRE(batch >= 0 && n >= 0, BX_ERR_INVALID, "bx_decode_clip: negative size");
  BX_REQUIRE(bx_aligned(anchors, 16) && bx_aligned(deltas, 16) && bx_aligned(out_boxes, 16), BX_ERR_INVALID,
             "bx_decode_clip: box tensors must be 16-byte aligned");
  if (batch == 0 || n == 0) return BX_OK;
  const long long total = static_cast<long long>(batch) * n;
  const int grid = static_cast<int>(bx_min_ll(bx_div_up(total, 256), 8ll * h->num_sms));
  decode_clip_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(anchors), anchors_batched, reinterpret_cast<const float4*>(deltas), batch, n,
      make_codec(means, stds, image_h, image_w), reinterpret_cast<float4*>(out_boxes));
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

extern "C" int bx_encode(bx_handle* h, const float* src, const float* dst, int n, const float means[4],
                         const float stds[4], float* out, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && src && dst && out && means && stds, BX_ERR_INVALID, "bx_encode: NULL argument");
  BX_REQUIRE(bx_aligned(src, 16) && bx_aligned(dst, 16) && bx_aligned(out, 16), BX_ERR_INVALID,
             "bx_encode: box tensors must be 16-byte aligned");
  if (n <= 0) return BX_OK;
  encode_kernel<<<min(bx_div_up(n, 256), 8 * h->num_sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<const float4*>(dst), n, make_codec(means, stds, 0, 0),
      reinterpret_cast<float4*>(out));
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

extern "C" int bx_clip_filter(bx_handle* h, const float* boxes, int n, float min_value, int image_h, int image_w,
                              float min_edge, float* out_boxes, int* out_idx, int* out_count, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && boxes && out_boxes && out_idx && out_count, BX_ERR_INVALID, "bx_clip_filter: NULL argument");
  BX_REQUIRE(n >= 0, BX_ERR_INVALID, "bx_clip_filter: negative size");
  BX_REQUIRE(bx_aligned(boxes, 16) && bx_aligned(out_boxes, 16), BX_ERR_INVALID, "bx_clip_filter: alignment");
  filter_compact_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(boxes), n, 0, min_value, static_cast<float>(image_w - 1),
      static_cast<float>(image_h - 1), min_edge, reinterpret_cast<float4*>(out_boxes), out_idx, out_count);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

extern "C" int bx_range_filter(bx_handle* h, const float* anchors, int n, int image_h, int image_w, int* out_idx,
                               int* out_count, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && anchors && out_idx && out_count, BX_ERR_INVALID, "bx_range_filter: NULL argument");
  BX_REQUIRE(n >= 0 && bx_aligned(anchors, 16), BX_ERR_INVALID, "bx_range_filter: bad size/alignment");
  filter_compact_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(anchors), n, 1, 0.0f, static_cast<float>(image_w - 1),
      static_cast<float>(image_h - 1), 0.0f, nullptr, out_idx, out_count);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

// internal (bx_prediction.cu): greedy NMS over precomputed boxes with precomputed order keys (0 = excluded candidate)
int bx_internal_nms_keys(bx_handle* h, const float* boxes, const uint32_t* keys, int batch, int n, int max_out,
                         float iou_threshold, float* out_boxes, int* out_idx, int* out_count, cudaStream_t st) {
  BX_REQUIRE(max_out <= kMaxPost, BX_ERR_UNSUPPORTED, "max_output_size %d > %d", max_out, kMaxPost);
  BX_REQUIRE(n < (1 << 22), BX_ERR_UNSUPPORTED, "n must be < 2^22");
  ProposalArgs a = {};
  a.boxes = reinterpret_cast<const float4*>(boxes);
  a.keys = keys;
  a.n = n;
  a.post_nms = max_out;
  a.thr = make_iou_thr(iou_threshold);
  a.out_boxes = reinterpret_cast<float4*>(out_boxes);
  a.out_idx = out_idx;
  a.out_count = out_count;
  return launch_proposals(h, a, batch, st);
}

extern "C" int bx_nms(bx_handle* h, const float* boxes, const float* scores, int batch, int n, int max_out,
                      float iou_threshold, int* out_idx, int* out_count, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && boxes && scores && out_idx && out_count, BX_ERR_INVALID, "bx_nms: NULL argument");
  BX_REQUIRE(batch >= 0 && n >= 0 && max_out >= 0, BX_ERR_INVALID, "bx_nms: negative size");
  BX_REQUIRE(iou_threshold >= 0.0f && iou_threshold <= 1.0f, BX_ERR_INVALID,
             "bx_nms: iou_threshold must be in [0, 1]");  // TF: InvalidArgument
  BX_REQUIRE(max_out <= kMaxPost, BX_ERR_UNSUPPORTED, "bx_nms: max_output_size %d > %d", max_out, kMaxPost);
  BX_REQUIRE(n < (1 << 22), BX_ERR_UNSUPPORTED, "bx_nms: n must be < 2^22");
  BX_REQUIRE(bx_aligned(boxes, 16), BX_ERR_INVALID, "bx_nms: boxes must be 16-byte aligned");
  if (batch == 0) return BX_OK;
  ProposalArgs a = {};
  a.boxes = reinterpret_cast<const float4*>(boxes);
  a.scores = scores;
  a.n = n;
  a.post_nms = max_out;
  a.thr = make_iou_thr(iou_threshold);
  a.out_idx = out_idx;
  a.out_count = out_count;
  void* top_ws = nullptr;
  if (n > kKeyCacheMax) {
    if (int rc = bx_ws_reserve(h, topset_ws_bytes(batch, kKeyCacheMax), static_cast<cudaStream_t>(stream))) return rc;
    top_ws = h->ws;
  }
  return launch_proposals(h, a, batch, static_cast<cudaStream_t>(stream), top_ws);
}

__global__ void __launch_bounds__(256) rpn_scores_kernel(const float* __restrict__ logits, int layout, int A,
                                                         long long total, float* __restrict__ out) {
  // logits of consecutive images are contiguous and n is a multiple of A, so the flat index works across the batch
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    if (layout == BX_RPN_CAFFE) {
      const long long cell = i / A;
      const int k = static_cast<int>(i - cell * A);
      const float* row = logits + cell * 2 * A;
      const float bg = row[k], fg = row[A + k];
      const float m = fmaxf(bg, fg);
      const float e0 = expf(bg - m), e1 = expf(fg - m);
      out[i] = e1 / (e0 + e1);
    } else {
      const float2 v = *reinterpret_cast<const float2*>(logits + 2 * i);
      const float m = fmaxf(v.x, v.y);
      const float e0 = expf(v.x - m), e1 = expf(v.y - m);
      out[i] = e1 / (e0 + e1);
    }
  }
}

static int rpn_scores_launch(bx_handle* h, const float* logits, int layout, int A, int batch, int n, float* out,
                             cudaStream_t st) {
  const long long total = static_cast<long long>(batch) * n;
  if (total == 0) return BX_OK;
  const int grid = static_cast<int>(bx_min_ll(bx_div_up(total, 256), 8ll * h->num_sms));
  rpn_scores_kernel<<<grid, 256, 0, st>>>(logits, layout, A, total, out);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

// ---- f2: anchors on the device (utils/anchor_generator.py:46-60 generate_by_anchor_base_tf, :137-162 make_anchors).
// Both reduce to (x*stride, y*stride, x*stride, y*stride) + a per-anchor offset quadruple: one exact fp32 add each.
constexpr int kMaxAnchorLevels = 5;
constexpr int kMaxAnchorsPerCell = 32;
struct AnchorGenArgs {
  int n_levels, a;
  int fw[kMaxAnchorLevels];
  float stride[kMaxAnchorLevels];
  long long first[kMaxAnchorLevels + 1];     // first anchor index of each level
  float4 off[kMaxAnchorLevels * kMaxAnchorsPerCell];
  float4* out;
};

__global__ void __launch_bounds__(256) generate_anchors_kernel(const __grid_constant__ AnchorGenArgs g) {
  const long long total = g.first[g.n_levels];
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    int l = 0;
    while (l + 1 < g.n_levels && i >= g.first[l + 1]) ++l;
    const long long r = i - g.first[l];
    const int k = static_cast<int>(r % g.a);
    const long long cell = r / g.a;
    const float sx = static_cast<float>(cell % g.fw[l]) * g.stride[l];
    const float sy = static_cast<float>(cell / g.fw[l]) * g.stride[l];
    const float4 o = g.off[l * g.a + k];
    g.out[i] = make_float4(sx + o.x, sy + o.y, sx + o.z, sy + o.w);
  }
}

extern "C" int bx_generate_anchors(bx_handle* h, int n_levels, const int* fh, const int* fw, const float* stride,
                                   int anchors_per_cell, const float* offsets, float* out_anchors, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && fh && fw && stride && offsets, BX_ERR_INVALID, "bx_generate_anchors: NULL argument");
  BX_REQUIRE(n_levels >= 1 && n_levels <= kMaxAnchorLevels, BX_ERR_UNSUPPORTED,
             "bx_generate_anchors: n_levels %d not in [1, %d]", n_levels, kMaxAnchorLevels);
  BX_REQUIRE(anchors_per_cell >= 1 && anchors_per_cell <= kMaxAnchorsPerCell, BX_ERR_UNSUPPORTED,
             "bx_generate_anchors: anchors_per_cell %d not in [1, %d]", anchors_per_cell, kMaxAnchorsPerCell);
  BX_REQUIRE(bx_aligned(out_anchors, 16), BX_ERR_INVALID, "bx_generate_anchors: output must be 16-byte aligned");
  AnchorGenArgs g = {};
  g.n_levels = n_levels;
  g.a = anchors_per_cell;
  g.first[0] = 0;
  for (int l = 0; l < n_levels; ++l) {
    BX_REQUIRE(fh[l] >= 0 && fw[l] >= 0 && fw[l] < (1 << 24) && fh[l] < (1 << 24), BX_ERR_INVALID,
               "bx_generate_anchors: bad feature-map shape at level %d", l);
    g.fw[l] = fw[l] > 0 ? fw[l] : 1;
    g.stride[l] = stride[l];
    g.first[l + 1] = g.first[l] + static_cast<long long>(fh[l]) * fw[l] * anchors_per_cell;
    for (int k = 0; k < anchors_per_cell; ++k) {
      const float* o = offsets + (static_cast<size_t>(l) * anchors_per_cell + k) * 4;
      g.off[l * anchors_per_cell + k] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  g.out = reinterpret_cast<float4*>(out_anchors);
  const long long total = g.first[n_levels];
  if (total == 0) return BX_OK;
  BX_REQUIRE(out_anchors, BX_ERR_INVALID, "bx_generate_anchors: NULL output");
  const int grid = static_cast<int>(bx_min_ll(bx_div_up(total, 256), 8ll * h->num_sms));
  generate_anchors_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(g);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

static int check_rpn(const char* who, const float* logits, int layout, int A, int n) {
  BX_REQUIRE(logits, BX_ERR_INVALID, "%s: NULL logits", who);
  BX_REQUIRE(layout == BX_RPN_CAFFE || layout == BX_RPN_PAIRS, BX_ERR_INVALID, "%s: bad rpn layout %d", who, layout);
  BX_REQUIRE(layout == BX_RPN_PAIRS || (A > 0 && n % A == 0), BX_ERR_INVALID,
             "%s: n (%d) must be a multiple of anchors_per_cell (%d)", who, n, A);
  BX_REQUIRE(bx_aligned(logits, 8), BX_ERR_INVALID, "%s: logits must be 8-byte aligned", who);
  return BX_OK;
}

extern "C" int bx_rpn_scores(bx_handle* h, const float* logits, int layout, int anchors_per_cell, int batch, int n,
                             float* out_scores, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && out_scores, BX_ERR_INVALID, "bx_rpn_scores: NULL argument");
  BX_REQUIRE(batch >= 0 && n >= 0, BX_ERR_INVALID, "bx_rpn_scores: negative size");
  if (int rc = check_rpn("bx_rpn_scores", logits, layout, anchors_per_cell, n)) return rc;
  return rpn_scores_launch(h, logits, layout, anchors_per_cell, batch, n, out_scores, static_cast<cudaStream_t>(stream));
}

static int proposals_impl(bx_handle* h, const float* anchors, const float* deltas, const float* scores,
                          const float* logits, int layout, int A, float* out_scores, int batch, int n,
                          const bx_proposal_params* p, float* out_boxes, int* out_idx, int* out_count, void* stream);

extern "C" int bx_proposals(bx_handle* h, const float* anchors, const float* deltas, const float* scores, int batch,
                            int n, const bx_proposal_params* p, float* out_boxes, int* out_idx, int* out_count,
                            void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(scores, BX_ERR_INVALID, "bx_proposals: NULL argument");
  return proposals_impl(h, anchors, deltas, scores, nullptr, 0, 0, nullptr, batch, n, p, out_boxes, out_idx, out_count,
                        stream);
}

extern "C" int bx_proposals_rpn(bx_handle* h, const float* anchors, const float* deltas, const float* logits,
                                int layout, int anchors_per_cell, int batch, int n, const bx_proposal_params* p,
                                float* out_boxes, int* out_idx, int* out_count, float* out_scores, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(batch >= 0 && n >= 0, BX_ERR_INVALID, "bx_proposals_rpn: negative size");
  if (int rc = check_rpn("bx_proposals_rpn", logits, layout, anchors_per_cell, n)) return rc;
  return proposals_impl(h, anchors, deltas, nullptr, logits, layout, anchors_per_cell, out_scores, batch, n, p, out_boxes,
                        out_idx, out_count, stream);
}

static int proposals_impl(bx_handle* h, const float* anchors, const float* deltas, const float* scores,
                          const float* logits, int layout, int A, float* out_scores, int batch, int n,
                          const bx_proposal_params* p, float* out_boxes, int* out_idx, int* out_count, void* stream) {
  BX_REQUIRE(h && anchors && deltas && p && out_boxes && out_idx && out_count, BX_ERR_INVALID,
             "bx_proposals: NULL argument");
  BX_REQUIRE(batch >= 0 && n >= 0, BX_ERR_INVALID, "bx_proposals: negative size");
  BX_REQUIRE(p->iou_threshold >= 0.0f && p->iou_threshold <= 1.0f, BX_ERR_INVALID,
             "bx_proposals: iou_threshold must be in [0, 1]");
  BX_REQUIRE(p->post_nms >= 0 && p->post_nms <= kMaxPost, BX_ERR_UNSUPPORTED, "bx_proposals: post_nms %d not in [0, %d]",
             p->post_nms, kMaxPost);
  BX_REQUIRE(p->image_h > 0 && p->image_w > 0, BX_ERR_INVALID, "bx_proposals: image shape must be positive");
  BX_REQUIRE(n < (1 << 22), BX_ERR_UNSUPPORTED, "bx_proposals: n must be < 2^22");
  BX_REQUIRE(bx_aligned(anchors, 16) && bx_aligned(deltas, 16) && bx_aligned(out_boxes, 16), BX_ERR_INVALID,
             "bx_proposals: box tensors must be 16-byte aligned");
  if (batch == 0) return BX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProposalArgs a = {};
  a.anchors = reinterpret_cast<const float4*>(anchors);
  a.deltas = reinterpret_cast<const float4*>(deltas);
  a.scores = scores;
  a.n = n;
  a.codec = make_codec(p->means, p->stds, p->image_h, p->image_w);
  a.pre_nms_top_k = p->pre_nms_top_k;
  a.post_nms = p->post_nms;
  a.thr = make_iou_thr(p->iou_threshold);
  a.out_boxes = reinterpret_cast<float4*>(out_boxes);
  a.out_idx = out_idx;
  a.out_count = out_count;
  const size_t total = static_cast<size_t>(batch) * n;
  const bool fuse = logits && n <= kKeyCacheMax && !(p->min_size > 0.0f);   // softmax inside the key pass
  const size_t ws_scores = ((logits && !fuse && !out_scores) ? total * sizeof(float) + 15 : 0) & ~size_t(15);
  const size_t ws_min = ((p->min_size > 0.0f) ? total * (sizeof(float4) + sizeof(uint32_t)) + 15 : 0) & ~size_t(15);
  const size_t ws_top = (n > kKeyCacheMax) ? topset_ws_bytes(batch, kKeyCacheMax) : 0;
  if (int rc = bx_ws_reserve(h, ws_scores + ws_min + ws_top, st)) return rc;
  if (fuse) {
    a.logits = logits;
    a.logit_layout = layout;
    a.anchors_per_cell = A;
    a.out_scores = out_scores;
  } else if (logits) {
    float* sc = out_scores ? out_scores : reinterpret_cast<float*>(h->ws);
    if (int rc = rpn_scores_launch(h, logits, layout, A, batch, n, sc, st)) return rc;
    scores = sc;
    a.scores = sc;
  }
  if (p->min_size > 0.0f) {
    // filter -> top-k -> NMS: decode everything once, mask the keys of undersized boxes
    float4* wboxes = reinterpret_cast<float4*>(static_cast<char*>(h->ws) + ws_scores);
    uint32_t* wkeys = reinterpret_cast<uint32_t*>(wboxes + total);
    const int grid = static_cast<int>(bx_min_ll(bx_div_up((long long)total, 256), 8ll * h->num_sms));
    decode_filter_keys_kernel<<<grid, 256, 0, st>>>(a.anchors, a.deltas, scores, batch, n, a.codec, p->min_size,
                                                    wboxes, wkeys);
    BX_LAUNCH_CHECK(h);
    a.boxes = wboxes;
    a.keys = wkeys;
  }
  return launch_proposals(h, a, batch, st, ws_top ? static_cast<char*>(h->ws) + ws_scores + ws_min : nullptr);
}
