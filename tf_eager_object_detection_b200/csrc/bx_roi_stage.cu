// RoI-stationary pooled extractor with the roi's feature footprint staged in shared memory by TMA (sm_100a).
//
// Replaces the same reference code as roi_pool2_kernel (bx_roi.cu): model/roi_pooling.py:15-42 RoiPoolingCropAndResize2
// under model/fpn/base_fpn_model.py:152-161 (_get_roi_features, level routing through `order` / `level`), and the
// pooled C4 / RoIAlign extractors (:45-90, :93-176) when their channel count allows it.
//
// Why: roi_pool2_kernel reads its 4 - 16 bilinear taps per output pixel straight from global memory; ncu shows it bound
// by the latency of those loads (long-scoreboard stalls, 35 % of the warp slots occupied at 80 registers, 52 % issue
// utilisation) although neither DRAM (footprints are compulsory misses: traffic = algorithmic) nor the L2 nor the LSU
// pipe is saturated.  Here one CTA owns (roi, 64-channel slice): it derives the roi's sample geometry, has TMA
// (cp.async.bulk.tensor.2d over the [B*h*w, C] view of the roi's level) copy the feature rows the roi touches into
// shared memory — 16-pixel x 64-channel boxes, one mbarrier — and then takes every tap from shared memory.  The copy of
// one CTA overlaps the arithmetic of the 3 - 4 other CTAs resident on the SM, taps arrive with shared-memory latency,
// and the tap registers no longer have to cover a DRAM round trip.  A roi whose footprint does not fit the per-CTA
// budget is processed in strips of pooled rows; a single pooled row that does not fit (very tall or wide rois clamped
// to P2) falls back to global taps inside the same kernel.  Arithmetic, op order and therefore bits are those of
// roi_pool2_kernel / the oracle.
#include <cuda.h>
#include <stdlib.h>

#include "bx_roi.cuh"

namespace bxroi {
namespace {

constexpr int kSliceCh = 64;                 // channels per CTA: 256 B per pixel, 16 lanes x float4
constexpr int kLanesPerPix = kSliceCh / 4;   // 16
constexpr int kBoxPx = 16;                   // pixels per TMA box (4 KB)
constexpr int kPixBytes = kSliceCh * 4;

struct StageMaps {
  CUtensorMap m[kMaxLevels];
};

struct StageArgs {
  RoiArgs r;
  int n_slices;
  int budget;        // bytes of shared memory for staged rows
  float neg_zero;
};

struct AxisEnt {     // one crop sample along one axis
  int lo, hi;        // tap indices in the (un-padded) map
  float lerp;
  int valid;
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void st_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void st_mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(s_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void st_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(s_u32(bar))
      : "memory");
}

// x-interpolation of one feature row at one sample column: left + (right - left) * lx (TF's `top` / `bottom`)
__device__ __forceinline__ ulonglong2 lerp_x(const ulonglong2 l, const ulonglong2 r, unsigned long long w, unsigned long long nz) {
  ulonglong2 o;
  o.x = f2_add(l.x, f2_mul(f2_sub(r.x, l.x), w, nz));
  o.y = f2_add(l.y, f2_mul(f2_sub(r.y, l.y), w, nz));
  return o;
}

// Pixel-column walker.  Thread = (pooled pixel column px, 4 channels): it owns BOTH sample columns of its pooled pixels,
// so the 2 x 2 window is folded in registers in the reference's order (s00, s01, s10, s11), and it advances row by row
// through the staged feature rows: each row is x-interpolated ONCE per sample column, the previous row's values stay in
// registers, and a sample row is finished when its bottom row has been reached.  TF computes `top` and `bottom` per
// sample, but a row that is the bottom of one sample and the top of the next yields the same number from the same
// operands, so the reuse is bit-identical; versus the per-pixel form (4 samples x 4 taps per output) this needs about
// half the instructions and shared-memory wavefronts.  kMono: the valid y samples of the strip are non-decreasing in
// their tap rows (always the case for y2 >= y1; checked by the producer); otherwise every sample interpolates its two
// rows afresh.
template <int POOL, bool kMono, typename Ld>
__device__ __forceinline__ void walk_strip_cols(const Ld& ld, const AxisEnt* __restrict__ ax_y, const AxisEnt x0, const AxisEnt x1,
                                                int py0, int py1, int P, int px, float ext, unsigned long long nz,
                                                float4* __restrict__ out, int c4) {
  const unsigned long long w0 = f2_splat(x0.lerp), w1 = f2_splat(x1.lerp);
  const ulonglong2 extv = make_ulonglong2(f2_splat(ext), f2_splat(ext));
  const bool b_same = (x1.lo == x0.lo) && (x1.hi == x0.hi);       // second sample column between the same two pixels
  const bool b_shared = (x1.lo == x0.hi);                         // ... or starting at the first column's right pixel
  auto row_lerp = [&](int row, ulonglong2& va, ulonglong2& vb) {
    if (x0.valid) {
      const ulonglong2 l = ld(row, x0.lo), rr = ld(row, x0.hi);
      va = lerp_x(l, rr, w0, nz);
      if (x1.valid) {
        const ulonglong2 l1 = b_same ? l : (b_shared ? rr : ld(row, x1.lo));
        const ulonglong2 r1 = b_same ? rr : ld(row, x1.hi);
        vb = lerp_x(l1, r1, w1, nz);
      }
    } else if (x1.valid) {
      vb = lerp_x(ld(row, x1.lo), ld(row, x1.hi), w1, nz);
    }
  };
  int cur_row = -(1 << 30);
  ulonglong2 curA = extv, curB = extv, prevA = extv, prevB = extv;
  ulonglong2 v00 = extv, v01 = extv;
  for (int s = 2 * py0; s < 2 * py1; ++s) {
    const AxisEnt ye = ax_y[s];
    ulonglong2 sa = extv, sb = extv;
    if (ye.valid) {
      bool top_is_cur;
      if (kMono) {
        if (cur_row < ye.lo - 1) cur_row = ye.lo - 1;             // rows nobody samples are skipped
        while (cur_row < ye.hi) {
          ++cur_row;
          prevA = curA; prevB = curB;
          row_lerp(cur_row, curA, curB);
        }
        top_is_cur = (ye.lo == cur_row);
      } else {
        row_lerp(ye.lo, prevA, prevB);
        row_lerp(ye.hi, curA, curB);
        top_is_cur = false;
      }
      const unsigned long long wy = f2_splat(ye.lerp);
      if (x0.valid) {
        const ulonglong2 top = top_is_cur ? curA : prevA;
        sa.x = f2_add(top.x, f2_mul(f2_sub(curA.x, top.x), wy, nz));
        sa.y = f2_add(top.y, f2_mul(f2_sub(curA.y, top.y), wy, nz));
      }
      if (x1.valid) {
        const ulonglong2 top = top_is_cur ? curB : prevB;
        sb.x = f2_add(top.x, f2_mul(f2_sub(curB.x, top.x), wy, nz));
        sb.y = f2_add(top.y, f2_mul(f2_sub(curB.y, top.y), wy, nz));
      }
    }
    if ((s & 1) == 0) {
      v00 = sa; v01 = sb;
    } else {
      const ulonglong2 res = pool2<POOL>(pool2<POOL>(pool2<POOL>(v00, v01), sa), sb);      // s00, s01, s10, s11
      float4 o = *reinterpret_cast<const float4*>(&res);
      if (POOL == BX_POOL_AVG2) o = make_float4(o.x / 4.0f, o.y / 4.0f, o.z / 4.0f, o.w / 4.0f);
      out[static_cast<size_t>((s >> 1) * P + px) * c4] = o;
    }
  }
}

// ---- persistent, warp-specialised kernel -------------------------------------------------------------------------
// One CTA per SM loops over rois (grid-stride).  Warp 0 is the PRODUCER: it derives the sample geometry of a roi (lanes
// 0..Q-1 the x samples, Q..2Q-1 the y samples), cuts the roi into strips of pooled rows whose feature rows fit a stage,
// and for every (strip, 64-channel slice) UNIT fills the next stage of a ring: sample tables + descriptor written to
// shared memory, the feature rows copied by TMA, completion signalled on the stage's `full` mbarrier.  The metadata of
// the next roi (order -> level / box_ind / roi: three dependent global loads) is fetched one roi ahead.  The other warps
// form G CONSUMER groups of Q*16 threads (one per (sample column, 4 channels)); group g takes the units u = g (mod G),
// waits for `full`, runs the column walker on the staged rows and releases the stage on its `empty` mbarrier.  Copy,
// geometry and arithmetic of different units overlap; nothing on the consumers' path touches global memory except the
// output stores (and the taps of units too large for a stage, which are flagged `staged = 0`).
#ifndef BX_STAGE_GROUPS
#define BX_STAGE_GROUPS 2
#endif
#ifndef BX_STAGE_STAGES
#define BX_STAGE_STAGES 4
#endif
constexpr int kGroups = BX_STAGE_GROUPS;          // consumer groups (each computes one unit at a time)
constexpr int kStages = BX_STAGE_STAGES;          // ring depth; a multiple of kGroups so a stage always meets the same group
static_assert(kStages % kGroups == 0, "stage -> group mapping must be fixed");
constexpr int kGroupThreads = 128;                // 8 pixel-column slots x 16 lanes
constexpr int kSlots = kGroupThreads / kLanesPerPix;
constexpr int kStageMaxThreads = 32 + kGroups * kGroupThreads;

struct UnitDesc {          // 48 bytes, written by producer lane 0
  int kind;                // 0 = compute, 1 = zero-fill (padded roi), 2 = done
  int j, slice;            // output roi row, channel slice
  int py0, py1;            // pooled rows of the strip
  int r_lo, x_min;         // first staged feature row / column
  int row_pitch;           // bytes per staged row
  int staged;              // 1: taps from the stage, 0: from global memory
  int img, lvl, mono;      // mono: valid y samples non-decreasing in their tap rows
};

struct StageCtl {
  UnitDesc d;
  AxisEnt ax_x[kMaxQ];
  AxisEnt ax_y[kMaxQ];
};

template <int POOL>
__global__ void __launch_bounds__(kStageMaxThreads, 1)
roi_stage_kernel(const __grid_constant__ StageMaps maps, const StageArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ StageCtl ctl[kStages];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  const RoiArgs& r = a.r;
  const int P = r.P, Q = r.Q, C = r.c;
  const int c4 = C >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warps_per_group = kGroupThreads / 32;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      st_mbar_init(&full_bar[s], 1);
      st_mbar_init(&empty_bar[s], warps_per_group);
    }
  }
  __syncthreads();

  if (warp == 0) {
    // ================================================================= producer
    const int grid = gridDim.x;
    auto load_src = [&](int j) { return (j < r.r) ? (r.order ? r.order[j] : j) : -1; };
    struct Meta { int src, lvl, img, zero; float4 roi; };
    auto load_meta = [&](int src) {
      Meta m;
      m.src = src; m.lvl = 0; m.img = 0; m.zero = 0; m.roi = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src < 0) return m;
      m.lvl = r.level ? r.level[src] - r.level_base : 0;
      m.img = r.box_ind ? r.box_ind[src] : 0;
      if (r.roi_counts) {
        m.img = src / r.rois_per_image;
        m.zero = (src % r.rois_per_image) >= r.roi_counts[m.img];
      }
      m.zero = m.zero || m.img < 0 || m.img >= r.b;
      m.roi = r.rois[src];
      return m;
    };
    int j = blockIdx.x;
    int src1 = load_src(j + grid);
    Meta m0 = load_meta(load_src(j));
    uint32_t unit = 0;                       // units issued so far: stage = unit % kStages, use count = unit / kStages
    auto acquire_stage = [&](uint32_t u) {
      const uint32_t st = u % kStages, n = u / kStages;
      st_mbar_wait(&empty_bar[st], (n & 1u) ^ 1u);      // first use of a stage passes at once
      return st;
    };
    for (; j < r.r; j += grid) {
      const int src2 = load_src(j + 2 * grid);
      const Meta m1 = load_meta(src1);
      // ---- geometry of roi j (lane < Q: x sample `lane`; Q <= lane < 2Q: y sample `lane - Q`)
      AxisEnt e;
      e.lo = 0; e.hi = 0; e.lerp = 0.f; e.valid = 0;
      const int fh = r.lv[m0.lvl].fh, fw = r.lv[m0.lvl].fw;
      if (!m0.zero) {
        const NormBox nb = roi_norm_box(r, m0.roi, fh, fw);
        const bool is_x = lane < Q;
        const int sidx = is_x ? lane : lane - Q;
        if (lane < 2 * Q) {
          const Axis v = is_x ? sample_axis(nb.x1, nb.x2, sidx, Q, nb.dimx, nb.pad) : sample_axis(nb.y1, nb.y2, sidx, Q, nb.dimy, nb.pad);
          e.lo = v.lo; e.hi = v.hi; e.lerp = v.lerp; e.valid = v.valid;
        }
      }
      // monotone y taps? (lane Q+s holds y sample s): every valid sample's rows >= those of the valid sample before it
      int mono = 1;
      {
        int last_lo = -(1 << 30), last_hi = -(1 << 30);
        for (int sy = 0; sy < Q; ++sy) {
          const int v_lo = __shfl_sync(0xFFFFFFFFu, e.lo, Q + sy), v_hi = __shfl_sync(0xFFFFFFFFu, e.hi, Q + sy);
          const int v_ok = __shfl_sync(0xFFFFFFFFu, e.valid, Q + sy);
          if (v_ok) {
            if (v_lo < last_lo || v_hi < last_hi) mono = 0;
            last_lo = v_lo; last_hi = v_hi;
          }
        }
      }
      const bool x_ok = (lane < Q) && e.valid;
      const int x_min = __reduce_min_sync(0xFFFFFFFFu, x_ok ? e.lo : (1 << 30));
      const int x_max = __reduce_max_sync(0xFFFFFFFFu, x_ok ? e.hi : -1);
      const int n_chunks = (x_max >= x_min) ? (x_max - x_min + kBoxPx) / kBoxPx : 0;
      const int row_pitch = n_chunks * kBoxPx * kPixBytes;
      const int rows_cap = row_pitch > 0 ? a.budget / row_pitch : 0;
      int py0 = 0;
      while (py0 < P) {
        // ---- next strip [py0, py1): rows touched by its valid y samples must fit a stage
        int r_lo = 1 << 30, r_hi = -1, py1 = py0, staged = n_chunks > 0 ? 1 : 0, kind = 0;
        if (m0.zero) {
          kind = 1; py1 = P; staged = 0;
        } else {
          for (; py1 < P; ++py1) {
            int lo = r_lo, hi = r_hi;
            for (int k = 0; k < 2; ++k) {
              const int from = Q + 2 * py1 + k;
              const int v_lo = __shfl_sync(0xFFFFFFFFu, e.lo, from), v_hi = __shfl_sync(0xFFFFFFFFu, e.hi, from);
              const int v_ok = __shfl_sync(0xFFFFFFFFu, e.valid, from);
              if (v_ok) { lo = min(lo, v_lo); hi = max(hi, v_hi); }
            }
            if (hi >= lo && hi - lo + 1 > rows_cap) break;
            r_lo = lo; r_hi = hi;
          }
          if (py1 == py0) { staged = 0; py1 = py0 + 1; }
        }
        const int n_rows = (staged && r_hi >= r_lo) ? r_hi - r_lo + 1 : 0;
        const int pix0 = (m0.img * fh + r_lo) * fw + x_min;
        for (int slice = 0; slice < a.n_slices; ++slice) {
          const uint32_t st = acquire_stage(unit);
          StageCtl& sc = ctl[st];
          if (lane < Q) sc.ax_x[lane] = e; else if (lane < 2 * Q) sc.ax_y[lane - Q] = e;
          if (lane == 0) {
            UnitDesc d;
            d.kind = kind; d.j = j; d.slice = slice; d.py0 = py0; d.py1 = py1; d.r_lo = r_lo; d.x_min = x_min;
            d.row_pitch = row_pitch; d.staged = staged; d.img = m0.img; d.lvl = m0.lvl; d.mono = mono;
            sc.d = d;
          }
          __syncwarp();
          if (lane == 0) {
            if (n_rows > 0) st_mbar_expect(&full_bar[st], static_cast<uint32_t>(n_rows) * static_cast<uint32_t>(row_pitch));
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&full_bar[st])) : "memory");
          }
          __syncwarp();
          unsigned char* dst = smem_raw + static_cast<size_t>(st) * a.budget;
          for (int q = lane; q < n_rows * n_chunks; q += 32) {
            const int rr = q / n_chunks, ch = q - rr * n_chunks;
            st_tma_2d(dst + static_cast<size_t>(rr) * row_pitch + ch * (kBoxPx * kPixBytes), &maps.m[m0.lvl], slice * kSliceCh,
                      pix0 + rr * fw + ch * kBoxPx, &full_bar[st]);
          }
          ++unit;
        }
        py0 = py1;
      }
      m0 = m1;
      src1 = src2;
    }
    // ---- one `done` unit per consumer group
    for (int g = 0; g < kGroups; ++g) {
      const uint32_t st = acquire_stage(unit);
      if (lane == 0) {
        ctl[st].d.kind = 2;
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&full_bar[st])) : "memory");
      }
      __syncwarp();
      ++unit;
    }
    return;
  }

  // =================================================================== consumers
  const int ct = tid - 32;
  const int group = ct / kGroupThreads;
  const int gt = ct - group * kGroupThreads;          // thread within the group
  const int cg = gt & (kLanesPerPix - 1), slot = gt / kLanesPerPix;
  const unsigned long long nz = f2_splat(a.neg_zero);
  const float ext = r.extrapolation;
  for (uint32_t unit = group;; unit += kGroups) {
    const uint32_t st = unit % kStages, n = unit / kStages;
    st_mbar_wait(&full_bar[st], n & 1u);
    const StageCtl& sc = ctl[st];
    const UnitDesc d = sc.d;
    if (d.kind == 2) break;
    float4* out = reinterpret_cast<float4*>(r.out) + (static_cast<size_t>(d.j) * P * P * C + d.slice * kSliceCh) / 4 + cg;
    if (d.kind == 1) {
      for (int pix = slot; pix < P * P; pix += kSlots) out[static_cast<size_t>(pix) * c4] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int px = slot; px < P; px += kSlots) {
        const AxisEnt x0 = sc.ax_x[2 * px], x1 = sc.ax_x[2 * px + 1];
        if (d.staged) {
          const unsigned char* sbase = smem_raw + static_cast<size_t>(st) * a.budget + cg * 16;
          const int r_lo = d.r_lo, x_min = d.x_min, row_pitch = d.row_pitch;
          auto ld = [&](int row, int col) {
            return *reinterpret_cast<const ulonglong2*>(sbase + (row - r_lo) * row_pitch + (col - x_min) * kPixBytes);
          };
          if (d.mono) walk_strip_cols<POOL, true>(ld, sc.ax_y, x0, x1, d.py0, d.py1, P, px, ext, nz, out, c4);
          else walk_strip_cols<POOL, false>(ld, sc.ax_y, x0, x1, d.py0, d.py1, P, px, ext, nz, out, c4);
        } else {
          const LevelFeat lf = r.lv[d.lvl];
          const ulonglong2* gfeat = reinterpret_cast<const ulonglong2*>(lf.feat) + static_cast<size_t>(d.img) * lf.fh * lf.fw * c4 +
                                    d.slice * kLanesPerPix + cg;
          const int fw = lf.fw;
          auto ld = [&](int row, int col) { return __ldg(gfeat + (static_cast<size_t>(row) * fw + col) * c4); };
          walk_strip_cols<POOL, false>(ld, sc.ax_y, x0, x1, d.py0, d.py1, P, px, ext, nz, out, c4);
        }
      }
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&empty_bar[st])) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn stage_encode_fn() {
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

}  // namespace

// returns BX_OK and sets *used = 1 when the staged kernel handled the launch (pooled crops only)
int roi_stage_launch(bx_handle* h, const RoiArgs& ra, int pool, cudaStream_t st, int* used) {
  *used = 0;
  // Opt-in (BX_ROI_STAGE=1): measured SLOWER than roi_pool2_kernel on the FPN workloads (cfg3 B=16: 1.2 - 1.4 ms vs 0.55 ms,
  // profiles/README.md "TMA-staged FPN extractor"): staging whole footprints (median 44 KB per roi and 64-channel slice)
  // leaves room for 4 units and 8 consumer warps per SM, too few to keep the issue slots busy.  Kept as the measured
  // north_star variant and as a second implementation the parity tests compare.
  const char* env_s = getenv("BX_ROI_STAGE");                                             // read per call: tests toggle it
  const int env = env_s ? atoi(env_s) : 0;
  if (!env || pool == BX_POOL_NONE) return BX_OK;
  if (ra.c % kSliceCh != 0 || !bx_aligned(ra.out, 16) || ra.Q != 2 * ra.P || ra.Q > kMaxQ) return BX_OK;
  EncodeTiledFn enc = stage_encode_fn();
  if (!enc) return BX_OK;
  StageMaps maps;
  for (int l = 0; l < ra.n_levels; ++l) {
    const LevelFeat& lf = ra.lv[l];
    if (!bx_aligned(lf.feat, 16)) return BX_OK;
    const unsigned long long npix = static_cast<unsigned long long>(ra.b) * lf.fh * lf.fw;
    if (npix >= (1ull << 31)) return BX_OK;
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ra.c), static_cast<cuuint64_t>(npix)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ra.c) * sizeof(float)};
    const cuuint32_t box[2] = {kSliceCh, kBoxPx};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(lf.feat), gdim, gstride, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BX_REQUIRE(cr == CUDA_SUCCESS, BX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
  }
  for (int l = ra.n_levels; l < kMaxLevels; ++l) maps.m[l] = maps.m[0];
  static const int budget_env = getenv("BX_ROI_STAGE_KB") ? atoi(getenv("BX_ROI_STAGE_KB")) : 0;
  if (2 * ra.Q > 32) return BX_OK;                                         // the producer warp holds the 2Q samples in its lanes
  const int threads = kStageMaxThreads;
  StageArgs a;
  a.r = ra;
  a.n_slices = ra.c / kSliceCh;
  a.budget = (budget_env > 0 ? budget_env : 208 / kStages) * 1024;   // bytes per stage; kStages stages per CTA, one CTA per SM
  a.neg_zero = -0.0f;
  const size_t smem = static_cast<size_t>(a.budget) * kStages;
  if (smem + 8 * 1024 > h->smem_optin) return BX_OK;
  int grid = h->num_sms;
  if (grid > ra.r) grid = ra.r;
  if (grid <= 0) return BX_OK;
  if (pool == BX_POOL_MAX2) {
    BX_CUDA(cudaFuncSetAttribute(roi_stage_kernel<BX_POOL_MAX2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_stage_kernel<BX_POOL_MAX2><<<grid, threads, smem, st>>>(maps, a);
  } else {
    BX_CUDA(cudaFuncSetAttribute(roi_stage_kernel<BX_POOL_AVG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_stage_kernel<BX_POOL_AVG2><<<grid, threads, smem, st>>>(maps, a);
  }
  BX_LAUNCH_CHECK(h);
  *used = 1;
  return BX_OK;
}

}  // namespace bxroi
