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
  int chunk;      // rois staged per pass (x-axis table + work list capacity)
  int nbox;       // TMA boxes per band
  int scan_all;   // 1: rois of any image may be anywhere (box_ind given) -> scan every roi
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

struct XParam {        // one crop sample column
  uint32_t packed;     // lo | hi << 12 | valid << 24
  float lerp;
};

// work entry: word 0 = roi_local | py << 12 | zero_fill << 20 ; then per sample row s: (top | bot << 12 | valid << 24), ly
template <int S>
struct WorkEntry {
  uint32_t head;
  uint32_t y[S];
  float ly[S];
};

template <int POOL>
__global__ void __launch_bounds__(kThreads, 1) roi_band_kernel(const __grid_constant__ CUtensorMap tmap, const BandArgs a) {
  constexpr int S = (POOL == BX_POOL_NONE) ? 1 : 2;
  using Entry = WorkEntry<S>;
  extern __shared__ __align__(128) unsigned char smem[];
  float* band = reinterpret_cast<float*>(smem);
  XParam* xtab = reinterpret_cast<XParam*>(smem + static_cast<size_t>(a.nbox) * kBoxBytes);
  Entry* work = reinterpret_cast<Entry*>(xtab + a.chunk * a.r.Q);
  __shared__ uint64_t mbar;
  __shared__ int n_work;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const RoiArgs& r = a.r;
  const int P = r.P, Q = r.Q, C = r.c;
  int u = blockIdx.x;
  const int band_i = u % a.n_bands; u /= a.n_bands;
  const int slice = u % a.n_slices;
  const int img = u / a.n_slices;
  const int fh = r.lv[0].fh, fw = r.lv[0].fw;
  const int r0 = band_i * a.rows_per_band;
  const int r1 = min(fh, r0 + a.rows_per_band);          // owned sample rows: top in [r0, r1)
  const int rows_loaded = min(r1 + 1, fh) - r0;            // + one halo row for the bottom taps
  const float* feat_img = r.lv[0].feat + static_cast<size_t>(img) * fh * fw * C;

  if (tid == 0) {
    mbar_init(&mbar, 1);
    mbar_expect_tx(&mbar, static_cast<uint32_t>(a.nbox) * kBoxBytes);
    const int row0 = (img * fh + r0) * fw;
    for (int bx = 0; bx < a.nbox; ++bx)
      tma_load_2d(smem + static_cast<size_t>(bx) * kBoxBytes, &tmap, slice * kSlice, row0 + bx * kBoxRows, &mbar);
  }

  // roi range this CTA has to look at
  int g_begin = 0, g_end = r.r;
  if (!a.scan_all && r.roi_counts) {
    g_begin = img * r.rois_per_image;
    g_end = g_begin + r.rois_per_image;
  }
  bool band_ready = false;

  for (int c0 = g_begin; c0 < g_end; c0 += a.chunk) {
    const int nroi = min(a.chunk, g_end - c0);
    if (tid == 0) n_work = 0;
    __syncthreads();
    // ---- phase 1a: x-axis table for every (roi, sample column)
    for (int idx = tid; idx < nroi * Q; idx += kThreads) {
      const int l = idx / Q, s = idx % Q;
      const NormBox nb = roi_norm_box(r, r.rois[c0 + l], fh, fw);
      const Axis ax = sample_axis(nb.x1, nb.x2, s, Q, nb.dimx, nb.pad);
      XParam xp;
      xp.packed = static_cast<uint32_t>(ax.lo) | (static_cast<uint32_t>(ax.hi) << 12) | (ax.valid ? (1u << 24) : 0u);
      xp.lerp = ax.lerp;
      xtab[idx] = xp;
    }
    // ---- phase 1b: output rows (roi, py) owned by this band
    for (int idx = tid; idx < nroi * P; idx += kThreads) {
      const int l = idx / P, py = idx % P;
      const int g = c0 + l;
      int rimg = 0, zero = 0;
      if (r.roi_counts) {
        rimg = g / r.rois_per_image;
        zero = (g % r.rois_per_image) >= r.roi_counts[rimg];
      } else if (r.box_ind) {
        rimg = r.box_ind[g];
      }
      if (rimg != img) {
        // a roi of another image; rois whose box_ind is out of range are zero-filled by the CTAs of image 0, band 0
        if (!(r.box_ind && (rimg < 0 || rimg >= r.b) && img == 0)) continue;
        zero = 1;
      }
      Entry e;
      e.head = static_cast<uint32_t>(l) | (static_cast<uint32_t>(py) << 12) | (zero ? (1u << 20) : 0u);
      int owner_row = -1;
      const NormBox nb = roi_norm_box(r, r.rois[g], fh, fw);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const Axis ay = sample_axis(nb.y1, nb.y2, py * S + s, Q, nb.dimy, nb.pad);
        const bool v = ay.valid && !zero;
        e.y[s] = static_cast<uint32_t>(ay.lo) | (static_cast<uint32_t>(ay.hi) << 12) | (v ? (1u << 24) : 0u);
        e.ly[s] = ay.lerp;
        if (v && owner_row < 0) owner_row = ay.lo;
      }
      const int owner = owner_row < 0 ? 0 : min(owner_row / a.rows_per_band, a.n_bands - 1);
      if (owner == band_i) work[atomicAdd(&n_work, 1)] = e;
    }
    __syncthreads();
    if (!band_ready) {
      mbar_wait(&mbar, 0);
      band_ready = true;
    }
    // ---- phase 2: one warp per output row; 8 lanes (32 channels) per pixel, 4 pixels per pass
    const int nw = n_work;
    const int q = lane & 7, sub = lane >> 3;
    for (int w = warp; w < nw; w += kThreads / 32) {
      const Entry e = work[w];
      const int l = e.head & 0xFFF, py = (e.head >> 12) & 0xFF;
      const bool zero = (e.head >> 20) & 1u;
      float* out_row = r.out + ((static_cast<size_t>(c0 + l) * P + py) * P) * C + slice * kSlice + q * 4;
      for (int px = sub; px < P; px += 4) {
        float acc[4];
#pragma unroll
        for (int sy = 0; sy < S; ++sy) {
          const uint32_t yp = e.y[sy];
          const int top = yp & 0xFFF, bot = (yp >> 12) & 0xFFF;
          const bool yv = (yp >> 24) & 1u;
          const float ly = e.ly[sy];
          // rows of this sample: in the staged band, or (second sample row of a pooled pair only) beyond it
          const bool in_band = (S == 1) || (top >= r0 && bot < r0 + rows_loaded);
#pragma unroll
          for (int sx = 0; sx < S; ++sx) {
            const XParam xp = xtab[l * Q + px * S + sx];
            const int lo = xp.packed & 0xFFF, hi = (xp.packed >> 12) & 0xFFF;
            const bool xv = (xp.packed >> 24) & 1u;
            float val[4];
            if (yv && xv) {
              float4 tl, tr, bl, br;
              if (in_band) {
                const float* bt = band + (static_cast<size_t>(top - r0) * fw) * kSlice + q * 4;
                const float* bb = band + (static_cast<size_t>(bot - r0) * fw) * kSlice + q * 4;
                tl = *reinterpret_cast<const float4*>(bt + lo * kSlice);
                tr = *reinterpret_cast<const float4*>(bt + hi * kSlice);
                bl = *reinterpret_cast<const float4*>(bb + lo * kSlice);
                br = *reinterpret_cast<const float4*>(bb + hi * kSlice);
              } else {
                const float* gt = feat_img + (static_cast<size_t>(top) * fw) * C + slice * kSlice + q * 4;
                const float* gb = feat_img + (static_cast<size_t>(bot) * fw) * C + slice * kSlice + q * 4;
                tl = __ldg(reinterpret_cast<const float4*>(gt + static_cast<size_t>(lo) * C));
                tr = __ldg(reinterpret_cast<const float4*>(gt + static_cast<size_t>(hi) * C));
                bl = __ldg(reinterpret_cast<const float4*>(gb + static_cast<size_t>(lo) * C));
                br = __ldg(reinterpret_cast<const float4*>(gb + static_cast<size_t>(hi) * C));
              }
              const float lx = xp.lerp;
              const float t0 = tl.x + (tr.x - tl.x) * lx, b0 = bl.x + (br.x - bl.x) * lx;
              const float t1 = tl.y + (tr.y - tl.y) * lx, b1 = bl.y + (br.y - bl.y) * lx;
              const float t2 = tl.z + (tr.z - tl.z) * lx, b2 = bl.z + (br.z - bl.z) * lx;
              const float t3 = tl.w + (tr.w - tl.w) * lx, b3 = bl.w + (br.w - bl.w) * lx;
              val[0] = t0 + (b0 - t0) * ly;
              val[1] = t1 + (b1 - t1) * ly;
              val[2] = t2 + (b2 - t2) * ly;
              val[3] = t3 + (b3 - t3) * ly;
            } else {
              const float ev = zero ? 0.0f : r.extrapolation;
              val[0] = val[1] = val[2] = val[3] = ev;
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              if (sy == 0 && sx == 0) acc[v] = val[v];
              else if (POOL == BX_POOL_MAX2) acc[v] = fmaxf(acc[v], val[v]);
              else acc[v] = acc[v] + val[v];
            }
          }
        }
        float4 o;
        if (POOL == BX_POOL_AVG2) o = make_float4(acc[0] / 4.0f, acc[1] / 4.0f, acc[2] / 4.0f, acc[3] / 4.0f);
        else o = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(out_row + static_cast<size_t>(px) * C) = o;
      }
    }
    __syncthreads();
  }
  if (!band_ready) mbar_wait(&mbar, 0);  // never leave with a TMA still in flight to this CTA's shared memory
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
int launch_band(bx_handle* h, const BandArgs& a, const CUtensorMap& tmap, size_t smem, int grid, cudaStream_t st) {
  BX_CUDA(cudaFuncSetAttribute(roi_band_kernel<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  roi_band_kernel<POOL><<<grid, kThreads, smem, st>>>(tmap, a);
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
  const int S = (pool == BX_POOL_NONE) ? 1 : 2;
  const size_t entry = sizeof(uint32_t) * (1 + 2 * S);
  const size_t budget = h->smem_optin - 1024;
  // rois per staging pass: everything of one image when that fits in ~1/5 of shared memory
  const int rois_img = ra.roi_counts ? ra.rois_per_image : ra.r;
  int chunk = rois_img < 4096 ? rois_img : 4096;
  const size_t per_roi = static_cast<size_t>(ra.Q) * sizeof(XParam) + static_cast<size_t>(ra.P) * entry;
  const size_t tab_cap = budget / 5;
  if (chunk * per_roi > tab_cap) chunk = static_cast<int>(tab_cap / per_roi);
  if (chunk < 32) return BX_OK;
  const size_t tab = ((chunk * per_roi + 127) / 128) * 128;
  // band height: as many rows (+1 halo) as fit, then balanced over the bands
  const size_t row_bytes = static_cast<size_t>(fw) * kSlice * 4;
  int max_rows_loaded = static_cast<int>(((budget - tab) / kBoxBytes) * kBoxBytes / row_bytes);
  if (max_rows_loaded < 3) return BX_OK;                          // map too wide for a useful band: direct kernel
  int rows_per_band = max_rows_loaded - 1;
  int n_bands = (fh + rows_per_band - 1) / rows_per_band;
  rows_per_band = (fh + n_bands - 1) / n_bands;
  const int rows_loaded = (rows_per_band + 1 < fh) ? rows_per_band + 1 : fh;
  const int nbox = static_cast<int>((static_cast<size_t>(rows_loaded) * fw + kBoxRows - 1) / kBoxRows);
  const size_t smem = static_cast<size_t>(nbox) * kBoxBytes + tab;
  if (smem > budget) return BX_OK;

  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ra.c), static_cast<cuuint64_t>(ra.b) * fh * fw};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ra.c) * sizeof(float)};
  const cuuint32_t box[2] = {kSlice, kBoxRows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ra.lv[0].feat), gdim, gstride,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BX_REQUIRE(cr == CUDA_SUCCESS, BX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);

  BandArgs a;
  a.r = ra;
  a.rows_per_band = rows_per_band;
  a.n_bands = n_bands;
  a.n_slices = ra.c / kSlice;
  a.chunk = chunk;
  a.nbox = nbox;
  a.scan_all = (ra.box_ind != nullptr && !ra.roi_counts) ? 1 : 0;
  const int grid = ra.b * a.n_slices * n_bands;
  int rc;
  if (pool == BX_POOL_NONE) rc = launch_band<BX_POOL_NONE>(h, a, tmap, smem, grid, st);
  else if (pool == BX_POOL_MAX2) rc = launch_band<BX_POOL_MAX2>(h, a, tmap, smem, grid, st);
  else rc = launch_band<BX_POOL_AVG2>(h, a, tmap, smem, grid, st);
  if (rc == BX_OK) *used = 1;
  return rc;
}

}  // namespace bxroi
