// Shared host/device helpers for libboxpath (sm_100a).  Compiled with -fmad=false: every fp32 expression on the
// path is evaluated in the reference's op order with IEEE add/mul/div, no FMA contraction (DESIGN.md "Numerics").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/boxpath.h"

#define BX_NUM_SMS_B200 148

struct bx_handle {
  int device;
  int num_sms;
  size_t smem_optin;     // max dynamic shared memory per block
  size_t smem_sm;        // shared memory per SM
  void* ws;              // device workspace (grown on demand, stream-ordered by the caller's stream)
  size_t ws_bytes;
  void* stage;           // device staging area of the *_host entry points
  size_t stage_bytes;
  void* plan;            // RoI-pooling plan blocks (bx_roi_band.cu)
  size_t plan_bytes;
  unsigned long long* dbg_ptr;   // BX_BAND_DEBUG: device timestamps of the last band launch
  long long dbg_count;
  int dbg_info[4];
  long long launches;
  long long band_launches;       // RoI launches served by the TMA band kernel ...
  long long band_fallbacks;      // ... and plain-crop launches that fell back to a gather kernel (bx_stats)
  cudaStream_t last_stream;      // stream of the handle's latest call (orders workspace frees, bx_destroy)
  int last_stream_valid;
  int deterministic;             // bx_set_deterministic: bit-reproducible kernels where the default is not (bx_roi_pool_grad)
  // optional event bracketing of the RoI-pooling kernel (bx_profile_roi)
  cudaEvent_t* prof_ev;   // 2 * prof_cap events
  int prof_cap;
  int prof_n;
  int prof_on;
};

void bx_set_error(const char* fmt, ...);
// ensure the handle's workspace / staging area / plan area has >= bytes (stream-ordered growth, never a device sync)
int bx_ws_reserve(bx_handle* h, size_t bytes, cudaStream_t st);
int bx_stage_reserve(bx_handle* h, size_t bytes, cudaStream_t st);
int bx_plan_reserve(bx_handle* h, size_t bytes, cudaStream_t st);

// First statement of every entry point that takes a handle: makes the handle's device current for the duration of the
// call (kernels, memsets and allocations land on h->device whatever the caller's current device is) and restores the
// caller's device on return; remembers the stream for stream-ordered workspace retirement.  NULL handle: no-op.
struct BxEnter {
  // `fn` = the entry point's name (the default argument is evaluated at the call site): with BX_NVTX=1 in the environment
  // the call is bracketed by an NVTX range of that name, so ncu --nvtx --nvtx-include "bx_roi_pool/" (or an nsys timeline)
  // attributes kernels to the C-ABI call that launched them.
  BxEnter(bx_handle* h, void* stream, const char* fn = __builtin_FUNCTION());
  ~BxEnter();
  BxEnter(const BxEnter&) = delete;
  BxEnter& operator=(const BxEnter&) = delete;
  int prev_;
  bool nvtx_;
};
int bx_internal_nms_keys(bx_handle* h, const float* boxes, const uint32_t* keys, int batch, int n, int max_out,
                         float iou_threshold, float* out_boxes, int* out_idx, int* out_count, cudaStream_t st);

#define BX_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      bx_set_error(__VA_ARGS__);     \
      return (code);                 \
    }                                \
  } while (0)

#define BX_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      bx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      return BX_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define BX_LAUNCH_CHECK(h)                                                                   \
  do {                                                                                       \
    (h)->launches++;                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) {                                                                \
      bx_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return BX_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

static inline bool bx_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static inline int bx_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long bx_min_ll(long long a, long long b) { return a < b ? a : b; }

#ifdef __CUDACC__
struct BoxCodec {
  float m0, m1, m2, m3, s0, s1, s2, s3;
  float max_x, max_y;  // W-1, H-1 (clip upper bounds)
  int clip;
};

// utils/bbox_transform.py:32-55 followed by utils/bbox_tf.py:71-74, same fp32 op order.
__device__ __forceinline__ float4 bx_decode_clip_one(const float4 a, const float4 t, const BoxCodec& k) {
  const float dx = t.x * k.s0 + k.m0;
  const float dy = t.y * k.s1 + k.m1;
  const float dw = t.z * k.s2 + k.m2;
  const float dh = t.w * k.s3 + k.m3;
  float w = a.z - a.x + 1.0f;
  float h = a.w - a.y + 1.0f;
  float cx = a.x + 0.5f * w;
  float cy = a.y + 0.5f * h;
  cx = cx + dx * w;
  cy = cy + dy * h;
  w = w * expf(dw);
  h = h * expf(dh);
  float4 o;
  o.x = cx - 0.5f * w;
  o.y = cy - 0.5f * h;
  o.z = o.x + w;
  o.w = o.y + h;
  if (k.clip) {
    o.x = fmaxf(fminf(o.x, k.max_x), 0.0f);
    o.y = fmaxf(fminf(o.y, k.max_y), 0.0f);
    o.z = fmaxf(fminf(o.z, k.max_x), 0.0f);
    o.w = fmaxf(fminf(o.w, k.max_y), 0.0f);
  }
  return o;
}

// utils/bbox_transform.py:4-29
__device__ __forceinline__ float4 bx_encode_one(const float4 b, const float4 g, const BoxCodec& k) {
  const float w = b.z - b.x + 1.0f, h = b.w - b.y + 1.0f;
  const float cx = b.x + 0.5f * w, cy = b.y + 0.5f * h;
  const float gw = g.z - g.x + 1.0f, gh = g.w - g.y + 1.0f;
  const float gcx = g.x + 0.5f * gw, gcy = g.y + 0.5f * gh;
  float4 d;
  d.x = ((gcx - cx) / w - k.m0) / k.s0;
  d.y = ((gcy - cy) / h - k.m1) / k.s1;
  d.z = (logf(gw / w) - k.m2) / k.s2;
  d.w = (logf(gh / h) - k.m3) / k.s3;
  return d;
}

// utils/bbox_tf.py:7-56 for one pair ("+1" convention)
__device__ __forceinline__ float bx_iou_plus1(const float4 a, const float area_a, const float4 b, const float area_b) {
  const float ih = fmaxf(0.0f, fminf(a.w, b.w) - fmaxf(a.y, b.y) + 1.0f);
  const float iw = fmaxf(0.0f, fminf(a.z, b.z) - fmaxf(a.x, b.x) + 1.0f);
  const float inter = ih * iw;
  const float uni = area_a + area_b - inter;
  return inter == 0.0f ? 0.0f : inter / uni;
}
__device__ __forceinline__ float bx_area_plus1(const float4 a) { return (a.w - a.y + 1.0f) * (a.z - a.x + 1.0f); }

// order-preserving map fp32 -> u32 (larger float -> larger key); 0 is reserved for "excluded"
__device__ __forceinline__ uint32_t bx_score_key(float s) {
  const uint32_t u = __float_as_uint(s);
  const uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return k == 0u ? 1u : k;
}
// Bitonic sort of v[0..pow2) in shared memory by the whole CTA (pow2 a power of two >= 64, blockDim.x a multiple of 32).
// Each thread keeps two adjacent elements in registers: compare-exchange distance 1 is inside the thread, 2..32 by
// shuffles inside the warp, so only the distances >= 64 go through shared memory with a CTA barrier (21 barriers for
// 2048 elements instead of 66).  kDesc: descending, else ascending.
template <bool kDesc>
__device__ __forceinline__ void bx_bitonic_reg_steps(unsigned long long& e0, unsigned long long& e1, int k, int jstart, int i0) {
  const bool first_big = kDesc ? ((i0 & k) == 0) : ((i0 & k) != 0);   // same for i0 and i0 + 1 (k >= 2)
  for (int j = jstart; j >= 2; j >>= 1) {
    const unsigned long long p0 = __shfl_xor_sync(0xFFFFFFFFu, e0, j >> 1), p1 = __shfl_xor_sync(0xFFFFFFFFu, e1, j >> 1);
    const bool take_max = (first_big == ((i0 & j) == 0));            // the lower index of a "big first" pair keeps the larger
    e0 = take_max ? max(e0, p0) : min(e0, p0);
    e1 = take_max ? max(e1, p1) : min(e1, p1);
  }
  const unsigned long long hi = max(e0, e1), lo = min(e0, e1);
  e0 = first_big ? hi : lo;
  e1 = first_big ? lo : hi;
}

template <bool kDesc>
__device__ void bx_bitonic_sort(unsigned long long* v, int pow2) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i0 = 2 * tid; i0 < pow2; i0 += 2 * nt) {                   // whole warps: pow2 is a multiple of 64
    unsigned long long e0 = v[i0], e1 = v[i0 + 1];
    for (int k = 2; k <= 64; k <<= 1) bx_bitonic_reg_steps<kDesc>(e0, e1, k, k >> 1, i0);
    v[i0] = e0;
    v[i0 + 1] = e1;
  }
  __syncthreads();
  for (int k = 128; k <= pow2; k <<= 1) {
    for (int j = k >> 1; j >= 64; j >>= 1) {
      for (int t = tid; t < (pow2 >> 1); t += nt) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));          // index with bit j clear
        const int p = i | j;
        const unsigned long long x = v[i], y = v[p];
        const bool first_big = kDesc ? ((i & k) == 0) : ((i & k) != 0);
        if ((x < y) == first_big) {
          v[i] = y;
          v[p] = x;
        }
      }
      __syncthreads();
    }
    for (int i0 = 2 * tid; i0 < pow2; i0 += 2 * nt) {
      unsigned long long e0 = v[i0], e1 = v[i0 + 1];
      bx_bitonic_reg_steps<kDesc>(e0, e1, k, 32, i0);
      v[i0] = e0;
      v[i0 + 1] = e1;
    }
    __syncthreads();
  }
}
#endif

