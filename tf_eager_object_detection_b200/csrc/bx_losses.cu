// Losses of the training step ("next" row f3): model/losses.py:4-28 as called at
// faster_rcnn/base_faster_rcnn_model.py:200-224 and fpn/base_fpn_model.py (same helpers).  Each entry computes the
// scalar loss and, in the same pass, its gradient w.r.t. the prediction, so the drop-in closes the training step around
// bx_anchor_target / bx_proposal_target without a framework autograd graph.  Reductions are deterministic: per-CTA
// partial sums in a fixed tree, the last CTA (ticket) adds the partials in index order in fp64.
#include "bx_common.cuh"

namespace {

constexpr int kThreads = 256;

struct ReduceWs {
  double* partial;        // [grid]
  unsigned int* ticket;   // zeroed before launch
  int* partial_count;     // [grid] (cls only)
};

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < kThreads / 32; ++i) t += s_red[i];
  __syncthreads();
  return t;   // valid on thread 0
}

__device__ __forceinline__ bool last_block(unsigned int* ticket) {
  __shared__ bool s_last;
  __threadfence();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// ---- smooth L1 (losses.py:16-28) ----------------------------------------------------------------------------------
struct SmoothL1Args {
  const float* pred; const float* target; const float* in_w; const float* out_w;
  long long total;        // n * d elements
  float sigma2, inv_denom;
  float* out_loss; float* out_grad;
  ReduceWs ws;
};

__global__ void __launch_bounds__(kThreads) smooth_l1_kernel(const SmoothL1Args a) {
  __shared__ double s_red[kThreads / 32];
  const float thr = 1.0f / a.sigma2, half = 0.5f / a.sigma2, hs2 = a.sigma2 / 2.0f;
  double acc = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < a.total;
       i += static_cast<long long>(kThreads) * gridDim.x) {
    const float iw = a.in_w[i], ow = a.out_w[i];
    const float d = iw * (a.pred[i] - a.target[i]);                       // losses.py:18-19
    const float ad = fabsf(d);
    const bool quad = ad < thr;                                           // losses.py:21
    const float per = quad ? (d * d) * hs2 : (ad - half);                 // losses.py:22
    acc += static_cast<double>(ow * per);                                 // losses.py:23
    if (a.out_grad) {
      const float dd = quad ? d * a.sigma2 : (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f));
      a.out_grad[i] = ow * dd * iw * a.inv_denom;
    }
  }
  const double t = block_sum(acc, s_red);
  if (threadIdx.x == 0) a.ws.partial[blockIdx.x] = t;
  if (last_block(a.ws.ticket) && threadIdx.x == 0) {
    double s = 0.0;
    for (unsigned int i = 0; i < gridDim.x; ++i) s += a.ws.partial[i];
    *a.out_loss = static_cast<float>(s * static_cast<double>(a.inv_denom));   // losses.py:24-27
  }
}

// ---- sparse softmax cross entropy (losses.py:4-13) ----------------------------------------------------------------
struct ClsArgs {
  const float* logits; const float* labels;
  int n, c;
  float weight;
  float* out_loss; int* out_count; float* out_grad;
  ReduceWs ws;
  int* d_count;           // total selected rows (device scalar the gradient kernel reads)
};

__global__ void __launch_bounds__(kThreads) cls_loss_kernel(const ClsArgs a) {
  __shared__ double s_red[kThreads / 32];
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  double acc = 0.0;
  int cnt = 0;
  for (int r = blockIdx.x * kThreads + threadIdx.x; r < a.n; r += kThreads * gridDim.x) {
    const float lab = a.labels[r];
    if (!(lab >= 0.0f)) continue;                                         // base_faster_rcnn_model.py:204
    const float* x = a.logits + static_cast<size_t>(r) * a.c;
    float m = x[0];
    for (int k = 1; k < a.c; ++k) m = fmaxf(m, x[k]);
    float se = 0.0f;
    for (int k = 0; k < a.c; ++k) se += expf(x[k] - m);
    const int li = static_cast<int>(lab);                                 // tf.to_int32 truncates
    // a label >= c is invalid: tf.losses.sparse_softmax_cross_entropy raises on the CPU and yields NaN on the GPU; the
    // loss (and that row's gradient) is NaN here, never a silently clamped class
    acc += (li < a.c) ? static_cast<double>((logf(se) - (x[li] - m)) * a.weight) : static_cast<double>(nanf(""));
    ++cnt;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
  const double t = block_sum(acc, s_red);                                 // has __syncthreads: s_cnt complete
  if (threadIdx.x == 0) {
    a.ws.partial[blockIdx.x] = t;
    a.ws.partial_count[blockIdx.x] = s_cnt;
  }
  if (last_block(a.ws.ticket) && threadIdx.x == 0) {
    double s = 0.0;
    int total = 0;
    for (unsigned int i = 0; i < gridDim.x; ++i) { s += a.ws.partial[i]; total += a.ws.partial_count[i]; }
    const bool present = total > 0 && a.weight != 0.0f;                   // SUM_BY_NONZERO_WEIGHTS, div_no_nan
    *a.out_loss = present ? static_cast<float>(s / static_cast<double>(total)) : 0.0f;
    *a.d_count = present ? total : 0;
    if (a.out_count) *a.out_count = total;
  }
}

__global__ void __launch_bounds__(kThreads) cls_grad_kernel(const ClsArgs a) {
  const int total = *a.d_count;
  const float scale = total > 0 ? a.weight / static_cast<float>(total) : 0.0f;
  for (int r = blockIdx.x * kThreads + threadIdx.x; r < a.n; r += kThreads * gridDim.x) {
    const float lab = a.labels[r];
    const float* x = a.logits + static_cast<size_t>(r) * a.c;
    float* g = a.out_grad + static_cast<size_t>(r) * a.c;
    if (!(lab >= 0.0f) || total == 0) {
      for (int k = 0; k < a.c; ++k) g[k] = 0.0f;
      continue;
    }
    float m = x[0];
    for (int k = 1; k < a.c; ++k) m = fmaxf(m, x[k]);
    float se = 0.0f;
    for (int k = 0; k < a.c; ++k) se += expf(x[k] - m);
    const int li = static_cast<int>(lab);
    if (li >= a.c) {
      for (int k = 0; k < a.c; ++k) g[k] = nanf("");
      continue;
    }
    for (int k = 0; k < a.c; ++k) g[k] = (expf(x[k] - m) / se - (k == li ? 1.0f : 0.0f)) * scale;
  }
}

int reduce_ws(bx_handle* h, int grid, ReduceWs* ws, int** d_count, cudaStream_t st) {
  const size_t bytes = static_cast<size_t>(grid) * (sizeof(double) + sizeof(int)) + 64;
  if (int rc = bx_ws_reserve(h, bytes, st)) return rc;
  char* p = static_cast<char*>(h->ws);
  ws->partial = reinterpret_cast<double*>(p);
  ws->ticket = reinterpret_cast<unsigned int*>(p + static_cast<size_t>(grid) * sizeof(double));
  *d_count = reinterpret_cast<int*>(ws->ticket + 1);
  ws->partial_count = reinterpret_cast<int*>(ws->ticket + 4);
  BX_CUDA(cudaMemsetAsync(ws->ticket, 0, 16, st));
  return BX_OK;
}

}  // namespace

extern "C" int bx_smooth_l1_loss(bx_handle* h, const float* pred, const float* target, const float* in_w,
                                 const float* out_w, long long n, int d, float sigma, int reduce_all, float* out_loss,
                                 float* out_grad, void* stream) {
  BxEnter guard(h, stream);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BX_REQUIRE(h && out_loss, BX_ERR_INVALID, "bx_smooth_l1_loss: null handle / output");
  BX_REQUIRE(n >= 0 && d >= 1 && sigma > 0.0f, BX_ERR_INVALID, "bx_smooth_l1_loss: n >= 0, d >= 1, sigma > 0 required");
  BX_REQUIRE(n == 0 || (pred && target && in_w && out_w), BX_ERR_INVALID, "bx_smooth_l1_loss: null input");
  const long long total = n * d;
  const int grid = static_cast<int>(bx_min_ll(2ll * h->num_sms, (total + kThreads - 1) / kThreads > 0
                                                                     ? (total + kThreads - 1) / kThreads : 1));
  SmoothL1Args a{pred, target, in_w, out_w, total, sigma * sigma,
                 reduce_all ? 1.0f : (n > 0 ? 1.0f / static_cast<float>(n) : 0.0f), out_loss, out_grad, {}};
  int* unused;
  if (int rc = reduce_ws(h, grid, &a.ws, &unused, st)) return rc;
  smooth_l1_kernel<<<grid, kThreads, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  return BX_OK;
}

extern "C" int bx_cls_loss(bx_handle* h, const float* logits, const float* labels, int n, int c, float weight,
                           float* out_loss, int* out_count, float* out_grad, void* stream) {
  BxEnter guard(h, stream);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BX_REQUIRE(h && out_loss, BX_ERR_INVALID, "bx_cls_loss: null handle / output");
  BX_REQUIRE(n >= 0 && c >= 1, BX_ERR_INVALID, "bx_cls_loss: n >= 0, c >= 1 required");
  BX_REQUIRE(n == 0 || (logits && labels), BX_ERR_INVALID, "bx_cls_loss: null input");
  const int blocks = (n + kThreads - 1) / kThreads;
  const int grid = blocks < 1 ? 1 : (blocks > 2 * h->num_sms ? 2 * h->num_sms : blocks);
  ClsArgs a{logits, labels, n, c, weight, out_loss, out_count, out_grad, {}, nullptr};
  if (int rc = reduce_ws(h, grid, &a.ws, &a.d_count, st)) return rc;
  cls_loss_kernel<<<grid, kThreads, 0, st>>>(a);
  BX_LAUNCH_CHECK(h);
  if (out_grad && n > 0) {
    cls_grad_kernel<<<grid, kThreads, 0, st>>>(a);
    BX_LAUNCH_CHECK(h);
  }
  return BX_OK;
}
