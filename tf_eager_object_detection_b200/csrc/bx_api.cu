// Handle, error reporting, workspace, DLPack validation and the composite C4 entry points of libboxpath.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "bx_common.cuh"
#include "bx_dlpack.h"

static thread_local char g_err[512] = "";

void bx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Workspace growth is stream-ordered (cudaMallocAsync / cudaFreeAsync on the caller's stream): no device-wide
// synchronisation, other streams and other handles keep running.  It happens only when a call needs more than any
// earlier call of the handle did (bx_reserve() sizes everything up front, e.g. before a CUDA-graph capture, during which
// growth is refused).  If the handle was last used on a different stream, the free is ordered behind that stream too.
static int reserve(bx_handle* h, void** p, size_t* cur, size_t bytes, cudaStream_t st) {
  if (*cur >= bytes) return BX_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) cudaGetLastError();
  BX_REQUIRE(cap == cudaStreamCaptureStatusNone, BX_ERR_UNSUPPORTED,
             "workspace growth (%zu -> %zu bytes) during CUDA-graph capture: run the call once, or bx_reserve(), before capturing",
             *cur, bytes);
  size_t want = bytes + (bytes >> 2);
  want = (want + 0xFFFFF) & ~static_cast<size_t>(0xFFFFF);
  if (*p) {
    if (h->last_stream_valid && h->last_stream != st) {          // the old block may still be in use over there
      cudaEvent_t ev;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
        if (cudaEventRecord(ev, h->last_stream) == cudaSuccess) cudaStreamWaitEvent(st, ev, 0);
        cudaEventDestroy(ev);
      }
      cudaGetLastError();                                        // a destroyed stream has nothing pending
    }
    BX_CUDA(cudaFreeAsync(*p, st));
    *p = nullptr;
    *cur = 0;
  }
  BX_CUDA(cudaMallocAsync(p, want, st));
  *cur = want;
  return BX_OK;
}

int bx_ws_reserve(bx_handle* h, size_t bytes, cudaStream_t st) { return reserve(h, &h->ws, &h->ws_bytes, bytes, st); }
int bx_stage_reserve(bx_handle* h, size_t bytes, cudaStream_t st) { return reserve(h, &h->stage, &h->stage_bytes, bytes, st); }
int bx_plan_reserve(bx_handle* h, size_t bytes, cudaStream_t st) { return reserve(h, &h->plan, &h->plan_bytes, bytes, st); }

static bool nvtx_enabled() {
  static const bool on = getenv("BX_NVTX") && atoi(getenv("BX_NVTX")) != 0;   // read once: ranges must pair up
  return on;
}

BxEnter::BxEnter(bx_handle* h, void* stream, const char* fn) : prev_(-1), nvtx_(false) {
  if (nvtx_enabled()) {
    nvtxRangePushA(fn);
    nvtx_ = true;
  }
  if (!h) return;
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur != h->device) {
    if (cudaSetDevice(h->device) == cudaSuccess) prev_ = cur;
  }
  h->last_stream = static_cast<cudaStream_t>(stream);
  h->last_stream_valid = 1;
}
BxEnter::~BxEnter() {
  if (prev_ >= 0) cudaSetDevice(prev_);
  if (nvtx_) nvtxRangePop();
}

extern "C" int bx_version(void) { return BX_VERSION; }
extern "C" const char* bx_last_error(void) { return g_err; }

extern "C" int bx_create(int device, bx_handle** out) {
  BX_REQUIRE(out, BX_ERR_INVALID, "bx_create: NULL out");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    bx_set_error("bx_create: no CUDA device available (%s); libboxpath has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return BX_ERR_CUDA;
  }
  BX_REQUIRE(device >= 0 && device < count, BX_ERR_INVALID, "bx_create: device %d not in [0, %d)", device, count);
  cudaDeviceProp prop;                               // no cudaSetDevice: the caller's current device stays what it was
  BX_CUDA(cudaGetDeviceProperties(&prop, device));
  BX_REQUIRE(prop.major >= 10, BX_ERR_UNSUPPORTED, "bx_create: device %d is sm_%d%d; libboxpath is built for sm_100a only",
             device, prop.major, prop.minor);
  bx_handle* h = new bx_handle();
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  h->smem_sm = prop.sharedMemPerMultiprocessor;
  *out = h;
  return BX_OK;
}

extern "C" int bx_destroy(bx_handle* h) {
  if (!h) return BX_OK;
  {
    BxEnter guard(h, h->last_stream_valid ? h->last_stream : nullptr);
    // stream-ordered frees behind the handle's last stream, then wait for them: after bx_destroy nothing of the handle
    // is in flight
    cudaStream_t st = h->last_stream_valid ? h->last_stream : nullptr;
    if (h->ws) cudaFreeAsync(h->ws, st);
    if (h->stage) cudaFreeAsync(h->stage, st);
    if (h->plan) cudaFreeAsync(h->plan, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) cudaGetLastError();
    bx_profile_roi(h, 0, 0);
  }
  delete h;
  return BX_OK;
}

extern "C" int bx_reserve(bx_handle* h, size_t workspace_bytes, size_t plan_bytes, size_t stage_bytes, void* stream) {
  BX_REQUIRE(h, BX_ERR_INVALID, "bx_reserve: NULL handle");
  BxEnter guard(h, stream);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = bx_ws_reserve(h, workspace_bytes, st)) return rc;
  if (int rc = bx_plan_reserve(h, plan_bytes, st)) return rc;
  return bx_stage_reserve(h, stage_bytes, st);
}

extern "C" int bx_stats(const bx_handle* h, long long* out, int n) {
  BX_REQUIRE(h && out && n >= 0, BX_ERR_INVALID, "bx_stats: NULL argument");
  const long long v[6] = {h->launches, h->band_launches, h->band_fallbacks, static_cast<long long>(h->ws_bytes),
                          static_cast<long long>(h->plan_bytes), static_cast<long long>(h->stage_bytes)};
  for (int i = 0; i < n && i < 6; ++i) out[i] = v[i];
  return BX_OK;
}

extern "C" int bx_set_deterministic(bx_handle* h, int on) {
  BX_REQUIRE(h, BX_ERR_INVALID, "bx_set_deterministic: NULL handle");
  h->deterministic = on != 0;
  return BX_OK;
}

extern "C" long long bx_launch_count(const bx_handle* h) { return h ? h->launches : 0; }

// measurement aid (not in the public header): copy the BX_BAND_DEBUG timestamps of the last band launch to the host
extern "C" long long bx_debug_band_dump(bx_handle* h, unsigned long long* out, long long max_ctas, int* info) {
  if (!h || !h->dbg_ptr) return 0;
  const long long n = h->dbg_count < max_ctas ? h->dbg_count : max_ctas;
  cudaDeviceSynchronize();
  cudaMemcpy(out, h->dbg_ptr, static_cast<size_t>(n) * 32, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 4; ++i) info[i] = h->dbg_info[i];
  return n;
}

extern "C" int bx_profile_roi(bx_handle* h, int enable, int capacity) {
  BX_REQUIRE(h, BX_ERR_INVALID, "bx_profile_roi: NULL handle");
  if (h->prof_ev) {
    for (int i = 0; i < 2 * h->prof_cap; ++i) cudaEventDestroy(h->prof_ev[i]);
    delete[] h->prof_ev;
    h->prof_ev = nullptr;
  }
  h->prof_cap = h->prof_n = h->prof_on = 0;
  if (!enable || capacity <= 0) return BX_OK;
  BX_CUDA(cudaSetDevice(h->device));
  h->prof_ev = new cudaEvent_t[2 * capacity];
  for (int i = 0; i < 2 * capacity; ++i) BX_CUDA(cudaEventCreate(&h->prof_ev[i]));
  h->prof_cap = capacity;
  h->prof_on = 1;
  return BX_OK;
}

extern "C" int bx_profile_read(bx_handle* h, float* ms_out, int max_records, int* n_out) {
  BX_REQUIRE(h && ms_out && n_out, BX_ERR_INVALID, "bx_profile_read: NULL argument");
  const int n = h->prof_n < max_records ? h->prof_n : max_records;
  for (int i = 0; i < n; ++i) {
    BX_CUDA(cudaEventSynchronize(h->prof_ev[2 * i + 1]));
    BX_CUDA(cudaEventElapsedTime(&ms_out[i], h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
  }
  *n_out = n;
  h->prof_n = 0;
  return BX_OK;
}

extern "C" int bx_dlpack_data(const void* dltensor, int device, int dtype_code, int bits, int ndim,
                              const int64_t* shape, int align_bytes, void** out_data) {
  BX_REQUIRE(dltensor && out_data, BX_ERR_INVALID, "bx_dlpack_data: NULL argument");
  const BxDLTensor* t = static_cast<const BxDLTensor*>(dltensor);
  BX_REQUIRE(t->device.device_type == 2 /*kDLCUDA*/, BX_ERR_DLPACK,
             "tensor is not on a CUDA device (DLDeviceType %d); libboxpath has no CPU path", t->device.device_type);
  BX_REQUIRE(device < 0 || t->device.device_id == device, BX_ERR_DLPACK, "tensor is on cuda:%d, handle is on cuda:%d",
             t->device.device_id, device);
  BX_REQUIRE(t->dtype.code == dtype_code && t->dtype.bits == bits && t->dtype.lanes == 1, BX_ERR_DLPACK,
             "tensor dtype (code %d, %d bits) != expected (code %d, %d bits)", t->dtype.code, t->dtype.bits, dtype_code, bits);
  BX_REQUIRE(t->ndim == ndim, BX_ERR_DLPACK, "tensor has %d dims, expected %d", t->ndim, ndim);
  int64_t expect_stride = 1;
  for (int d = ndim - 1; d >= 0; --d) {
    BX_REQUIRE(!shape || shape[d] < 0 || t->shape[d] == shape[d], BX_ERR_DLPACK, "tensor dim %d is %lld, expected %lld", d,
               (long long)t->shape[d], (long long)(shape ? shape[d] : -1));
    if (t->strides && t->shape[d] > 1)
      BX_REQUIRE(t->strides[d] == expect_stride, BX_ERR_DLPACK, "tensor is not C-contiguous (dim %d stride %lld)", d,
                 (long long)t->strides[d]);
    expect_stride *= t->shape[d];
  }
  void* p = static_cast<char*>(t->data) + t->byte_offset;
  BX_REQUIRE(align_bytes <= 1 || expect_stride == 0 || bx_aligned(p, align_bytes), BX_ERR_DLPACK,
             "tensor data is not %d-byte aligned", align_bytes);
  *out_data = p;
  return BX_OK;
}

extern "C" int bx_c4_proposal_roi(bx_handle* h, const float* anchors, const float* deltas, const float* scores,
                                  const float* feat, int batch, int n, int fh, int fw, int c,
                                  const bx_proposal_params* p, float stride, int pool_size, int pool, float* out_rois,
                                  int* out_idx, int* out_count, float* out_feat, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(p, BX_ERR_INVALID, "bx_c4_proposal_roi: NULL params");
  int rc = bx_proposals(h, anchors, deltas, scores, batch, n, p, out_rois, out_idx, out_count, stream);
  if (rc) return rc;
  return bx_roi_pool(h, BX_ROI_STRIDE_NORM, pool, pool_size, feat, batch, fh, fw, c, out_rois, nullptr, out_count,
                     batch * p->post_nms, stride, p->image_h, p->image_w, out_feat, stream);
}

extern "C" int bx_c4_proposal_roi_host(bx_handle* h, const float* anchors_dev, const float* deltas_host,
                                       const float* scores_host, const float* feat_host, int batch, int n, int fh,
                                       int fw, int c, const bx_proposal_params* p, float stride, int pool_size,
                                       int pool, float* out_rois_host, int* out_idx_host, int* out_count_host,
                                       float* out_feat_host, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && p && deltas_host && scores_host && feat_host && out_rois_host && out_idx_host && out_count_host &&
                 out_feat_host, BX_ERR_INVALID, "bx_c4_proposal_roi_host: NULL argument");
  BX_REQUIRE(batch > 0 && n > 0 && fh > 0 && fw > 0 && c > 0 && pool_size > 0, BX_ERR_INVALID,
             "bx_c4_proposal_roi_host: bad size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  const size_t b_deltas = up(sizeof(float) * 4 * batch * (size_t)n), b_scores = up(sizeof(float) * batch * (size_t)n);
  const size_t b_feat = up(sizeof(float) * (size_t)batch * fh * fw * c);
  const size_t k = static_cast<size_t>(batch) * p->post_nms;
  const size_t b_rois = up(sizeof(float) * 4 * k), b_idx = up(sizeof(int) * k), b_cnt = up(sizeof(int) * batch);
  const size_t b_out = up(sizeof(float) * k * pool_size * pool_size * c);
  int rc = bx_stage_reserve(h, b_deltas + b_scores + b_feat + b_rois + b_idx + b_cnt + b_out, st);
  if (rc) return rc;
  char* base = static_cast<char*>(h->stage);
  float* d_deltas = reinterpret_cast<float*>(base); base += b_deltas;
  float* d_scores = reinterpret_cast<float*>(base); base += b_scores;
  float* d_feat = reinterpret_cast<float*>(base); base += b_feat;
  float* d_rois = reinterpret_cast<float*>(base); base += b_rois;
  int* d_idx = reinterpret_cast<int*>(base); base += b_idx;
  int* d_cnt = reinterpret_cast<int*>(base); base += b_cnt;
  float* d_out = reinterpret_cast<float*>(base);
  BX_CUDA(cudaMemcpyAsync(d_deltas, deltas_host, sizeof(float) * 4 * batch * (size_t)n, cudaMemcpyHostToDevice, st));
  BX_CUDA(cudaMemcpyAsync(d_scores, scores_host, sizeof(float) * batch * (size_t)n, cudaMemcpyHostToDevice, st));
  BX_CUDA(cudaMemcpyAsync(d_feat, feat_host, sizeof(float) * (size_t)batch * fh * fw * c, cudaMemcpyHostToDevice, st));
  rc = bx_c4_proposal_roi(h, anchors_dev, d_deltas, d_scores, d_feat, batch, n, fh, fw, c, p, stride, pool_size, pool,
                          d_rois, d_idx, d_cnt, d_out, stream);
  if (rc) return rc;
  BX_CUDA(cudaMemcpyAsync(out_rois_host, d_rois, sizeof(float) * 4 * k, cudaMemcpyDeviceToHost, st));
  BX_CUDA(cudaMemcpyAsync(out_idx_host, d_idx, sizeof(int) * k, cudaMemcpyDeviceToHost, st));
  BX_CUDA(cudaMemcpyAsync(out_count_host, d_cnt, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
  BX_CUDA(cudaMemcpyAsync(out_feat_host, d_out, sizeof(float) * k * pool_size * pool_size * c, cudaMemcpyDeviceToHost, st));
  return BX_OK;
}

// ---- multi-GPU exchange: detection records all-gather over the host framework's NCCL communicator
#include <dlfcn.h>
namespace {
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*nccl_group_fn)(void);
typedef const char* (*nccl_errstr_fn)(int);
struct NcclApi {
  nccl_allgather_fn all_gather = nullptr;
  nccl_group_fn group_start = nullptr, group_end = nullptr;
  nccl_errstr_fn err = nullptr;
  bool tried = false;
};
NcclApi g_nccl;

// Only a libnccl that is ALREADY loaded is acceptable: the communicator was created by that copy.
bool nccl_resolve() {
  if (g_nccl.tried) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_NOLOAD);
  void* src = lib ? lib : RTLD_DEFAULT;
  g_nccl.all_gather = reinterpret_cast<nccl_allgather_fn>(dlsym(src, "ncclAllGather"));
  g_nccl.group_start = reinterpret_cast<nccl_group_fn>(dlsym(src, "ncclGroupStart"));
  g_nccl.group_end = reinterpret_cast<nccl_group_fn>(dlsym(src, "ncclGroupEnd"));
  g_nccl.err = reinterpret_cast<nccl_errstr_fn>(dlsym(src, "ncclGetErrorString"));
  g_nccl.tried = g_nccl.all_gather && g_nccl.group_start && g_nccl.group_end;   // a failure is retried next call
  return g_nccl.tried;
}
}  // namespace

extern "C" int bx_allgather_detections(bx_handle* h, void* nccl_comm, const float* records, const int* counts,
                                       int b_local, int kmax, int fields, int world, float* out_records,
                                       int* out_counts, void* stream) {
  BxEnter guard(h, stream);
  BX_REQUIRE(h && nccl_comm && out_records && out_counts, BX_ERR_INVALID, "bx_allgather_detections: NULL argument");
  BX_REQUIRE(b_local >= 0 && kmax >= 0 && fields > 0 && world >= 1, BX_ERR_INVALID, "bx_allgather_detections: bad size");
  BX_REQUIRE(b_local == 0 || (records && counts), BX_ERR_INVALID, "bx_allgather_detections: NULL input");
  BX_REQUIRE(nccl_resolve(), BX_ERR_UNSUPPORTED,
             "bx_allgather_detections: no NCCL library is loaded in this process (the communicator's own libnccl is required)");
  if (b_local == 0) return BX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int kFloat32 = 7, kInt32 = 2;                      // ncclDataType_t
  int rc = g_nccl.group_start();
  if (rc == 0) rc = g_nccl.all_gather(records, out_records, static_cast<size_t>(b_local) * kmax * fields, kFloat32, nccl_comm, st);
  if (rc == 0) rc = g_nccl.all_gather(counts, out_counts, static_cast<size_t>(b_local), kInt32, nccl_comm, st);
  const int rc_end = g_nccl.group_end();
  if (rc == 0) rc = rc_end;
  BX_REQUIRE(rc == 0, BX_ERR_CUDA, "bx_allgather_detections: NCCL error %d (%s)", rc, g_nccl.err ? g_nccl.err(rc) : "?");
  h->launches++;
  return BX_OK;
}
