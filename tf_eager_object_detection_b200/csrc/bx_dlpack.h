// Minimal DLPack ABI (dlpack.h, DLTensor layout — stable since v0.2) so bx_dlpack_data can validate tensors that the
// Python side extracts from `__dlpack__()` capsules.  Only the struct layout is restated; no DLPack code is used.
#pragma once
#include <stdint.h>

extern "C" {
typedef struct { int32_t device_type; int32_t device_id; } BxDLDevice;       // kDLCPU=1, kDLCUDA=2, kDLCUDAHost=3, kDLCUDAManaged=13
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } BxDLDataType; // kDLInt=0, kDLUInt=1, kDLFloat=2
typedef struct {
  void* data;
  BxDLDevice device;
  int32_t ndim;
  BxDLDataType dtype;
  int64_t* shape;
  int64_t* strides;   // NULL = compact row-major
  uint64_t byte_offset;
} BxDLTensor;
}
