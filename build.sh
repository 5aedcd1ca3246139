#!/bin/bash
# Build libboxpath.so (sm_100a only) in-tree.  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
SRC=tf_eager_object_detection_b200/csrc
mkdir -p tf_eager_object_detection_b200/lib
nvcc -shared -Xcompiler -fPIC -O3 -lineinfo -fmad=false -std=c++17 --extended-lambda \
  -gencode arch=compute_100a,code=sm_100a ${BX_NVCC_EXTRA} \
  -o tf_eager_object_detection_b200/lib/libboxpath.so \
  $SRC/bx_api.cu $SRC/bx_proposals.cu $SRC/bx_roi.cu $SRC/bx_roi_band.cu $SRC/bx_roi_stage.cu $SRC/bx_roi_grad.cu $SRC/bx_targets.cu $SRC/bx_prediction.cu $SRC/bx_losses.cu -ldl
