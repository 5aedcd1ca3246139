"""B200-native box-processing hot path (proposal + NMS + RoI pooling + IoU targets).

Drop-in for the box-processing classes/functions of irvingzhang0512/tf_eager_object_detection
(see DESIGN.md for the path and its boundary).  All compute runs in hand-written sm_100a CUDA
kernels behind the C-ABI library `lib/libboxpath.so` (header: `include/boxpath.h`); there is no
CPU fallback — calling any op without the built library or without a CUDA device raises.
"""
__version__ = '0.1.0'
