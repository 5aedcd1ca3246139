"""Mirror of the reference's `object_detection/model/roi_pooling.py` — same class names and call signatures."""
from . import _lib, ops

__all__ = ['RoiPoolingCropAndResize', 'RoiPoolingRoiAlign', 'RoiPoolingCropAndResize2', 'crop_and_resize']


class RoiPoolingCropAndResize2:
    """model/roi_pooling.py:8-42 (FPN): boxes normalised by image H, W; crop 2Px2P; 2x2 max pool."""

    def __init__(self, pool_size):
        self._pool_size = pool_size

    def call(self, inputs, training=None, mask=None):
        shared_layers, rois, image_shape = inputs
        return ops.roi_pool_autograd(_lib.ROI_IMAGE_NORM, _lib.POOL_MAX2, self._pool_size, shared_layers, rois,
                            image_shape=image_shape)

    __call__ = call


class RoiPoolingCropAndResize:
    """model/roi_pooling.py:45-90 (C4): rois / stride normalised by (h-1), (w-1); max_pooling_flag picks
    crop 2Px2P + 2x2 max pool (VGG16) or crop PxP (ResNet)."""

    def __init__(self, pool_size, max_pooling_flag=True):
        self._pool_size = pool_size
        self._max_pooling_flag = max_pooling_flag

    def call(self, inputs, training=None, mask=None):
        shared_layers, rois, extractor_stride = inputs
        pool = _lib.POOL_MAX2 if self._max_pooling_flag else _lib.POOL_NONE
        return ops.roi_pool_autograd(_lib.ROI_STRIDE_NORM, pool, self._pool_size, shared_layers, rois,
                            stride=float(extractor_stride))

    __call__ = call


class RoiPoolingRoiAlign:
    """model/roi_pooling.py:158-176 (-> roi_align :140-155 -> crop_and_resize(pad_border=True) :93-137)."""

    def __init__(self, pool_size):
        self._pool_size = pool_size

    def call(self, inputs, training=None, mask=None):
        shared_layers, rois, extractor_stride = inputs
        return ops.roi_pool_autograd(_lib.ROI_ALIGN_PAD, _lib.POOL_AVG2, self._pool_size, shared_layers, rois,
                            stride=float(extractor_stride))

    __call__ = call


def crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0):
    """tf.image.crop_and_resize(image, boxes, box_ind, crop_size) — the op behind roi_pooling.py:37,79,86,134."""
    return ops.crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value)
