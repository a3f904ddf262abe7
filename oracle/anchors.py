"""Anchor generation (oracle only).

``generate_anchors_reference``  -> ``luminoth/utils/anchors.py:4-52``
``fasterrcnn_anchors``          -> ``luminoth/models/fasterrcnn/fasterrcnn.py:261-308``
   Quirk Q1: ``np_float64_reference + tf_int32_shifts`` converts the numpy
   reference to the tensor's dtype (int32, truncation toward zero), pinned by
   ``fasterrcnn_test.py:285-302``.
``ssd_*``                       -> ``luminoth/models/ssd/utils.py:5-145`` and
   ``luminoth/models/ssd/ssd.py:112-129``.
"""
import numpy as np

from .bbox import clip_boxes_np


def generate_anchors_reference(base_size, aspect_ratios, scales):
    scales_grid, ratios_grid = np.meshgrid(scales, aspect_ratios)
    base_scales = scales_grid.reshape(-1)
    base_ratios = ratios_grid.reshape(-1)
    sq = np.sqrt(base_ratios)
    heights = base_scales * sq * base_size
    widths = base_scales / sq * base_size
    anchors = np.column_stack([
        0 - (widths - 1) / 2, 0 - (heights - 1) / 2,
        0 + (widths - 1) / 2, 0 + (heights - 1) / 2,
    ])
    real_h = (anchors[:, 3] - anchors[:, 1]).astype(np.int64)
    real_w = (anchors[:, 2] - anchors[:, 0]).astype(np.int64)
    if (real_w == 0).any() or (real_h == 0).any():
        raise ValueError(
            'base_size {} is too small for aspect_ratios and scales.'.format(base_size))
    return anchors


def fasterrcnn_anchors(feat_h, feat_w, base_size, ratios, scales, stride):
    """(feat_h*feat_w*A, 4) int32, index = (y*W + x)*A + a."""
    ref = generate_anchors_reference(base_size, np.array(ratios), np.array(scales))
    ref_i = np.trunc(ref).astype(np.int32)          # Q1
    sx = np.arange(feat_w, dtype=np.int32) * np.int32(stride)
    sy = np.arange(feat_h, dtype=np.int32) * np.int32(stride)
    sx, sy = np.meshgrid(sx, sy)
    sx = sx.reshape(-1); sy = sy.reshape(-1)
    shifts = np.stack([sx, sy, sx, sy], axis=0).T
    all_anchors = ref_i[None, :, :] + shifts[:, None, :]
    return all_anchors.reshape(-1, 4).astype(np.int32)


# ---------------------------------------------------------------- SSD
def ssd_anchor_reference(ratios, scales, num_anchors, fmap_shape):
    heights = np.zeros(num_anchors)
    widths = np.zeros(num_anchors)
    if len(scales) > 1:
        widths[0] = heights[0] = np.sqrt(scales[0] * scales[1]) * fmap_shape[0]
    else:
        heights[0] = scales[0] * fmap_shape[0] * 0.99
        widths[0] = scales[0] * fmap_shape[1] * 0.99
    ratios = ratios[:num_anchors - 1]
    heights[1:] = scales[0] / np.sqrt(ratios) * fmap_shape[0]
    widths[1:] = scales[0] * np.sqrt(ratios) * fmap_shape[1]
    c = 0.5
    return np.column_stack([c - widths / 2, c - heights / 2,
                            c + widths / 2, c + heights / 2])


def ssd_anchors_per_fmap(fmap_shape, ref):
    sx = np.arange(fmap_shape[1]); sy = np.arange(fmap_shape[0])
    sx, sy = np.meshgrid(sx, sy)
    sx = sx.reshape(-1); sy = sy.reshape(-1)
    shifts = np.stack([sx, sy, sx, sy], axis=0).T
    return (ref[None, :, :] + shifts[:, None, :]).reshape(-1, 4)


def ssd_adjust_bboxes(b, old_h, old_w, new_h, new_w):
    x0 = b[:, 0] / old_w * new_w
    y0 = b[:, 1] / old_h * new_h
    x1 = b[:, 2] / old_w * new_w
    y1 = b[:, 3] / old_h * new_h
    return np.stack([x0, y0, x1, y1], axis=1)


def ssd_anchors(fmap_shapes, min_scale, max_scale, ratios, anchors_per_point,
                image_shape):
    """All SSD anchors in image pixels, float32 (n,4); float64 until the final
    ``tf.convert_to_tensor(dtype=float32)`` like ``ssd.py:128-129``."""
    scales = np.linspace(min_scale, max_scale, len(fmap_shapes))
    ratios = np.array(ratios)
    out = []
    for i, shp in enumerate(fmap_shapes):
        ref = ssd_anchor_reference(ratios, scales[i:i + 2], anchors_per_point[i], shp)
        raw = ssd_anchors_per_fmap(shp, ref)
        scaled = ssd_adjust_bboxes(raw, shp[0], shp[1], image_shape[0], image_shape[1])
        out.append(clip_boxes_np(scaled, image_shape))
    return np.concatenate(out, axis=0).astype(np.float32)
