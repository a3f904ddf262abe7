"""Box arithmetic of the hot path (oracle only).

Follows ``luminoth/utils/bbox_transform_tf.py``: ``get_width_upright`` :4-16,
``encode`` :19-38 (training-only; kept because the reference's proposal tests
build their inputs with it), ``decode`` :41-66, ``clip_boxes`` :69-99,
``change_order`` :102-126; and numpy ``clip_boxes``
``luminoth/utils/bbox_transform.py:105-122`` (used by the SSD anchors).
Every expression keeps the reference's left-to-right evaluation order so fp32
rounding is reproduced step for step (no fused multiply-add).
"""
import numpy as np


def get_width_upright(b):
    b = b.astype(np.float32)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    one, half = np.float32(1.), np.float32(.5)
    width = x2 - x1 + one
    height = y2 - y1 + one
    urx = x1 + half * width
    ury = y1 + half * height
    return width, height, urx, ury


def encode(bboxes, gt_boxes, variances=None):
    bw, bh, bx, by = get_width_upright(bboxes)
    gw, gh, gx, gy = get_width_upright(gt_boxes)
    if variances is None:
        variances = [1., 1.]
    v0, v1 = np.float32(variances[0]), np.float32(variances[1])
    dx = (gx - bx) / (bw * v0)
    dy = (gy - by) / (bh * v0)
    dw = np.log(gw / bw) / v1
    dh = np.log(gh / bh) / v1
    return np.stack([dx, dy, dw, dh], axis=1).astype(np.float32)


def decode(roi, deltas, variances=None):
    w, h, urx, ury = get_width_upright(roi)
    deltas = deltas.astype(np.float32)
    dx, dy, dw, dh = deltas[:, 0], deltas[:, 1], deltas[:, 2], deltas[:, 3]
    if variances is None:
        variances = [1., 1.]
    v0, v1 = np.float32(variances[0]), np.float32(variances[1])
    half, one = np.float32(.5), np.float32(1.)
    px = dx * w * v0 + urx
    py = dy * h * v0 + ury
    pw = np.exp(dw * v1) * w
    ph = np.exp(dh * v1) * h
    x1 = px - half * pw
    y1 = py - half * ph
    # "This -1. extra is different from reference implementation."
    x2 = px + half * pw - one
    y2 = py + half * ph - one
    return np.stack([x1, y1, x2, y2], axis=1).astype(np.float32)


def clip_boxes(bboxes, imshape):
    """imshape = (height, width)."""
    b = bboxes.astype(np.float32)
    height = np.float32(imshape[0])
    width = np.float32(imshape[1])
    one, zero = np.float32(1.), np.float32(0.)
    x1 = np.maximum(np.minimum(b[:, 0], width - one), zero)
    x2 = np.maximum(np.minimum(b[:, 2], width - one), zero)
    y1 = np.maximum(np.minimum(b[:, 1], height - one), zero)
    y2 = np.maximum(np.minimum(b[:, 3], height - one), zero)
    return np.stack([x1, y1, x2, y2], axis=1)


def change_order(b):
    return np.stack([b[:, 1], b[:, 0], b[:, 3], b[:, 2]], axis=1)


def clip_boxes_np(boxes, image_shape):
    """numpy twin (``bbox_transform.py:105-122``); NB it mutates in place in
    the reference -- here a copy is returned, callers that relied on the
    mutation are vacuous asserts."""
    boxes = np.array(boxes, copy=True)
    mw = image_shape[1] - 1
    mh = image_shape[0] - 1
    boxes[:, 0] = np.maximum(np.minimum(boxes[:, 0], mw), 0)
    boxes[:, 1] = np.maximum(np.minimum(boxes[:, 1], mh), 0)
    boxes[:, 2] = np.maximum(np.minimum(boxes[:, 2], mw), 0)
    boxes[:, 3] = np.maximum(np.minimum(boxes[:, 3], mh), 0)
    return boxes
