"""TensorFlow-1.x op semantics restated on torch-CPU / numpy (oracle only).

TensorFlow is not vendored in the reference (``setup.py:107``
``tensorflow>=1.5``, unpinned); the kernels restated here are the TF 1.x CPU
kernels the hot path calls (call sites cited per function).  All arrays are
NHWC like the reference graph; fp32 unless the caller passes float64 arrays
(used only to bound fp32 rounding noise in tests).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# padding arithmetic (TF `SAME` / `VALID`)
# --------------------------------------------------------------------------
def same_pads(in_size, k, stride, rate=1):
    """TF SAME: out=ceil(in/stride); extra padding goes to the END."""
    k_eff = k + (k - 1) * (rate - 1)
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k_eff - in_size, 0)
    before = total // 2
    return out, before, total - before


def valid_out(in_size, k, stride, rate=1):
    k_eff = k + (k - 1) * (rate - 1)
    return (in_size - k_eff) // stride + 1


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def conv2d(x, w, stride=1, padding='SAME', rate=1, bias=None):
    """tf.nn.conv2d / slim.conv2d / snt.Conv2D.

    x: (N,H,W,Cin) ; w: TF layout (kh,kw,Cin,Cout) ; returns (N,H',W',Cout).
    Call sites: slim nets via ``base_network.py:143-151``; ``rpn.py:69-90``;
    ``ssd.py:83-96``; ``feature_extractor.py:28-37``.
    """
    kh, kw = w.shape[0], w.shape[1]
    w = w.astype(x.dtype, copy=False)
    if bias is not None:
        bias = bias.astype(x.dtype, copy=False)
    xt = _t(x).permute(0, 3, 1, 2)
    wt = _t(w).permute(3, 2, 0, 1).contiguous()
    if padding == 'SAME':
        _, pt, pb = same_pads(x.shape[1], kh, stride, rate)
        _, pl, pr = same_pads(x.shape[2], kw, stride, rate)
        xt = F.pad(xt, (pl, pr, pt, pb))
    elif padding != 'VALID':
        raise ValueError(padding)
    bt = _t(bias) if bias is not None else None
    y = F.conv2d(xt, wt, bt, stride=stride, dilation=rate)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def conv2d_same(x, w, stride, rate=1):
    """slim ``resnet_utils.conv2d_same``: stride 1 -> SAME; otherwise explicit
    symmetric-ish zero pad (pad_beg = (k_eff-1)//2) followed by VALID."""
    k = w.shape[0]
    if stride == 1:
        return conv2d(x, w, 1, 'SAME', rate)
    k_eff = k + (k - 1) * (rate - 1)
    pad_total = k_eff - 1
    pb = pad_total // 2
    pe = pad_total - pb
    xp = np.pad(x, ((0, 0), (pb, pe), (pb, pe), (0, 0)))
    return conv2d(xp, w, stride, 'VALID', rate)


def batch_norm_inference(x, gamma, beta, mean, var, eps):
    """slim.batch_norm(is_training=False): fused inference BN
    y = (x - mean) * (gamma * rsqrt(var + eps)) + beta, per channel."""
    dt = x.dtype
    inv = (gamma.astype(dt) / np.sqrt(var.astype(dt) + dt.type(eps))).astype(dt)
    return ((x - mean.astype(dt)) * inv + beta.astype(dt)).astype(dt)


def relu(x):
    return np.maximum(x, 0)


def relu6(x):
    return np.minimum(np.maximum(x, 0), 6)


def max_pool(x, k, stride, padding):
    """tf.nn.max_pool / slim.max_pool2d.  SAME pads with -inf (TF ignores the
    padded cells).  Call sites: resnet pool1 (3x3/2 SAME), VGG pools (2x2/2
    VALID, ``truncated_vgg.py:101-113``), pool5 3x3/1 SAME
    (``feature_extractor.py:92-95``), ROI 2x2/2 VALID (``roi_pool.py:82-86``)."""
    xt = _t(x).permute(0, 3, 1, 2)
    if padding == 'SAME':
        _, pt, pb = same_pads(x.shape[1], k, stride)
        _, pl, pr = same_pads(x.shape[2], k, stride)
        xt = F.pad(xt, (pl, pr, pt, pb), value=float('-inf'))
    y = F.max_pool2d(xt, k, stride)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def softmax(x):
    """tf.nn.softmax over the last axis (max-subtracted)."""
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis=-1, keepdims=True)).astype(x.dtype)


def l2_normalize(x, axis, eps):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), eps))."""
    ss = (x * x).sum(axis=axis, keepdims=True)
    return (x / np.sqrt(np.maximum(ss, x.dtype.type(eps)))).astype(x.dtype)


def top_k(values, k):
    """tf.nn.top_k(sorted=True): descending, equal values -> lower index first."""
    values = np.asarray(values)
    order = np.argsort(-values, kind='stable')[:k]
    return values[order], order.astype(np.int32)


def crop_and_resize(image, boxes, box_ind, crop_h, crop_w, extrapolation=0.0):
    """tf.image.crop_and_resize, bilinear (``roi_pool.py:75-78``).

    image (N,H,W,C); boxes (R,4) normalised [y1,x1,y2,x2]; returns
    (R,crop_h,crop_w,C).  Arithmetic order follows the TF CPU kernel.
    """
    dt = image.dtype
    f = dt.type
    N, H, W, C = image.shape
    R = boxes.shape[0]
    out = np.empty((R, crop_h, crop_w, C), dtype=dt)
    boxes = boxes.astype(dt)
    for b in range(R):
        y1, x1, y2, x2 = boxes[b]
        img = image[int(box_ind[b])]
        if crop_h > 1:
            hs = (y2 - y1) * f(H - 1) / f(crop_h - 1)
            in_y = y1 * f(H - 1) + np.arange(crop_h, dtype=dt) * hs
        else:
            in_y = np.full((1,), f(0.5) * (y1 + y2) * f(H - 1), dtype=dt)
        if crop_w > 1:
            ws = (x2 - x1) * f(W - 1) / f(crop_w - 1)
            in_x = x1 * f(W - 1) + np.arange(crop_w, dtype=dt) * ws
        else:
            in_x = np.full((1,), f(0.5) * (x1 + x2) * f(W - 1), dtype=dt)
        y_ok = ~((in_y < 0) | (in_y > f(H - 1)))
        x_ok = ~((in_x < 0) | (in_x > f(W - 1)))
        yc = np.where(y_ok, in_y, f(0))
        xc = np.where(x_ok, in_x, f(0))
        top = np.floor(yc).astype(np.int64)
        bot = np.ceil(yc).astype(np.int64)
        yl = (yc - top.astype(dt)).astype(dt)
        lef = np.floor(xc).astype(np.int64)
        rig = np.ceil(xc).astype(np.int64)
        xl = (xc - lef.astype(dt)).astype(dt)
        tl = img[top][:, lef]
        tr = img[top][:, rig]
        bl = img[bot][:, lef]
        br = img[bot][:, rig]
        xlb = xl[None, :, None]
        topv = tl + (tr - tl) * xlb
        botv = bl + (br - bl) * xlb
        val = topv + (botv - topv) * yl[:, None, None]
        ok = (y_ok[:, None] & x_ok[None, :])[:, :, None]
        out[b] = np.where(ok, val, f(extrapolation))
    return out


def iou_tf(box_i, boxes_j):
    """IoU exactly as tf.image.non_max_suppression computes it (fp32, no +1,
    coordinates min/max-normalised, 0 if either area <= 0)."""
    dt = boxes_j.dtype
    ymin_i = min(box_i[0], box_i[2]); xmin_i = min(box_i[1], box_i[3])
    ymax_i = max(box_i[0], box_i[2]); xmax_i = max(box_i[1], box_i[3])
    ymin_j = np.minimum(boxes_j[:, 0], boxes_j[:, 2])
    xmin_j = np.minimum(boxes_j[:, 1], boxes_j[:, 3])
    ymax_j = np.maximum(boxes_j[:, 0], boxes_j[:, 2])
    xmax_j = np.maximum(boxes_j[:, 1], boxes_j[:, 3])
    area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i)
    area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
    iy0 = np.maximum(ymin_i, ymin_j)
    ix0 = np.maximum(xmin_i, xmin_j)
    iy1 = np.minimum(ymax_i, ymax_j)
    ix1 = np.minimum(xmax_i, xmax_j)
    inter = np.maximum(iy1 - iy0, dt.type(0)) * np.maximum(ix1 - ix0, dt.type(0))
    with np.errstate(divide='ignore', invalid='ignore'):
        iou = inter / (area_i + area_j - inter)
    bad = (area_j <= 0) | (area_i <= 0)
    return np.where(bad, dt.type(0), iou).astype(dt)


def non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    """tf.image.non_max_suppression (greedy, score desc, suppress iff
    IoU > thr strictly).  boxes (n,4) [y1,x1,y2,x2].  Returns indices into the
    input in selection order.  Equal scores: lower index first (canonical;
    TF1 minors differ).  Call sites
    ``rpn_proposal.py:152-157``, ``rcnn_proposal.py:114-117``,
    ``ssd/proposal.py:123-126``."""
    boxes = np.asarray(boxes)
    scores = np.asarray(scores)
    n = boxes.shape[0]
    if n == 0 or max_output_size <= 0:
        return np.zeros((0,), dtype=np.int32)
    order = np.argsort(-scores, kind='stable')
    sb = boxes[order]
    thr = boxes.dtype.type(iou_threshold)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if len(keep) >= max_output_size:
            break
        if i + 1 < n:
            iou = iou_tf(sb[i], sb[i + 1:])
            suppressed[i + 1:] |= iou > thr
    return np.asarray(keep, dtype=np.int32)


def resize_bilinear(image, new_h, new_w):
    """tf.image.resize_images(..., BILINEAR), TF1 legacy (align_corners=False,
    no half-pixel centres): src = dst * (in/out).  ``utils/image.py:94-97,127-130``."""
    dt = image.dtype
    H, W, C = image.shape
    if (H, W) == (new_h, new_w):
        return image.copy()
    hs = dt.type(H) / dt.type(new_h)
    ws = dt.type(W) / dt.type(new_w)
    ys = np.arange(new_h, dtype=dt) * hs
    xs = np.arange(new_w, dtype=dt) * ws
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, H - 1)
    x1 = np.minimum(x0 + 1, W - 1)
    yl = (ys - y0.astype(dt))[:, None, None]
    xl = (xs - x0.astype(dt))[None, :, None]
    tl = image[y0][:, x0]; tr = image[y0][:, x1]
    bl = image[y1][:, x0]; br = image[y1][:, x1]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return (top + (bot - top) * yl).astype(dt)


def to_int32(x):
    """tf.to_int32: truncation toward zero."""
    return int(math.trunc(float(x)))
