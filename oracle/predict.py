"""``PredictorNetwork.predict_image`` end to end on the CPU (oracle only).

``luminoth/utils/predicting.py:109-148`` (fetch, rescale, int(round()),
round(prob, 4), stable sort desc) and the inference preprocessing
``luminoth/datasets/object_detection_dataset.py:71-83,202-234`` ->
``luminoth/utils/image.py:38-114`` (``resize_image``) / :117-147
(``resize_image_fixed``).
"""
import numpy as np

from . import tf_ops as T
from . import fasterrcnn, ssd


def preprocess(image, config):
    """Returns (resized float32 image, scale_factor) -- scalar for the
    aspect-preserving path, (sh, sw) tuple for the fixed-size path."""
    image = np.asarray(image, np.float32)
    f32 = np.float32
    ip = config['dataset']['image_preprocessing']
    h, w = f32(image.shape[0]), f32(image.shape[1])
    if ip.get('fixed_height') and ip.get('fixed_width'):
        nh, nw = ip['fixed_height'], ip['fixed_width']
        scale = (f32(nh) / h, f32(nw) / w)
        return T.resize_bilinear(image, int(nh), int(nw)), scale
    mn, mx = ip.get('min_size'), ip.get('max_size')
    up = max(f32(mn) / min(h, w), f32(1.)) if mn is not None else f32(1.)
    down = min(f32(mx) / max(h, w), f32(1.)) if mx is not None else f32(1.)
    scale = f32(up * down)
    nh, nw = T.to_int32(h * scale), T.to_int32(w * scale)
    return T.resize_bilinear(image, nh, nw), scale


def network_outputs(image, wts, config):
    """objects (K,4) f32, labels (K,) i32, probs (K,) f32, scale_factor."""
    resized, scale = preprocess(image, config)
    mtype = config['model']['type']
    if mtype == 'ssd':
        pred = ssd.forward(resized, wts, config)['classification_prediction']
        return pred['objects'], pred['labels'], pred['probs'], scale
    if mtype == 'fasterrcnn':
        out = fasterrcnn.forward(resized, wts, config)
        if config['model']['network'].get('with_rcnn', False):
            p = out['classification_prediction']
            return p['objects'], p['labels'], p['probs'], scale
        rp = out['rpn_prediction']
        return rp['proposals'], np.zeros(rp['scores'].shape, np.int32), rp['scores'], scale
    raise ValueError("Model type '{}' not supported".format(mtype))


def finalize_predictions(objects, labels, probs, scale_factor, class_labels=None):
    """``predicting.py:114-148`` -- shared by the oracle and (restated
    independently) by the product wrapper."""
    objects = np.array(objects, np.float32, copy=True)
    labels = np.asarray(labels).tolist()
    probs = np.asarray(probs, np.float32).tolist()
    if class_labels is not None:
        labels = [class_labels[l] for l in labels]
    if isinstance(scale_factor, tuple):
        objects /= np.array([scale_factor[1], scale_factor[0],
                             scale_factor[1], scale_factor[0]], np.float32)
    else:
        objects /= np.float32(scale_factor)
    objs = [[int(round(c)) for c in o] for o in objects.tolist()]
    return sorted([{'bbox': o, 'label': l, 'prob': round(p, 4)}
                   for o, l, p in zip(objs, labels, probs)],
                  key=lambda x: x['prob'], reverse=True)


def predict_image(image, wts, config, class_labels=None):
    o, l, p, s = network_outputs(image, wts, config)
    return finalize_predictions(o, l, p, s, class_labels)
