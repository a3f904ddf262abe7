"""``PredictorNetwork`` -- drop-in for ``luminoth/utils/predicting.py:10-148``.

Same constructor and ``predict_image(image) -> [{'bbox', 'label', 'prob'}]``
(sorted by probability, integer pixel boxes in the ORIGINAL image, probs
rounded to 4 decimals).  ``predict_batch(images)`` is the batched extension
(the reference is hard-wired to batch 1, ``fasterrcnn.py:101-103``).

What changed underneath: graph build + ``session.run`` is one call into the
sm_100a engine (``lumi_predict``).  The aspect-preserving resize of
``datasets/object_detection_dataset.py:71-83`` / ``utils/image.py:38-147`` is
still done on the host here (SURVEY section 8f item 2 moves it to the GPU);
it is the identity at the benchmark shapes.
"""
import json
import math
import os
import warnings

import numpy as np

from .config import get_model_type
from .engine import Engine


def _resize_bilinear_legacy(image, new_h, new_w):
    """tf.image.resize_images(BILINEAR), TF1 legacy kernel (align_corners=False,
    src = dst * in/out, no half-pixel offset) -- ``utils/image.py:94-97``."""
    image = np.asarray(image, np.float32)
    H, W = image.shape[:2]
    if (H, W) == (new_h, new_w):
        return image
    hs = np.float32(H) / np.float32(new_h)
    ws = np.float32(W) / np.float32(new_w)
    ys = np.arange(new_h, dtype=np.float32) * hs
    xs = np.arange(new_w, dtype=np.float32) * ws
    y0 = np.floor(ys).astype(np.int64); x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, H - 1); x1 = np.minimum(x0 + 1, W - 1)
    yl = (ys - y0.astype(np.float32))[:, None, None]
    xl = (xs - x0.astype(np.float32))[None, :, None]
    top = image[y0][:, x0] + (image[y0][:, x1] - image[y0][:, x0]) * xl
    bot = image[y1][:, x0] + (image[y1][:, x1] - image[y1][:, x0]) * xl
    return (top + (bot - top) * yl).astype(np.float32)


def preprocess_image(image, config):
    """Inference preprocessing; returns (uint8-compatible float image, scale_factor)."""
    image = np.asarray(image)
    if image.ndim != 3 or image.shape[2] != 3:
        raise ValueError('expected an (H, W, 3) RGB image')
    ip = config['dataset']['image_preprocessing']
    f32 = np.float32
    h, w = f32(image.shape[0]), f32(image.shape[1])
    if ip.get('fixed_height') and ip.get('fixed_width'):
        nh, nw = int(ip['fixed_height']), int(ip['fixed_width'])
        return _resize_bilinear_legacy(image, nh, nw), (f32(nh) / h, f32(nw) / w)
    mn, mx = ip.get('min_size'), ip.get('max_size')
    up = max(f32(mn) / min(h, w), f32(1.)) if mn is not None else f32(1.)
    down = min(f32(mx) / max(h, w), f32(1.)) if mx is not None else f32(1.)
    scale = f32(up * down)
    nh, nw = int(math.trunc(float(h * scale))), int(math.trunc(float(w * scale)))
    return _resize_bilinear_legacy(image, nh, nw), scale


def format_predictions(objects, labels, probs, scale_factor, class_labels=None):
    """``predicting.py:114-148``: rescale, int(round()), round(prob, 4), sort desc."""
    objects = np.array(objects, np.float32, copy=True)
    labels = np.asarray(labels).tolist()
    probs = np.asarray(probs, np.float32).tolist()
    if class_labels is not None:
        labels = [class_labels[label] for label in labels]
    if isinstance(scale_factor, tuple):
        objects /= np.array([scale_factor[1], scale_factor[0], scale_factor[1], scale_factor[0]], np.float32)
    else:
        objects /= np.float32(scale_factor)
    objects = [[int(round(coord)) for coord in obj] for obj in objects.tolist()]
    return sorted([{'bbox': obj, 'label': label, 'prob': round(prob, 4)}
                   for obj, label, prob in zip(objects, labels, probs)],
                  key=lambda x: x['prob'], reverse=True)


def load_checkpoint_weights(engine, job_dir):
    """The ``Saver.restore`` of ``predicting.py:51-63`` without TensorFlow: reads the latest Saver-V2 bundle under
    ``job_dir`` and returns {tf variable name: array} for exactly the variables the engine's plan uses (optimizer
    slots, ``global_step`` and never-executed layers such as ResNet-50's block4 are ignored, like a Saver built from
    the inference graph would).  Raises ``ValueError`` when the directory holds no checkpoint, a variable is
    missing or a shape differs."""
    from . import tf_checkpoint as tfc
    prefix = tfc.latest_checkpoint(job_dir)                # ValueError('Could not find checkpoint in ...')
    reader = tfc.BundleReader(prefix)
    weights, missing = {}, []
    for name, shape in engine.weight_specs():
        if not reader.has_tensor(name):
            missing.append(name)
            continue
        if tuple(reader.shape(name)) != tuple(shape):
            raise ValueError("checkpoint variable '%s' has shape %s, the model expects %s"
                             % (name, tuple(reader.shape(name)), tuple(shape)))
        weights[name] = reader.get_tensor(name).astype(np.float32, copy=False)
    if missing:
        raise ValueError('checkpoint %s lacks %d model variables, e.g. %s' % (prefix, len(missing), missing[:3]))
    return weights


class PredictorNetwork(object):
    """Instantiates a network in order to get predictions from it.

    ``weights``: optional dict {tf variable name: array} (TF layouts).  Without
    it, and without a checkpoint, the model is randomly initialised with a
    warning -- exactly the reference's "prediction without checkpoint is just
    used for testing" branch (``predicting.py:64-72``).
    """

    def __init__(self, config, weights=None, device=0, max_batch=1):
        self.class_labels = None
        if config.dataset.dir:
            classes_file = os.path.join(config.dataset.dir, 'classes.json')
            if os.path.exists(classes_file):
                with open(classes_file) as f:
                    self.class_labels = json.load(f)
        config.dataset.data_augmentation = None
        get_model_type(config.model.type)                 # ValueError on unknown model types
        self.config = config
        self.engine = Engine(config, device=device, max_batch=max_batch)
        if weights is None:
            if config.train.job_dir:
                job_dir = config.train.job_dir
                if config.train.run_name:
                    job_dir = os.path.join(job_dir, config.train.run_name)
                # predicting.py:51-63: latest checkpoint of <job_dir>/<run_name>, restored by variable name
                weights = load_checkpoint_weights(self.engine, job_dir)
            else:
                warnings.warn('Could not load checkpoint. Using initialized model.')
                from .synth import make_weights
                weights = make_weights(config, seed=config.train.seed or 0, profile='reference')
        self.engine.load_weights(weights).finalize()

    # -- reference API
    def predict_image(self, image):
        return self.predict_batch([image])[0]

    # -- batched extension
    def predict_batch(self, images):
        pre = [preprocess_image(np.asarray(im), self.config) for im in images]
        shapes = {p[0].shape for p in pre}
        if len(shapes) != 1:
            raise ValueError('predict_batch needs images that preprocess to one size; got %s' % sorted(shapes))
        # the reference feeds the resized float image; the engine ingests uint8 pixels
        # (exact for unresized uint8 inputs -- the identity path at the benchmark shapes)
        batch = np.stack([np.clip(np.rint(p[0]), 0, 255).astype(np.uint8) for p in pre])
        out = []
        for s in range(0, len(batch), self.engine.max_batch):
            chunk = batch[s:s + self.engine.max_batch]
            boxes, scores, labels, counts = self.engine.predict_raw(chunk)
            for i in range(len(chunk)):
                k = int(counts[i])
                out.append(format_predictions(boxes[i, :k], labels[i, :k], scores[i, :k], pre[s + i][1],
                                              self.class_labels))
        return out
