"""Generates the golden fixtures under tests/golden/ from the CPU oracle.

The reference itself cannot be imported here (TensorFlow 1.x is not
installable), so the vectors come from the oracle --
which is pinned on the reference's own known-answer tests
(tests/test_oracle_reference_goldens.py).  They freeze the oracle's behaviour
(a CPU test re-derives them) and give the GPU tests fixed files to compare the
CUDA path against.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from luminoth_b200 import default_config, synth          # noqa: E402
from oracle import fasterrcnn as ofr, ssd as ossd, tf_ops as T   # noqa: E402
from oracle.anchors import fasterrcnn_anchors, ssd_anchors      # noqa: E402


def rpn_chain():
    rng = np.random.default_rng(2024)
    anchors = fasterrcnn_anchors(8, 10, 256, [0.5, 1, 2], [0.25, 0.5, 1, 2], 16).astype(np.float32)
    na = anchors.shape[0]
    prob = T.softmax(rng.standard_normal((na, 2)).astype(np.float32))
    pred = (rng.standard_normal((na, 4)) * 0.25).astype(np.float32)
    pred[:, 2:] = 0            # exp(0) == 1: decode is bit-reproducible on every platform
    cfg = {'pre_nms_top_n': 600, 'post_nms_top_n': 120, 'nms_threshold': 0.7, 'min_prob_threshold': 0.05,
           'clip_after_nms': False, 'filter_outside_anchors': False, 'apply_nms': True}
    out = ofr.rpn_proposal(prob, pred, anchors, (128, 160), cfg)
    return dict(anchors=anchors, cls_prob=prob, bbox_pred=pred, im_shape=np.array([128, 160]),
                proposals=out['proposals'], scores=out['scores'], **{'cfg_' + k: np.array(v) for k, v in cfg.items()})


def class_chain():
    rng = np.random.default_rng(7)
    r, nc = 150, 6
    c = rng.uniform(0, 500, (r, 2)); s = rng.uniform(16, 200, (r, 2))
    props = np.concatenate([c, c + s], 1).astype(np.float32)
    deltas = (rng.standard_normal((r, 4 * nc)) * 0.5).astype(np.float32)
    deltas.reshape(r, nc, 4)[:, :, 2:] = 0
    prob = T.softmax((rng.standard_normal((r, nc + 1)) * 2).astype(np.float32))
    cfg = {'class_max_detections': 20, 'class_nms_threshold': 0.5, 'total_max_detections': 50, 'min_prob_threshold': 0.05}
    out = ofr.rcnn_proposal(props, deltas, prob, (600, 1024), nc, cfg, variances=[0.1, 0.2])
    return dict(proposals=props, deltas=deltas, cls_prob=prob, objects=out['objects'], labels=out['proposal_label'],
                probs=out['proposal_label_prob'])


def roi_case():
    rng = np.random.default_rng(11)
    fmap = rng.standard_normal((1, 12, 16, 32)).astype(np.float32)
    rois = np.array([[0, 0, 255, 191], [10, 20, 100, 90], [200, 100, 255, 191], [30, 30, 30, 30], [5, 150, 250, 160]],
                    np.float32)
    out = ofr.roi_pool(rois, fmap, (192, 256), 7, 7)['roi_pool']
    return dict(fmap=fmap, rois=rois, pooled=out)


def frcnn_tiny():
    cfg = default_config('fasterrcnn', ['model.base_network.architecture=resnet_v1_50', 'model.network.num_classes=5',
                                        'model.rpn.proposals.post_nms_top_n=60',
                                        'model.rcnn.proposals.min_prob_threshold=0.05'])
    wts = synth.make_weights(cfg, seed=3)
    img = synth.make_images(1, 96, 128, seed=4)[0]
    out = ofr.forward(img, wts, cfg)
    p = out['classification_prediction']
    return dict(image=img, objects=p['objects'], labels=p['labels'], probs=p['probs'],
                feature_map=out['conv_feature_map'][0].astype(np.float32),
                proposals=out['rpn_prediction']['proposals'])


def anchors_case():
    return dict(frcnn_38x64=fasterrcnn_anchors(38, 64, 256, [0.5, 1, 2], [0.25, 0.5, 1, 2], 16)[:2400],
                ssd300=ssd_anchors([(37, 37), (18, 18), (9, 9), (5, 5), (3, 3), (1, 1)], 0.1, 0.88,
                                   [1, 0.5, 2, 0.333, 3], [4, 6, 6, 6, 4, 4], [300, 300, 3]))


CASES = {'rpn_chain': rpn_chain, 'class_chain': class_chain, 'roi_pool': roi_case, 'frcnn_r50_tiny': frcnn_tiny,
         'anchors': anchors_case}


def main():
    for name, fn in CASES.items():
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **fn())
        print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
