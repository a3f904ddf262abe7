"""End-to-end parity through the C ABI engine vs the CPU oracle (`-m gpu`).

Stage taps are compared first (feature map, RPN head, proposals, ROI pool,
class probabilities), then the final detections: identical class assignment
after NMS and boxes within 1e-3 px (north star).  Same synthetic weights and
images on both sides.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from luminoth_b200 import default_config, synth
from luminoth_b200.engine import Engine
from luminoth_b200.predicting import PredictorNetwork
from oracle import fasterrcnn as ofr
from oracle import ssd as ossd
from oracle import predict as opredict


def rel_err(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(1e-12, np.abs(b).max()))


def match_detections(boxes, labels, probs, ref, box_atol=1e-3, prob_atol=1e-5):
    """Detections as sets: each engine row must have a reference row with the same
    label, prob within prob_atol and box within box_atol (order may differ only among
    near-equal probabilities)."""
    rb, rl, rp = ref['objects'], ref['labels'], ref['probs']
    assert len(boxes) == len(rb), 'detection count %d vs oracle %d' % (len(boxes), len(rb))
    used = np.zeros(len(rb), bool)
    for i in range(len(boxes)):
        cand = np.where((rl == labels[i]) & ~used & (np.abs(rp - probs[i]) <= prob_atol))[0]
        ok = [j for j in cand if np.abs(rb[j] - boxes[i]).max() <= box_atol]
        assert ok, 'row %d (label %d prob %.6f box %s) has no oracle match' % (i, labels[i], probs[i], boxes[i])
        used[ok[0]] = True
    # order: probabilities must be non-increasing like tf.nn.top_k
    assert (np.diff(probs) <= 1e-7).all()


def frcnn_cfg(arch, extra=()):
    return default_config('fasterrcnn', ['model.base_network.architecture=' + arch, 'model.network.num_classes=20',
                                         'model.rpn.proposals.post_nms_top_n=200',
                                         'model.rcnn.proposals.min_prob_threshold=0.05'] + list(extra))


@pytest.mark.parametrize('arch,impl', [('resnet_v1_50', 'simt'), ('resnet_v1_50', 'tc'), ('resnet_v1_101', 'tc')])
def test_fasterrcnn_stages_and_detections(arch, impl):
    cfg = frcnn_cfg(arch)
    wts = synth.make_weights(cfg, seed=1)
    h, w = 224, 320
    imgs = synth.make_images(2, h, w, seed=2)
    eng = Engine(cfg, max_batch=2, max_h=h, max_w=w)
    eng.load_weights(wts).finalize()
    eng.set_conv_impl(impl)
    boxes, scores, labels, counts = eng.predict_raw(imgs)
    fmap = eng.get_tensor('conv_feature_map')
    heads = eng.get_tensor('rpn_heads')
    props = eng.get_tensor('proposals')
    pcnt = eng.get_tensor('proposal_counts').astype(int)
    anchors = eng.get_tensor('all_anchors')
    pooled = eng.get_tensor('roi_pool')
    cls_prob = eng.get_tensor('rcnn_cls_prob')
    for i in range(2):
        ref = ofr.forward(imgs[i].astype(np.float32), wts, cfg)
        assert rel_err(fmap[i], ref['conv_feature_map'][0]) < 2e-5, 'feature map'
        np.testing.assert_array_equal(anchors, ref['all_anchors'].astype(np.float32))
        A = 12
        rh = heads[i].reshape(-1, 6 * A)
        np.testing.assert_allclose(rh[:, :2 * A].reshape(-1, 2), ref['rpn']['rpn_cls_score'], atol=2e-5)
        np.testing.assert_allclose(rh[:, 2 * A:].reshape(-1, 4), ref['rpn']['rpn_bbox_pred'], atol=2e-5)
        rp = ref['rpn_prediction']['proposals']
        assert pcnt[i] == rp.shape[0], 'proposal count'
        np.testing.assert_allclose(props[i, :pcnt[i]], rp, atol=1e-3)
        k = int(counts[i])
        r0 = i * 200
        assert rel_err(pooled[r0:r0 + pcnt[i]], ref['roi']['roi_pool']) < 2e-5, 'roi_pool'
        np.testing.assert_allclose(cls_prob[i, :pcnt[i]], ref['rcnn']['cls_prob'], atol=2e-5)
        match_detections(boxes[i, :k], labels[i, :k], scores[i, :k], ref['classification_prediction'])
    eng.close()


def test_fasterrcnn_rpn_only_mode():
    """with_rcnn: False -> proposals as objects, labels 0 (predicting.py:85-92)."""
    cfg = frcnn_cfg('resnet_v1_50', ['model.network.with_rcnn=False'])
    wts = synth.make_weights(cfg, seed=3)
    imgs = synth.make_images(1, 160, 192, seed=4)
    eng = Engine(cfg, max_batch=1, max_h=160, max_w=192)
    eng.load_weights(wts).finalize()
    boxes, scores, labels, counts = eng.predict_raw(imgs)
    ref = ofr.forward(imgs[0].astype(np.float32), wts, cfg)['rpn_prediction']
    k = int(counts[0])
    assert k == ref['proposals'].shape[0]
    np.testing.assert_allclose(boxes[0, :k], ref['proposals'], atol=1e-3)
    np.testing.assert_allclose(scores[0, :k], ref['scores'], atol=1e-5)
    assert (labels[0, :k] == 0).all()
    eng.close()


@pytest.mark.parametrize('impl', ['simt', 'tc'])
def test_ssd_stages_and_detections(impl):
    cfg = default_config('ssd', ['model.proposals.min_prob_threshold=0.2'])
    wts = synth.make_weights(cfg, seed=5)
    imgs = synth.make_images(2, 300, 300, seed=6)
    eng = Engine(cfg, max_batch=2)
    eng.load_weights(wts).finalize()
    eng.set_conv_impl(impl)
    boxes, scores, labels, counts = eng.predict_raw(imgs)
    loc = eng.get_tensor('loc_pred'); prob = eng.get_tensor('cls_prob'); anchors = eng.get_tensor('all_anchors')
    for i in range(2):
        ref = ossd.forward(imgs[i].astype(np.float32), wts, cfg)
        for j, fm in enumerate(ref['feature_maps'].values()):
            assert rel_err(eng.get_tensor('fmap_%d' % j)[i], fm[0]) < 2e-5, 'fmap %d' % j
        np.testing.assert_array_equal(anchors, ref['all_anchors'])
        np.testing.assert_allclose(loc[i], ref['loc_pred'], atol=3e-5 * max(1, np.abs(ref['loc_pred']).max()))
        np.testing.assert_allclose(prob[i], ref['cls_prob'], atol=2e-5)
        k = int(counts[i])
        match_detections(boxes[i, :k], labels[i, :k], scores[i, :k], ref['classification_prediction'])
    eng.close()


def test_predictor_network_drop_in_schema():
    """PredictorNetwork(config).predict_image(image) -> [{'bbox','label','prob'}] == oracle predict_image.
    600x640 input: the aspect-preserving resize is the identity (scale_factor 1.0)."""
    cfg = frcnn_cfg('resnet_v1_50')
    wts = synth.make_weights(cfg, seed=7)
    img = synth.make_images(1, 600, 640, seed=8)[0]
    net = PredictorNetwork(cfg, weights=wts)
    got = net.predict_image(img)
    ref = opredict.predict_image(img, wts, cfg)
    assert isinstance(got, list) and len(got) == len(ref) and len(got) > 0
    for d in got:
        assert set(d) == {'bbox', 'label', 'prob'} and len(d['bbox']) == 4 and all(isinstance(c, int) for c in d['bbox'])
    assert [d['prob'] for d in got] == sorted([d['prob'] for d in got], reverse=True)
    key = lambda d: (d['label'], d['bbox'])
    for g, r in zip(sorted(got, key=key), sorted(ref, key=key)):
        assert g['label'] == r['label'] and g['bbox'] == r['bbox'] and abs(g['prob'] - r['prob']) <= 1.01e-4
    net.engine.close()


def test_engine_fails_loudly_on_missing_weight():
    cfg = frcnn_cfg('resnet_v1_50')
    wts = synth.make_weights(cfg, seed=1)
    eng = Engine(cfg, max_batch=1, max_h=64, max_w=64)
    name = 'fasterrcnn/rpn/conv/w'
    for n, _ in eng.weight_specs():
        if n != name:
            eng.set_weight(n, wts[n])
    with pytest.raises(ValueError):
        eng.finalize()
    with pytest.raises(ValueError):
        eng.set_weight(name, np.zeros((1, 1, 4, 4), np.float32))     # wrong shape
    eng.close()
