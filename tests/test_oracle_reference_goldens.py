"""Pins the CPU oracle against every known-answer test the reference holds for
the hot path.  Each test cites the reference test it
ports (paths under /root/reference/luminoth/); inputs and expected values are
the reference's own, restated -- TensorFlow cannot be imported here."""
import numpy as np
import pytest

from oracle import tf_ops as T
from oracle.anchors import generate_anchors_reference, fasterrcnn_anchors
from oracle.bbox import encode, decode, clip_boxes, clip_boxes_np
from oracle.fasterrcnn import rpn_proposal, rcnn_proposal, roi_pool
from oracle.predict import preprocess
from oracle.resnet import subtract_means


# ---------------------------------------------------------------- anchors
def _wh(a):
    return np.column_stack((a[:, 2] - a[:, 0] + 1, a[:, 3] - a[:, 1] + 1))


def test_anchor_reference_exact():
    """utils/anchors_test.py:17-94"""
    a = generate_anchors_reference(256, [1.], [1.])
    assert a.shape == (1, 4)
    np.testing.assert_array_equal(a[0], [-127.5, -127.5, 127.5, 127.5])
    scales = np.array([0.5, 1., 2., 4.])
    a = generate_anchors_reference(256, [1.], scales)
    assert a.shape == (4, 4)
    wh = _wh(a)
    assert (wh[:, 0] == wh[:, 1]).all()
    np.testing.assert_array_equal(wh[:, 0], 256 * scales)
    np.testing.assert_array_equal(a, [[-63.5, -63.5, 63.5, 63.5],
                                      [-127.5, -127.5, 127.5, 127.5],
                                      [-255.5, -255.5, 255.5, 255.5],
                                      [-511.5, -511.5, 511.5, 511.5]])
    scales = np.array([0.5, 1., 2.]); ratios = np.array([0.5, 1., 2.])
    a = generate_anchors_reference(256, ratios, scales)
    assert a.shape == (9, 4)
    wh = _wh(a)
    np.testing.assert_allclose(wh[:, 1] / wh[:, 0], [0.5] * 3 + [1.] * 3 + [2.] * 3)
    np.testing.assert_allclose(np.sqrt(wh[:, 1] * wh[:, 0] / 256 ** 2),
                               [0.5, 1., 2.] * 3)


def test_anchor_reference_invalid():
    """utils/anchors_test.py:96-110"""
    with pytest.raises(ValueError):
        generate_anchors_reference(1, [0.5], [0.5])


def test_fasterrcnn_anchor_grid_int32_truncation():
    """models/fasterrcnn/fasterrcnn_test.py:256-302 (quirk Q1: int32 anchors)."""
    a = fasterrcnn_anchors(32, 32, 16, [0.5, 1, 2], [0.5, 1, 2], 1)
    assert a.shape == (9216, 4) and a.dtype == np.int32
    w = a[:, 2] - a[:, 0]; h = a[:, 3] - a[:, 1]
    np.testing.assert_array_equal(np.unique(w), np.unique(h))
    assert np.unique(w * h).shape[0] == 6
    assert a[:, 0].min() == -22 and a[:, 0].max() == 29
    assert a[:, 1].min() == -22 and a[:, 1].max() == 29
    assert a[:, 2].min() == 2 and a[:, 2].max() == 53
    assert a[:, 3].min() == 2 and a[:, 3].max() == 53
    for col in range(4):
        u = np.unique(a[:, col])
        np.testing.assert_array_equal(np.diff(u), 1)


def test_default_config_anchor_reference_values():
    """truncated default reference (cross-check)."""
    a = fasterrcnn_anchors(1, 1, 256, [0.5, 1, 2], [0.25, 0.5, 1, 2], 16)
    exp = [[-44, -22, 44, 22], [-90, -44, 90, 44], [-180, -90, 180, 90], [-361, -180, 361, 180],
           [-31, -31, 31, 31], [-63, -63, 63, 63], [-127, -127, 127, 127], [-255, -255, 255, 255],
           [-22, -44, 22, 44], [-44, -90, 44, 90], [-90, -180, 90, 180], [-180, -361, 180, 361]]
    np.testing.assert_array_equal(a, exp)


# ---------------------------------------------------------------- bbox transform
def _gt_boxes(rng, n, image_size, min_size=10):
    """utils/test/gt_boxes.py:4-45 (own RNG)."""
    mx = image_size - min_size
    sizes = rng.integers(min_size, mx, size=(n, 2))
    lt = rng.integers(0, mx, size=(n, 2))
    rb = np.minimum(sizes + lt, image_size - 1)
    return np.column_stack((lt, rb))


def test_clip_boxes_known_answer():
    """utils/bbox_transform_test.py:128-146"""
    boxes = np.array([[-1, 10, 20, 20], [10, -1, 20, 20], [10, 10, 60, 20],
                      [10, 10, 20, 50], [10, 10, 20, 20], [60, 50, 60, 50]])
    exp = [[0, 10, 20, 20], [10, 0, 20, 20], [10, 10, 59, 20],
           [10, 10, 20, 49], [10, 10, 20, 20], [59, 49, 59, 49]]
    np.testing.assert_array_equal(clip_boxes(boxes, (50, 60)), exp)
    np.testing.assert_array_equal(clip_boxes_np(boxes, (50, 60)), exp)


def test_encode_decode_round_trip():
    """utils/bbox_transform_test.py:97-126,148-152"""
    rng = np.random.default_rng(0)
    p = _gt_boxes(rng, 3, 100)
    d = encode(p, p)
    np.testing.assert_array_equal(d, np.zeros((3, 4)))
    np.testing.assert_allclose(decode(p, d), p)
    for n in range(1, 2000, 117):
        gt = _gt_boxes(rng, n, 800); pr = _gt_boxes(rng, n, 800)
        np.testing.assert_allclose(decode(pr, encode(pr, gt)), gt, atol=1e-4 * 800)


# ---------------------------------------------------------------- mean subtraction
def test_subtract_channels():
    """models/base/base_network_test.py:26-44"""
    res = subtract_means(np.ones([1, 2, 2, 3], np.float32) * 255)
    np.testing.assert_allclose(res, np.ones([1, 2, 2, 3]) * [255 - 123.68, 255 - 116.78, 255 - 103.94],
                               rtol=1e-6)


# ---------------------------------------------------------------- ROI pooling
def _quadrant_map():
    a = np.ones((5, 5)); m = np.block([[a * 1, a * 2], [a * 3, a * 4]])
    return m[None, :, :, None].astype(np.float32)


def _roi(props):
    return roi_pool(np.array(props, np.float32), _quadrant_map(), (10, 10), 2, 2, 'VALID')


def test_roi_pool_basic():
    """models/fasterrcnn/roi_pool_test.py:56-111"""
    r = _roi([[1, 1, 4, 4], [6, 1, 9, 4], [1, 6, 4, 9], [6, 6, 9, 9]])
    assert r['crops'].shape == (4, 4, 4, 1) and r['roi_pool'].shape == (4, 2, 2, 1)
    for i in range(4):
        np.testing.assert_array_equal(r['roi_pool'][i, :, :, 0], np.ones((2, 2)) * (i + 1))


def test_roi_pool_without_interpolation():
    """models/fasterrcnn/roi_pool_test.py:113-175"""
    r = _roi([[3, 1, 6, 4], [1, 3, 4, 7], [5, 3, 9, 7], [3, 6, 6, 9]])['roi_pool'][..., 0]
    np.testing.assert_array_equal(r[0], [[1, 2], [1, 2]])
    np.testing.assert_array_equal(r[1], [[1, 1], [3, 3]])
    np.testing.assert_array_equal(r[2], [[2, 2], [4, 4]])
    np.testing.assert_array_equal(r[3], [[3, 4], [3, 4]])


def test_roi_pool_with_interpolation_bounds():
    """models/fasterrcnn/roi_pool_test.py:177-239"""
    r = _roi([[4, 1, 7, 4], [1, 4, 4, 8], [5, 4, 9, 8], [4, 6, 7, 9]])
    lo = [1, 1, 2, 3]; hi = [2, 3, 4, 4]
    for i in range(4):
        assert (r['roi_pool'][i] >= lo[i]).all() and (r['crops'][i] <= hi[i]).all()


# ---------------------------------------------------------------- RPN proposals
RPN_CFG = {'pre_nms_top_n': 4, 'post_nms_top_n': 3, 'nms_threshold': 1, 'min_size': 0,
           'clip_after_nms': False, 'filter_outside_anchors': False, 'apply_nms': True,
           'min_prob_threshold': 0.0}


def _rpn(anchors, prob, cfg, gt=None, pred=None):
    anchors = np.array(anchors, np.float32)
    if pred is None:
        pred = encode(anchors, np.array(gt, np.float32))
    return rpn_proposal(np.array(prob, np.float32), np.array(pred, np.float32), anchors, (40, 40), cfg)


def test_rpn_nms_threshold():
    """models/fasterrcnn/rpn_proposal_test.py:61-172"""
    gt = [[10, 10, 26, 36], [10, 10, 20, 22], [10, 11, 20, 21], [19, 30, 33, 38]]
    anchors = [[11, 13, 34, 31], [10, 10, 20, 22], [11, 13, 34, 28], [21, 29, 34, 37]]
    prob = [[0.8, 0.2], [0.1, 0.9], [0.4, 0.6], [0.2, 0.8]]
    cfg = dict(RPN_CFG, post_nms_top_n=4, nms_threshold=0.0)
    r = _rpn(anchors, prob, cfg, gt=gt)
    assert r['proposals'].shape == (2, 4)
    np.testing.assert_allclose(r['scores'], [0.9, 0.8])
    for thr in (0.3, 0.6, 0.8):
        r = _rpn(anchors, prob, dict(cfg, nms_threshold=thr), gt=gt)
        assert r['proposals'].shape == (3, 4)
        np.testing.assert_allclose(r['scores'], [0.9, 0.8, 0.2])
    r = _rpn(anchors, prob, dict(cfg, nms_threshold=1.0), gt=gt)
    assert r['proposals'].shape == (4, 4)


def test_rpn_outsiders_and_topn():
    """models/fasterrcnn/rpn_proposal_test.py:174-296"""
    gt = [[10, 10, 20, 22], [10, 10, 20, 22], [10, 10, 20, 50], [10, 10, 20, 22]]
    anchors = [[11, 13, 34, 31], [10, 10, 20, 22], [11, 13, 34, 40], [7, 13, 34, 30]]
    prob = [[0.3, 0.7], [0.4, 0.6], [0.9, 0.1], [0.8, 0.2]]
    r = _rpn(anchors, prob, RPN_CFG, gt=gt)
    assert r['proposals'].shape == (3, 4) and r['unsorted_proposals'].shape == (4, 4)
    np.testing.assert_allclose(r['scores'], [0.7, 0.6, 0.2])
    r = _rpn(anchors, prob, dict(RPN_CFG, post_nms_top_n=2), gt=gt)
    assert r['proposals'].shape == (2, 4) and r['unsorted_proposals'].shape == (4, 4)
    np.testing.assert_allclose(r['scores'], [0.7, 0.6])
    np.testing.assert_allclose(r['sorted_top_scores'], [0.7, 0.6, 0.2, 0.1])
    r = _rpn(anchors, prob, dict(RPN_CFG, post_nms_top_n=3, pre_nms_top_n=2), gt=gt)
    assert r['proposals'].shape == (2, 4) and r['sorted_top_proposals'].shape == (2, 4)
    np.testing.assert_allclose(r['scores'], [0.7, 0.6])
    np.testing.assert_allclose(r['sorted_top_scores'], [0.7, 0.6])
    r = _rpn(anchors, prob, dict(RPN_CFG, post_nms_top_n=1, pre_nms_top_n=2), gt=gt)
    assert r['proposals'].shape == (1, 4) and r['sorted_top_proposals'].shape == (2, 4)
    np.testing.assert_allclose(r['scores'], [0.7])
    np.testing.assert_allclose(r['sorted_top_scores'], [0.7, 0.6])


def test_rpn_negative_area():
    """models/fasterrcnn/rpn_proposal_test.py:298-364"""
    gt = [[10, 10, 20, 3], [10, 10, 20, 22], [10, 10, 8, 22], [10, 10, 20, 22]]
    anchors = [[11, 13, 12, 16], [10, 10, 20, 22], [11, 13, 12, 19], [7, 13, 34, 30]]
    prob = [[0.3, 0.7], [0.4, 0.6], [0.9, 0.1], [0.8, 0.2]]
    r = _rpn(anchors, prob, RPN_CFG, gt=gt)
    assert r['proposals'].shape == (2, 4) and r['unsorted_proposals'].shape == (2, 4)
    anchors = [[11, 13, 12, 16], [10, 10, 9, 9], [11, 13, 12, 28], [7, 13, 34, 30]]
    r = _rpn(anchors, prob, RPN_CFG, pred=np.zeros((4, 4)))
    assert r['unsorted_proposals'].shape == (3, 4)


def test_rpn_clipping_of_proposals():
    """models/fasterrcnn/rpn_proposal_test.py:366-454"""
    anchors = [[-20, -10, 12, 6], [2, -10, 20, 20], [0, 0, 12, 16], [2, -10, 20, 2]]
    prob = [[0.3, 0.7], [0.4, 0.6], [0.3, 0.7], [0.1, 0.9]]
    before = _rpn(anchors, prob, dict(RPN_CFG, clip_after_nms=False), pred=np.zeros((4, 4)))
    np.testing.assert_array_equal(before['unsorted_proposals'],
                                  clip_boxes(before['proposals_unclipped'], (40, 40)))
    assert (before['proposals'] >= 0).all() and (before['proposals'] < 40).all()
    after = _rpn(anchors, prob, dict(RPN_CFG, clip_after_nms=True), pred=np.zeros((4, 4)))
    np.testing.assert_array_equal(after['unsorted_proposals'], after['proposals_unclipped'])
    assert (after['proposals'] >= 0).all() and (after['proposals'] < 40).all()


def test_rpn_filter_outside_anchors():
    """models/fasterrcnn/rpn_proposal_test.py:456-500"""
    gt = [[0, 0, 10, 12], [10, 10, 20, 22], [10, 10, 20, 22], [30, 25, 39, 39], [30, 25, 39, 39]]
    anchors = [[-20, -10, 12, 6], [2, 10, 20, 20], [0, 0, 50, 16], [2, -10, 20, 50], [25, 30, 27, 33]]
    prob = [[0.3, 0.7], [0.4, 0.6], [0.3, 0.7], [0.1, 0.9], [0.2, 0.8]]
    pred = encode(np.array(anchors, np.float32), np.array(gt, np.float32))
    r = _rpn(anchors, prob, dict(RPN_CFG, filter_outside_anchors=False), pred=pred)
    assert r['all_proposals'].shape == (5, 4)
    r = _rpn(anchors, prob, dict(RPN_CFG, filter_outside_anchors=True), pred=pred)
    assert r['all_proposals'].shape == (2, 4)


# ---------------------------------------------------------------- RCNN proposals
RCNN_CFG = {'class_max_detections': 100, 'class_nms_threshold': 0.6,
            'total_max_detections': 300, 'min_prob_threshold': 0.0}


def _bbox_pred(proposed, gt_per_class):
    proposed = np.array(proposed, np.float32)
    return np.concatenate([encode(proposed, np.array(g * len(proposed), np.float32)
                                  if len(g) == 1 else np.array(g, np.float32))
                           for g in gt_per_class], axis=1)


def _rcnn(proposed, gt_per_class, cls_prob, shape=(900, 1440), cfg=RCNN_CFG, nc=3, bbox_pred=None):
    if bbox_pred is None:
        bbox_pred = _bbox_pred(proposed, gt_per_class)
    return rcnn_proposal(np.array(proposed, np.float32), bbox_pred, np.array(cls_prob, np.float32),
                         shape, nc, cfg)


def test_rcnn_no_background_class():
    """models/fasterrcnn/rcnn_proposal_test.py:75-116"""
    r = _rcnn([(85, 500, 730, 590), (50, 500, 70, 530), (700, 570, 740, 598)],
              [[(101, 101, 201, 249)], [(200, 502, 209, 532)], [(86, 571, 743, 599)]],
              [(0., .3, .3, .4), (.8, 0., 0., 2.), (.35, .3, .2, .15)])
    assert len(r['objects']) == 3
    assert set(r['proposal_label'].tolist()) == {0, 1, 2}


def test_rcnn_nms_filter():
    """models/fasterrcnn/rcnn_proposal_test.py:118-148"""
    r = _rcnn([(85, 500, 730, 590), (50, 500, 740, 570), (700, 570, 740, 598)],
              [[(101, 101, 201, 249)], [(200, 502, 209, 532)], [(86, 571, 743, 599)]],
              [(0., .1, .3, .6), (.1, .2, .25, .45), (.2, .3, .25, .25)])
    assert len(r['objects']) == 3


def test_rcnn_image_clipping():
    """models/fasterrcnn/rcnn_proposal_test.py:150-196"""
    args = ([(1300, 800, 1435, 870), (10, 1, 30, 7), (2, 870, 80, 898)],
            [[(1320, 815, 1455, 912)], [(5, -8, 31, 8)], [(-120, 910, 78, 1040)]],
            [(0., 1., 0., 0.), (.2, .25, .3, .25), (.45, 0., 0., .55)])
    for shape in ((1440, 900), (900, 1440)):
        o = _rcnn(*args, shape=shape)['objects']
        assert (o >= 0).all()
        assert (o[:, 0] < shape[1]).all() and (o[:, 2] < shape[1]).all()
        assert (o[:, 1] < shape[0]).all() and (o[:, 3] < shape[0]).all()


def test_rcnn_bbox_pred_exact():
    """models/fasterrcnn/rcnn_proposal_test.py:198-242 (exact boxes + order, atol 1e-3)"""
    gt = [[(0, 0, 1, 1)], [(5, 5, 10, 10)], [(15, 15, 20, 20)]]
    cls_prob = np.array([(0., 1., 0., 0.), (.2, .25, .3, .25), (.45, 0., 0., .55)], np.float32)
    r = _rcnn([(200, 315, 400, 370), (56, 0, 106, 4), (15, 15, 20, 20)], gt, cls_prob)
    objects = np.array([g[0] for g in gt], np.float32)
    order = cls_prob[:, 1:].max(axis=1).argsort()[::-1]
    np.testing.assert_allclose(r['objects'], objects[order], atol=1e-3)


def test_rcnn_limits():
    """models/fasterrcnn/rcnn_proposal_test.py:244-292"""
    cfg = dict(RCNN_CFG, class_max_detections=2, total_max_detections=3)
    boxes = [(0, 0, 1, 1), (5, 5, 10, 10), (15, 15, 20, 20), (25, 25, 30, 30), (35, 35, 40, 40),
             (38, 40, 65, 65), (70, 50, 90, 90), (95, 95, 100, 100), (105, 105, 110, 110)]
    prob = [(0., 1., 0.), (0., .2, .8), (0., .45, .55), (0., .55, .45), (1., 0., 0.),
            (1., 0., 0.), (0., .95, .05), (1., 0., 0.), (0., .495, .505)]
    r = _rcnn(boxes, None, prob, cfg=cfg, nc=2, bbox_pred=np.zeros((9, 8), np.float32))
    l = r['proposal_label']
    assert (l == 0).sum() <= 2 and (l == 1).sum() <= 2 and l.shape[0] <= 3


# ---------------------------------------------------------------- resize
def _resize(h, w, **ip):
    cfg = {'dataset': {'image_preprocessing': ip}}
    img, scale = preprocess(np.zeros((h, w, 3), np.float32), cfg)
    return img.shape, scale


def test_resize_only_image():
    """utils/image_test.py:118-226"""
    assert _resize(100, 1024) == ((100, 1024, 3), 1.0)
    assert _resize(100, 1024, min_size=0, max_size=2000) == ((100, 1024, 3), 1.0)
    s, sc = _resize(100, 1024, max_size=1000); assert s == (97, 1000, 3) and int(sc * 100) == 97
    s, sc = _resize(100, 1024, min_size=120); assert s == (120, 1228, 3) and int(sc * 100) == 120
    assert _resize(100, 1024, max_size=512) == ((50, 512, 3), 0.5)
    assert _resize(100, 1024, min_size=200) == ((200, 2048, 3), 2.0)
    for change in (1.1, 1.6):
        s, sc = _resize(100, 200, min_size=int(100 * change), max_size=round(200 / change))
        assert s == (100, 200, 3) and int(sc) == 1
    assert _resize(100, 200, min_size=600, max_size=1000) == ((600, 1200, 3), 6.0)
    assert _resize(2000, 600, min_size=600, max_size=1000) == ((1000, 300, 3), 0.5)


# ---------------------------------------------------------------- TF-op cross checks
def test_nms_matches_bruteforce_greedy():
    rng = np.random.default_rng(1)
    for n in (1, 7, 64, 257):
        c = rng.uniform(0, 100, (n, 2)); s = rng.uniform(5, 40, (n, 2))
        boxes = np.concatenate([c, c + s], 1).astype(np.float32)
        scores = rng.permutation(n).astype(np.float32)
        sel = T.non_max_suppression(boxes, scores, n, 0.5)
        order = np.argsort(-scores, kind='stable'); kept = []
        for i in order:
            if all(not (T.iou_tf(boxes[i], boxes[j:j + 1])[0] > np.float32(0.5)) for j in kept):
                kept.append(i)
        np.testing.assert_array_equal(sel, kept)


def test_top_k_ties_lower_index_first():
    v, i = T.top_k(np.array([1., 3., 3., 2., 3.], np.float32), 4)
    np.testing.assert_array_equal(i, [1, 2, 4, 3])


def test_same_padding_rules():
    assert T.same_pads(300, 3, 2) == (150, 0, 1)       # resnet pool1: extra pad at the end
    assert T.same_pads(38, 3, 1) == (38, 1, 1)
    assert T.same_pads(18, 3, 1, rate=6) == (18, 6, 6)
    assert T.valid_out(75, 2, 2) == 37                   # VGG pool3 (quirk Q8)
