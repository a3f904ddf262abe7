"""The committed fixtures under tests/golden/ (made by tests/golden/make_golden.py from the pinned oracle).
CPU: the oracle still reproduces them.  GPU (`-m gpu`): the CUDA path, through the C ABI, matches them."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def test_oracle_reproduces_golden_fixtures():
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(GOLD, 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for name in ('rpn_chain', 'class_chain', 'roi_pool', 'anchors'):
        want, got = load(name), mg.CASES[name]()
        assert set(want.files) == set(got)
        for k in want.files:
            np.testing.assert_array_equal(np.asarray(got[k]), want[k], err_msg='%s/%s' % (name, k))
    want, got = load('frcnn_r50_tiny'), mg.CASES['frcnn_r50_tiny']()      # conv stack: BLAS blocking may differ
    np.testing.assert_array_equal(got['labels'], want['labels'])
    np.testing.assert_allclose(got['objects'], want['objects'], atol=2e-3)
    np.testing.assert_allclose(got['probs'], want['probs'], atol=1e-5)


def test_reference_anchor_values_in_golden():
    """Quirk Q1 (int32 truncation) frozen in the fixture: first cell = the truncated reference itself."""
    a = load('anchors')['frcnn_38x64']
    assert a.dtype == np.int32
    np.testing.assert_array_equal(a[0], [-44, -22, 44, 22])
    np.testing.assert_array_equal(a[12], [-44 + 16, -22, 44 + 16, 22])
    assert load('anchors')['ssd300'].shape == (8096, 4)


@pytest.mark.gpu
def test_gpu_rpn_chain_matches_golden():
    import gpu_ops
    g = load('rpn_chain')
    cfg = {k[4:]: g[k].item() for k in g.files if k.startswith('cfg_')}
    p, s = gpu_ops.rpn_proposals(g['cls_prob'], g['bbox_pred'], g['anchors'], tuple(g['im_shape']), cfg)
    np.testing.assert_array_equal(s, g['scores'])
    np.testing.assert_array_equal(p, g['proposals'])


@pytest.mark.gpu
def test_gpu_class_chain_matches_golden():
    import gpu_ops
    g = load('class_chain')
    cfg = {'class_max_detections': 20, 'class_nms_threshold': 0.5, 'total_max_detections': 50, 'min_prob_threshold': 0.05}
    obj, lab, prob = gpu_ops.class_detections(g['proposals'], g['deltas'], g['cls_prob'], (600, 1024), 6, cfg, [0.1, 0.2])
    np.testing.assert_array_equal(lab, g['labels'])
    np.testing.assert_array_equal(prob, g['probs'])
    np.testing.assert_array_equal(obj, g['objects'])


@pytest.mark.gpu
def test_gpu_roi_pool_matches_golden():
    import gpu_ops
    g = load('roi_pool')
    y = gpu_ops.roi_pool(g['fmap'], g['rois'], (192, 256), 7, 7)
    np.testing.assert_allclose(y, g['pooled'], atol=2e-6 * max(1.0, float(np.abs(g['pooled']).max())))


@pytest.mark.gpu
def test_gpu_frcnn_tiny_matches_golden():
    from luminoth_b200 import default_config, synth
    from luminoth_b200.engine import Engine
    g = load('frcnn_r50_tiny')
    cfg = default_config('fasterrcnn', ['model.base_network.architecture=resnet_v1_50', 'model.network.num_classes=5',
                                        'model.rpn.proposals.post_nms_top_n=60',
                                        'model.rcnn.proposals.min_prob_threshold=0.05'])
    eng = Engine(cfg, max_batch=1, max_h=96, max_w=128)
    eng.load_weights(synth.make_weights(cfg, seed=3)).finalize()
    boxes, scores, labels, counts = eng.predict_raw(g['image'][None])
    k = int(counts[0])
    assert k == len(g['probs'])
    assert sorted(labels[0, :k].tolist()) == sorted(g['labels'].tolist())       # identical class assignment
    oe = np.lexsort((boxes[0, :k, 1], boxes[0, :k, 0], labels[0, :k]))
    og = np.lexsort((g['objects'][:, 1], g['objects'][:, 0], g['labels']))
    np.testing.assert_allclose(boxes[0, :k][oe], g['objects'][og], atol=5e-3)     # fp32 noise floor, see DESIGN section 3
    np.testing.assert_allclose(np.sort(scores[0, :k]), np.sort(g['probs']), atol=2e-5)
    fm = eng.get_tensor('conv_feature_map')[0]
    assert np.abs(fm - g['feature_map']).max() <= 2e-5 * np.abs(g['feature_map']).max()
    eng.close()
