"""Host-side pieces of ``lumi eval``, CPU only: TFRecord / SequenceExample codec, ground-truth
scaling, IoU and the AP/AR computation -- pinned on the reference's own known-answer tests where it has them
(``utils/bbox_overlap_test.py``, ``utils/image_test.py:223-244``) and on hand-computed PR curves otherwise."""
import io
import os

import numpy as np
import pytest

from luminoth_b200 import default_config
from luminoth_b200 import eval as E


def _png(img):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(img).save(b, format='PNG')
    return b.getvalue()


def test_tfrecord_sequence_example_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    recs = []
    for i in range(4):
        img = rng.integers(0, 256, (40 + i, 60, 3), dtype=np.uint8)
        recs.append({'width': 60, 'height': 40 + i, 'depth': 3, 'filename': 'im%d.png' % i, 'image_raw': _png(img),
                     'gt_boxes': [{'label': i % 3, 'xmin': 1, 'ymin': 2, 'xmax': 30 + i, 'ymax': 20},
                                  {'label': 1, 'xmin': 5, 'ymin': 5, 'xmax': 50, 'ymax': 35}][:1 + i % 2], 'pixels': img})
    E.write_tfrecord(str(tmp_path / 'val.tfrecords'), [E.make_sequence_example(r) for r in recs])
    got = list(E.read_split(str(tmp_path), 'val'))
    assert len(got) == 4
    for (im, bb, fn), r in zip(got, recs):
        assert fn == r['filename']
        np.testing.assert_array_equal(im, r['pixels'])
        assert bb.dtype == np.int32
        assert bb.tolist() == [[g['xmin'], g['ymin'], g['xmax'], g['ymax'], g['label']] for g in r['gt_boxes']]
    with pytest.raises(ValueError, match='does not exist'):
        list(E.read_split(str(tmp_path), 'train'))
    # a flipped payload byte is caught by the record checksum
    raw = bytearray(open(tmp_path / 'val.tfrecords', 'rb').read())
    raw[40] ^= 0xFF
    open(tmp_path / 'bad.tfrecords', 'wb').write(bytes(raw))
    with pytest.raises(ValueError):
        list(E.read_tfrecord(str(tmp_path / 'bad.tfrecords')))


def test_sequence_example_wire_format_known_bytes():
    """Hand-assembled protobuf bytes (field numbers of tensorflow/core/example/{example,feature}.proto): context
    feature 'width' = Int64List[7] (packed), feature list 'label' with two steps [3], [4] (unpacked varints)."""
    int7 = b'\x1a\x03\x0a\x01\x07'                                # Feature{int64_list{value: packed [7]}}
    ctx = b'\x0a' + bytes([2 + 5 + 2 + len(int7)]) + b'\x0a\x05width' + b'\x12' + bytes([len(int7)]) + int7
    step = lambda v: b'\x0a\x04\x1a\x02\x08' + bytes([v])         # FeatureList.feature = Feature{int64_list{value: v}}
    fl_val = step(3) + step(4)
    fl = b'\x0a' + bytes([2 + 5 + 2 + len(fl_val)]) + b'\x0a\x05label' + b'\x12' + bytes([len(fl_val)]) + fl_val
    buf = b'\x0a' + bytes([len(ctx)]) + ctx + b'\x12' + bytes([len(fl)]) + fl
    context, lists = E.parse_sequence_example(buf)
    assert context == {'width': [7]}
    assert lists == {'label': [[3], [4]]}


def test_bbox_overlap_reference_known_answers():
    """``utils/bbox_overlap_test.py:44-88``."""
    np.testing.assert_array_equal(E.bbox_overlap([[0, 0, 10, 10]], [[11, 11, 20, 20]]), [[0.]])
    np.testing.assert_array_equal(E.bbox_overlap([[0, 0, 10, 10], [5, 5, 10, 10]], [[11, 11, 20, 20], [15, 15, 20, 20]]),
                                  [[0., 0.], [0., 0.]])
    np.testing.assert_array_equal(E.bbox_overlap([[0, 0, 10, 10]], [[0, 0, 10, 10]]), [[1.]])
    np.testing.assert_array_equal(E.bbox_overlap([[0, 0, 10, 10], [11, 11, 20, 20]], [[0, 0, 10, 10], [11, 11, 20, 20]]),
                                  [[1., 0.], [0., 1.]])
    np.testing.assert_array_equal(E.bbox_overlap([[10, 0, 9, 10]], [[0, 0, 10, 10]]), [[0.]])
    np.testing.assert_array_equal(E.bbox_overlap([[10, 0, 7, 10]], [[0, 0, 10, 10]]), [[0.]])
    np.testing.assert_array_equal(E.bbox_overlap([[10, 0, 7, 10]], [[10, 0, 7, 10]]), [[0.]])


def test_ground_truth_scaling_reference_known_answers():
    """``utils/image_test.py:223-244`` (resize_image with boxes): 100x100 image, max_size 50 / 25."""
    cfg = default_config('fasterrcnn', [])
    cfg.dataset.image_preprocessing.min_size = None
    cfg.dataset.image_preprocessing.max_size = 50
    assert E.scaled_ground_truth((100, 100, 3), np.array([[0, 0, 10, 10, -1]]), cfg).tolist() == [[0, 0, 5, 5, -1]]
    assert E.scaled_ground_truth((100, 100, 3), np.array([[10, 10, 90, 90, -1]]), cfg).tolist() == [[5, 5, 45, 45, -1]]
    cfg.dataset.image_preprocessing.max_size = 25
    assert E.scaled_ground_truth((100, 100, 3), np.array([[0, 0, 99, 99, -1]]), cfg).tolist() == [[0, 0, 24, 24, -1]]
    ssd = default_config('ssd', [])
    assert E.scaled_ground_truth((600, 150, 3), np.array([[10, 20, 100, 400, 2]]), ssd).tolist() == [[20, 10, 200, 200, 2]]


def test_calculate_metrics_hand_computed():
    """Two images, one class.  Image 0: gt A; detections d1 (score .9, IoU 1 with A), d2 (.8, no overlap).
    Image 1: gt B, C; detection d3 (.7, IoU exactly 0.7 with B: a TP at 0.50 but not at 0.75).
    Ranked at IoU 0.50: d1 TP, d2 FP, d3 TP -> recall [1/3, 1/3, 2/3], precision [1, 1/2, 2/3] -> interpolated
    [1, 2/3, 2/3]; AP@0.5 = (34 * 1 + 33 * 2/3) / 101 (recall levels 0..0.33 -> p 1, 0.34..0.66 -> 2/3, above: none)."""
    out = {'bboxes': [np.array([[0, 0, 9, 9], [50, 50, 59, 59]], float), np.array([[0, 0, 9, 6]], float)],
           'classes': [np.array([0, 0]), np.array([0])], 'scores': [np.array([.9, .8]), np.array([.7])],
           'gt_bboxes': [np.array([[0, 0, 9, 9]]), np.array([[0, 0, 9, 9], [30, 30, 39, 39]])],
           'gt_classes': [np.array([0]), np.array([0, 0])]}
    assert abs(E.bbox_overlap([[0, 0, 9, 6]], [[0, 0, 9, 9]])[0, 0] - 0.7) < 1e-12
    ap, ar = E.calculate_metrics(out, 1)
    assert ap.shape == (1, 10) and ar.shape == (1, 10)
    want50 = (34 * 1.0 + 33 * (2.0 / 3.0)) / 101
    np.testing.assert_allclose(ap[0, 0], want50, rtol=1e-12)
    np.testing.assert_allclose(ar[0, 0], 2.0 / 3.0)
    # at IoU 0.75 only d1 matches: recall [1/3,1/3,1/3], precision [1, .5, 1/3] -> AP = 34/101, AR = 1/3
    np.testing.assert_allclose(ap[0, 5], 34 / 101, rtol=1e-12)
    np.testing.assert_allclose(ar[0, 5], 1.0 / 3.0)
    m = E.summarize_metrics(ap, ar)
    assert set(m) == {'AP@0.50', 'AP@0.75', 'AP@[0.50:0.95]', 'AR@[0.50:0.95]'}
    # a class without ground truth and without detections contributes zeros (eval.py:560-566)
    ap2, ar2 = E.calculate_metrics(out, 2)
    assert ap2[1].sum() == 0 and ar2[1].sum() == 0


def test_eval_config_mutations_and_checkpoint_listing(tmp_path):
    """``eval.py:61-76`` and ``get_checkpoints`` :222-275."""
    cfg = E.prepare_eval_config(default_config('fasterrcnn', []), 'val', 77)
    assert cfg.model.rcnn.proposals.total_max_detections == 77 and cfg.model.rcnn.proposals.min_prob_threshold == 0.0
    assert cfg.dataset.split == 'val' and cfg.dataset.data_augmentation == []
    cfg = E.prepare_eval_config(default_config('fasterrcnn', ['model.network.with_rcnn=False']), 'val', 55)
    assert cfg.model.rpn.proposals.post_nms_top_n == 55
    cfg = E.prepare_eval_config(default_config('ssd', []), 'test', 33)
    assert cfg.model.proposals.total_max_detections == 33 and cfg.model.proposals.min_prob_threshold == 0.0
    with pytest.raises(ValueError, match='Could not find checkpoint'):
        E.get_checkpoints(str(tmp_path))
    (tmp_path / 'checkpoint').write_text('model_checkpoint_path: "model.ckpt-300"\n'
                                         'all_model_checkpoint_paths: "model.ckpt-100"\n'
                                         'all_model_checkpoint_paths: "model.ckpt-300"\n'
                                         'all_model_checkpoint_paths: "model.ckpt-200"\n')
    steps = lambda **kw: [c['global_step'] for c in E.get_checkpoints(str(tmp_path), **kw)]
    assert steps() == [100, 200, 300]
    assert steps(last_only=True) == [300]
    assert steps(from_global_step=100) == [200, 300]
    assert os.path.basename(E.get_checkpoints(str(tmp_path), last_only=True)[0]['file']) == 'model.ckpt-300'
