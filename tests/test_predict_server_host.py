"""Host logic of ``lumi predict`` / ``lumi server web``, CPU only, with a stand-in network:
file resolution and class filters (``predict.py:20-63``), frame batching of ``predict_video`` (:100-171), the HTTP
surface of ``tools/server/web.py:31-56`` and the request micro-batcher."""
import io
import json
import threading
import time
import urllib.error
import urllib.request

import numpy as np
import pytest

from luminoth_b200 import default_config
from luminoth_b200 import predict as P
from luminoth_b200.server import LumiServer, MicroBatcher, parse_multipart_image


class FakeEngine(object):
    max_batch = 4
    device = 0

    def close(self):
        pass


class FakeNetwork(object):
    """predict_batch answers one object per image whose bbox encodes the image's first pixel and shape."""

    def __init__(self, config=None):
        self.engine = FakeEngine()
        self.calls = []

    def predict_batch(self, images):
        self.calls.append(len(images))
        return [[{'bbox': [int(im[0, 0, 0]), 0, im.shape[1], im.shape[0]], 'label': 'thing', 'prob': 0.9},
                 {'bbox': [0, 0, 1, 1], 'label': 'other', 'prob': 0.2}] for im in images]

    def predict_image(self, image):
        return self.predict_batch([np.asarray(image)])[0]


def test_file_resolution_and_filters(tmp_path):
    for n in ('a.jpg', 'b.PNG', 'c.txt', 'd.mp4'):
        (tmp_path / n).write_bytes(b'x')
    got = P.resolve_files((str(tmp_path),))
    assert sorted(p.split('/')[-1] for p in got) == ['a.jpg', 'b.PNG', 'd.mp4']
    assert P.resolve_files((str(tmp_path / 'a.jpg'), str(tmp_path / 'missing.jpg'), str(tmp_path / 'c.txt'))) == [str(tmp_path / 'a.jpg')]
    assert P.get_file_type('x.JPEG') == 'image' and P.get_file_type('x.avi') == 'video' and P.get_file_type('x.gif') is None
    objs = [{'label': 'cat'}, {'label': 'dog'}, {'label': 'car'}]
    assert P.filter_classes(objs, only_classes=['cat', 'car']) == [{'label': 'cat'}, {'label': 'car'}]
    assert P.filter_classes(objs, ignore_classes=['dog']) == [{'label': 'cat'}, {'label': 'car'}]
    assert P.filter_classes(objs) == objs


def test_predict_images_batches_a_directory(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    paths = []
    for i in range(5):
        img = rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)
        img[0, 0, 0] = i
        p = tmp_path / ('im%d.png' % i)
        Image.fromarray(img).save(p)
        paths.append(str(p))
    (tmp_path / 'broken.jpg').write_bytes(b'not an image')
    paths.insert(2, str(tmp_path / 'broken.jpg'))
    net = FakeNetwork()
    out = P.predict_images(net, paths, only_classes=['thing'], save_dir=str(tmp_path), echo=lambda m: None)
    assert [p for p, _ in out] == paths
    assert out[2][1] is None                                   # unreadable file: skipped like predict.py:75-79
    firsts = [o[0]['bbox'][0] for p, o in out if o is not None]
    assert firsts == [0, 1, 2, 3, 4]                           # order kept
    assert all(len(o) == 1 and o[0]['label'] == 'thing' for _, o in out if o is not None)
    assert net.calls == [5]                                    # ONE batched call for the directory
    assert (tmp_path / 'pred_im0.png').exists()


def test_predict_video_batches_frames(tmp_path):
    import cv2
    path = str(tmp_path / 'clip.avi')
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'MJPG'), 10.0, (64, 48))
    for i in range(10):
        w.write(np.full((48, 64, 3), 20 * i, np.uint8))
    w.release()
    net = FakeNetwork()
    frames = P.predict_video(net, path, ignore_classes=['other'], save_path=str(tmp_path / 'out.avi'), echo=lambda m: None)
    assert [f['frame'] for f in frames] == list(range(10))
    assert net.calls == [4, 4, 2]                              # max_batch frames per engine call
    assert all(len(f['objects']) == 1 and f['objects'][0]['bbox'][2:] == [64, 48] for f in frames)
    assert (tmp_path / 'out.mp4').exists()                     # output container hard-coded to mp4 (predict.py:104)


def test_micro_batcher_groups_concurrent_requests():
    seen = []

    def predict_batch(images):
        seen.append(len(images))
        time.sleep(0.02)
        return [int(im) * 2 for im in images]

    b = MicroBatcher(predict_batch, max_batch=4, batch_window_ms=50.0)
    out = [None] * 10
    ts = [threading.Thread(target=lambda i=i: out.__setitem__(i, b.predict(i))) for i in range(10)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    b.close()
    assert out == [2 * i for i in range(10)]
    assert sum(seen) == 10 and max(seen) <= 4 and len(seen) <= 5      # grouped, never above max_batch

    def boom(images):
        raise RuntimeError('engine failure')
    b = MicroBatcher(boom, max_batch=2, batch_window_ms=1.0)
    with pytest.raises(RuntimeError, match='engine failure'):
        b.predict(1)
    b.close()


def _multipart(field, filename, data, boundary='XBOUNDARYX'):
    body = (('--%s\r\nContent-Disposition: form-data; name="%s"; filename="%s"\r\n'
             'Content-Type: application/octet-stream\r\n\r\n' % (boundary, field, filename)).encode() + data +
            ('\r\n--%s--\r\n' % boundary).encode())
    return 'multipart/form-data; boundary=%s' % boundary, body


def _post(url, ctype, body):
    req = urllib.request.Request(url, data=body, method='POST', headers={'Content-Type': ctype})
    try:
        with urllib.request.urlopen(req, timeout=30) as r:
            return r.status, json.loads(r.read())
    except urllib.error.HTTPError as e:
        return e.code, json.loads(e.read())


def test_http_surface_matches_the_reference_server():
    from PIL import Image
    cfg = default_config('fasterrcnn', [])
    srv = LumiServer(cfg, port=0, network_factory=lambda c: FakeNetwork(c), max_batch=4, batch_window_ms=20.0)
    assert srv.config.model.rcnn.proposals.min_prob_threshold == 0.01          # web.py:97-100
    srv.start()
    base = 'http://127.0.0.1:%d' % srv.port
    try:
        with pytest.raises(urllib.error.HTTPError) as ei:
            urllib.request.urlopen(base + '/api/fasterrcnn/predict/', timeout=30)
        assert ei.value.code == 400 and json.loads(ei.value.read()) == {'error': 'Use POST method to send image.'}
        assert urllib.request.urlopen(base + '/', timeout=30).status == 200
        img = np.zeros((24, 32, 3), np.uint8)
        img[0, 0, 0] = 7
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format='PNG')
        ctype, body = _multipart('image', 'x.png', buf.getvalue())
        code, out = _post(base + '/api/fasterrcnn/predict/', ctype, body)
        assert code == 200 and out['objects'][0]['bbox'] == [7, 0, 32, 24] and len(out['objects']) == 2
        code, out = _post(base + '/api/fasterrcnn/predict/?total=1', ctype, body)
        assert code == 200 and len(out['objects']) == 1
        ctype2, body2 = _multipart('file', 'x.png', buf.getvalue())
        assert _post(base + '/api/m/predict/', ctype2, body2) == (400, {'error': 'Missing image'})
        ctype3, body3 = _multipart('image', 'x.png', b'definitely not an image')
        assert _post(base + '/api/m/predict/', ctype3, body3) == (400, {'error': 'Incompatible file type'})
        # concurrent requests share engine calls
        results = []
        ts = [threading.Thread(target=lambda: results.append(_post(base + '/api/m/predict/', ctype, body)[0])) for _ in range(8)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert results == [200] * 8
        assert sum(srv.batcher.batches) == 10 and max(srv.batcher.batches) <= 4
    finally:
        srv.close()


def test_multipart_parser():
    ctype, body = _multipart('image', 'a.jpg', b'\x00\x01binary\r\n--notboundary')
    assert parse_multipart_image(ctype, body) == b'\x00\x01binary\r\n--notboundary'
    with pytest.raises(ValueError):
        parse_multipart_image('application/json', b'{}')
