"""Diagnostic: engine with the fused fp32 feature-map copy vs the separate conversion pass (GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from luminoth_b200 import synth  # noqa: E402
from luminoth_b200.config import default_config  # noqa: E402
from luminoth_b200.engine import Engine  # noqa: E402

cfg = default_config('fasterrcnn', ['model.base_network.architecture=resnet_v1_50', 'model.network.num_classes=80'])
wts = synth.make_weights(cfg, seed=0, profile='peaky')
imgs = synth.make_images(8, 600, 1024, seed=0)


def make(fused):
    os.environ['LUMI_FMAP_F32_FUSED'] = str(fused)
    e = Engine(cfg, max_batch=8, max_h=600, max_w=1024)
    e.load_weights(wts).finalize()
    return e


def summary(tag, a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    bad = ~np.isfinite(a) | ~np.isfinite(b)
    print('%-34s shape %-22s max|d| %.3e  at %s   non-finite %d   max|ref| %.3e' % (
        tag, a.shape, np.nanmax(d) if d.size else 0.0, np.unravel_index(np.nanargmax(d), d.shape) if d.size else '-',
        int(bad.sum()), np.nanmax(np.abs(b)) if b.size else 0.0))


engs = {f: make(f) for f in (1, 0)}
prod = {}
for f, e in engs.items():
    runs = [e.predict_raw(imgs) for _ in range(3)]
    prod[f] = runs[0]
    for k in (1, 2):
        same = all(np.array_equal(x, y) for x, y in zip(runs[0], runs[k]))
        print('fused=%d production run %d identical to run 0: %s   counts %s' % (f, k, same, runs[k][3].tolist()))
summary('production boxes fused vs separate', prod[1][0], prod[0][0])
print('counts fused', prod[1][3].tolist(), 'separate', prod[0][3].tolist())
taps = {}
for f, e in engs.items():
    e.set_debug_taps(True)
    out = e.predict_raw(imgs)
    taps[f] = dict(out=out, fmap=e.get_tensor('conv_feature_map'), roi=e.get_tensor('roi_pool'),
                   feat=e.get_tensor('rcnn_features'), props=e.get_tensor('proposals'))
    print('fused=%d debug counts %s' % (f, out[3].tolist()))
    summary('  debug boxes vs production (same engine)', out[0], prod[f][0])
for k in ('fmap', 'props', 'roi', 'feat'):
    summary('debug tap %s fused vs separate' % k, taps[1][k], taps[0][k])
d = np.abs(taps[1]['feat'].astype(np.float64) - taps[0]['feat']).reshape(8, -1, taps[1]['feat'].shape[-1])
print('rcnn_features: per image max diff', d.max(axis=(1, 2)))
worst = np.argsort(d.max(axis=2).ravel())[::-1][:8]
print('worst (image, roi):', [(int(w // d.shape[1]), int(w % d.shape[1]), float(d.max(axis=2).ravel()[w])) for w in worst])
