"""Run-to-run determinism of the production path under the conv switches (GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from luminoth_b200 import synth  # noqa: E402
from luminoth_b200.config import default_config  # noqa: E402
from luminoth_b200.engine import Engine  # noqa: E402

cfg = default_config('fasterrcnn', ['model.base_network.architecture=resnet_v1_50', 'model.network.num_classes=80'])
wts = synth.make_weights(cfg, seed=0, profile='peaky')
imgs = synth.make_images(8, 600, 1024, seed=33)
REPS = int(os.environ.get('REPS', '8'))


def valid_rows(out):
    boxes, scores, labels, counts = out
    return [np.concatenate([boxes[i, :counts[i]].ravel(), scores[i, :counts[i]], labels[i, :counts[i]].astype(np.float32)])
            for i in range(len(counts))]


def run(tag, env, pipeline=True):
    for k in ('LUMI_CONV_2CTA', 'LUMI_CONV_STREAMK', 'LUMI_GRAPHS', 'LUMI_CONV_EPI16', 'LUMI_NMS_LAZY'):
        os.environ.pop(k, None)
    os.environ.update(env)
    e = Engine(cfg, max_batch=8, max_h=600, max_w=1024)
    e.load_weights(wts).finalize()
    if not pipeline:
        e.set_pipeline(False)
    outs = [valid_rows(e.predict_raw(imgs)) for _ in range(REPS)]
    e.close()
    per_image = []
    for i in range(8):
        distinct = []
        for o in outs:
            if not any(len(o[i]) == len(d) and np.array_equal(o[i], d) for d in distinct):
                distinct.append(o[i])
        per_image.append(len(distinct))
    worst = 0.0
    for o in outs[1:]:
        for i in range(8):
            if len(o[i]) == len(outs[0][i]):
                worst = max(worst, float(np.abs(o[i] - outs[0][i]).max()))
            else:
                worst = float('inf')
    print('%-44s distinct results per image %s   max |diff| to run 0: %.3e' % (tag, per_image, worst), flush=True)
    return outs[0]


ref = run('pairs off, stream-K off, graphs off', {'LUMI_CONV_2CTA': '0', 'LUMI_CONV_STREAMK': '0', 'LUMI_GRAPHS': '0'})
run('default', {})
run('default, graphs off', {'LUMI_GRAPHS': '0'})
run('default, pipeline off', {}, pipeline=False)
run('pairs off', {'LUMI_CONV_2CTA': '0'})
run('pairs >= 64', {'LUMI_CONV_2CTA': '64'})
run('stream-K off', {'LUMI_CONV_STREAMK': '0'})
run('pairs off, stream-K off', {'LUMI_CONV_2CTA': '0', 'LUMI_CONV_STREAMK': '0'})
run('pairs >= 9, stream-K off, pipeline off', {'LUMI_CONV_STREAMK': '0'}, pipeline=False)
run('16-warp epilogue off', {'LUMI_CONV_EPI16': '0'})
run('lazy NMS off', {'LUMI_NMS_LAZY': '0'})
