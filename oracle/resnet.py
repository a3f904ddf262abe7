"""slim ``resnet_v1_{50,101}`` trunk to ``block3`` + the R101 ``block4`` tail
(oracle only).

Reference wiring: ``luminoth/models/base/base_network.py:69-101`` (slim net,
``num_classes=None, global_pool=False, output_stride=16``), :153-177 (RGB
mean subtraction), ``luminoth/models/base/truncated_base_network.py:8-16``
(endpoint ``block3``), :56-95 (tail: only for ``resnet_v1_101``, block4 =
3 x bottleneck(2048, 512, stride 1), rate 1).  The arithmetic itself is
``tf.contrib.slim.nets.resnet_v1`` (third-party, not in /root/reference):
restated from its published definition -- see SURVEY.md Appendix A.

Weights are a dict keyed by TF variable name (TF layouts), e.g.
``truncated_base_network/resnet_v1_50/block1/unit_1/bottleneck_v1/conv1/weights``.
"""
import numpy as np

from . import tf_ops as T

RGB_MEANS = np.array([123.68, 116.78, 103.94])
BN_EPS = 1e-5

UNITS = {'resnet_v1_50': (3, 4, 6, 3), 'resnet_v1_101': (3, 4, 23, 3),
         'resnet_v1_152': (3, 8, 36, 3)}
BASE_DEPTH = (64, 128, 256, 512)
BLOCK_STRIDE = (2, 2, 2, 1)          # carried by the LAST unit of each block


def _bn(x, wts, scope):
    p = scope + '/BatchNorm/'
    return T.batch_norm_inference(x, wts[p + 'gamma'], wts[p + 'beta'],
                                  wts[p + 'moving_mean'],
                                  wts[p + 'moving_variance'], BN_EPS)


def bottleneck(x, wts, scope, depth, depth_bottleneck, stride, rate=1):
    s = scope + '/bottleneck_v1'
    if x.shape[-1] == depth:
        shortcut = x if stride == 1 else x[:, ::stride, ::stride, :]
    else:
        shortcut = T.conv2d(x, wts[s + '/shortcut/weights'], stride, 'SAME')
        shortcut = _bn(shortcut, wts, s + '/shortcut')
    r = T.conv2d(x, wts[s + '/conv1/weights'], 1, 'SAME')
    r = T.relu(_bn(r, wts, s + '/conv1'))
    r = T.conv2d_same(r, wts[s + '/conv2/weights'], stride, rate)
    r = T.relu(_bn(r, wts, s + '/conv2'))
    r = T.conv2d(r, wts[s + '/conv3/weights'], 1, 'SAME')
    r = _bn(r, wts, s + '/conv3')
    return T.relu(shortcut + r)


def subtract_means(images):
    return (images - RGB_MEANS.astype(images.dtype)).astype(images.dtype)


def trunk(images, wts, arch, scope='truncated_base_network', output_stride=16,
          endpoints=None):
    """images (N,H,W,3) float RGB 0..255 -> block3 feature map (N,H/16,W/16,1024)."""
    root = '%s/%s' % (scope, arch)
    x = subtract_means(images)
    x = T.conv2d_same(x, wts[root + '/conv1/weights'], 2)
    x = T.relu(_bn(x, wts, root + '/conv1'))
    x = T.max_pool(x, 3, 2, 'SAME')
    if endpoints is not None:
        endpoints['pool1'] = x
    target = output_stride // 4
    current, rate = 1, 1
    for b in range(3):                      # block1..block3 (endpoint block3)
        n_units = UNITS[arch][b]
        for u in range(n_units):
            unit_stride = BLOCK_STRIDE[b] if u == n_units - 1 else 1
            sc = '%s/block%d/unit_%d' % (root, b + 1, u + 1)
            if current == target:
                x = bottleneck(x, wts, sc, BASE_DEPTH[b] * 4, BASE_DEPTH[b], 1, rate)
                rate *= unit_stride
            else:
                x = bottleneck(x, wts, sc, BASE_DEPTH[b] * 4, BASE_DEPTH[b],
                               unit_stride, 1)
                current *= unit_stride
        if endpoints is not None:
            endpoints['block%d' % (b + 1)] = x
    return x


def tail(pooled, wts, arch, scope='truncated_base_network', use_tail=True):
    """``_build_tail``: identity unless arch == resnet_v1_101 (quirk Q5)."""
    if not use_tail or arch != 'resnet_v1_101':
        return pooled
    root = '%s/%s' % (scope, arch)
    x = pooled
    for u in range(3):
        x = bottleneck(x, wts, '%s/block4/unit_%d' % (root, u + 1), 2048, 512, 1, 1)
    return x
