"""Bring-up probe of the halo-patch conv kernel: one 16 x 8 tile, one tap at a time, error map per output pixel.
Run on the GPU box: LUMI_HALO_BASEOFF={0,1} python scripts/halo_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import gpu_ops  # noqa: E402

rng = np.random.default_rng(0)
x = rng.standard_normal((1, 16, 8, 64)).astype(np.float32)
print('LUMI_HALO_BASEOFF =', os.environ.get('LUMI_HALO_BASEOFF', '0'))
for tap in list(range(9)) + [-1]:
    w = np.zeros((3, 3, 64, 64), np.float32)
    if tap < 0:
        w[:] = rng.standard_normal(w.shape) * 0.05
    else:
        w[tap // 3, tap % 3] = np.eye(64, dtype=np.float32)          # output = the input shifted by the tap
    ref = gpu_ops.conv2d(x, w, 1, 1, 'SAME', None, None, None, 0, 'tc_split')
    got = gpu_ops.conv2d(x, w, 1, 1, 'SAME', None, None, None, 0, 'tc_split_halo')
    err = np.abs(got - ref).max(axis=-1)[0]                          # [16, 8]
    print('tap %2d  max err %.3e   wrong pixels %d / 128' % (tap, err.max(), int((err > 1e-4).sum())))
    if err.max() > 1e-4:
        for row in err:
            print('   ' + ''.join('x' if e > 1e-4 else '.' for e in row))
