"""tcgen05.mma issue-rate probe (GPU box): python scripts/mma_probe.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch  # noqa: F401,E402  (loads the CUDA runtime the library binds to)
from luminoth_b200 import engine as E  # noqa: E402

lib = E.load_library()
torch.zeros(1, device='cuda')
names = {0: 'one accumulator', 1: 'D1/D2/D2 (conv)', 2: 'three accumulators', 3: 'four accumulators'}
print('%-22s %5s %8s %5s  %12s %16s' % ('pattern', 'N', 'A view', 'fill', 'clk per MMA', 'fill B/clk/SM'))
for n, modes in ((128, (0, 1, 2, 3)), (256, (0, 1))):
    for mode in modes:
        for shifted in (0, 1):
            for fill in (0, 1):
                if shifted and mode not in (0, 1):
                    continue
                c, f = ctypes.c_double(), ctypes.c_double()
                for _ in range(2):       # second run = warm
                    rc = lib.lumi_op_mma_probe(mode, n, 2000, shifted, fill, ctypes.byref(c), ctypes.byref(f))
                assert rc == 0, rc
                print('%-22s %5d %8s %5s  %12.1f %16.1f' % (names[mode], n, 'shifted' if shifted else 'aligned',
                                                            'yes' if fill else 'no', c.value, f.value))
