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


def run(mode, n, shifted=0, fill=0, warps=0, gap=0, sync=0, ring=4, mmas=12, flags=0):
    c, f, r = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    for _ in range(2):       # second run = warm
        rc = lib.lumi_op_mma_probe(mode, n, 2000, shifted, fill, warps, gap, sync, ring, mmas, flags, ctypes.byref(c),
                                   ctypes.byref(f), ctypes.byref(r))
    assert rc == 0, rc
    print('%-20s %4d %8s %5s %6d %6d  %10.1f %14.1f %14.1f   sync %d ring %d mmas %2d flags %d -> %7.1f clk / stage' % (
        names[mode], n, 'shifted' if shifted else 'aligned', 'yes' if fill else 'no', warps, gap, c.value, f.value,
        r.value, sync, ring, mmas, flags, c.value * mmas))


print('%-20s %4s %8s %5s %6s %6s  %10s %14s %14s' % ('pattern', 'N', 'A view', 'fill', 'ldtm w', 'gap', 'clk / MMA',
                                                   'fill B/clk/SM', 'ldtm B/clk/SM'))
for n, modes in (() if '--quick' in sys.argv else ((128, (0, 1, 2, 3)), (256, (0, 1)))):
    for mode in modes:
        for shifted in (0, 1):
            for fill in (0, 1):
                if shifted and mode not in (0, 1):
                    continue
                run(mode, n, shifted, fill)
print()
for warps in (() if '--quick' in sys.argv else (1, 2, 4, 8)):
    for gap in (0, 256, 1024, 4096):
        run(1, 128, 0, 0, warps, gap)
run(1, 128, 0, 1, 8, 0)
run(1, 128, 0, 1, 8, 1024)
print()
for mmas in (12, 4):
    run(1, 128, mmas=mmas)
    run(1, 128, sync=1, mmas=mmas)
    for ring in (2, 3, 4, 6, 8):
        run(1, 128, sync=2, ring=ring, mmas=mmas)
run(1, 128, sync=2, ring=4, warps=8)
run(1, 128, sync=2, ring=3, warps=8, gap=1024)
print()
for flags in (0, 1, 2, 3, 4, 5, 7):
    run(1, 128, sync=2, ring=4, flags=flags)
run(1, 128, sync=2, ring=4, flags=5, warps=8)
run(1, 128, sync=0, flags=4)
run(1, 128, sync=1, flags=4)
