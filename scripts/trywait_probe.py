"""Do the lanes of one mbarrier.try_wait warp instruction ever see different answers?  (GPU box)"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch  # noqa: F401,E402
from luminoth_b200 import engine as E  # noqa: E402
lib = E.load_library(); torch.zeros(1, device='cuda')
for rounds in (20000, 200000):
    d, s, a = ctypes.c_uint(), ctypes.c_uint(), ctypes.c_uint()
    rc = lib.lumi_op_trywait_probe(rounds, ctypes.byref(d), ctypes.byref(s), ctypes.byref(a))
    print('rounds per SM %d: rc %d, rounds with per-lane attempt counts differing (all SMs): %d, max spread %d, mean attempts per round %d'
          % (rounds, rc, d.value, s.value, a.value))
