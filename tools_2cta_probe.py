import ctypes, os, sys, torch
sys.path.insert(0, os.getcwd())
from medtok_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
res = (ctypes.c_double * 2)()
_lib.check(L.mtvq_microbench(7, 4, 1, res, _lib.stream_ptr(dev)), "2cta probe")
print("2-CTA MMA probe: wrong elements", res[0], "of", res[1])
