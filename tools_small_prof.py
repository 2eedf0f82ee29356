import os, sys, torch
sys.path.insert(0, os.getcwd())
from medtok_b200 import _lib
from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
dev = torch.device("cuda:0")
for k in (7000, 21000):
    x = torch.randn(256, 64, device=dev); table = torch.randn(k, 64, device=dev); ws = _lib.Workspace()
    with torch.no_grad():
        for _ in range(4):
            _SoftQuantizeFn.apply(x, table, 0, k, 5, 0.25, True, ws)
torch.cuda.synchronize()
