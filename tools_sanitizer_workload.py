import os, sys, torch
sys.path.insert(0, os.getcwd())
from medtok_b200 import _lib
from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, VectorQuantizer as LookupVQ
from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
dev = torch.device("cuda:0")
torch.manual_seed(0)
# hard forward (training) at an awkward shape: n % 128 != 0, k % 128 != 0
q = NormEMAVectorQuantizer(777, 64, 0.25).to(dev).train()
z = torch.randn(3001, 64, 1, 1, device=dev)
with torch.no_grad():
    for _ in range(2):
        q(z)
    q.eval(); q(z)
# soft forward + backward, odd sizes
x = torch.randn(1001, 128, device=dev, requires_grad=True)
table = torch.randn(1500, 128, device=dev, requires_grad=True)
out, vq, cm, idx, w = _SoftQuantizeFn.apply(x, table, 0, 1500, 5, 0.25, True, _lib.Workspace())
(out.sum() + vq + cm).backward()
# lookup + usage
lq = LookupVQ(300, 32, 0.25).to(dev).eval()
lq((50, 1, 1), torch.randint(0, 300, (50,), device=dev), torch.randn(50, 32, 1, 1, device=dev))
torch.cuda.synchronize()
print("sanitizer workload done")
