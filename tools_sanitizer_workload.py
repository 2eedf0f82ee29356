import os, sys, torch
sys.path.insert(0, os.getcwd())
from medtok_b200 import _lib
from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, VectorQuantizer as LookupVQ
from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
"""Workload for compute-sanitizer (memcheck / synccheck / racecheck): every kernel family at awkward shapes, both
routes (the small shapes below take the small-batch route; MTVQ route 3 forces the tensor-core kernels for them too)."""
dev = torch.device("cuda:0")
torch.manual_seed(0)
force_tensor = os.environ.get("SANITIZE_TENSOR_ROUTE") == "1"
_lib.check(_lib.lib().mtvq_debug_set_soft_route(3 if force_tensor else 4), "route")
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
# hard backward, k-means (inner-product assignment + centre update), soft top-1 / top-8
zg = torch.randn(300, 64, 1, 1, device=dev, requires_grad=True)
q.train()
zq, loss, _ = q(zg)
(zq.sum() + loss).backward()
from medtok_b200.norm_ema_quantizer import kmeans
kmeans(torch.nn.functional.normalize(torch.randn(900, 32, device=dev), dim=-1), 40, 2, use_cosine_sim=True, verbose=False)
with torch.no_grad():
    for kk, nn_, kc in ((1, 257, 333), (8, 130, 1000)):
        _SoftQuantizeFn.apply(torch.randn(nn_, 64, device=dev), torch.randn(kc, 64, device=dev), 0, kc, kk, 0.25, True,
                              _lib.Workspace())
# two-pass hard route (GROUP -> group sort -> REFINE with fused rows -> finish_flagged): forced, at the smallest size that
# gives four-row-tile blocks; ragged tail (n % 128 != 0), padding codes (k % 128 != 0), duplicated codes (near-ties
# across chunks, rows listed twice, COLLECT fallback), training and eval
if os.environ.get("SANITIZE_TWO_PASS", "1") == "1":
    _lib.check(_lib.lib().mtvq_debug_set_hard_route(2), "route")
    n2, k2 = 151700, 300
    q2 = NormEMAVectorQuantizer(k2, 32, 0.25).to(dev).train()
    w2 = torch.nn.functional.normalize(torch.randn(k2, 32, device=dev), dim=-1)
    w2[200:230] = w2[3]
    q2.embedding.weight.data.copy_(w2)
    z2 = torch.randn(n2, 32, 1, 1, device=dev)
    z2[:2000, :, 0, 0] = w2[3] + 1e-3 * torch.randn(2000, 32, device=dev)
    with torch.no_grad():
        q2(z2)
        q2.eval(); q2(z2)
    _lib.check(_lib.lib().mtvq_debug_set_hard_route(-1), "route")
torch.cuda.synchronize()
print("sanitizer workload done", "(tensor route forced)" if force_tensor else "(automatic route)")
