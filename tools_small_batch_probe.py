import os, sys, torch
sys.path.insert(0, os.getcwd())
from medtok_b200 import _lib
from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, hard_assign
from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
dev = torch.device("cuda:0")
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
d, k = 64, 8192
q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
w = torch.nn.functional.normalize(torch.randn(k, d, device=dev), dim=-1)
table = torch.randn(21000, d, device=dev)
for n in (256, 1024, 4096, 16384):
    z = torch.randn(n, d, 1, 1, device=dev)
    x = torch.randn(n, d, device=dev)
    ws = _lib.Workspace()
    with torch.no_grad():
        t_hard = timed(lambda: q(z))
        t_exact = timed(lambda: hard_assign(z.view(n, d), w, exact=True))
        t_tc = timed(lambda: hard_assign(z.view(n, d), w))
        t_soft = timed(lambda: _SoftQuantizeFn.apply(x, table, 0, 7000, 5, 0.25, True, ws))
        # plain PyTorch on the same GPU, reference formula (hard): for orientation only
        def ref():
            zf = torch.nn.functional.normalize(z.view(n, d), dim=-1)
            dist = zf.pow(2).sum(1, keepdim=True) + w.pow(2).sum(1) - 2 * zf @ w.t()
            return dist.argmin(1)
        t_ref = timed(ref)
    print(f"n={n}: hard fwd+EMA {t_hard:.0f} us | assign tc {t_tc:.0f} us, exact {t_exact:.0f} us | soft(K'=7000) {t_soft:.0f} us | torch dist+argmin only {t_ref:.0f} us", flush=True)
