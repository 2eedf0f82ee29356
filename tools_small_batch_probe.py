"""Small-batch timing (run under gpurun): the sizes MedTok's training loop actually calls the quantizer with
(batch 256-1024 per GPU, K' = 7000 specific / 21000 shared, D = 64, k = 5; train_MedTok.py:387, tokenizer.py:199).
For each shape: the soft forward through the small-batch route (small.cu), through the tensor-core route forced for
the same size, and eager PyTorch evaluating the reference formula on the same GPU; same for the hard assignment.
Writes gpurun_out/small_batch.json.  Diagnostic tool, not part of the product or the tests."""
import json
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from medtok_b200 import _lib  # noqa: E402
from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, hard_assign  # noqa: E402
from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def device_us(fn, reps=20):
    """sum of the kernels' device time per call (CUPTI through torch.profiler): what the call costs the GPU, as
    opposed to `timed`, which at these sizes mostly measures how fast the host can enqueue"""
    from torch.profiler import profile, ProfilerActivity
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    tot, names = 0.0, []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = e.cuda_time_total
        if t > 0:
            tot += t
            names.append((e.key[:40], round(t / reps, 1)))
    return tot / reps, sorted(names, key=lambda x: -x[1])[:6]


def route(tensor):
    _lib.check(_lib.lib().mtvq_debug_set_soft_route(3 if tensor else 4), "route")


rows = []
d, kk = 64, 5
for k in (7000, 21000):
    table = torch.randn(k, d, device=dev)
    for n in (256, 1024, 4096):
        x = torch.randn(n, d, device=dev)
        ws = _lib.Workspace()
        with torch.no_grad():
            route(False)
            t_small = timed(lambda: _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, ws))
            o1 = _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, ws)
            route(True)
            t_tc = timed(lambda: _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, ws))
            o2 = _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, ws)
            route(False)
            same = bool(torch.equal(o1[3], o2[3]) and torch.equal(o1[0], o2[0]) and torch.equal(o1[4], o2[4]))

            def ref():   # vector_quantization_soft_one_new.py:195-214, eager
                xn = torch.nn.functional.normalize(x, p=2, dim=-1)
                cn = torch.nn.functional.normalize(table, p=2, dim=-1)
                dist = torch.sum(xn ** 2, dim=1, keepdim=True) + torch.sum(cn ** 2, dim=1) - 2 * xn @ cn.t()
                v, i = torch.topk(dist, kk, dim=-1, largest=False)
                wgt = torch.softmax(-v, dim=-1)
                zq = torch.sum(wgt.unsqueeze(-1) * cn[i], dim=1)
                vq = torch.mean((zq - x) ** 2)
                return x + (zq - x), vq
            t_ref = timed(ref)
            dv_small, top_small = device_us(lambda: _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, ws))
            route(True)
            dv_tc, _ = device_us(lambda: _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, ws))
            route(False)
            dv_ref, _ = device_us(ref)
        rows.append({"path": "soft top-5 forward", "n": n, "d": d, "k": k, "us_auto_route": t_small, "us_tensor_route": t_tc,
                     "us_eager_torch": t_ref, "device_us_auto_route": dv_small, "device_us_tensor_route": dv_tc,
                     "device_us_eager_torch": dv_ref, "kernels_auto_route": top_small, "routes_bit_equal": same})
        print(rows[-1], flush=True)
for k in (8192,):
    w = torch.nn.functional.normalize(torch.randn(k, d, device=dev), dim=-1)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    for n in (256, 1024, 4096):
        z = torch.randn(n, d, 1, 1, device=dev)
        with torch.no_grad():
            route(False)
            t_a = timed(lambda: hard_assign(z.view(n, d), w))
            t_f = timed(lambda: q(z))
            route(True)
            t_at = timed(lambda: hard_assign(z.view(n, d), w))
            t_ft = timed(lambda: q(z))
            route(False)

            def ref():
                zf = torch.nn.functional.normalize(z.view(n, d), dim=-1)
                dist = zf.pow(2).sum(1, keepdim=True) + w.pow(2).sum(1) - 2 * zf @ w.t()
                return dist.argmin(1)
            t_ref = timed(ref)
        rows.append({"path": "hard", "n": n, "d": d, "k": k, "us_assign_auto": t_a, "us_assign_tensor": t_at,
                     "us_forward_ema_auto": t_f, "us_forward_ema_tensor": t_ft, "us_eager_torch_dist_argmin": t_ref})
        print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/small_batch.json", "w"), indent=1)
