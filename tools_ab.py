"""A/B timing of several builds of libmedtok_vq.so on one box (run under gpurun):

    python tools_ab.py [--shape n,d,k]... [--soft] medtok_b200/_build/libmedtok_vq_r1.so medtok_b200/libmedtok_vq.so ...

Each library runs in its own subprocess (MEDTOK_VQ_LIB override, own timeout, so a hanging experimental kernel cannot
take the box down): hard training forward at the given shapes (default BASELINE config 2), device-resident inputs,
CUDA-event timed; the indices are compared with the exact CUDA-core scan.  One JSON line per (library, shape) on stdout and
in gpurun_out/ab.jsonl.  Diagnostic tool, not part of the product or the tests."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def child(shapes, soft, steps):
    import torch
    sys.path.insert(0, ROOT)
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, hard_assign
    dev = torch.device("cuda:0")
    for (n, d, k) in shapes:
        g = torch.Generator().manual_seed(0)
        zs = [torch.randn(n, d, generator=g).to(dev).view(n, d, 1, 1) for _ in range(2 if n * d * 4 > (100 << 20) else 3)]
        w = torch.nn.functional.normalize(torch.randn(k, d, generator=torch.Generator().manual_seed(1)), dim=-1).to(dev)
        q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
        q.embedding.weight.data.copy_(w)
        with torch.no_grad():
            for i in range(3):
                q(zs[i % len(zs)])
            torch.cuda.synchronize()
            kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in kev:
                a.record(); b.record()
            torch.cuda.synchronize()
            for i in range(steps):
                _lib.set_profile_events(*kev[i])
                sev[i][0].record()
                q(zs[i % len(zs)])
                sev[i][1].record()
            torch.cuda.synchronize()
            kms = sorted(a.elapsed_time(b) for a, b in kev)[steps // 2]
            sms = sorted(a.elapsed_time(b) for a, b in sev)[steps // 2]
            cnt = _lib.debug_counters(q._ws.buf, n, d, k)
            # parity of the tensor path against the exact CUDA-core scan on the current codebook
            wq = q.embedding.weight.data.clone()
            rows = zs[0].view(n, d)[: min(n, 1 << 18)]
            i1, _ = hard_assign(rows, wq)
            i2, _ = hard_assign(rows, wq, exact=True)
            same = bool(torch.equal(i1, i2))
        kern = None
        if os.environ.get("AB_KERNELS") == "1":
            from torch.profiler import profile, ProfilerActivity
            with torch.no_grad():
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    for i in range(10):
                        q(zs[i % len(zs)])
                    torch.cuda.synchronize()
            kern = sorted(((e.key[:48], round(getattr(e, "device_time_total", 0) / 10.0, 1)) for e in prof.key_averages()
                           if getattr(e, "device_time_total", 0) > 0), key=lambda t: -t[1])
        print(json.dumps({"lib": os.path.basename(_lib.SO_PATH), "path": "hard+EMA", "n": n, "d": d, "k": k, "step_ms": sms,
                          "kernels_us": kern,
                          "coarse_ms": kms, "tflops_step": 2.0 * n * k * d / (sms * 1e-3) / 1e12,
                          "tflops_coarse": 2.0 * n * k * d / (kms * 1e-3) / 1e12, "near_tie": cnt["near_tie_rows"],
                          "idx_equal_exact": same}), flush=True)
        del q, zs
        torch.cuda.empty_cache()
    if soft:
        from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
        for (n, d, k) in ((1 << 20, 128, 16384), (1 << 20, 64, 8192)):
            xs = [torch.randn(n, d, device=dev) for _ in range(2)]
            table = torch.randn(k, d, device=dev)
            ws = _lib.Workspace()
            with torch.no_grad():
                for i in range(3):
                    _SoftQuantizeFn.apply(xs[i % 2], table, 0, k, 5, 0.25, True, ws)
                torch.cuda.synchronize()
                ts = []
                for i in range(steps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); _SoftQuantizeFn.apply(xs[i % 2], table, 0, k, 5, 0.25, True, ws); b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
            ms = sorted(ts)[len(ts) // 2]
            print(json.dumps({"lib": os.path.basename(_lib.SO_PATH), "path": "soft top-5", "n": n, "d": d, "k": k, "step_ms": ms,
                              "tflops_step": 2.0 * n * k * d / (ms * 1e-3) / 1e12}), flush=True)


def main():
    args = sys.argv[1:]
    shapes, libs, soft, steps, tmo = [], [], False, 30, 240
    i = 0
    while i < len(args):
        if args[i] == "--shape":
            shapes.append(tuple(int(x) for x in args[i + 1].split(","))); i += 2
        elif args[i] == "--soft":
            soft = True; i += 1
        elif args[i] == "--steps":
            steps = int(args[i + 1]); i += 2
        elif args[i] == "--timeout":
            tmo = int(args[i + 1]); i += 2
        elif args[i] == "--child":
            child(json.loads(args[i + 1]), args[i + 2] == "1", int(args[i + 3])); return
        else:
            libs.append(args[i]); i += 1
    shapes = shapes or [(1 << 20, 64, 8192)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab.jsonl"), "a") as log:
        for lib in libs or [os.path.join(ROOT, "medtok_b200", "libmedtok_vq.so")]:
            env = dict(os.environ, MEDTOK_VQ_LIB=os.path.abspath(lib))
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(shapes), "1" if soft else "0",
                                    str(steps)], env=env, capture_output=True, text=True, timeout=tmo)
                out = r.stdout + (("\nSTDERR " + r.stderr[-2000:]) if r.returncode else "")
            except subprocess.TimeoutExpired as e:
                out = json.dumps({"lib": os.path.basename(lib), "error": f"timeout after {tmo}s",
                                  "partial": (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")})
            print(out, flush=True)
            log.write(out + "\n")


if __name__ == "__main__":
    main()
