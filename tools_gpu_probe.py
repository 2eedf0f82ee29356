"""First-contact GPU probe (run under gpurun; the `micro` part needs the probe build of the library:
python -m medtok_b200.build --variant probe -DMTVQ_WITH_MICROBENCH, then
MEDTOK_VQ_LIB=medtok_b200/_build/variants/libmedtok_vq_probe.so python tools_gpu_probe.py micro): micro-benchmarks that size the epilogue, a raw
tensor-core score check, and a coarse timing of the hard forward at BASELINE config 2.
Writes gpurun_out/probe.json.  Diagnostic tool, not part of the product or the tests."""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from medtok_b200 import _lib  # noqa: E402

out = {}
dev = torch.device("cuda:0")
L = _lib.lib()
st = _lib.stream_ptr(dev)
props = torch.cuda.get_device_properties(0)
out["device"] = {"name": props.name, "sms": props.multi_processor_count, "cc": [props.major, props.minor]}

what = sys.argv[1] if len(sys.argv) > 1 else "all"

if what in ("all", "micro"):
    res = (ctypes.c_double * 2)()
    mb = {}
    for which, label in ((0, "tmem_ld_GBps"), (1, "fmnmx_Gscores")):
        for warps in (4, 8, 16, 32):
            _lib.check(L.mtvq_microbench(which, warps, 2000, res, st), "microbench")
            mb[f"{label}_w{warps}"] = res[1] / (res[0] * 1e-3) / 1e9
    for which, tag in ((2, "alone"), (3, "with_tmem_ld"), (4, "with_sts")):
        for ncols in (128, 256):
            _lib.check(L.mtvq_microbench(which, ncols // 8, 20000, res, st), "microbench umma")
            mb[f"umma_TFLOPs_n{ncols}_{tag}"] = res[1] / (res[0] * 1e-3) / 1e12
    for which, tag in ((5, "free_running"), (6, "commit_wait_each_tile")):
        for ni in (1, 2, 4):
            _lib.check(L.mtvq_microbench(which, ni, 40000, res, st), "microbench multi issuer")
            mb[f"umma_TFLOPs_{ni}issuers_{tag}"] = res[1] / (res[0] * 1e-3) / 1e12
    out["microbench"] = mb
    print(json.dumps(mb, indent=1))

if what in ("all", "scores"):
    torch.manual_seed(3)
    n, d, k = 700, 64, 1000
    z = torch.randn(n, d, device=dev)
    w = torch.nn.functional.normalize(torch.randn(k, d, device=dev), dim=-1)
    ws = _lib.Workspace().get(n, d, k, 0, dev)
    kpad = ctypes.c_int(0)
    _lib.check(L.mtvq_debug_coarse_scores(_lib.ptr(z), n, d, 0, _lib.ptr(w), k, None, 0, ctypes.byref(kpad),
                                          _lib.ptr(ws), ws.numel(), st), "kpad")
    scores = torch.full((n, kpad.value), float("nan"), device=dev)
    _lib.check(L.mtvq_debug_coarse_scores(_lib.ptr(z), n, d, 0, _lib.ptr(w), k, _lib.ptr(scores), n,
                                          ctypes.byref(kpad), _lib.ptr(ws), ws.numel(), st), "dump")
    torch.cuda.synchronize()
    zf = torch.nn.functional.normalize(z, dim=-1)
    ref = zf.half().double() @ w.half().double().t() - 0.5 * (w.double() ** 2).sum(1)
    got = scores[:, :k].double()
    err = (got - ref).abs()
    out["scores"] = {"swap": os.environ.get("MTVQ_DESC_SWAP", "0"), "max_err": err.max().item(),
                     "nan": int(torch.isnan(got).sum()), "frac_ok": float((err < 1e-4).float().mean()),
                     "got00": got[0, :4].tolist(), "ref00": ref[0, :4].tolist(),
                     "got_row5": got[5, 128:132].tolist(), "ref_row5": ref[5, 128:132].tolist()}
    print(json.dumps(out["scores"], indent=1))

if what in ("all", "time"):
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    n, d, k = 1 << 20, 64, 8192
    torch.manual_seed(0)
    z = torch.randn(n, d, 1, 1, device=dev)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    w0 = q.embedding.weight.data.clone()
    with torch.no_grad():
        for _ in range(3):
            q.embedding.weight.data.copy_(w0)
            q(z)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record()
        for i in range(10):
            q(z)
            ev[i + 1].record()
        torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(10)]
    out["time_c2_ms"] = ms
    out["c2_counters"] = _lib.debug_counters(q._ws.buf, n, d, k)
    print("hard forward C2 ms:", ms, out["c2_counters"])

if what == "kern":
    # time only the dominant kernel (coarse ARGMAX) through the profiling hook; MTVQ_DBG selects experiments
    from medtok_b200.norm_ema_quantizer import hard_assign
    n, d, k = 1 << 20, int(os.environ.get("PROBE_D", "64")), int(os.environ.get("PROBE_K", "8192"))
    torch.manual_seed(0)
    z = torch.randn(n, d, device=dev)
    w = torch.nn.functional.normalize(torch.randn(k, d, device=dev), dim=-1)
    wsx = _lib.Workspace()
    ms = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record(); torch.cuda.synchronize()
        _lib.set_profile_events(a, b)
        hard_assign(z, w, workspace=wsx)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    # slot-loop trace of CTA 0 / slot 0 (clock64 stamps)
    tr = torch.zeros(512 * 8, dtype=torch.int64, device=dev)
    _lib.check(L.mtvq_debug_set_trace(_lib.ptr(tr)), "set_trace")
    hard_assign(z, w, workspace=wsx)
    torch.cuda.synchronize()
    _lib.check(L.mtvq_debug_set_trace(None), "set_trace")
    t = tr.view(512, 8).cpu().numpy()[8:200]
    import numpy as np
    def med(a): return float(np.median(a))
    out["trace"] = {
        "issue_span (got slot -> commits issued)": med(t[:, 1] - t[:, 0]),
        "mma_done_seen (got slot -> epilogue sees tile)": med(t[:, 2] - t[:, 0]),
        "epi_to_release (sees tile -> slot released)": med(t[:, 3] - t[:, 2]),
        "release_to_reissue (slot released -> issuer got slot again)": med(t[1:, 0] - t[:-1, 3]),
        "last_granule (release -> tile reduced)": med(t[:, 4] - t[:, 3]),
        "slot_loop (issuer got slot -> next got slot)": med(t[1:, 0] - t[:-1, 0]),
    }
    print("TRACE", json.dumps(out["trace"]))
    out["kern_ms"] = ms
    out["dbg"] = os.environ.get("MTVQ_DBG", "0")
    print("DBG", out["dbg"], "d", d, "k", k, "coarse kernel ms:", [round(x, 3) for x in ms], "TFLOP/s:",
          round(2.0 * n * k * d / (min(ms) * 1e-3) / 1e12, 1), _lib.debug_counters(wsx.buf, n, d, k))

if what == "soft":
    # soft-one quantisation at BASELINE config 3 shape: N = 2^20 rows, D = 128, K' = 16384 codes per codebook, k = 5
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    n, d, k, kk = int(os.environ.get("PROBE_N", str(1 << 20))), 128, 16384, 5
    torch.manual_seed(0)
    x = torch.randn(n, d, device=dev)
    table = torch.randn(k, d, device=dev)
    wsx = _lib.Workspace()
    ms = []
    with torch.no_grad():
        for i in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out_, vq, commit, idx, wts = _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, wsx)
            b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
    cnt = _lib.debug_counters(wsx.buf, n, d, k, kk)
    flops = 2.0 * n * k * d
    out["soft_c3"] = {"ms": ms, "TFLOPs_algorithmic": flops / (min(ms) * 1e-3) / 1e12, "emb_per_s": n / (min(ms) * 1e-3), "counters": cnt}
    print("SOFT", json.dumps(out["soft_c3"]))

os.makedirs("gpurun_out", exist_ok=True)
tag = what + ("_swap" if os.environ.get("MTVQ_DESC_SWAP") == "1" else "")
json.dump(out, open(f"gpurun_out/probe_{tag}.json", "w"), indent=1)

if what == "pcie":
    # pinned host <-> device copy bandwidth of this box (what bounds bench.py's e2e leg)
    res = {}
    for mb in (16, 64, 256):
        nbytes = mb << 20
        h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        for direction in ("h2d", "d2h"):
            best = 1e9
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if direction == "h2d":
                    dbuf.copy_(h, non_blocking=True)
                else:
                    h.copy_(dbuf, non_blocking=True)
                b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            res[f"{direction}_{mb}MiB_GBps"] = nbytes / (best * 1e-3) / 1e9
    # both directions at once on two streams
    h1 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); h2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(256 << 20, dtype=torch.uint8, device=dev); d2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    res["duplex_256MiB_each_GBps_per_dir"] = (256 << 20) / (time.perf_counter() - t0) / 1e9
    out["pcie"] = res
    print("PCIE", json.dumps(res))
    json.dump(out, open("gpurun_out/probe_pcie.json", "w"), indent=1)

if what == "e2e":
    # where the host-buffer pipeline's time goes: sustained H2D (1 stream / 2 streams x half batches), then the pipeline
    from medtok_b200.host_pipeline import HostStream
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    n, d, k = 1 << 20, 64, 8192
    res = {}
    host = [torch.randn(n, d).pin_memory() for _ in range(2)]
    dbuf = [torch.empty(n, d, device=dev) for _ in range(2)]
    def timed(fn, reps=20):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(reps); torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3
    def one_stream(reps):
        for i in range(reps):
            dbuf[i % 2].copy_(host[i % 2], non_blocking=True)
    streams = [torch.cuda.Stream() for _ in range(4)]
    def multi(ns):
        def run(reps):
            per = n // ns
            for i in range(reps):
                for s in range(ns):
                    with torch.cuda.stream(streams[s]):
                        dbuf[i % 2][s * per:(s + 1) * per].copy_(host[i % 2][s * per:(s + 1) * per], non_blocking=True)
        return run
    res["h2d_ms_1stream"] = timed(one_stream)
    res["h2d_ms_2streams"] = timed(multi(2))
    res["h2d_ms_4streams"] = timed(multi(4))
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    with torch.no_grad():
        for depth in (2, 3):
            hs = HostStream(q, n, d, dev, depth=depth)
            hs.run(host[i % 2] for i in range(3)); hs.finish()
            def pipe(reps):
                hs.run(host[i % 2] for i in range(reps)); hs.finish()
            res[f"pipeline_ms_depth{depth}"] = timed(pipe)
    out["e2e"] = res
    print("E2E", json.dumps(res))
    json.dump(out, open("gpurun_out/probe_e2e.json", "w"), indent=1)

if what == "kernels":
    # per-kernel device times of the hard training forward at C2, kernels running back to back (CUPTI via
    # torch.profiler; unlike an ncu launch list the clocks stay up between the small kernels)
    from torch.profiler import profile, ProfilerActivity
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    n, d, k = int(os.environ.get("PROBE_N", str(1 << 20))), int(os.environ.get("PROBE_D", "64")), int(os.environ.get("PROBE_K", "8192"))
    torch.manual_seed(0)
    zs = [torch.randn(n, d, 1, 1, device=dev) for _ in range(2)]
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    with torch.no_grad():
        for i in range(3):
            q(zs[i % 2])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(10):
                q(zs[i % 2])
            torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = e.cuda_time_total
        if t > 0:
            rows.append((e.key[:70], t / 10.0, e.count / 10.0))
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    for name, us, cnt in rows:
        print(f"{name:70s} {us:9.1f} us/step  x{cnt:.1f}")
    print(f"{'total device time':70s} {tot:9.1f} us/step")
    out["kernels"] = [{"name": r[0], "us_per_step": r[1], "launches_per_step": r[2]} for r in rows]
    json.dump(out, open("gpurun_out/probe_kernels.json", "w"), indent=1)

if what == "stats":
    # the statistics kernels in isolation (mtvq_code_sums: hist -> scan -> scatter -> segsum), repeated
    from torch.profiler import profile, ProfilerActivity
    from medtok_b200.norm_ema_quantizer import hard_assign, code_statistics
    n, d, k = 1 << 20, 64, 8192
    torch.manual_seed(0)
    z = torch.randn(n, d, device=dev)
    w = torch.nn.functional.normalize(torch.randn(k, d, device=dev), dim=-1)
    wsx = _lib.Workspace()
    idx, den = hard_assign(z, w, workspace=wsx)
    for _ in range(2):
        code_statistics(z, den, w, idx, workspace=wsx)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            code_statistics(z, den, w, idx, workspace=wsx)
        torch.cuda.synchronize()
    for e in sorted(prof.key_averages(), key=lambda e: -getattr(e, "device_time_total", 0)):
        t = getattr(e, "device_time_total", 0)
        if t > 0:
            print(f"{e.key[:70]:70s} {t / 5.0:9.1f} us/call x{e.count / 5.0:.1f}")

if what == "softkernels":
    # per-kernel device times of one soft-one quantisation at the C3 shape (CUPTI via torch.profiler)
    from torch.profiler import profile, ProfilerActivity
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    n, d, k, kk = int(os.environ.get("PROBE_N", str(1 << 20))), int(os.environ.get("PROBE_D", "128")), int(os.environ.get("PROBE_K", "16384")), 5
    torch.manual_seed(0)
    x = torch.randn(n, d, device=dev)
    table = torch.randn(k, d, device=dev)
    wsx = _lib.Workspace()
    with torch.no_grad():
        for i in range(2):
            _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, wsx)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(5):
                _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, wsx)
            torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = e.cuda_time_total
        if t > 0:
            rows.append((e.key[:70], t / 5.0, e.count / 5.0))
    rows.sort(key=lambda r: -r[1])
    for name, us, cnt in rows:
        print(f"{name:70s} {us:9.1f} us/call  x{cnt:.1f}")
    cnt = _lib.debug_counters(wsx.buf, n, d, k, kk)
    print("counters", cnt)
    out["soft_kernels"] = [{"name": r[0], "us_per_call": r[1], "launches_per_call": r[2]} for r in rows]
    out["soft_counters"] = cnt
    json.dump(out, open("gpurun_out/probe_softkernels.json", "w"), indent=1)

if what == "softbwd":
    # soft-one forward + backward (training step of the quantisation core) at the C3 shape
    from torch.profiler import profile, ProfilerActivity
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    n, d, k, kk = int(os.environ.get("PROBE_N", str(1 << 20))), int(os.environ.get("PROBE_D", "128")), int(os.environ.get("PROBE_K", "16384")), 5
    torch.manual_seed(0)
    x = torch.randn(n, d, device=dev, requires_grad=True)
    table = torch.randn(k, d, device=dev, requires_grad=True)
    wsx = _lib.Workspace()
    def step():
        out, vq, commit, idx, wts = _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, wsx)
        (out.sum() * 1e-3 + vq + commit).backward()
        x.grad = None; table.grad = None
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None) or 0
        if t > 0:
            rows.append((e.key[:80], t / 3.0, e.count / 3.0))
    rows.sort(key=lambda r: -r[1])
    for name, us, cnt in rows[:14]:
        print(f"{name:80s} {us:9.1f} us/step  x{cnt:.1f}")
    print("total", sum(r[1] for r in rows))
