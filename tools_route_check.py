"""Two-pass vs single-pass hard route on one box (run under gpurun): results and timing.

    python tools_route_check.py [--shape n,d,k]... [--steps S] [--kernels]

For every shape the training forward (hard argmin + norm-EMA update) runs from the same initial state with
mtvq_debug_set_hard_route(1) (single tensor pass with score records) and (2) (record-free pass -> rows sorted by chunk ->
refine pass fused with the row work): indices, z_q and cluster_size must be bit-equal, the loss and the updated codebook
equal to summation-order tolerance; the indices are also compared with the exact CUDA-core scan.  Then both routes are
timed (CUDA events, alternating input batches).  One JSON line per (shape, route) on stdout and in
gpurun_out/route_check.jsonl.  Diagnostic tool, not part of the product or the tests."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from medtok_b200 import _lib  # noqa: E402
from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, hard_assign  # noqa: E402


def set_route(r):
    _lib.check(_lib.lib().mtvq_debug_set_hard_route(r), "route")


def make_inputs(n, d, k, kind, dev):
    g = torch.Generator().manual_seed(0)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=torch.Generator().manual_seed(1)), dim=-1)
    if kind == "clustered":      # rows near codes, Zipf-skewed popularity, a block of duplicated codes
        w[k // 2:k // 2 + 40] = w[7]
        ranks = torch.arange(1, k + 1, dtype=torch.float64)
        prob = (1.0 / ranks) / (1.0 / ranks).sum()
        zs = []
        for s in range(2):
            gi = torch.Generator().manual_seed(10 + s)
            c = torch.multinomial(prob, n, replacement=True, generator=gi)
            zs.append((w[c] + 0.05 * torch.randn(n, d, generator=gi)).to(dev))
    else:
        zs = [torch.randn(n, d, generator=g).to(dev) for _ in range(2)]
    return zs, w.to(dev)


def run_route(route, zs, w, n, d, k, steps):
    dev = zs[0].device
    set_route(route)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    q.embedding.weight.data.copy_(w)
    with torch.no_grad():
        z_q, loss, idx = q(zs[0].view(n, d, 1, 1))
        out = dict(idx=idx.clone(), z_q=z_q.reshape(n, d).clone(), loss=loss.clone(), w=q.embedding.weight.data.clone(),
                   cs=q.cluster_size.clone())
        cnt = _lib.debug_counters(q._ws.buf, n, d, k)
        for i in range(3):
            q(zs[i % 2].view(n, d, 1, 1))
        torch.cuda.synchronize()
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in kev:          # create the event handles
            a.record(); b.record()
        torch.cuda.synchronize()
        for i in range(steps):
            _lib.set_profile_events(*kev[i])
            sev[i][0].record()
            q(zs[i % 2].view(n, d, 1, 1))
            sev[i][1].record()
        torch.cuda.synchronize()
        kms = sorted(a.elapsed_time(b) for a, b in kev)[steps // 2]
        sms = sorted(a.elapsed_time(b) for a, b in sev)[steps // 2]
    kern = None
    if "--trace" in sys.argv and route == 2:
        import numpy as np
        tr = torch.zeros(512 * 8, dtype=torch.int64, device=dev)
        _lib.check(_lib.lib().mtvq_debug_set_trace(_lib.ptr(tr)), "set_trace")
        with torch.no_grad():
            q(zs[0].view(n, d, 1, 1))
        torch.cuda.synchronize()
        _lib.check(_lib.lib().mtvq_debug_set_trace(None), "set_trace")
        t = tr.view(512, 8).cpu().numpy()[4:44].astype(np.int64)
        names = ["issue->landed", "landed->a_full", "a_full->acc_seen", "acc_seen->dec_pub", "dec_pub->rows_seen",
                 "rows_seen->rows_done", "rows_done->sums_done"]
        med = {nm: float(np.median(t[:, i + 1] - t[:, i])) for i, nm in enumerate(names)}
        med["tile_period(done)"] = float(np.median(np.diff(t[:, 7])))
        med["tile_period(issue)"] = float(np.median(np.diff(t[:, 0])))
        print("TRACE", json.dumps(med), flush=True)
        full = tr.view(512, 8).cpu().numpy().astype(np.int64)
        base = full[0, 2]
        print("TRACE rows0-7", (full[:8] - base).tolist(), flush=True)
        print("TRACE rows40-44", (full[40:45] - base).tolist(), flush=True)
        dd = full[256 + 20:256 + 44]
        dn = ["acc->hits", "hits->sorted", "sorted->pub", "pub->bookkept", "bookkept->joined", "joined->sums_start", "sums_start->sums_done"]
        print("TRACE decider", json.dumps({nm: float(np.median(dd[:, i + 1] - dd[:, i])) for i, nm in enumerate(dn)}),
              "period", float(np.median(np.diff(dd[:, 0]))), flush=True)
    if "--kernels" in sys.argv:
        from torch.profiler import profile, ProfilerActivity
        with torch.no_grad():
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for i in range(10):
                    q(zs[i % 2].view(n, d, 1, 1))
                torch.cuda.synchronize()
        kern = sorted(((e.key[:40], round(getattr(e, "device_time_total", 0) / 10.0, 1)) for e in prof.key_averages()
                       if getattr(e, "device_time_total", 0) > 0), key=lambda t: -t[1])
    return out, dict(step_ms=sms, coarse_ms=kms, near_tie=cnt["near_tie_rows"], kernels_us=kern)


def main():
    args = sys.argv[1:]
    shapes, steps, only = [], 30, 0
    i = 0
    while i < len(args):
        if args[i] == "--shape":
            p = args[i + 1].split(",")
            shapes.append((int(p[0]), int(p[1]), int(p[2]), p[3] if len(p) > 3 else "random")); i += 2
        elif args[i] == "--steps":
            steps = int(args[i + 1]); i += 2
        elif args[i] == "--only-route":
            only = int(args[i + 1]); i += 2
        else:
            i += 1
    shapes = shapes or [(1 << 20, 64, 8192, "random"), (300000, 32, 1000, "random"), (1 << 19, 64, 4096, "clustered")]
    dev = torch.device("cuda:0")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "route_check.jsonl"), "a")
    for (n, d, k, kind) in shapes:
        zs, w = make_inputs(n, d, k, kind, dev)
        res = {}
        for route in ((only,) if only else (1, 2)):
            res[route], t = run_route(route, zs, w, n, d, k, steps)
            line = dict(route=route, n=n, d=d, k=k, kind=kind, **t,
                        tflops_step=2.0 * n * k * d / (t["step_ms"] * 1e-3) / 1e12,
                        tflops_coarse=2.0 * n * k * d / (t["coarse_ms"] * 1e-3) / 1e12)
            print(json.dumps(line), flush=True)
            log.write(json.dumps(line) + "\n")
        if only:
            continue
        a, b = res[1], res[2]
        set_route(1)
        i_exact, _ = hard_assign(zs[0], w, exact=True)
        cmp = dict(n=n, d=d, k=k, kind=kind,
                   idx_equal=bool(torch.equal(a["idx"], b["idx"])),
                   idx_diff_rows=int((a["idx"] != b["idx"]).sum()),
                   idx_equal_exact=bool(torch.equal(b["idx"], i_exact)),
                   zq_equal=bool(torch.equal(a["z_q"], b["z_q"])),
                   zq_maxdiff=float((a["z_q"] - b["z_q"]).abs().max()),
                   cs_equal=bool(torch.equal(a["cs"], b["cs"])),
                   loss=(float(a["loss"]), float(b["loss"])),
                   w_maxdiff=float((a["w"] - b["w"]).abs().max()))
        print(json.dumps(cmp), flush=True)
        log.write(json.dumps(cmp) + "\n")
        del zs, res
        torch.cuda.empty_cache()
    set_route(-1)


if __name__ == "__main__":
    main()
