"""BASELINE config 5 on one GPU (run under gpurun): hard forward + EMA over N x K x D, plus the soft top-5 forward,
device-resident inputs, CUDA-event timed.  Writes gpurun_out/sweep.json and a markdown table (profiles/ gets a copy).
    python tools_sweep.py [--quick]
Rows per call are capped so that z, z_q and the workspace stay below ~60 GB."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from medtok_b200 import _lib  # noqa: E402
from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer  # noqa: E402
from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn  # noqa: E402

dev = torch.device("cuda:0")
quick = "--quick" in sys.argv
PEAK = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops_sustained"] if os.path.exists("MEASURED_PEAKS.json") else 1400.0


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
ns = [1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24] + ([] if quick else [1 << 26])
ks = [4096, 8192, 16384, 32768, 65536]
for d in (64, 256):
    for k in ks:
        w = torch.nn.functional.normalize(torch.randn(k, d, device=dev), dim=-1)
        q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
        for n in ns:
            if n * d * 4 * 2 > 40e9 or (quick and n * k * d > 2e15):
                continue
            if n * k * d > 6e15:          # keep a single sweep point below ~10 s
                continue
            z = torch.randn(n, d, 1, 1, device=dev)
            q.embedding.weight.data.copy_(w)
            with torch.no_grad():
                ms = timed(lambda: q(z), 3 if n >= (1 << 24) else 7)
            cnt = _lib.debug_counters(q._ws.buf, n, d, k)
            tf = 2.0 * n * k * d / (ms * 1e-3) / 1e12
            rows.append({"path": "hard+EMA", "n": n, "d": d, "k": k, "ms": ms, "emb_per_s": n / (ms * 1e-3), "tflops": tf,
                         "frac_of_peak": tf / PEAK, "near_tie_rows": cnt["near_tie_rows"], "full_scan_rows": cnt["full_scan_rows"]})
            print(rows[-1], flush=True)
            del z
            torch.cuda.empty_cache()
        del q
for d, k in ((64, 8192), (128, 16384), (128, 49152), (256, 16384)):
    table = torch.randn(k, d, device=dev)
    for n in (1 << 16, 1 << 18, 1 << 20) + (() if quick else (1 << 22,)):
        x = torch.randn(n, d, device=dev)
        ws = _lib.Workspace()
        with torch.no_grad():
            ms = timed(lambda: _SoftQuantizeFn.apply(x, table, 0, k, 5, 0.25, True, ws), 5)
        cnt = _lib.debug_counters(ws.buf, n, d, k, 5)
        tf = 2.0 * n * k * d / (ms * 1e-3) / 1e12
        rows.append({"path": "soft top-5", "n": n, "d": d, "k": k, "ms": ms, "emb_per_s": n / (ms * 1e-3), "tflops": tf,
                     "frac_of_peak": tf / PEAK, "full_scan_rows": cnt["full_scan_rows"]})
        print(rows[-1], flush=True)
        del x, ws
        torch.cuda.empty_cache()

os.makedirs("gpurun_out", exist_ok=True)
json.dump({"peak_tflops": PEAK, "rows": rows}, open("gpurun_out/sweep.json", "w"), indent=1)
with open("gpurun_out/sweep.md", "w") as f:
    f.write("| path | N | D | K | ms | embeddings/s | algorithmic TFLOP/s | fraction of %.0f TF/s |\n|---|---|---|---|---|---|---|---|\n" % PEAK)
    for r in rows:
        f.write(f"| {r['path']} | 2^{r['n'].bit_length() - 1} | {r['d']} | {r['k']} | {r['ms']:.3f} | {r['emb_per_s']:.3e} | "
                f"{r['tflops']:.0f} | {r['frac_of_peak']:.2f} |\n")
