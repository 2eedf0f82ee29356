#!/usr/bin/env python
"""bench.py -- headline benchmark of the MedTok vector-quantizer hot path on B200.

Metric (BASELINE.json): quantized embeddings/sec, forward + EMA update.
Workload (BASELINE config 2): per GPU N = 2^20 synthetic code embeddings, D = 64,
K = 8192, hard argmin + norm-EMA codebook update through the drop-in NormEMAVectorQuantizer module
(training forward).  Weak scaling: every rank quantises its own 2^20 rows against a replicated codebook
and the packed per-code statistics are summed with one NCCL all-reduce before the identical EMA step.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (ATen port), rank 0 only

One JSON line on stdout (rank 0).  `value`  = rows all ranks quantised / device time (inputs resident in HBM),
`e2e` = the same metric from pinned HOST buffers through the streamed public API (H2D + forward + D2H of the
code indices and loss inside the timed region), `roofline` = the dominant kernel (tcgen05 coarse-score pass)
timed live with CUDA events on its own stream, `cpu_baseline` = the oracle's ATen port on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS = 1 << 20
DIM = 64
CODES = 8192
BETA = 0.25
DECAY = 0.99
METRIC = "quantized embeddings/sec (fwd+EMA)"
UNIT = "embeddings/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (measured)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


def _profile_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None"""
    path = os.path.join(ROOT, "profiles", "coarse_kernel_summary.json")
    try:
        return float(json.load(open(path))["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, maxc, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                clocks.append(float(f[1]))
                maxc = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        clocks.sort()
        # the sampler also sees the idle tail: report the median of the upper half (samples under load)
        load = clocks[len(clocks) // 2:] if clocks else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": maxc, "reasons": sorted(reasons), "samples": len(clocks)}


def synth_inputs(n, d, k, seed, device=None):
    """SURVEY 8d synthetic inputs: z = randn(N, D) (the module l2-normalises), codebook = l2norm(randn(K, D))."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, d, generator=g)
    g2 = torch.Generator().manual_seed(1)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g2), p=2, dim=-1)
    if device is not None:
        z, w = z.to(device), w.to(device)
    return z, w


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU path (pure PyTorch -> its ATen port in oracle/torch_port.py)
# --------------------------------------------------------------------------------------------------
def cpu_port_throughput(sample_rows, steps=1, warmup=0, budget_s=20.0, d=DIM, k=CODES):
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    z, w = synth_inputs(sample_rows, d, k, seed=0)
    z4 = z.view(sample_rows, d, 1, 1)
    cs = torch.zeros(k)
    times = []
    for i in range(warmup + steps):
        wi = w.clone()
        t0 = time.perf_counter()
        torch_port.hard_forward(z4, wi, cs, BETA, DECAY, training=True, update=True, chunk_rows=4096)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if sum(times) > budget_s and len(times) >= 1:
            break
    per = sum(times) / len(times)
    return sample_rows / per, per, threads, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 32768
    val, per, threads, done = cpu_port_throughput(sample, steps=args.steps, warmup=min(args.warmup, 1), budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": min(args.warmup, 1), "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"hard argmin + norm-EMA update, D={DIM}, K={CODES} (BASELINE config 2 shape)",
                   "note": "reference's CPU path = ATen ops of norm_ema_quantizer.py:166-218 (oracle/torch_port.py), "
                           f"each step a bounded sample of {sample} rows in 4096-row chunks"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} rows x {done} steps, all host threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the quantizer has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)      # BEFORE constructing the module (it binds all_reduce)

    from medtok_b200 import _lib
    from medtok_b200.host_pipeline import HostStream
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer

    n, d, k = args.rows or N_ROWS, args.dim or DIM, args.codes or CODES      # defaults = BASELINE config 2
    # two different input batches (2 x 256 MiB > 126 MB L2), alternated so no step re-reads cached rows
    zs = [synth_inputs(n, d, k, seed=100 * rank + i, device=dev)[0].view(n, d, 1, 1) for i in range(2)]
    _, w0 = synth_inputs(8, d, k, seed=0, device=dev)
    q = NormEMAVectorQuantizer(k, d, BETA, DECAY).to(dev).train()
    q.embedding.weight.data.copy_(w0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    steps, warmup = args.steps, max(args.warmup, 3)
    with torch.no_grad():
        for i in range(warmup):
            q(zs[i % 2])
        barrier()
        _lib.debug_counters(q._ws.buf, n, d, k)             # resets the host-side launch counter
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in kev:                                     # materialise the cudaEvent handles
            a.record(); b.record()
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            _lib.set_profile_events(*kev[i])
            q(zs[i % 2])
        t1.record()
        barrier()
        elapsed_ms = t0.elapsed_time(t1)
        counters = _lib.debug_counters(q._ws.buf, n, d, k)
        kernel_ms = sorted(a.elapsed_time(b) for a, b in kev)

        # ---- end-to-end: pinned host batches -> streamed upload / forward / download of indices + loss ----
        host = [zs[i].view(n, d).cpu().pin_memory() for i in range(2)]
        hs = HostStream(q, n, d, dev)
        hs.run(host[i % 2] for i in range(2))               # warm the pipeline
        hs.finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hs.run(host[i % 2] for i in range(steps))
        hs.copy_out.synchronize()
        e1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None      # sampled across both timed regions (device-resident + e2e)
        e2e_ms = e0.elapsed_time(e1)

    tm = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)           # device-timed, max over ranks
    elapsed_ms, e2e_ms = tm.tolist()

    if rank == 0:
        value = world * n * steps / (elapsed_ms * 1e-3)
        e2e = world * n * steps / (e2e_ms * 1e-3)
        flops = 2.0 * n * k * d                              # algorithmic FLOPs of one codebook pass
        kms = kernel_ms[len(kernel_ms) // 2]
        achieved = flops / (kms * 1e-3) / 1e12
        peak, peak_src = _peaks()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": (f"BASELINE config 2: hard argmin + norm-EMA update, N={n} rows/GPU, D={d}, K={k}"
                                    if (n, d, k) == (N_ROWS, DIM, CODES) else
                                    f"config-5 sweep point: hard argmin + norm-EMA update, N={n} rows/GPU, D={d}, K={k}"),
                       "arithmetic": "results are the reference's fp32 results (exact re-scoring in reference order); the "
                                     "tcgen05 candidate pass uses fp16 operands with fp32 accumulation",
                       "rows_per_gpu": n, "dim": d, "codes": k, "parallelism": f"rows sharded x{world}, codebook replicated",
                       "l2": "two alternating 256 MiB input batches (> 126 MB L2) between timed iterations",
                       "near_tie_rows_last_step": counters["near_tie_rows"],
                       "full_scan_rows_last_step": counters["full_scan_rows"]},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": _profile_traffic(),
                         "kernel": "mtvq::coarse_kernel<RT=4, ARGMAX> (tcgen05 coarse-score pass)",
                         "kernel_ms_median": kms, "kernel_share_of_step": kms / (elapsed_ms / steps),
                         "peak_source": peak_src, "flops_per_launch": flops},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes * world,
                    "d2h_bytes_per_step": hs.d2h_bytes * world,
                    "note": "pinned host rows in, int64 code indices + loss out, 2-deep upload/compute/download pipeline"},
            "gpu_launches": counters["launches"],
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            # bounded CPU sample of the same workload on this box's host cores (rank 0, N=1 only)
            val, per, threads, done = cpu_port_throughput(32768, steps=2, warmup=0, budget_s=25.0, d=d, k=k)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{done} x 32768 rows (D={d}, K={k}), ATen port of the reference forward"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# BASELINE config 3 (not the headline line; `--workload c3`): soft-one top-5 quantisation of N = 2^20 rows, D = 128,
# against one K' = 16384-row codebook slice -- the quantisation core of VectorQuantizer.specific_embedding
# (vector_quantization_soft_one_new.py:195-214): normalise, distances, top-k, softmax, weighted sum, STE, losses.
# --------------------------------------------------------------------------------------------------
def run_soft(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from medtok_b200 import _lib
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    n, d, k, kk = 1 << 20, 128, 16384, 5
    g = torch.Generator().manual_seed(7 + rank)
    xs = [torch.randn(n, d, generator=g).to(dev) for _ in range(2)]          # 2 x 512 MiB > L2
    table = torch.randn(k, d, generator=torch.Generator().manual_seed(1)).to(dev)
    wsx = _lib.Workspace()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    steps, warmup = args.steps, max(args.warmup, 3)
    with torch.no_grad():
        for i in range(warmup):
            _SoftQuantizeFn.apply(xs[i % 2], table, 0, k, kk, BETA, True, wsx)
        barrier()
        _lib.debug_counters(wsx.buf, n, d, k, kk)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in kev:
            a.record(); b.record()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            _lib.set_profile_events(*kev[i])
            _SoftQuantizeFn.apply(xs[i % 2], table, 0, k, kk, BETA, True, wsx)
        t1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        elapsed_ms = t0.elapsed_time(t1)
        counters = _lib.debug_counters(wsx.buf, n, d, k, kk)
        kernel_ms = sorted(a.elapsed_time(b) for a, b in kev)
    tm = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    elapsed_ms = tm.item()
    if rank == 0:
        flops = 2.0 * n * k * d
        kms = kernel_ms[len(kernel_ms) // 2]
        peak, peak_src = _peaks()
        print(json.dumps({
            "metric": "quantized embeddings/sec (soft top-5 forward)", "value": world * n * steps / (elapsed_ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"BASELINE config 3: soft-one top-{kk} quantisation, N={n} rows/GPU, D={d}, K'={k}",
                       "arithmetic": "fp32 results (exact re-scoring); fp16-operand / fp32-accumulate tcgen05 candidate pass",
                       "l2": "two alternating 512 MiB input batches", "full_scan_rows_last_step": counters["full_scan_rows"]},
            "roofline": {"bound": "tensor", "achieved": flops / (kms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (kms * 1e-3) / 1e12 / peak, "traffic": None,
                         "kernel": "mtvq::coarse_kernel<RT=2, TOPK> (single-pass tcgen05 top-k candidate pass)",
                         "kernel_ms_median": kms, "kernel_share_of_step": kms / (elapsed_ms / steps),
                         "peak_source": peak_src, "flops_per_launch": flops},
            "gpu_launches": counters["launches"], "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (config-5 sweep points; default 2^20)")
    ap.add_argument("--dim", type=int, default=0, help="embedding dim (default 64)")
    ap.add_argument("--codes", type=int, default=0, help="codebook size (default 8192)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2 = BASELINE config 2 (headline: hard argmin + EMA); c3 = config 3 (soft top-5 forward)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c3":
        run_soft(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
