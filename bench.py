#!/usr/bin/env python
"""bench.py -- headline benchmark of the MedTok vector-quantizer hot path on B200.

Metric (BASELINE.json): quantized embeddings/sec, forward + EMA update.
Workload (BASELINE config 2): per GPU N = 2^20 synthetic code embeddings, D = 64,
K = 8192, hard argmin + norm-EMA codebook update through the drop-in NormEMAVectorQuantizer module
(training forward).  Weak scaling: every rank quantises its own 2^20 rows against a replicated codebook
and the packed per-code statistics are exchanged (fused NVLink exchange / one NCCL all-reduce) before the
identical EMA step.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path, rank 0 only

One JSON line on stdout (rank 0):
  value      rows all ranks quantised / device time of the K driver-controlled steps (inputs resident in HBM)
  roofline   the dominant kernel (tcgen05 coarse-score pass) timed live with CUDA events on its own stream, against BOTH
             measured peaks (burst and sustained) with the one that matches the timed region's clock regime as `peak`;
             `sustained_leg` repeats the measurement over >= 2 s of back-to-back steps
  e2e        the same metric from pinned HOST buffers through the streamed public API (H2D of the rows, forward,
             D2H of z_q + code indices + loss inside the timed region)
  parity_ok  (N > 1) a side problem run before timing: replicas bit-identical and equal to a single-process run
  extra      BASELINE configs 3 (reference-faithful 4-pass soft quantisation), 4 (two-modality vocabulary sweep) and
             three config-5 points, each with its own roofline and a bounded CPU sample
  cpu_baseline  the reference's CPU path on a bounded sample (N = 1 only)
"""
import argparse
import json
import math
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
    """measured peaks of this pool's B200s (driver-written), else the fallbacks B200_PROFILING.md states"""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p["bf16_tflops_sustained"]), "hbm_gbs": float(p["hbm_gbs"]),
                "source": "MEASURED_PEAKS.json (measured on this pool: cuBLAS bf16 best-of-10 / 4 s back to back)"}
    return {"burst": 1650.0, "sustained": 1400.0, "hbm_gbs": 6400.0, "source": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


def _pick_peak(peaks, region_s, sm_mhz, sm_max_mhz):
    """burst figure for a short region at max clocks, sustained figure for a long one or when the clocks dropped"""
    at_max = sm_mhz is not None and sm_max_mhz and sm_mhz >= 0.97 * sm_max_mhz
    if region_s < 1.0 and at_max:
        return peaks["burst"], f"burst ({region_s * 1e3:.0f} ms timed region at {sm_mhz:.0f} of {sm_max_mhz:.0f} MHz)"
    return peaks["sustained"], f"sustained ({region_s:.2f} s timed region, median SM clock {sm_mhz} MHz)"


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
            self.samples.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def summary(self, t_lo=None, t_hi=None):
        clocks, maxc, reasons, power = [], None, set(), []
        for ts, s in self.samples:
            if (t_lo is not None and ts < t_lo) or (t_hi is not None and ts > t_hi):
                continue
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                clocks.append(float(f[1]))
                maxc = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        clocks.sort()
        # a window can include idle samples at its edges: report the median of the upper half (samples under load)
        load = clocks[len(clocks) // 2:] if clocks else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": maxc, "reasons": sorted(reasons), "samples": len(clocks),
                "power_w_max": max(power) if power else None}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()


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
# CPU arm: the reference's own CPU path.  Preferred: the UNMODIFIED reference modules staged under oracle/_ref
# (oracle/build_ref.py) -- kind "reference"; otherwise their ATen port oracle/torch_port.py -- kind "port".
# --------------------------------------------------------------------------------------------------
def _reference_modules():
    try:
        from oracle import ref_loader
        if ref_loader.available():
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                return ref_loader.load()
    except Exception:
        pass
    return None


def cpu_hard_throughput(sample_rows, steps=1, warmup=0, budget_s=20.0, d=DIM, k=CODES, prefer_reference=True):
    """training forward (+EMA) of the hard quantizer on `sample_rows` rows, all host threads -> (emb/s, s/step, threads, steps, kind)"""
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    z, w = synth_inputs(sample_rows, d, k, seed=0)
    z4 = z.view(sample_rows, d, 1, 1)
    mods = _reference_modules() if prefer_reference else None
    kind = "port"
    q = None
    if mods is not None:
        try:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                q = mods[0].NormEMAVectorQuantizer(n_embed=k, embedding_dim=d, beta=BETA, decay=DECAY).train()
            kind = "reference"
        except Exception:
            q = None
    cs = torch.zeros(k)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            if q is not None:
                q.embedding.weight.data.copy_(w)
                q(z4)                                    # unmodified norm_ema_quantizer.py:166-218 (materialises N x K)
            else:
                torch_port.hard_forward(z4, w.clone(), cs, BETA, DECAY, training=True, update=True, chunk_rows=4096)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if sum(times) > budget_s and len(times) >= 1:
            break
    per = sum(times) / len(times)
    return sample_rows / per, per, threads, len(times), kind


def cpu_soft_throughput(sample_rows, d, slices, kk=5, budget_s=10.0):
    """the 4-pass soft quantisation (two specific + two shared passes) on a bounded sample, ATen port"""
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(sample_rows, d, generator=g)
    n_e = max(hi for _, hi in slices)
    table = torch.randn(n_e, d, generator=g)
    t0 = time.perf_counter()
    done = 0
    for lo, hi in slices:
        torch_port.soft_core(x, table[lo:hi], kk, BETA, training=True, chunk_rows=2048)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) * len(slices) / done
    return sample_rows / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 32768
    val, per, threads, done, kind = cpu_hard_throughput(sample, steps=args.steps, warmup=min(args.warmup, 1), budget_s=150.0)
    what = ("UNMODIFIED reference module (MedTok/norm_ema_quantizer.py NormEMAVectorQuantizer.forward, staged under oracle/_ref)"
            if kind == "reference" else "ATen port of norm_ema_quantizer.py:166-218 (oracle/torch_port.py), 4096-row chunks")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": min(args.warmup, 1), "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE config 2: hard argmin + norm-EMA update, N={N_ROWS} rows/GPU, D={DIM}, K={CODES}",
                   "note": f"reference's CPU path = {what}; each step a bounded sample of {sample} rows of that workload "
                           "(the CPU cost is linear in N)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{sample} rows x {done} steps, all host threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores nearest to its GPU (NVML's affinity mask), so that the pinned host buffers of the
    e2e leg are first-touched on the GPU's own NUMA node; with 8 ranks streaming 0.5 GB per step each, remote-socket
    buffers halve the PCIe throughput.  Returns the number of cores bound to, or None when NVML / affinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class Ctx:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the quantizer has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa_cores = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)      # BEFORE constructing the module (it binds all_reduce)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, vals):
        t = torch.tensor(vals, device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()


def time_hard(ctx, q, zs, steps, warmup, kernel_every=1):
    """`steps` training forwards, device-timed between barriers; returns (total ms, sorted coarse-kernel ms, counters)"""
    from medtok_b200 import _lib
    n, d = zs[0].shape[0], zs[0].shape[1]
    with torch.no_grad():
        for i in range(warmup):
            q(zs[i % len(zs)])
        ctx.barrier()
        _lib.debug_counters(q._ws.buf, n, d, q.num_tokens)             # resets the host-side launch counter
        nk = (steps + kernel_every - 1) // kernel_every
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nk)]
        for a, b in kev:                                                 # materialise the cudaEvent handles
            a.record(); b.record()
        torch.cuda.synchronize(ctx.dev)
        ctx.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            if i % kernel_every == 0:
                _lib.set_profile_events(*kev[i // kernel_every])
            q(zs[i % len(zs)])
        t1.record()
        ctx.barrier()
        counters = _lib.debug_counters(q._ws.buf, n, d, q.num_tokens)
    return t0.elapsed_time(t1), sorted(a.elapsed_time(b) for a, b in kev), counters


def dist_parity_check(ctx):
    """world > 1: a small side problem through the SAME module path (fused NVLink exchange or packed all-reduce):
    every rank's codebook / cluster_size must be bit-identical after three training steps, the code indices must equal
    those of a single-process module run on the concatenated batch, and the replicated codebook must equal that
    single-process codebook (<= 1 ulp of a unit-norm row: fp64 partial sums are added in a different order)."""
    import hashlib
    import torch.nn as nn
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    dist, dev = ctx.dist, ctx.dev
    n, d, k, steps = 16384, 64, 1024, 3
    _, w0 = synth_inputs(8, d, k, seed=0, device=dev)
    q = NormEMAVectorQuantizer(k, d, BETA, DECAY).to(dev).train()
    q.embedding.weight.data.copy_(w0)
    local = NormEMAVectorQuantizer(k, d, BETA, DECAY).to(dev).train()
    local.all_reduce_fn = nn.Identity()                                  # single-process semantics
    local.embedding.weight.data.copy_(w0)
    idx_ok = True
    with torch.no_grad():
        for s in range(steps):
            z = synth_inputs(n, d, k, seed=7000 + 31 * s + ctx.rank, device=dev)[0]
            _, _, idx = q(z.view(n, d, 1, 1))
            allz = [torch.empty_like(z) for _ in range(ctx.world)]
            dist.all_gather(allz, z)
            _, _, idx1 = local(torch.cat(allz).view(ctx.world * n, d, 1, 1))
            idx_ok = idx_ok and bool(torch.equal(idx, idx1[ctx.rank * n:(ctx.rank + 1) * n]))
        state = torch.cat([q.embedding.weight.data.reshape(-1), q.cluster_size.reshape(-1)])
        states = [torch.empty_like(state) for _ in range(ctx.world)]
        dist.all_gather(states, state)
        replicas_identical = all(bool(torch.equal(states[0], s_)) for s_ in states[1:])
        ref = torch.cat([local.embedding.weight.data.reshape(-1), local.cluster_size.reshape(-1)])
        max_diff = float((state - ref).abs().max())
    flags = torch.tensor([float(idx_ok), float(replicas_identical), -max_diff], device=dev, dtype=torch.float64)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    idx_ok, replicas_identical, max_diff = bool(flags[0] > 0), bool(flags[1] > 0), float(-flags[2])
    checksum = hashlib.sha256(states[0].cpu().numpy().tobytes()).hexdigest()[:16]
    return {"parity_ok": bool(idx_ok and replicas_identical and max_diff <= 1.2e-7),
            "parity": {"side_problem": f"{steps} training steps, {n} rows/rank, D={d}, K={k}, through the module's exchange route "
                                       f"({'fused NVLink exchange in the finalize kernel' if q._peers is not None else 'one packed all-reduce'})",
                       "replicas_bit_identical": replicas_identical, "indices_equal_single_process": idx_ok,
                       "max_abs_diff_vs_single_process": max_diff, "codebook_checksum_sha256_16": checksum}}


def extra_workloads(ctx, peaks, cpu):
    """BASELINE configs 3, 4 and three config-5 points, measured in this process (a few steps each)."""
    from medtok_b200 import _lib
    from medtok_b200.distributed import shard_rows
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    dev, world = ctx.dev, ctx.world
    out = {}

    def roof(flops, ms, kms=None):
        tf = flops / (ms * 1e-3) / 1e12
        r = {"bound": "tensor", "achieved": tf, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": tf / peaks["sustained"],
             "frac_burst": tf / peaks["burst"], "frac_sustained": tf / peaks["sustained"], "flops_per_step_per_gpu": flops,
             "basis": "whole step (every kernel of the step), algorithmic 2*N*K*D per codebook pass"}
        if kms is not None:
            r["tensor_pass_ms"] = kms
        return r

    # ---- config 3, reference-faithful: n_e = 49152 -> text rows [:16384], graph rows [-16384:], shared = all rows, x2 ----
    try:
        n, d, n_e, kk = 1 << 20, 128, 49152, 5
        region = n_e // 3
        slices = [(0, region), (n_e - region, n_e), (0, n_e), (0, n_e)]
        g = torch.Generator().manual_seed(7 + ctx.rank)
        xs = [torch.randn(n, d, generator=g).to(dev) for _ in range(2)]          # 2 x 512 MiB > L2
        table = torch.randn(n_e, d, generator=torch.Generator().manual_seed(1)).to(dev)
        ws = _lib.Workspace()
        steps = 3

        def step(i):
            for j, (lo, hi) in enumerate(slices):
                _SoftQuantizeFn.apply(xs[(i + j) % 2], table, lo, hi, kk, BETA, True, ws)
        with torch.no_grad():
            step(0)
            ctx.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(steps):
                step(i)
            t1.record()
            ctx.barrier()
        ms = ctx.max_over_ranks([t0.elapsed_time(t1)])[0] / steps
        flops = sum(2.0 * n * (hi - lo) * d for lo, hi in slices)
        out["c3_faithful"] = {"workload": f"BASELINE config 3 as the reference lays it out (soft:147-217): n_e={n_e}, D={d}, k={kk}, "
                                          f"N={n} rows/GPU; per step 2 specific passes (K'={region}) + 2 shared passes (K'={n_e})",
                              "metric": "quantized embeddings/sec (4-pass soft top-5 forward)", "value": world * n / (ms * 1e-3),
                              "unit": UNIT, "ms_per_step": ms, "steps": steps, "roofline": roof(flops, ms)}
        del xs, table, ws
        torch.cuda.empty_cache()
        if cpu:
            v, dt, th = cpu_soft_throughput(1024, d, slices, kk, budget_s=12.0)
            out["c3_faithful"]["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": th, "kind": "port",
                                                  "sample": f"1024 rows through the same 4 passes (ATen port of soft:195-214), {dt:.1f} s"}
    except Exception as e:                                                   # never lose the headline line to an extra
        out["c3_faithful"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- config 4: >= 600K codes x 2 modalities, K = 21000, rows sharded over the ranks ----
    try:
        n_tot, d, k = 655360, 64, 21000
        lo, hi = shard_rows(n_tot, world, ctx.rank)
        n = hi - lo
        zs = [synth_inputs(n, d, k, seed=400 + 10 * ctx.rank + m, device=dev)[0].view(n, d, 1, 1) for m in range(2)]
        _, w0 = synth_inputs(8, d, k, seed=0, device=dev)
        qs = []
        for m in range(2):                                                   # one quantizer per modality (text / graph)
            q = NormEMAVectorQuantizer(k, d, BETA, DECAY).to(dev).train()
            q.embedding.weight.data.copy_(w0)
            qs.append(q)
        steps = 5
        with torch.no_grad():
            for _ in range(2):
                for m in range(2):
                    qs[m](zs[m])
            ctx.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(steps):
                for m in range(2):
                    qs[m](zs[m])
            t1.record()
            ctx.barrier()
        ms = ctx.max_over_ranks([t0.elapsed_time(t1)])[0] / steps
        flops = 2 * 2.0 * n * k * d
        out["c4"] = {"workload": f"BASELINE config 4: vocabulary sweep, {n_tot} codes x 2 modalities (text, graph), D={d}, K={k}, "
                                 f"rows sharded x{world} ({n} rows/GPU/modality), statistics exchanged + EMA finalize per modality",
                     "metric": METRIC, "value": 2 * n_tot / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
                     "scaling": "strong", "roofline": roof(flops, ms),
                     "note": "both 42 MiB modality batches stay L2-resident between steps (the sweep is smaller than L2 x 2)"}
        del zs, qs
        torch.cuda.empty_cache()
        if cpu:
            v, per, th, done, kind = cpu_hard_throughput(8192, steps=1, budget_s=10.0, d=d, k=k, prefer_reference=False)
            out["c4"]["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": th, "kind": kind,
                                         "sample": f"{done} x 8192 rows (D={d}, K={k}), ATen port of the reference forward"}
    except Exception as e:
        out["c4"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- config 5: three sweep points (weak scaling: rows per GPU fixed) ----
    pts = []
    for (n, d, k) in ((1 << 22, 64, 4096), (1 << 20, 64, 65536), (1 << 20, 256, 16384)):
        try:
            zs = [synth_inputs(n, d, k, seed=500 + 3 * ctx.rank + i, device=dev)[0].view(n, d, 1, 1) for i in range(2)]
            _, w0 = synth_inputs(8, d, k, seed=0, device=dev)
            q = NormEMAVectorQuantizer(k, d, BETA, DECAY).to(dev).train()
            q.embedding.weight.data.copy_(w0)
            steps = 5
            ms_tot, kms, _ = time_hard(ctx, q, zs, steps, 2)
            ms = ctx.max_over_ranks([ms_tot])[0] / steps
            p = {"workload": f"config-5 point: hard argmin + norm-EMA update, N={n} rows/GPU, D={d}, K={k}, {world} GPU(s)",
                 "metric": METRIC, "value": world * n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
                 "roofline": roof(2.0 * n * k * d, ms, kms[len(kms) // 2])}
            del zs, q
            torch.cuda.empty_cache()
            if cpu:
                v, per, th, done, kind = cpu_hard_throughput(4096 if k > 16384 else 8192, steps=1, budget_s=8.0, d=d, k=k,
                                                             prefer_reference=False)
                p["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": th, "kind": kind,
                                     "sample": f"{done} x {4096 if k > 16384 else 8192} rows (D={d}, K={k}), ATen port"}
            pts.append(p)
        except Exception as e:
            pts.append({"workload": f"config-5 point N={n}, D={d}, K={k}", "error": f"{type(e).__name__}: {e}"})
    out["c5"] = pts
    return out


def sustained_steps(max_over_ranks, elapsed_ms, steps, sustain_s):
    """Step count of the sustained leg.  It must be the SAME on every rank -- each step is a rendezvous in the exchange,
    so a rank that ran one step more than a peer would wait for statistics that never come -- hence it is derived from
    the max-over-ranks time of the driver leg, not from this rank's own event pair (whose microsecond skew can straddle
    a ceil() boundary: at 8 ranks that hung one r02b run inside the exchange)."""
    worst = max_over_ranks([elapsed_ms])[0]
    return max(steps, int(math.ceil(sustain_s * 1e3 / max(worst / steps, 1e-3))))


def _note(rank, msg):
    """progress notes on stderr (stdout carries the one JSON line)"""
    if os.environ.get("BENCH_PROGRESS", "1") == "1":
        print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)


def run_ours(args):
    # a rank waiting inside the fused exchange for statistics that never come traps after this many seconds (library
    # default 600): the ranks of a bench run are never more than seconds apart, so fail fast instead of hanging
    os.environ.setdefault("MEDTOK_EXCHANGE_TIMEOUT_S", "180")
    ctx = Ctx()
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    from medtok_b200 import _lib
    from medtok_b200.host_pipeline import HostStream
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer

    peaks = _peaks()
    parity = dist_parity_check(ctx) if world > 1 else {}
    _note(rank, f"parity side problem done: {parity}")

    n, d, k = args.rows or N_ROWS, args.dim or DIM, args.codes or CODES      # defaults = BASELINE config 2
    # two different input batches (2 x 256 MiB > 126 MB L2), alternated so no step re-reads cached rows
    zs = [synth_inputs(n, d, k, seed=100 * rank + i, device=dev)[0].view(n, d, 1, 1) for i in range(2)]
    _, w0 = synth_inputs(8, d, k, seed=0, device=dev)
    q = NormEMAVectorQuantizer(k, d, BETA, DECAY).to(dev).train()
    q.embedding.weight.data.copy_(w0)

    steps, warmup = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- leg 1: the driver-controlled K steps ----
    m0 = sampler.mark()
    elapsed_ms, kernel_ms, counters = time_hard(ctx, q, zs, steps, warmup)
    _note(rank, f"timed leg done: {elapsed_ms / steps:.4f} ms/step")
    m1 = sampler.mark()
    # ---- leg 2: >= 2 s of back-to-back steps (the sustained clock / power regime) ----
    sus_steps = sustained_steps(ctx.max_over_ranks, elapsed_ms, steps, args.sustain_s)
    m2 = sampler.mark()
    sus_ms, sus_kernel_ms, _ = time_hard(ctx, q, zs, sus_steps, 0, kernel_every=16)
    _note(rank, f"sustained leg done: {sus_ms / sus_steps:.4f} ms/step")
    m3 = sampler.mark()

    # ---- end-to-end: pinned host batches -> streamed upload / forward / download of z_q + indices + loss ----
    with torch.no_grad():
        host = [zs[i].view(n, d).cpu().pin_memory() for i in range(2)]
        hs = HostStream(q, n, d, dev, return_zq=True)
        hs.run(host[i % 2] for i in range(2))               # warm the pipeline
        hs.finish()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hs.run(host[i % 2] for i in range(steps))
        hs.copy_out.synchronize()
        e1.record()
        ctx.barrier()
        e2e_ms = e0.elapsed_time(e1)
    _note(rank, f"e2e leg done: {e2e_ms / steps:.4f} ms/step")
    m4 = sampler.mark()
    del host

    elapsed_ms, e2e_ms, sus_ms = ctx.max_over_ranks([elapsed_ms, e2e_ms, sus_ms])
    h2d, d2h = hs.h2d_bytes, hs.d2h_bytes
    del hs, zs
    torch.cuda.empty_cache()
    extra = None
    if not args.no_extras and (args.rows, args.dim, args.codes) == (0, 0, 0):
        extra = extra_workloads(ctx, peaks, cpu=(world == 1 and not args.no_cpu))
    sampler.stop()

    if rank == 0:
        clk1, clk2, clk_all = sampler.summary(m0, m1), sampler.summary(m2, m3), sampler.summary(m0, m4)
        value = world * n * steps / (elapsed_ms * 1e-3)
        e2e = world * n * steps / (e2e_ms * 1e-3)
        flops = 2.0 * n * k * d                              # algorithmic FLOPs of one codebook pass
        kms = kernel_ms[len(kernel_ms) // 2]
        achieved = flops / (kms * 1e-3) / 1e12
        step_tf = flops / (elapsed_ms / steps * 1e-3) / 1e12
        peak, why = _pick_peak(peaks, elapsed_ms * 1e-3, clk1["sm_mhz"], clk1["sm_max_mhz"])
        skms = sus_kernel_ms[len(sus_kernel_ms) // 2]
        s_ach = flops / (skms * 1e-3) / 1e12
        s_step_tf = flops / (sus_ms / sus_steps * 1e-3) / 1e12
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
                         "peak_choice": why, "peak_burst": peaks["burst"], "peak_sustained": peaks["sustained"],
                         "frac_burst": achieved / peaks["burst"], "frac_sustained": achieved / peaks["sustained"],
                         "step_frac": step_tf / peak, "step_frac_burst": step_tf / peaks["burst"],
                         "step_frac_sustained": step_tf / peaks["sustained"], "step_achieved": step_tf,
                         "kernel": ("mtvq::coarse_kernel<RT=4, GROUP> (record-free tcgen05 coarse-score pass of the two-pass route; "
                                    "the per-chunk REFINE pass, 1/64 of its FLOPs, is part of the step, not of this figure)"
                                    if getattr(q, "last_route", 1) == 2 else
                                    "mtvq::coarse_kernel<RT=4, ARGMAX> (tcgen05 coarse-score pass)"),
                         "route": ("two-pass" if getattr(q, "last_route", 1) == 2 else
                                   "single-pass" + (" (groups of more than MEDTOK_TWO_PASS_MAX_WORLD=4 ranks: DESIGN.md section 7)"
                                                    if world > 4 else "")),
                         "kernel_ms_median": kms, "kernel_share_of_step": kms / (elapsed_ms / steps),
                         "peak_source": peaks["source"], "flops_per_launch": flops,
                         "sustained_leg": {"seconds": sus_ms * 1e-3, "steps": sus_steps, "ms_per_step": sus_ms / sus_steps,
                                           "kernel_ms_median": skms, "kernel_launches_timed": len(sus_kernel_ms),
                                           "achieved": s_ach, "frac_of_sustained_peak": s_ach / peaks["sustained"],
                                           "frac_of_burst_peak": s_ach / peaks["burst"], "step_achieved": s_step_tf,
                                           "step_frac_of_sustained_peak": s_step_tf / peaks["sustained"],
                                           "step_frac_of_burst_peak": s_step_tf / peaks["burst"],
                                           "value": world * n * sus_steps / (sus_ms * 1e-3), "clocks": clk2}},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "note": "pinned host rows in; z_q (fp32 rows), int64 code indices and the loss out -- the reference's "
                            "whole return tuple; 2-deep upload / compute / download pipeline"
                            + (f"; each rank bound to the {ctx.numa_cores} cores of its GPU's NUMA node" if ctx.numa_cores else "")},
            "gpu_launches": counters["launches"],
            "clocks": {**clk_all, "driver_leg": clk1},
        }
        line.update(parity)
        if extra is not None:
            line["extra"] = extra
        if world == 1 and not args.no_cpu:
            # bounded CPU sample of the same workload on this box's host cores (rank 0, N=1 only)
            val, per, threads, done, kind = cpu_hard_throughput(32768, steps=2, warmup=0, budget_s=25.0, d=d, k=k)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"{done} x 32768 rows (D={d}, K={k}), "
                                              + ("unmodified reference module" if kind == "reference" else "ATen port of the reference forward")}
        print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# BASELINE config 3, single codebook slice (`--workload c3`): soft-one top-5 quantisation of N = 2^20 rows, D = 128,
# against one K' = 16384-row codebook slice -- the quantisation core of VectorQuantizer.specific_embedding
# (vector_quantization_soft_one_new.py:195-214).  The reference-faithful 4-pass layout is `extra.c3_faithful` above.
# --------------------------------------------------------------------------------------------------
def run_soft(args):
    ctx = Ctx()
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    from medtok_b200 import _lib
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    n, d, k, kk = 1 << 20, 128, 16384, 5
    g = torch.Generator().manual_seed(7 + rank)
    xs = [torch.randn(n, d, generator=g).to(dev) for _ in range(2)]          # 2 x 512 MiB > L2
    table = torch.randn(k, d, generator=torch.Generator().manual_seed(1)).to(dev)
    wsx = _lib.Workspace()
    peaks = _peaks()
    steps, warmup = args.steps, max(args.warmup, 3)
    with torch.no_grad():
        for i in range(warmup):
            _SoftQuantizeFn.apply(xs[i % 2], table, 0, k, kk, BETA, True, wsx)
        ctx.barrier()
        _lib.debug_counters(wsx.buf, n, d, k, kk)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in kev:
            a.record(); b.record()
        sampler = ClockSampler(ctx.local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        ctx.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            _lib.set_profile_events(*kev[i])
            _SoftQuantizeFn.apply(xs[i % 2], table, 0, k, kk, BETA, True, wsx)
        t1.record()
        ctx.barrier()
        sampler.stop()
        clocks = sampler.summary() if rank == 0 else None
        elapsed_ms = t0.elapsed_time(t1)
        counters = _lib.debug_counters(wsx.buf, n, d, k, kk)
        kernel_ms = sorted(a.elapsed_time(b) for a, b in kev)
    elapsed_ms = ctx.max_over_ranks([elapsed_ms])[0]
    if rank == 0:
        flops = 2.0 * n * k * d
        kms = kernel_ms[len(kernel_ms) // 2]
        peak, why = _pick_peak(peaks, elapsed_ms * 1e-3, clocks["sm_mhz"], clocks["sm_max_mhz"])
        ach = flops / (kms * 1e-3) / 1e12
        print(json.dumps({
            "metric": "quantized embeddings/sec (soft top-5 forward)", "value": world * n * steps / (elapsed_ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"BASELINE config 3 (one slice): soft-one top-{kk} quantisation, N={n} rows/GPU, D={d}, K'={k}",
                       "arithmetic": "fp32 results (exact re-scoring); fp16-operand / fp32-accumulate tcgen05 candidate pass",
                       "l2": "two alternating 512 MiB input batches", "full_scan_rows_last_step": counters["full_scan_rows"]},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                         "peak_choice": why, "frac_burst": ach / peaks["burst"], "frac_sustained": ach / peaks["sustained"],
                         "kernel": "mtvq::coarse_kernel<RT=2, TOPK> (single-pass tcgen05 top-k candidate pass)",
                         "kernel_ms_median": kms, "kernel_share_of_step": kms / (elapsed_ms / steps),
                         "peak_source": peaks["source"], "flops_per_launch": flops},
            "gpu_launches": counters["launches"], "clocks": clocks}))
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3 / 4 / 5 extras")
    ap.add_argument("--sustain-s", type=float, default=2.2, help="length of the sustained leg in seconds")
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (config-5 sweep points; default 2^20)")
    ap.add_argument("--dim", type=int, default=0, help="embedding dim (default 64)")
    ap.add_argument("--codes", type=int, default=0, help="codebook size (default 8192)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2 = BASELINE config 2 (headline: hard argmin + EMA); c3 = config 3, one slice (soft top-5 forward)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c3":
        run_soft(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
