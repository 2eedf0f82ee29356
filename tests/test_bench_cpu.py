"""Host logic of bench.py that must hold on every rank (no GPU): the sustained leg's step count is a collective
decision -- each step of the multi-GPU path is a rendezvous in the statistics exchange, so two ranks that derived
different counts from their own (microsecond-skewed) timings would park the job inside the exchange."""
import math
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, local_ms, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def max_over_ranks(vals):                       # bench.Ctx.max_over_ranks on CPU tensors
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    q.put((rank, bench.sustained_steps(max_over_ranks, local_ms[rank], 20, 2.2)))
    dist.barrier()
    dist.destroy_process_group()


def test_sustained_leg_step_count_is_the_same_on_every_rank():
    # 20 steps; the two ranks' event pairs straddle the ceil() boundary at 2200 / 1575 ms per step
    boundary = 20 * 2200.0 / 1575
    local_ms = [boundary - 0.004, boundary + 0.004]
    own = [max(20, math.ceil(2200.0 / (t / 20))) for t in local_ms]
    assert own[0] != own[1]                         # each rank on its own clock would disagree ...
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29757, local_ms, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1] == min(own)             # ... the collective count is the slowest rank's


def test_sustained_leg_never_shorter_than_the_driver_leg():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.sustained_steps(lambda v: v, 5000.0, 20, 2.2) == 20          # 250 ms steps: 20 steps already exceed 2.2 s
    assert bench.sustained_steps(lambda v: v, 26.0, 20, 2.2) == math.ceil(2200 / 1.3)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU path on the host cores) needs no GPU: one step of its
    bounded sample, checked for the keys the driver reads."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["metric"] == "quantized embeddings/sec (fwd+EMA)" and line["unit"] == "embeddings/s"
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert "BASELINE config 2" in line["config"]["workload"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    import subprocess
    env = dict(os.environ, RANK="3", WORLD_SIZE="8", LOCAL_RANK="3")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: on a box without CUDA the product arm must fail loudly, not time something else"""
    import subprocess
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "CUDA" in out.stderr
