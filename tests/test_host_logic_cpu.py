"""Host-side control flow of NormEMAVectorQuantizer._run without a GPU: the compute entry points of the C ABI are
replaced by a recorder (they return 0 and do nothing -- NOT a fallback: nothing is computed and no result is checked),
the host-only entry points (sizes, route query / pin) go to the real library.  What is checked is the sequence of
library calls and collectives a forward issues: flags, the exchange route, and the single-pass pin of process groups
with more than MEDTOK_TWO_PASS_MAX_WORLD ranks (a branch the 1-GPU test box never takes)."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HOST_ONLY = {"mtvq_abi_version", "mtvq_last_error", "mtvq_workspace_bytes", "mtvq_stats_len", "mtvq_hard_route",
             "mtvq_debug_set_hard_route", "mtvq_debug_get_hard_route"}


class Recorder:
    def __init__(self, real):
        self.real = real
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("mtvq_"):
            raise AttributeError(name)

        def call(*args):
            self.calls.append((name, args))
            return getattr(self.real, name)(*args) if name in HOST_ONLY else 0
        return call

    def names(self):
        return [n for n, _ in self.calls if n not in ("mtvq_workspace_bytes", "mtvq_stats_len")]


def _patched_module():
    sys.path.insert(0, ROOT)
    from medtok_b200 import _lib
    rec = Recorder(_lib.lib())
    _lib.lib = lambda: rec
    _lib.require_cuda = lambda t, name: None
    _lib.stream_ptr = lambda device: None
    from medtok_b200 import norm_ema_quantizer
    return _lib, rec, norm_ema_quantizer


def _forward_calls(queue, env):
    os.environ.update(env)
    _lib, rec, mod = _patched_module()
    q = mod.NormEMAVectorQuantizer(512, 64, 0.25).train()
    rec.calls.clear()
    with torch.no_grad():
        q(torch.randn(300, 64, 1, 1))
    flags = [a[9] for n, a in rec.calls if n == "mtvq_hard_forward"]
    queue.put((rec.names(), flags, q.last_route, rec.real.mtvq_debug_get_hard_route()))


def _dist_worker(rank, world, port, env, queue):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _forward_calls(queue, env)
    dist.barrier()
    dist.destroy_process_group()


def _spawn(world, port, env):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    if world == 1:
        procs = [ctx.Process(target=_forward_calls, args=(queue, env))]
    else:
        procs = [ctx.Process(target=_dist_worker, args=(r, world, port, env, queue)) for r in range(world)]
    for p in procs:
        p.start()
    res = [queue.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_single_process_forward_is_one_library_call():
    (names, flags, last_route, pinned), = _spawn(1, 0, {})
    assert names == ["mtvq_hard_forward", "mtvq_hard_route"]            # finalize fused into the forward: no exchange
    assert flags == [1 | 2]                                              # TRAINING | UPDATE, no SKIP_FINALIZE
    assert pinned == -1


def test_two_ranks_exchange_once_and_finalize():
    for names, flags, last_route, pinned in _spawn(2, 29771, {}):
        assert names == ["mtvq_debug_get_hard_route", "mtvq_hard_forward", "mtvq_hard_route", "mtvq_ema_finalize"]
        assert flags == [1 | 2 | 4]                                      # statistics only; the EMA follows the exchange
        assert pinned == -1


def test_groups_above_the_validated_size_pin_the_single_pass_route():
    """MEDTOK_TWO_PASS_MAX_WORLD (default 4): here 1, so that two gloo ranks take the branch"""
    for names, flags, last_route, pinned in _spawn(2, 29773, {"MEDTOK_TWO_PASS_MAX_WORLD": "1"}):
        assert names == ["mtvq_debug_get_hard_route", "mtvq_debug_set_hard_route", "mtvq_hard_forward",
                         "mtvq_debug_set_hard_route", "mtvq_ema_finalize"]
        assert last_route == 1
        assert pinned == -1                                              # the pin does not outlive the call


def test_an_explicit_route_is_never_overridden():
    for names, flags, last_route, pinned in _spawn(2, 29775, {"MEDTOK_TWO_PASS_MAX_WORLD": "1", "MTVQ_HARD_ROUTE": "2"}):
        assert "mtvq_debug_set_hard_route" not in names
