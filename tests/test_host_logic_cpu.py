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


# ------------------------------------------------------------------------------------------------------------------
# Soft quantizer: structure of the forward dict against the UNMODIFIED reference run live on CPU (same inputs): every
# key the reference returns, with the same container types, tensor shapes and dtypes; plus the token keys
# tokenizer.py:235-238 reads.  Values are NOT compared here (the recorder computes nothing; tests/test_gpu_soft.py does).
# ------------------------------------------------------------------------------------------------------------------
def _structure(v):
    if isinstance(v, torch.Tensor):
        return ("tensor", tuple(v.shape), str(v.dtype))
    if isinstance(v, (tuple, list)):
        return ("tuple", tuple(_structure(x) for x in v))
    if v is None:
        return ("none",)
    return (type(v).__name__,)


def _soft_inputs(e_dim, bsz=6, tokens=7, nodes=15):
    g = torch.Generator().manual_seed(5)
    z = torch.randn(bsz, 2 * e_dim, generator=g)
    z_aug = torch.randn(bsz, 2 * e_dim, generator=g)
    text = torch.randn(bsz, tokens, e_dim, generator=g)
    graph = torch.randn(nodes, e_dim, generator=g)
    mask = torch.ones(bsz, tokens, dtype=torch.long)
    mask[1, 4:] = 0
    mask[3, 2:] = 0
    batch = torch.tensor([0, 0, 0, 1, 1, 2, 2, 2, 3, 3, 4, 4, 5, 5, 5])
    return z, text, graph, mask, batch, z_aug


def _soft_structure(queue, training):
    import contextlib
    import io
    _lib, rec, _ = _patched_module()
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    e_dim, n_e, k = 32, 96, 5
    with contextlib.redirect_stdout(io.StringIO()):
        ours = VectorQuantizer(n_e=n_e, e_dim=e_dim, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True, show_usage=True,
                               split=[e_dim, e_dim], k=k).train(training)
    z, text, graph, mask, batch, z_aug = _soft_inputs(e_dim)
    out = ours(z, text, graph, mask, batch, z_aug=z_aug if training else None)
    result = {"ours": {key: _structure(v) for key, v in out.items()}, "calls": rec.names(), "ref": None}
    if training:                                   # the backward reaches the fused kernel and hands gradients on
        loss = out["specific_embedding_text"].sum() + out["text_specific_loss"][0] + out["shared_embed_loss"][1]
        loss.backward()
        result["grads"] = {"codebook": _structure(ours.codebook.weight.grad), "proj_text": _structure(ours.proj_text.weight.grad)}
        result["calls_bwd"] = [n for n in rec.names() if "backward" in n]
    from oracle import ref_loader
    if ref_loader.available():
        with contextlib.redirect_stdout(io.StringIO()):
            _, soft = ref_loader.load()
            ref = soft.VectorQuantizer(n_e=n_e, e_dim=e_dim, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True,
                                       show_usage=True, split=[e_dim, e_dim], k=k).train(training)
        del ref._buffers["codebook_used"]          # harness accommodation of SURVEY Appendix B-3 (oracle/make_golden.py)
        ref.register_buffer("codebook_used", torch.zeros(300000))
        rout = ref(z, text, graph, mask, batch, z_aug=z_aug if training else None)
        result["ref"] = {key: _structure(v) for key, v in rout.items()}
    queue.put(result)


def _run_soft(training):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    p = ctx.Process(target=_soft_structure, args=(queue, training))
    p.start()
    res = queue.get(timeout=300)
    p.join(timeout=60)
    assert p.exitcode == 0
    return res


def test_soft_forward_dict_has_the_reference_structure_train():
    res = _run_soft(True)
    ours, ref = res["ours"], res["ref"]
    # 2 shared passes + 2 specific + 2 augmented specific passes, each followed by its usage-window update where the
    # reference has one (shared: one update for both modalities)
    assert res["calls"].count("mtvq_soft_forward") == 6 and res["calls"].count("mtvq_codebook_usage") == 5
    assert res["calls_bwd"] and set(res["calls_bwd"]) == {"mtvq_soft_backward_fused"}
    assert res["grads"]["codebook"] == ("tensor", (96, 32), "torch.float32")
    assert res["grads"]["proj_text"] == ("tensor", (32, 32), "torch.float32")
    for key in ("text_tokens", "graph_tokens", "shared_text_tokens", "shared_graph_tokens"):
        assert ours[key] == ("tensor", (6, 5), "torch.int64")
        assert ours[key + "_weights"] == ("tensor", (6, 5), "torch.float32")
    if ref is not None:
        for key, st in ref.items():
            assert ours[key] == st, f"{key}: ours {ours[key]} vs reference {st}"


def test_soft_forward_dict_has_the_reference_structure_eval():
    res = _run_soft(False)
    ours, ref = res["ours"], res["ref"]
    assert res["calls"].count("mtvq_soft_forward") == 4
    assert ours["specific_embedding_text_aug"] == ("none",) and ours["specific_embedding_graph_aug"] == ("none",)
    if ref is not None:
        for key, st in ref.items():
            assert ours[key] == st, f"{key}: ours {ours[key]} vs reference {st}"


# ------------------------------------------------------------------------------------------------------------------
# Hard quantizers: return tuple of forward against the UNMODIFIED reference run live on CPU (structure only).
# ------------------------------------------------------------------------------------------------------------------
def _hard_structure(queue):
    import contextlib
    import io
    _lib, rec, mod = _patched_module()
    from oracle import ref_loader
    hard = None
    if ref_loader.available():
        with contextlib.redirect_stdout(io.StringIO()):
            hard, _ = ref_loader.load()
    out = []
    g = torch.Generator().manual_seed(3)
    for shape, training in (((40, 64, 1, 1), True), ((5, 32, 4, 3), True), ((40, 64, 1, 1), False)):
        z = torch.randn(*shape, generator=g)
        k, d = 128, shape[1]
        ours = mod.NormEMAVectorQuantizer(k, d, 0.25).train(training)
        with torch.no_grad():
            o = ours(z)
        r = None
        if hard is not None:
            ref = hard.NormEMAVectorQuantizer(n_embed=k, embedding_dim=d, beta=0.25).train(training)
            with torch.no_grad():
                r = ref(z)
        out.append(("hard", shape, training, _structure(o), None if r is None else _structure(r),
                    sorted(ours.state_dict().keys()), None if r is None else sorted(ref.state_dict().keys())))
    # lookup-only quantizer: code indices in, codebook rows out (reference :222-268)
    for bhw, d, with_z, training in (((6, 4, 3), 32, True, True), ((6, 4, 3), 32, False, False), ((10, 1, 1), 64, True, False)):
        k = 50
        idx = torch.randint(0, k, (bhw[0] * bhw[1] * bhw[2],), generator=g)
        z = torch.randn(bhw[0], d, bhw[1], bhw[2], generator=g) if with_z else None
        ours = mod.VectorQuantizer(k, d, 0.25).train(training)
        with torch.no_grad():
            o = ours(bhw, idx, z)
        r = None
        if hard is not None:
            ref = hard.VectorQuantizer(n_embed=k, embedding_dim=d, beta=0.25).train(training)
            with torch.no_grad():
                r = ref(bhw, idx, z)
        out.append(("lookup", bhw, training, _structure(o), None if r is None else _structure(r),
                    sorted(ours.state_dict().keys()), None if r is None else sorted(ref.state_dict().keys())))
    queue.put((out, rec.names()))


def test_hard_quantizers_return_the_reference_structure():
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    p = ctx.Process(target=_hard_structure, args=(queue,))
    p.start()
    cases, calls = queue.get(timeout=300)
    p.join(timeout=60)
    assert p.exitcode == 0
    assert calls.count("mtvq_hard_forward") == 3 and calls.count("mtvq_lookup") == 3
    for kind, shape, training, ours, ref, keys, ref_keys in cases:
        assert ours[0] == "tuple" and len(ours[1]) == 3
        if kind == "hard":
            b, c, h, w = shape
            assert ours[1][0] == ("tensor", (b, c, h, w), "torch.float32")
            assert ours[1][1] == ("tensor", (), "torch.float32")
            assert ours[1][2] == ("tensor", (b * h * w,), "torch.int64")
        if ref is not None:
            assert ours == ref, f"{kind} {shape} training={training}: ours {ours} vs reference {ref}"
            assert keys == ref_keys
