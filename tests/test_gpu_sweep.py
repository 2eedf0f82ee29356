"""BASELINE config 4 shapes on one GPU: the vocabulary quantised shard by shard (what 8 ranks do side by side) must
equal one pass, at a codebook size that is not a multiple of the 128-code tile; and the tokenization sweep writer fed by
the soft quantizer's eval-mode forward."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("assign_route")]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def test_config4_sharded_vocabulary_equals_one_pass(dev):
    """N = 655 360 code embeddings, D = 64, K = 21 000: eight row shards (SURVEY 8e partitioning) accumulated into one
    packed statistics buffer and ONE finalize == the single-call training forward; tensor path == exact kernel."""
    from medtok_b200 import _lib
    from medtok_b200.distributed import shard_rows
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, hard_assign
    n, d, k, world = 655_360, 64, 21_000, 8
    g = torch.Generator().manual_seed(44)
    z = torch.randn(n, d, generator=g).to(dev)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1).to(dev)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    q.embedding.weight.data.copy_(w)
    with torch.no_grad():
        z_q1, loss1, idx1 = q(z.view(n, d, 1, 1))
    idx_x, _ = hard_assign(z, w, exact=True)
    assert torch.equal(idx1, idx_x), "tensor-core path != exact CUDA-core kernel"
    assert int(q.last_stats[:k].sum().item()) == n

    L = _lib.lib()
    w2, cs2 = w.clone(), torch.zeros(k, device=dev)
    stats = torch.zeros(_lib.stats_len(d, k), dtype=torch.float64, device=dev)
    idx2 = torch.empty(n, dtype=torch.int64, device=dev)
    z_q2 = torch.empty(n, d, device=dev)
    loss = torch.empty(1, device=dev)
    wsb = _lib.Workspace()
    flags = _lib.F_TRAINING | _lib.F_UPDATE | _lib.F_SKIP_FINALIZE | _lib.F_ACCUMULATE
    for r in range(world):
        lo, hi = shard_rows(n, world, r)
        ws = wsb.get(hi - lo, d, k, 0, dev)
        _lib.check(L.mtvq_hard_forward(_lib.ptr(z[lo:hi]), hi - lo, d, 0, _lib.ptr(w2), k, _lib.ptr(cs2), 0.25, 0.99, flags,
                                       _lib.ptr(z_q2[lo:hi]), _lib.ptr(idx2[lo:hi]), _lib.ptr(loss), _lib.ptr(stats),
                                       _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "shard")
    _lib.check(L.mtvq_ema_finalize(_lib.ptr(w2), _lib.ptr(cs2), _lib.ptr(stats), k, d, 0.25, 0.99, 1, 1, _lib.ptr(loss),
                                   _lib.stream_ptr(dev)), "finalize")
    assert torch.equal(idx1, idx2) and torch.equal(z_q1.view(n, d), z_q2)
    assert torch.equal(q.cluster_size, cs2)
    assert (q.embedding.weight - w2).abs().max().item() <= 1.2e-7
    assert abs(float(loss) - float(loss1)) <= 1e-6 * float(loss1)


def test_token_sweep_from_soft_quantizer(dev, tmp_path):
    """tokens_all / weights_all / embeddings_all written by TokenSweep == the module's eval-mode outputs, row for row, in
    vocabulary order although the batches arrive shuffled."""
    from medtok_b200.tokenize_sweep import TokenSweep, compose_eval_outputs
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    torch.manual_seed(5)
    n_codes, e_dim, k = 640, 64, 5
    sq = VectorQuantizer(n_e=3072, e_dim=e_dim, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True, show_usage=True,
                         split=[e_dim, e_dim], k=k).to(dev).eval()
    sweep = TokenSweep(n_codes, 4 * e_dim, k)
    order = torch.randperm(n_codes)
    ref = {}
    with torch.no_grad():
        for lo in range(0, n_codes, 96):
            ci = order[lo:lo + 96]
            b = ci.numel()
            z = torch.randn(b, 2 * e_dim, device=dev)
            text = torch.randn(b, 7, e_dim, device=dev)
            mask = torch.ones(b, 7, dtype=torch.long, device=dev)
            nodes = torch.randn(3 * b, e_dim, device=dev)
            batch = torch.arange(b, device=dev).repeat_interleave(3)
            out = sq(z, text, nodes, mask, batch)
            e, t, w = compose_eval_outputs(out)
            assert t.shape == (b, 4, k) and w.shape == (b, 4, k) and e.shape == (b, 4 * e_dim)
            sweep.add(ci, e, t, w)
            for j, c in enumerate(ci.tolist()):
                ref[c] = (e[j].cpu(), t[j].cpu(), w[j].cpu())
    assert sweep.finish()
    sweep.save(str(tmp_path))
    e_all = np.load(os.path.join(tmp_path, "embeddings_all.npy"))
    t_all = np.load(os.path.join(tmp_path, "tokens_all.npy"))
    w_all = np.load(os.path.join(tmp_path, "weights_all.npy"))
    assert e_all.shape == (n_codes, 4 * e_dim) and t_all.dtype == np.int64 and w_all.dtype == np.float32
    for c in (0, 1, 333, n_codes - 1):
        assert np.array_equal(e_all[c], ref[c][0].numpy()) and np.array_equal(t_all[c], ref[c][1].numpy())
        assert np.array_equal(w_all[c], ref[c][2].numpy())
    assert np.allclose(w_all.sum(-1), 1.0, atol=1e-5)
