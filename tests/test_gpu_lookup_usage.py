"""GPU parity tests of the lookup-only VectorQuantizer (norm_ema_quantizer.py:222-268), the sliding usage window
(vector_quantization_soft_one_new.py:219-236) and batched statistics accumulation (MTVQ_F_ACCUMULATE):
CUDA path vs golden outputs of the unmodified reference and vs the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("assign_route")]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", golden_names("lookup_"))
def test_lookup_module_matches_reference_golden(name, dev):
    from medtok_b200.norm_ema_quantizer import VectorQuantizer
    meta, a = load_golden(name)
    k, d = a["weight"].shape
    q = VectorQuantizer(n_embed=k, embedding_dim=d, beta=meta["beta"]).to(dev)
    assert sorted(q.state_dict().keys()) == meta["state_dict_keys"]
    assert q.decay == 0.0
    q.train(meta["training"])
    q.embedding.weight.data.copy_(torch.from_numpy(a["weight"]).to(dev))
    q.cluster_size.fill_(3.0)
    idx = torch.from_numpy(a["idx"]).to(dev)
    z = torch.from_numpy(a["z"]).to(dev) if meta["with_z"] else None
    z_q, loss, out_idx = q(tuple(meta["bhw"]), idx, z)
    assert out_idx is idx and z_q.shape == a["z_q"].shape and loss.dim() == 0
    assert np.array_equal(z_q.cpu().numpy(), a["z_q"])                       # bit-exact gather
    assert np.array_equal(q.cluster_size.cpu().numpy(), a["cluster_size"])   # bit-exact counts (eval) / untouched (train)
    ref = float(a["loss"][0])
    assert abs(float(loss) - ref) <= 1e-5 * abs(ref) + 1e-12                 # fp32 mean order


def test_lookup_large_matches_oracle(oracle, dev):
    """N = 300 000 rows, skewed codes: counts exact, z_q exact, loss rel 1e-6 (fp64 accumulation here)."""
    from medtok_b200.norm_ema_quantizer import VectorQuantizer
    n, d, k = 300_000, 64, 8192
    g = torch.Generator().manual_seed(4)
    idx = (torch.rand(n, generator=g) ** 3 * k).long().clamp_(0, k - 1)
    z = torch.randn(n, d, 1, 1, generator=g)
    q = VectorQuantizer(k, d, 0.25).to(dev).eval()
    zq_ref, loss_ref, counts = oracle.lookup_forward((n, 1, 1), idx.numpy(), q.embedding.weight.cpu().numpy(), 0.25,
                                                     z.numpy(), training=False)
    z_q, loss, _ = q((n, 1, 1), idx.to(dev), z.to(dev))
    assert np.array_equal(z_q.cpu().numpy(), zq_ref)
    assert np.array_equal(q.cluster_size.cpu().numpy(), counts)
    assert abs(float(loss) - loss_ref) <= 1e-6 * loss_ref


def test_usage_window_matches_reference_golden(dev):
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    meta, a = load_golden("usage_window_ne3000")
    q = VectorQuantizer(n_e=meta["n_e"], e_dim=16, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True, show_usage=True,
                        split=[16, 16]).to(dev)
    for i in range(meta["calls"]):
        usage = q.codebook_usage(torch.from_numpy(a[f"idx_{i}"]).to(dev), types="shared")
        assert usage == a["usage"][i], f"call {i}: {usage} != {a['usage'][i]}"
    assert np.array_equal(q.codebook_used.cpu().numpy(), a["window"])
    assert "codebook_used" in q.state_dict() and q.state_dict()["codebook_used"].shape == (300000,)


def test_usage_window_overflow_and_oracle(oracle, dev):
    """More than 300 000 new entries in one call keeps the last 300 000 (the reference raises there, Appendix B-4);
    n_e not a multiple of 32."""
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    n_e = 49152 + 7
    q = VectorQuantizer(n_e=n_e, e_dim=16, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True, show_usage=True,
                        split=[16, 16]).to(dev)
    g = torch.Generator().manual_seed(8)
    win = np.zeros(300000, np.float32)
    for rows in (1000, 70001, 12):
        idx = torch.randint(0, n_e, (rows, 5), generator=g)
        win, ref = oracle.usage_window(win, idx.numpy(), n_e)
        assert q.codebook_usage(idx.to(dev)) == ref
    assert np.array_equal(q.codebook_used.cpu().numpy(), win)


def test_accumulated_batches_equal_one_pass(dev):
    """A vocabulary quantised in ragged batches with MTVQ_F_ACCUMULATE | MTVQ_F_SKIP_FINALIZE and ONE finalize equals
    the single-call forward: identical indices / z_q / counts, identical updated codebook (fp64 partial sums)."""
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    n, d, k = 50_000, 64, 2048
    g = torch.Generator().manual_seed(21)
    z = torch.randn(n, d, generator=g).to(dev)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1).to(dev)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    q.embedding.weight.data.copy_(w)
    with torch.no_grad():
        z_q1, loss1, idx1 = q(z.view(n, d, 1, 1))
    w1, cs1 = q.embedding.weight.data.clone(), q.cluster_size.clone()

    L = _lib.lib()
    w2, cs2 = w.clone(), torch.zeros(k, device=dev)
    stats = torch.zeros(_lib.stats_len(d, k), dtype=torch.float64, device=dev)
    z_q2 = torch.empty(n, d, device=dev)
    idx2 = torch.empty(n, dtype=torch.int64, device=dev)
    loss = torch.empty(1, device=dev)
    wsb = _lib.Workspace()
    flags = _lib.F_TRAINING | _lib.F_UPDATE | _lib.F_SKIP_FINALIZE | _lib.F_ACCUMULATE
    lo = 0
    for rows in (12345, 1, 20000, 0, 17654):
        zb = z[lo:lo + rows].contiguous()
        ws = wsb.get(max(rows, 1), d, k, 0, dev)
        _lib.check(L.mtvq_hard_forward(_lib.ptr(zb), rows, d, 0, _lib.ptr(w2), k, _lib.ptr(cs2), 0.25, 0.99, flags,
                                       _lib.ptr(z_q2[lo:lo + rows]), _lib.ptr(idx2[lo:lo + rows]), _lib.ptr(loss),
                                       _lib.ptr(stats), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "batch")
        lo += rows
    assert lo == n
    _lib.check(L.mtvq_ema_finalize(_lib.ptr(w2), _lib.ptr(cs2), _lib.ptr(stats), k, d, 0.25, 0.99, 1, 1, _lib.ptr(loss),
                                   _lib.stream_ptr(dev)), "finalize")
    assert torch.equal(idx1, idx2) and torch.equal(z_q1.view(n, d), z_q2)
    assert torch.equal(cs1, cs2)
    assert torch.equal(stats[:k], q.last_stats[:k])
    assert (w1 - w2).abs().max().item() <= 1.2e-7
    assert abs(float(loss) - float(loss1)) <= 1e-6 * float(loss1)


def test_cvector_quantiser_matches_reference_golden(dev):
    """CVectorQuantiser (norm_ema_quantizer.py:272-403, anchor='closest'): indices / counts / z_q bit-exact, embed_prob and
    the re-initialised codebook within tolerance, over two training steps and one eval step."""
    from medtok_b200.norm_ema_quantizer import CVectorQuantiser
    meta, a = load_golden("cvq_closest_n600_d32_k96")
    k, d = a["weight0"].shape
    q = CVectorQuantiser(n_embed=k, embedding_dim=d, beta=meta["beta"], kmeans_init=False, anchor='closest').to(dev)
    assert sorted(q.state_dict().keys()) == meta["state_dict_keys"]
    q.embedding.weight.data.copy_(torch.from_numpy(a["weight0"]).to(dev))
    for s, mode in enumerate(meta["modes"]):
        q.train(mode == "train")
        with torch.no_grad():
            z_q, loss, idx = q(torch.from_numpy(a[f"z_{s}"]).to(dev))
        assert np.array_equal(idx.cpu().numpy(), a[f"idx_{s}"]), f"step {s}"
        assert np.array_equal(q.cluster_size.cpu().numpy(), a[f"cluster_size_{s}"])
        assert np.array_equal(z_q.cpu().numpy(), a[f"z_q_{s}"])
        ref = float(a[f"loss_{s}"][0])
        assert abs(float(loss) - ref) <= 1e-5 * ref
        assert np.max(np.abs(q.embed_prob.cpu().numpy() - a[f"embed_prob_{s}"])) <= 1e-8
        assert np.max(np.abs(q.embedding.weight.cpu().numpy() - a[f"weight_{s}"])) <= 2e-6
        # continue from the reference's own state so that every step is checked on its own
        q.embedding.weight.data.copy_(torch.from_numpy(a[f"weight_{s}"]).to(dev))
        q.embed_prob.copy_(torch.from_numpy(a[f"embed_prob_{s}"]).to(dev))


@pytest.mark.parametrize("anchor", ["random", "probrandom"])
def test_cvector_quantiser_random_anchors(dev, anchor):
    """anchor='random' (feature pool, norm_ema_quantizer.py:405-460) and 'probrandom' (:387-390) draw samples, so they
    are checked through what is deterministic: the forward part equals the norm-EMA forward of the same inputs (indices,
    z_q, loss, cluster_size), embed_prob follows its EMA, and every re-initialised row is the reference's blend
    w * (1 - dec) + feat * dec of the norm-EMA-updated row with SOME normalised sample of the batch."""
    from medtok_b200.norm_ema_quantizer import CVectorQuantiser, NormEMAVectorQuantizer
    g = torch.Generator().manual_seed(8)
    n, d, k = 700, 32, 96
    z = torch.randn(n, d, 1, 1, generator=g).to(dev)
    w0 = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1).to(dev)
    q = CVectorQuantiser(n_embed=k, embedding_dim=d, beta=0.25, kmeans_init=False, anchor=anchor).to(dev).train()
    base = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    q.embedding.weight.data.copy_(w0)
    base.embedding.weight.data.copy_(w0)
    torch.manual_seed(0)
    with torch.no_grad():
        z_q, loss, idx = q(z)
        z_q0, loss0, idx0 = base(z)
    assert torch.equal(idx, idx0) and torch.equal(z_q, z_q0) and float(loss) == float(loss0)
    assert torch.equal(q.cluster_size, base.cluster_size)
    bins = torch.bincount(idx, minlength=k).float()
    assert torch.allclose(q.embed_prob, 0.01 * bins / n, atol=1e-9)
    dec = torch.exp(-(q.embed_prob * k * 10) / (1 - 0.99) - 1e-3).unsqueeze(1)
    zf = torch.nn.functional.normalize(z.view(n, d), dim=-1)
    w_ema = base.embedding.weight.data
    feat = (q.embedding.weight.data - w_ema * (1 - dec)) / dec             # the sample each row was pulled towards
    live = dec.squeeze(1) > 1e-3                                            # rows whose blend weight is not negligible
    assert live.any()
    nearest = torch.cdist(feat[live], zf).min(dim=1).values
    assert float(nearest.max()) <= 1e-3, "a re-initialised row is not a blend with a sample of the batch"

