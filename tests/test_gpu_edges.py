"""Edge cases of both paths on the GPU: empty and tiny inputs, half-precision inputs (autocast policy), codebooks
smaller than a tile, NaN rows, codebook slicing with n_e % 3 != 0, full-size soft forward properties (config 3), and the
fused backward against PyTorch autograd of the reference formula."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("assign_route")]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def test_empty_and_single_row_inputs(dev):
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    q = NormEMAVectorQuantizer(256, 32, 0.25).to(dev).train()
    w0, cs0 = q.embedding.weight.data.clone(), q.cluster_size.clone()
    z_q, loss, idx = q(torch.empty(0, 32, 1, 1, device=dev))
    assert z_q.shape == (0, 32, 1, 1) and idx.shape == (0,)
    # reference with an empty batch: bins = 0 everywhere -> cluster_size decays, every row keeps its (renormalised) value
    assert torch.allclose(q.cluster_size, cs0 * 0.99)
    assert (q.embedding.weight - w0).abs().max().item() <= 2e-7
    z_q, loss, idx = q(torch.randn(1, 32, 1, 1, device=dev))
    assert idx.shape == (1,) and 0 <= int(idx) < 256 and torch.isfinite(loss)
    out, vq, commit, sidx, wts = _SoftQuantizeFn.apply(torch.empty(0, 32, device=dev), torch.randn(64, 32, device=dev), 0, 64,
                                                      5, 0.25, True, _lib.Workspace())
    assert out.shape == (0, 32) and sidx.shape == (0, 5)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_inputs_are_widened(dev, dtype):
    """fp16 / bf16 rows (autocast) are widened to fp32; indices equal those of the fp32 copy of the same values."""
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    g = torch.Generator().manual_seed(2)
    z = torch.randn(3000, 64, 1, 1, generator=g).to(dev).to(dtype)
    q = NormEMAVectorQuantizer(2048, 64, 0.25).to(dev).eval()
    z_q, loss, idx = q(z)
    z_q32, loss32, idx32 = q(z.float())
    assert z_q.dtype == dtype and torch.equal(idx, idx32)
    assert torch.equal(z_q.float(), z_q32.to(dtype).float())


def test_small_codebooks_and_odd_slices(oracle, dev):
    """K smaller than one 128-code tile (hard), and a soft table with n_e % 3 != 0: text = rows [:n_e//3], graph = rows
    [-(n_e//3):] (vector_quantization_soft_one_new.py:195-200), indices slice-relative."""
    from medtok_b200.norm_ema_quantizer import hard_assign
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    g = torch.Generator().manual_seed(9)
    z = torch.randn(500, 16, generator=g)
    w = torch.nn.functional.normalize(torch.randn(7, 16, generator=g), dim=-1)
    idx, _ = hard_assign(z.to(dev), w.to(dev))
    zf, _ = oracle.C().l2norm_rows(z.numpy())
    idx_ref, _ = oracle.C().hard_assign(zf, w.numpy())
    assert np.array_equal(idx.cpu().numpy(), idx_ref)

    n_e, e_dim = 1000, 32                                   # region = 333, one row belongs to neither specific slice
    sq = VectorQuantizer(n_e=n_e, e_dim=e_dim, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True, show_usage=False,
                         split=[e_dim, e_dim], k=5).to(dev).eval()
    x = torch.randn(64, e_dim, generator=g).to(dev)
    table = sq.codebook.weight.data
    with torch.no_grad():
        for types, lo, hi in (("text", 0, 333), ("graph", n_e - 333, n_e)):
            z_q, _, _ = sq.specific_embedding(x, types=types)
            tok, wts = sq._last_specific
            proj = (sq.proj_text if types == "text" else sq.proj_graph)(x)
            r = oracle.soft_core(proj.cpu().numpy(), table[lo:hi].cpu().numpy(), 5, 0.25, training=False)
            assert np.array_equal(tok.cpu().numpy(), r.idx) and int(tok.max()) < 333
            assert np.max(np.abs(wts.cpu().numpy() - r.weights)) <= 1e-6


def test_nan_rows_do_not_disturb_their_neighbours(dev):
    """A NaN row must not corrupt other rows (its own index is unspecified, like an argmin over NaNs)."""
    from medtok_b200.norm_ema_quantizer import hard_assign
    g = torch.Generator().manual_seed(12)
    z = torch.randn(1000, 64, generator=g)
    w = torch.nn.functional.normalize(torch.randn(1024, 64, generator=g), dim=-1)
    idx_clean, _ = hard_assign(z.to(dev), w.to(dev), exact=True)
    z[[3, 500, 999]] = float("nan")
    idx, _ = hard_assign(z.to(dev), w.to(dev))
    keep = torch.ones(1000, dtype=torch.bool)
    keep[[3, 500, 999]] = False
    assert torch.equal(idx.cpu()[keep], idx_clean.cpu()[keep])
    assert int(idx.min()) >= 0 and int(idx.max()) < 1024


def test_full_size_config3_properties(dev):
    """BASELINE config 3 (N = 2^20, D = 128, K' = 16384, k = 5): size-independent properties of the soft forward plus
    bit-identity with the exact CUDA-core scan on a 4096-row sample."""
    from medtok_b200 import _lib
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    n, d, k, kk = 1 << 20, 128, 16384, 5
    torch.manual_seed(0)
    x = torch.randn(n, d, device=dev)
    table = torch.randn(k, d, device=dev)
    ws = _lib.Workspace()
    with torch.no_grad():
        out, vq, commit, idx, wts = _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, ws)
    cnt = _lib.debug_counters(ws.buf, n, d, k, kk)
    assert cnt["full_scan_rows"] == 0
    assert int(idx.min()) >= 0 and int(idx.max()) < k
    assert (idx.sort(dim=1).values.diff(dim=1) != 0).all()                      # k distinct codes per row
    assert torch.allclose(wts.sum(-1), torch.ones(n, device=dev), atol=1e-6)
    assert (wts.diff(dim=1) <= 0).all()                                         # ascending distance = descending weight
    cn = torch.nn.functional.normalize(table, dim=-1)
    xn = torch.nn.functional.normalize(x, dim=-1)
    sample = torch.arange(0, n, 4096, device=dev)
    zq_ref = (wts[sample].unsqueeze(-1) * cn[idx[sample]]).sum(1)
    assert (out[sample] - zq_ref).abs().max().item() <= 2e-6                   # out = x + (zq - x)
    dist = (2 - 2 * xn[sample] @ cn.t())
    kth = dist.topk(kk, dim=1, largest=False).values[:, -1]
    d_sel = dist.gather(1, idx[sample])
    assert (d_sel <= kth.unsqueeze(1) + 1e-5).all()                             # every chosen code is within the true top-k
    ref_vq = ((out - x) ** 2).mean().item()
    assert abs(float(vq) - ref_vq) <= 1e-4 * ref_vq and abs(float(commit) - 0.25 * float(vq)) <= 1e-7
    try:                                                                        # exact scan route on the sample rows
        _lib.check(_lib.lib().mtvq_debug_set_soft_route(2), "route")
        with torch.no_grad():
            out2, _, _, idx2, wts2 = _SoftQuantizeFn.apply(x[:4096].contiguous(), table, 0, k, kk, 0.25, True, _lib.Workspace())
    finally:
        _lib.lib().mtvq_debug_set_soft_route(0)
    assert torch.equal(idx[:4096], idx2) and torch.equal(wts[:4096], wts2)


def test_fused_backward_matches_autograd_of_reference_formula(dev):
    """Gradients of (out, vq, commit) w.r.t. x and the raw table == PyTorch autograd of soft :196-214 on the same top-k."""
    from medtok_b200 import _lib
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    for (n, d, k, kk) in ((700, 64, 900, 5), (333, 128, 2000, 5), (257, 32, 300, 3), (129, 256, 512, 8)):
        g = torch.Generator().manual_seed(n)
        x = (torch.randn(n, d, generator=g) * 1.3).to(dev).requires_grad_(True)
        table = torch.randn(k, d, generator=g).to(dev).requires_grad_(True)
        gout = torch.randn(n, d, generator=g).to(dev)
        out, vq, commit, idx, wts = _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, _lib.Workspace())
        ((out * gout).sum() + 3.0 * vq + 2.0 * commit).backward()
        x2 = x.detach().clone().requires_grad_(True)
        t2 = table.detach().clone().requires_grad_(True)
        xn = torch.nn.functional.normalize(x2, dim=-1)
        cn = torch.nn.functional.normalize(t2, dim=-1)
        dist = (xn ** 2).sum(1, keepdim=True) + (cn ** 2).sum(1) - 2 * xn @ cn.t()
        vals = dist.gather(1, idx)                                             # same top-k as the kernel chose
        w = torch.softmax(-vals, dim=-1)
        zq = (w.unsqueeze(-1) * cn[idx]).sum(1)
        vq_r = ((zq - x2.detach()) ** 2).mean()
        cm_r = 0.25 * ((zq.detach() - x2) ** 2).mean()
        out_r = x2 + (zq - x2).detach()
        ((out_r * gout).sum() + 3.0 * vq_r + 2.0 * cm_r).backward()
        assert torch.allclose(x.grad, x2.grad, rtol=2e-4, atol=2e-7), (n, d, (x.grad - x2.grad).abs().max())
        assert torch.allclose(table.grad, t2.grad, rtol=2e-4, atol=2e-7), (n, d, (table.grad - t2.grad).abs().max())
