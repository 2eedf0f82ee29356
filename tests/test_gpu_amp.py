"""Mixed-precision drop-in behaviour (train_MedTok.py:185,212,240-246 runs the quantizer under
`torch.cuda.amp.autocast` + `GradScaler`): the modules accept half-precision activations, return the caller's dtype,
keep every loss finite through scaler.scale(...).backward() / unscale_ / step, and give the gradients of the fp32 run up
to the half-precision rounding of their inputs."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("assign_route")]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _soft_inputs(dev, b=48, t=9, e=64):
    g = torch.Generator().manual_seed(5)
    z = torch.randn(b, 2 * e, generator=g).to(dev)
    z_aug = (z.cpu() + 0.05 * torch.randn(b, 2 * e, generator=g)).to(dev)
    text = torch.randn(b, t, e, generator=g).to(dev)
    mask = torch.ones(b, t, dtype=torch.int64)
    mask[:, 6:] = 0
    batch = torch.arange(b).repeat_interleave(4)
    nodes = torch.randn(batch.numel(), e, generator=g).to(dev)
    return z, text, nodes, mask.to(dev), batch.to(dev), z_aug


def _total_loss(out):
    # train_MedTok.py:215-217: the six codebook loss terms (+ a use of the embeddings so that the STE path carries gradient)
    loss = (out['shared_embed_loss'][0] + out['shared_embed_loss'][1] + out['text_specific_loss'][0]
            + out['text_specific_loss'][1] + out['graph_specific_loss'][0] + out['graph_specific_loss'][1])
    return loss + 0.1 * (out['specific_embedding_text'].float().pow(2).mean() + out['shared_text_embedding'].float().pow(2).mean())


@pytest.mark.parametrize("amp_dtype", [torch.float16, torch.bfloat16])
def test_soft_quantizer_under_autocast_and_gradscaler(dev, amp_dtype):
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    torch.manual_seed(0)
    sq = VectorQuantizer(n_e=1536, e_dim=64, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True, show_usage=True,
                         split=[64, 64], k=5).to(dev).train()
    for m in sq.modules():                                   # deterministic comparison: no attention / residual dropout
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    args = _soft_inputs(dev)
    # fp32 run
    sq.zero_grad(set_to_none=True)
    _total_loss(sq(*args)).backward()
    ref = {n: p.grad.detach().clone() for n, p in sq.named_parameters() if p.grad is not None}
    assert "codebook.weight" in ref and "proj_text.weight" in ref
    # autocast + GradScaler run (train_MedTok.py:212-246)
    opt = torch.optim.AdamW(sq.parameters(), lr=1e-4)
    scaler = torch.amp.GradScaler("cuda", enabled=(amp_dtype == torch.float16))
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=amp_dtype):
        out = sq(*args)
        loss = _total_loss(out)
    assert out['specific_embedding_text'].dtype in (amp_dtype, torch.float32)
    assert out['text_tokens'].dtype == torch.int64 and torch.isfinite(loss)
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    torch.nn.utils.clip_grad_norm_(sq.parameters(), 1.0)     # as train_MedTok.py:243-244
    w_before = sq.codebook.weight.detach().clone()
    scaler.step(opt)
    scaler.update()
    assert not torch.equal(w_before, sq.codebook.weight)      # the step was not skipped: every gradient was finite
    # gradients of the unclipped AMP run against fp32 (recompute without clipping)
    opt.zero_grad(set_to_none=True)
    sq.codebook.weight.data.copy_(w_before)
    with torch.autocast("cuda", dtype=amp_dtype):
        loss = _total_loss(sq(*args))
    loss.backward()
    tol = 0.05 if amp_dtype == torch.float16 else 0.2        # relative L2 error: half-precision activations in front
    for name in ("codebook.weight", "proj_text.weight", "proj_graph.weight"):
        g = dict(sq.named_parameters())[name].grad
        assert g is not None and torch.isfinite(g).all(), name
        if name == "codebook.weight":
            continue      # the optimizer moved the other parameters by lr * sign: top-k sets may differ for a few rows
        rel = float((g.float() - ref[name]).norm() / ref[name].norm().clamp_min(1e-12))
        assert rel <= tol, f"{name}: relative gradient error {rel:.3f} under autocast({amp_dtype})"


def test_hard_quantizer_accepts_half_inputs_under_autocast(dev):
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    g = torch.Generator().manual_seed(3)
    n, d, k = 600, 64, 512
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    w0 = q.embedding.weight.detach().clone()
    z32 = torch.randn(n, d, 1, 1, generator=g).to(dev)
    z16 = z32.half().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        z_q, loss, idx = q(z16)
    assert z_q.dtype == torch.float16 and loss.dtype == torch.float32 and idx.dtype == torch.int64
    scaler = torch.amp.GradScaler("cuda")
    scaler.scale(loss + z_q.float().pow(2).mean()).backward()
    assert z16.grad is not None and z16.grad.dtype == torch.float16 and torch.isfinite(z16.grad).all()
    # same codes as the fp32 module fed the SAME (half-rounded) values
    q2 = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    q2.embedding.weight.data.copy_(w0)
    with torch.no_grad():
        _, _, idx2 = q2(z16.detach().float())
    assert torch.equal(idx, idx2)
