"""GPU parity tests of the hard (norm-EMA) quantizer: CUDA path (through the drop-in module and the
C ABI) vs golden outputs of the unmodified reference and vs the CPU oracle on seeded inputs."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("assign_route")]

HARD = [n for n in golden_names("hard_") if "gloo" not in n]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _module(a, meta, dev):
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    k, d = a["weight0"].shape
    q = NormEMAVectorQuantizer(n_embed=k, embedding_dim=d, beta=meta["beta"], decay=meta["decay"]).to(dev)
    q.embedding.update = meta["update"]
    q.train(meta["training"])
    return q


@pytest.mark.parametrize("name", HARD)
def test_module_matches_reference_golden(name, dev):
    meta, a = load_golden(name)
    q = _module(a, meta, dev)
    assert sorted(q.state_dict().keys()) == meta["state_dict_keys"]
    z = torch.from_numpy(a["z"]).to(dev)
    w_prev, cs_prev = a["weight0"], np.zeros(a["weight0"].shape[0], np.float32)
    for s in range(meta["steps"]):
        q.embedding.weight.data.copy_(torch.from_numpy(w_prev).to(dev))
        q.cluster_size.copy_(torch.from_numpy(cs_prev).to(dev))
        with torch.no_grad():
            z_q, loss, idx = q(z)
        assert idx.dtype == torch.int64 and z_q.shape == z.shape and loss.dim() == 0
        # bit-exact: code indices, cluster_size (counts), quantised output
        assert np.array_equal(idx.cpu().numpy(), a[f"idx_{s}"]), f"{name}: index mismatch at step {s}"
        assert np.array_equal(q.cluster_size.cpu().numpy(), a[f"cluster_size_{s}"])
        assert np.array_equal(z_q.cpu().numpy(), a[f"z_q_{s}"])
        # stated tolerances: loss rel 1e-5; updated codebook abs 1e-6 (fp64 vs fp32-GEMM embed_sum order)
        ref_loss = float(a[f"loss_{s}"][0])
        assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss) + 1e-12
        assert np.max(np.abs(q.embedding.weight.cpu().numpy() - a[f"weight_{s}"])) <= 1e-6
        w_prev, cs_prev = a[f"weight_{s}"], a[f"cluster_size_{s}"]


def test_coarse_scores_match_fp16_model(dev):
    """The tensor-core pass computes zf.w - ||w||^2/2 from fp16 operands: check the raw TMEM scores
    against the same fp16-rounded operands multiplied in fp32 by torch (validates the UMMA descriptors,
    operand layout and TMEM lane/column mapping)."""
    from medtok_b200 import _lib
    torch.manual_seed(3)
    for (n, d, k) in ((700, 64, 1000), (300, 128, 384), (130, 256, 256), (256, 32, 512)):
        z = torch.randn(n, d, device=dev)
        w = torch.nn.functional.normalize(torch.randn(k, d, device=dev), dim=-1) * 1.3
        ws = _lib.Workspace().get(n, d, k, 0, dev)
        kpad = ctypes.c_int(0)
        L = _lib.lib()
        _lib.check(L.mtvq_debug_coarse_scores(_lib.ptr(z), n, d, 0, _lib.ptr(w), k, None, 0, ctypes.byref(kpad),
                                              _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "kpad query")
        scores = torch.full((n, kpad.value), float("nan"), device=dev)
        _lib.check(L.mtvq_debug_coarse_scores(_lib.ptr(z), n, d, 0, _lib.ptr(w), k, _lib.ptr(scores), n,
                                              ctypes.byref(kpad), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                   "mtvq_debug_coarse_scores")
        torch.cuda.synchronize()
        zf = torch.nn.functional.normalize(z, dim=-1)
        ref = zf.half().double() @ w.half().double().t() - 0.5 * (w.double() ** 2).sum(1)
        got = scores[:, :k].double()
        err = (got - ref).abs().max().item()
        # the loader forms the fp16 operand as fp16(x * (1/den)); against fp16(x / den) an element can land on
        # the neighbouring fp16 value (<= 2^-11 |z_i w_i| each), far inside the kernel's error bound eps ~ 1e-3
        assert err < 2e-4, f"(n={n}, d={d}, k={k}) coarse score mismatch, max err {err}"
        assert (got - ref).abs().mean().item() < 2e-6
        assert (scores[:, k:] < -1e4).all()            # padding codes can never win


@pytest.mark.parametrize("n,d,k", [(20000, 64, 8192), (5000, 128, 4096), (3000, 256, 1024), (9999, 32, 777),
                                   (1000, 64, 100), (777, 64, 300), (3, 16, 5), (70000, 48, 2000),
                                   (1, 64, 8192), (255, 64, 21000), (256, 64, 7000), (1024, 64, 8192)])
def test_assign_matches_oracle_and_exact_kernel(oracle, dev, n, d, k):
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import hard_assign
    g = torch.Generator().manual_seed(n + d + k)
    z = torch.randn(n, d, generator=g)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1)
    zf, den_ref = oracle.C().l2norm_rows(z.numpy())
    idx_ref, _ = oracle.C().hard_assign(zf, w.numpy())
    wsx = _lib.Workspace()
    idx, den = hard_assign(z.to(dev), w.to(dev), workspace=wsx)
    cnt = _lib.debug_counters(wsx.buf, n, d, k)
    idx_x, den_x = hard_assign(z.to(dev), w.to(dev), exact=True)
    assert np.array_equal(den.cpu().numpy(), den_ref)
    assert np.array_equal(den_x.cpu().numpy(), den_ref)
    assert np.array_equal(idx_x.cpu().numpy(), idx_ref), "exact CUDA-core kernel != oracle"
    assert np.array_equal(idx.cpu().numpy(), idx_ref), "tensor-core path != oracle"
    assert cnt["full_scan_rows"] == 0 and cnt["near_tie_rows"] < 0.3 * n


@pytest.mark.single_route
@pytest.mark.parametrize("n,d,k", [(60000, 256, 384), (120000, 128, 640), (80000, 64, 1100), (50000, 32, 256)])
def test_many_blocks_per_cta_match_exact_kernel(dev, n, d, k):
    """Several row blocks per CTA for every tile-shape variant (RT = 1, 2, 4; A-region ring sizes 1, 2, 4):
    the tensor-core path must equal the exact CUDA-core kernel on every row."""
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import hard_assign
    g = torch.Generator().manual_seed(7 * n + d)
    z = torch.randn(n, d, generator=g).to(dev)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1).to(dev)
    wsx = _lib.Workspace()
    idx, den = hard_assign(z, w, workspace=wsx)
    cnt = _lib.debug_counters(wsx.buf, n, d, k)
    idx_x, den_x = hard_assign(z, w, exact=True)
    assert torch.equal(den, den_x)
    bad = (idx != idx_x).nonzero().flatten()
    assert bad.numel() == 0, f"{bad.numel()} rows differ, first {bad[:8].tolist()}"
    assert cnt["full_scan_rows"] == 0


def test_degenerate_codebooks(oracle, dev, assign_route):
    """all-zero codebook (k-means pending: every code ties -> index 0) and massive duplication
    (candidate list overflows -> exact full scan fallback)."""
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import hard_assign
    g = torch.Generator().manual_seed(5)
    z = torch.randn(300, 64, generator=g)
    w0 = torch.zeros(512, 64)
    idx, _ = hard_assign(z.to(dev), w0.to(dev))
    assert (idx == 0).all()
    w = torch.nn.functional.normalize(torch.randn(8, 64, generator=g), dim=-1).repeat(64, 1)   # each row 64 times
    wsx = _lib.Workspace()
    idx, _ = hard_assign(z.to(dev), w.to(dev), workspace=wsx)
    zf, _ = oracle.C().l2norm_rows(z.numpy())
    idx_ref, _ = oracle.C().hard_assign(zf, w.numpy())
    assert np.array_equal(idx.cpu().numpy(), idx_ref)
    assert (idx < 8).all()
    # tensor route: every candidate list overflows -> exact full scan; the small-batch route scores everything exactly
    assert _lib.debug_counters(wsx.buf, 300, 64, 512)["full_scan_rows"] == (300 if assign_route == "tensor" else 0)


def test_forward_matches_oracle_random(oracle, dev):
    """Training forward on a mid-size seeded problem vs the oracle (statistics included)."""
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    n, d, k = 30000, 64, 4096
    g = torch.Generator().manual_seed(99)
    z = torch.randn(n, d, generator=g)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1)
    cs0 = torch.rand(k, generator=g)
    r = oracle.hard_forward(z.numpy(), w.numpy(), cs0.numpy(), 0.25, 0.99)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    q.embedding.weight.data.copy_(w.to(dev))
    q.cluster_size.copy_(cs0.to(dev))
    with torch.no_grad():
        z_q, loss, idx = q(z.view(n, d, 1, 1).to(dev))
    assert np.array_equal(idx.cpu().numpy(), r.idx)
    assert np.array_equal(q.cluster_size.cpu().numpy(), r.cluster_size)
    assert np.array_equal(z_q.view(n, d).cpu().numpy(), r.z_q)
    st = q.last_stats.cpu().numpy()
    assert np.array_equal(st[:k], r.bins)
    assert np.allclose(st[k:k + k * d].reshape(k, d), r.embed_sum, rtol=0, atol=1e-9)
    assert abs(st[k + k * d] - r.sse) <= 1e-9 * r.sse and st[k + k * d + 1] == n
    assert abs(float(loss) - r.loss) <= 1e-6 * r.loss
    assert np.max(np.abs(q.embedding.weight.cpu().numpy() - r.weight)) <= 1.2e-7   # <= 1 ulp of a unit-norm row


def test_gradient_is_straight_through_plus_commitment(dev):
    """autograd contract (reference :212-214): d/dz of (z_q.sum*g + loss) == torch autograd of the
    reference formula evaluated with the same indices."""
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    n, d, k = 257, 64, 512
    g = torch.Generator().manual_seed(17)
    z = torch.randn(n, d, 1, 1, generator=g).to(dev).requires_grad_(True)
    gout = torch.randn(n, d, 1, 1, generator=g).to(dev)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    w0 = q.embedding.weight.data.clone()
    z_q, loss, idx = q(z)
    ((z_q * gout).sum() + 3.0 * loss).backward()
    z2 = z.detach().clone().requires_grad_(True)
    zf = torch.nn.functional.normalize(z2.view(n, d), dim=-1)
    zq_ref = w0[idx]
    loss_ref = 0.25 * torch.nn.functional.mse_loss(zq_ref, zf)
    out_ref = zf + (zq_ref - zf).detach()
    ((out_ref * gout.view(n, d)).sum() + 3.0 * loss_ref).backward()
    assert torch.allclose(z.grad, z2.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.single_route
def test_full_size_config2_matches_oracle_on_every_row(oracle, dev):
    """BASELINE config 2 (N=2^20, D=64, K=8192), SURVEY 8(d): the training forward against the CPU oracle on the FULL N
    (OpenMP C restatement, ~20-40 s of host time): code indices exact on all 2^20 rows, cluster_size exact, z_q exact,
    updated codebook <= 1e-6, loss rel 1e-5; plus the exact CUDA-core kernel on every row."""
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, hard_assign
    n, d, k = 1 << 20, 64, 8192
    torch.manual_seed(0)
    z = torch.randn(n, d)
    torch.manual_seed(1)
    w = torch.nn.functional.normalize(torch.randn(k, d), dim=-1)
    cs0 = torch.rand(k, generator=torch.Generator().manual_seed(2))
    zd, wd = z.to(dev), w.to(dev)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    q.embedding.weight.data.copy_(wd)
    q.cluster_size.copy_(cs0.to(dev))
    with torch.no_grad():
        z_q, loss, idx = q(zd.view(n, d, 1, 1))
    cnt = _lib.debug_counters(q._ws.buf, n, d, k)
    idx_x, _ = hard_assign(zd, wd, exact=True)
    mism = (idx != idx_x).nonzero().flatten()
    assert mism.numel() == 0, f"{mism.numel()} rows differ from the exact kernel, first {mism[:8].tolist()}"
    r = oracle.hard_forward(z.numpy(), w.numpy(), cs0.numpy(), 0.25, 0.99)
    bad = np.nonzero(idx.cpu().numpy() != r.idx)[0]
    assert bad.size == 0, f"{bad.size} of {n} rows differ from the CPU oracle, first {bad[:8].tolist()}"
    assert np.array_equal(q.cluster_size.cpu().numpy(), r.cluster_size)
    assert np.array_equal(z_q.view(n, d).cpu().numpy(), r.z_q)
    st = q.last_stats.cpu().numpy()
    assert np.array_equal(st[:k], r.bins) and st[:k].sum() == n
    assert np.allclose(st[k:k + k * d].reshape(k, d), r.embed_sum, rtol=0, atol=1e-8)
    assert abs(float(loss) - r.loss) <= 1e-5 * r.loss
    assert np.max(np.abs(q.embedding.weight.cpu().numpy() - r.weight)) <= 1e-6
    wn = q.embedding.weight.norm(dim=-1)
    assert (wn - 1).abs().max().item() < 1e-6                              # rows stay unit norm
    assert cnt["full_scan_rows"] == 0
    print("config2 near-tie rows:", cnt)


@pytest.mark.single_route
def test_full_size_clustered_zipf_properties(dev):
    """SURVEY 8d second distribution at config-2 size: z = l2norm(C[j] + 0.1 randn), j ~ Zipf(1.5) over K (heavily skewed
    bins, thousands of dead codes, many near-ties around the popular codes): tensor-core path == exact CUDA-core kernel on
    every row, histogram / dead-code / unit-norm invariants of the EMA update."""
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer, hard_assign
    n, d, k = 1 << 20, 64, 8192
    g = torch.Generator().manual_seed(11)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1)
    ranks = torch.arange(1, k + 1, dtype=torch.float64)
    pz = (1.0 / ranks ** 1.5)
    j = torch.multinomial(pz / pz.sum(), n, replacement=True, generator=g)
    z = torch.nn.functional.normalize(w[j] + 0.1 * torch.randn(n, d, generator=g), dim=-1)
    zd, wd = z.to(dev), w.to(dev)
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
    q.embedding.weight.data.copy_(wd)
    with torch.no_grad():
        z_q, loss, idx = q(zd.view(n, d, 1, 1))
    cnt = _lib.debug_counters(q._ws.buf, n, d, k)
    idx_x, _ = hard_assign(zd, wd, exact=True)
    mism = (idx != idx_x).nonzero().flatten()
    assert mism.numel() == 0, f"{mism.numel()} rows differ from the exact kernel, first {mism[:8].tolist()}"
    bins = torch.bincount(idx, minlength=k).double()
    assert torch.equal(q.last_stats[:k], bins) and int(bins.sum()) == n
    assert (idx.cpu() == j).float().mean().item() > 0.95                     # rows land on their own cluster centre
    dead = bins == 0
    assert int(dead.sum()) > 100                                             # the Zipf tail leaves many codes unused
    assert torch.allclose(q.embedding.weight[dead], wd[dead], atol=2e-7)     # dead codes keep their row
    assert (q.embedding.weight.norm(dim=-1) - 1).abs().max().item() < 1e-6
    assert torch.equal(q.cluster_size, (0.01 * bins).float()) or torch.allclose(q.cluster_size, (0.01 * bins).float(), rtol=1e-6)
    ref_loss = 0.25 * ((wd[idx] - zd).double() ** 2).mean().item()
    assert abs(float(loss) - ref_loss) <= 1e-5 * ref_loss
    print("clustered/Zipf counters:", cnt, "dead codes:", int(dead.sum()))


def test_kmeans_matches_reference_golden(dev):
    """k-means codebook init (norm_ema_quantizer.py:24-57, cosine branch) on the CUDA kernels against the golden of the
    unmodified reference: cluster sizes of the last iteration exact, centres within 1e-6 (fp64 vs fp32 scatter_add)."""
    from medtok_b200.norm_ema_quantizer import kmeans
    meta, a = load_golden("kmeans_n3000_d64_k256")
    samples = torch.from_numpy(a["samples"]).to(dev)
    means, bins = kmeans(samples, a["init"].shape[0], meta["iters"], use_cosine_sim=True,
                         means=torch.from_numpy(a["init"]).to(dev), verbose=False)
    assert bins.dtype == torch.int64 and np.array_equal(bins.cpu().numpy(), a["bins"])
    assert np.max(np.abs(means.cpu().numpy() - a["means"])) <= 1e-6


def test_kmeans_init_through_the_module(dev, oracle):
    """kmeans_init=True: the first forward initialises the codebook from the (normalised) batch (:85-93,173); with the
    same initial centres the oracle's k-means gives the same codebook, and `initted` flips."""
    from medtok_b200 import norm_ema_quantizer as m
    g = torch.Generator().manual_seed(11)
    z = torch.randn(2048, 64, generator=g)
    q = m.NormEMAVectorQuantizer(128, 64, 0.25, kmeans_init=True).to(dev).eval()
    assert float(q.embedding.initted) == 0.0
    zf = torch.nn.functional.normalize(z, dim=-1)
    picked = {}
    orig = m.sample_vectors

    def pick(samples, num):
        picked["init"] = samples[:num].clone()
        return picked["init"]
    m.sample_vectors = pick
    try:
        with torch.no_grad():
            q(z.view(2048, 64, 1, 1).to(dev))
    finally:
        m.sample_vectors = orig
    assert float(q.embedding.initted) == 1.0
    # the module normalises the batch with torch on the device; feed the oracle the same normalised rows
    zf_dev = torch.nn.functional.normalize(z.to(dev), dim=-1).cpu().numpy()
    means, bins = oracle.kmeans_cosine(zf_dev, picked["init"].cpu().numpy(), 10)
    # the forward that follows the init does not update the codebook in eval mode
    assert np.max(np.abs(q.embedding.weight.cpu().numpy() - means)) <= 1e-6
    assert np.array_equal(q.embedding.cluster_size.cpu().numpy(), bins.astype(np.float32))


@pytest.mark.single_route
def test_training_forward_is_reproducible_run_to_run(dev):
    """The per-code sums are accumulated with fp64 atomics (order-dependent at ~1e-16 relative), then rounded to fp32
    before the EMA: two runs of the same three training steps must give the same indices / cluster sizes bit for bit and
    codebooks that agree to the last fp32 bit except where an fp64 sum sits within 1e-16 of a rounding boundary --
    stated tolerance 1 ulp of a unit-norm row (1.2e-7); in practice the runs are bit-equal."""
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    n, d, k = 200_000, 64, 4096
    g = torch.Generator().manual_seed(21)
    zs = [torch.randn(n, d, 1, 1, generator=g).to(dev) for _ in range(3)]
    w0 = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1).to(dev)
    runs = []
    for _ in range(2):
        q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
        q.embedding.weight.data.copy_(w0)
        with torch.no_grad():
            for z in zs:
                z_q, loss, idx = q(z)
        runs.append((q.embedding.weight.detach().clone(), q.cluster_size.clone(), idx.clone(), z_q.clone()))
    assert torch.equal(runs[0][2], runs[1][2]) and torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][3], runs[1][3])
    assert float((runs[0][0] - runs[1][0]).abs().max()) <= 1.2e-7
    print("codebooks bit-equal across runs:", bool(torch.equal(runs[0][0], runs[1][0])))

