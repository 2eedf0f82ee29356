"""GPU parity tests of the two-pass hard route (record-free GROUP tensor pass -> rows sorted by their best 128-code chunk
-> REFINE pass fused with the row work; medtok_b200/csrc/assign_tc.cu, group.cu): forced with
mtvq_debug_set_hard_route(2) at sizes the CPU oracle finishes in seconds, compared with the oracle, with the exact
CUDA-core scan and with the single-pass route (route 1) from the same initial state.  The full-size cases of
test_gpu_hard.py take this route automatically (N >= 2^18, D <= 64)."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture()
def route():
    from medtok_b200 import _lib

    def set_route(r):
        _lib.check(_lib.lib().mtvq_debug_set_hard_route(r), "route")
    yield set_route
    _lib.lib().mtvq_debug_set_hard_route(-1)


def _forward(dev, z, w, cs0, training=True, update=True):
    from medtok_b200 import _lib
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    n, d = z.shape
    k = w.shape[0]
    q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train(training)
    q.embedding.update = update
    q.embedding.weight.data.copy_(w.to(dev))
    q.cluster_size.copy_(cs0.to(dev))
    with torch.no_grad():
        z_q, loss, idx = q(z.view(n, d, 1, 1).to(dev))
    cnt = _lib.debug_counters(q._ws.buf, n, d, k)
    return dict(idx=idx.cpu().numpy(), z_q=z_q.view(n, d).cpu().numpy(), loss=float(loss), w=q.embedding.weight.cpu().numpy(),
                cs=q.cluster_size.cpu().numpy(), stats=q.last_stats.cpu().numpy(), cnt=cnt)


# (rows, dim, codes): 4-row-tile blocks need >= 2 * 148 * 512 rows; K % 128 != 0 keeps the bias K-step and padding codes on
SHAPES = [(160000, 64, 1000), (200000, 32, 4096), (155000, 16, 300), (180001, 48, 2048)]


@pytest.mark.parametrize("n,d,k", SHAPES)
def test_two_pass_matches_oracle_and_single_pass(oracle, dev, route, n, d, k):
    g = torch.Generator().manual_seed(n + d + k)
    z = torch.randn(n, d, generator=g)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1)
    cs0 = torch.rand(k, generator=g)
    route(2)
    b = _forward(dev, z, w, cs0)
    route(1)
    a = _forward(dev, z, w, cs0)
    r = oracle.hard_forward(z.numpy(), w.numpy(), cs0.numpy(), 0.25, 0.99)
    # bit-exact: indices, counts, quantised rows -- against the oracle and between the routes
    assert np.array_equal(b["idx"], r.idx) and np.array_equal(a["idx"], r.idx)
    assert np.array_equal(b["cs"], r.cluster_size)
    assert np.array_equal(b["z_q"], r.z_q) and np.array_equal(a["z_q"], b["z_q"])
    st = b["stats"]
    assert np.array_equal(st[:k], r.bins) and st[k + k * d + 1] == n
    assert np.allclose(st[k:k + k * d].reshape(k, d), r.embed_sum, rtol=0, atol=1e-9)     # fp64 sums, other order
    assert abs(st[k + k * d] - r.sse) <= 1e-7 * r.sse          # a lane's four squares are added in fp32 first
    assert abs(b["loss"] - r.loss) <= 1e-6 * r.loss
    assert np.max(np.abs(b["w"] - r.weight)) <= 1.2e-7          # <= 1 ulp of a unit-norm row
    assert b["cnt"]["near_tie_rows"] > 0                        # the near-tie machinery was exercised


def test_two_pass_eval_and_assign_only(oracle, dev, route):
    """No codebook update (eval forward: the REFINE pass only decides, the streaming row kernels do the rest) and the
    assignment-only entry point with a caller-provided norm buffer."""
    from medtok_b200.norm_ema_quantizer import hard_assign
    n, d, k = 170000, 64, 777
    g = torch.Generator().manual_seed(5)
    z = torch.randn(n, d, generator=g)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1)
    cs0 = torch.rand(k, generator=g)
    zf, den = oracle.C().l2norm_rows(z.numpy())
    idx_ref, _ = oracle.C().hard_assign(zf, w.numpy())
    route(2)
    e = _forward(dev, z, w, cs0, training=False)
    assert np.array_equal(e["idx"], idx_ref)
    assert np.array_equal(e["w"], w.numpy())                    # eval: codebook untouched
    idx, denom = hard_assign(z.to(dev), w.to(dev))
    assert np.array_equal(idx.cpu().numpy(), idx_ref)
    assert np.array_equal(denom.cpu().numpy(), den)
    route(1)
    e1 = _forward(dev, z, w, cs0, training=False)
    assert np.array_equal(e1["z_q"], e["z_q"]) and e1["loss"] == e["loss"] and np.array_equal(e1["cs"], e["cs"])


def test_two_pass_degenerate_inputs(dev, route):
    """Duplicated codes (every row near-ties across chunks -> listed twice / COLLECT / exact scan), a block of identical
    rows, zero and NaN rows: the two-pass route must agree with the exact CUDA-core scan row for row, and the statistics
    with the single-pass route."""
    from medtok_b200.norm_ema_quantizer import hard_assign
    n, d, k = 160000, 64, 1024
    g = torch.Generator().manual_seed(17)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1)
    w[300:340] = w[5]              # 41 copies of one code spread over several chunks
    w[900] = w[130]                # an exact duplicate in another chunk
    z = torch.randn(n, d, generator=g)
    z[:4096] = w[5] + 1e-3 * torch.randn(4096, d, generator=g)    # rows whose best codes tie exactly
    z[5000:5200] = z[5000]
    z[6000:6010] = 0.0
    z[7000, 3] = float("nan")
    zd, wd = z.to(dev), w.to(dev)
    route(2)
    i2, _ = hard_assign(zd, wd)
    ix, _ = hard_assign(zd, wd, exact=True)
    assert torch.equal(i2, ix)
    cs0 = torch.zeros(k)
    b = _forward(dev, z, w, cs0)
    route(1)
    a = _forward(dev, z, w, cs0)
    ok = ~np.isnan(a["z_q"]).any(axis=1)
    assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["cs"], b["cs"])
    assert np.array_equal(a["z_q"][ok], b["z_q"][ok])
    finite = np.isfinite(a["w"]).all(axis=1) & np.isfinite(b["w"]).all(axis=1)
    assert np.array_equal(np.isfinite(a["w"]).all(axis=1), np.isfinite(b["w"]).all(axis=1))
    assert np.max(np.abs(a["w"][finite] - b["w"][finite])) <= 1.2e-7


def test_two_pass_accumulates_like_single_pass(dev, route):
    """MTVQ_F_SKIP_FINALIZE statistics (the multi-GPU path's packed buffer) of the two routes agree: counts exactly, sums
    to fp64 summation-order accuracy."""
    import ctypes
    from medtok_b200 import _lib
    n, d, k = 200000, 64, 2048
    g = torch.Generator().manual_seed(23)
    z = torch.randn(n, d, generator=g).to(dev)
    w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1).to(dev)
    L = _lib.lib()
    out = {}
    for r in (1, 2):
        route(r)
        ws = _lib.Workspace().get(n, d, k, 0, dev)
        z_q = torch.empty_like(z)
        idx = torch.empty(n, dtype=torch.int64, device=dev)
        loss = torch.empty(1, device=dev)
        stats = torch.empty(_lib.stats_len(d, k), dtype=torch.float64, device=dev)
        flags = _lib.F_TRAINING | _lib.F_UPDATE | _lib.F_SKIP_FINALIZE
        wc = w.clone()
        _lib.check(L.mtvq_hard_forward(_lib.ptr(z), n, d, _lib.NORM_CONTIGUOUS, _lib.ptr(wc), k, None, 0.25, 0.99, flags,
                                       _lib.ptr(z_q), _lib.ptr(idx), _lib.ptr(loss), _lib.ptr(stats), _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr(dev)), "hard_forward")
        torch.cuda.synchronize()
        assert torch.equal(wc, w)                                   # SKIP_FINALIZE leaves the codebook alone
        out[r] = (idx.cpu(), z_q.cpu(), stats.cpu(), float(loss))
    assert torch.equal(out[1][0], out[2][0]) and torch.equal(out[1][1], out[2][1])
    s1, s2 = out[1][2].numpy(), out[2][2].numpy()
    assert np.array_equal(s1[:k], s2[:k]) and s1[-1] == s2[-1] == n
    assert np.allclose(s1[k:k + k * d], s2[k:k + k * d], rtol=0, atol=1e-10)
    assert abs(s1[k + k * d] - s2[k + k * d]) <= 1e-7 * s1[k + k * d]
    assert abs(out[1][3] - out[2][3]) <= 1e-6 * out[1][3]


def test_two_pass_strided_input_layout(oracle, dev, route):
    """'b c h w' input with h * w > 1 (reference :169-171: the rows are a permuted view and torch normalises the strided
    view with a sequential sum): the GROUP loader's strided norm, the norm carried in the sorted entries and the row work
    of the REFINE pass must reproduce the oracle's strided-norm results bit for bit."""
    from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer
    b, c, h, w_ = 600, 32, 16, 16                      # 153600 rows
    k = 520
    g = torch.Generator().manual_seed(31)
    z = torch.randn(b, c, h, w_, generator=g)
    w = torch.nn.functional.normalize(torch.randn(k, c, generator=g), dim=-1)
    rows = z.permute(0, 2, 3, 1).reshape(-1, c).contiguous()
    r = oracle.hard_forward(rows.numpy(), w.numpy(), np.zeros(k, np.float32), 0.25, 0.99, strided_norm=True)
    route(2)
    q = NormEMAVectorQuantizer(k, c, 0.25).to(dev).train()
    q.embedding.weight.data.copy_(w.to(dev))
    with torch.no_grad():
        z_q, loss, idx = q(z.to(dev))
    assert z_q.shape == z.shape
    assert np.array_equal(idx.cpu().numpy(), r.idx)
    assert np.array_equal(z_q.permute(0, 2, 3, 1).reshape(-1, c).cpu().numpy(), r.z_q)
    assert np.array_equal(q.cluster_size.cpu().numpy(), r.cluster_size)
    assert abs(float(loss) - r.loss) <= 1e-6 * r.loss
    assert np.max(np.abs(q.embedding.weight.cpu().numpy() - r.weight)) <= 1.2e-7


def test_route_choice_is_reported(dev):
    """mtvq_hard_route: the automatic rule (D <= 64, N >= 2^18, K x D >= 2^19) and the overrides."""
    from medtok_b200 import _lib
    L = _lib.lib()
    assert L.mtvq_hard_route(1 << 20, 64, 8192) == 2
    assert L.mtvq_hard_route(1 << 20, 64, 4096) == 1          # tensor pass too short for the longer tail to pay
    assert L.mtvq_hard_route(1 << 20, 128, 8192) == 1         # D > 64: two row tiles per block, single pass
    assert L.mtvq_hard_route(100000, 64, 8192) == 1
    _lib.check(L.mtvq_debug_set_hard_route(2), "route")
    try:
        assert L.mtvq_hard_route(160000, 32, 600) == 2
        assert L.mtvq_hard_route(1000, 32, 600) == 1          # not enough rows for four-row-tile blocks
    finally:
        L.mtvq_debug_set_hard_route(-1)
    _lib.check(L.mtvq_debug_set_hard_route(1), "route")
    try:
        assert L.mtvq_hard_route(1 << 20, 64, 8192) == 1
    finally:
        L.mtvq_debug_set_hard_route(-1)
