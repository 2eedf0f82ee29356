"""Pins the CPU oracle against the UNMODIFIED reference modules run live (oracle/ref_loader.py: /root/reference in the
authoring container, the byte-identical staged copies under oracle/_ref elsewhere) on seeded random shapes the committed
fixtures do not cover: a sweep of (N, D, K) incl. K % 3 != 0, N % 128 != 0, D = 16 ... 256, several consecutive training
steps carrying the EMA state, h*w > 1 layouts, duplicated codes / duplicated rows / zero rows.  CPU only; skipped when the
reference files are not present.  Same gates as tests/test_oracle_golden.py: indices, cluster sizes and z_q bit-exact,
loss rel 1e-5, updated codebook abs 1e-6 (the reference's embed_sum is an fp32 GEMM over N whose order is opaque)."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference modules not present")


@pytest.fixture(scope="module")
def ref():
    with contextlib.redirect_stdout(io.StringIO()):
        return ref_loader.load()


def _rows(z4):
    return np.ascontiguousarray(np.transpose(z4, (0, 2, 3, 1)).reshape(-1, z4.shape[1]))


def _unrows(zq, shape):
    b, c, h, w = shape
    return np.transpose(zq.reshape(b, h, w, c), (0, 3, 1, 2))


HARD_SHAPES = [
    # b, d, h, w, k, steps, training
    (1, 64, 1, 1, 16, 1, True),            # a single row
    (127, 16, 1, 1, 100, 2, True),
    (129, 32, 1, 1, 1000, 3, True),        # K % 3 != 0, K % 128 != 0
    (333, 48, 1, 1, 517, 2, True),
    (512, 64, 1, 1, 4096, 2, True),
    (200, 96, 1, 1, 300, 2, True),
    (257, 128, 1, 1, 2048, 2, True),
    (100, 256, 1, 1, 640, 2, True),
    (7, 64, 5, 3, 256, 2, True),           # 'b c h w' with h*w > 1 (strided normalisation)
    (16, 32, 4, 4, 999, 2, True),
    (600, 64, 1, 1, 1024, 1, False),       # eval: usage EMA only
    (11, 128, 2, 2, 130, 1, False),
]


@pytest.mark.parametrize("case", range(len(HARD_SHAPES)))
def test_hard_forward_random_shapes(oracle, ref, case):
    hard, _ = ref
    b, d, h, w_, k, steps, training = HARD_SHAPES[case]
    g = torch.Generator().manual_seed(9000 + case)
    weight = hard.l2norm(torch.randn(k, d, generator=g))
    q = hard.NormEMAVectorQuantizer(n_embed=k, embedding_dim=d, beta=0.25, decay=0.99)
    q.embedding.weight.data.copy_(weight)
    q.train(training)
    w = weight.numpy().copy()
    cs = np.zeros(k, np.float32)
    for s in range(steps):
        z4 = torch.randn(b, d, h, w_, generator=g)
        with torch.no_grad():
            z_q, loss, idx = q(z4)
        r = oracle.hard_forward(_rows(z4.numpy()), w, cs, 0.25, 0.99, training=training, update=True,
                                strided_norm=h * w_ > 1)
        assert np.array_equal(r.idx, idx.numpy()), f"case {case} step {s}: code indices"
        assert np.array_equal(r.cluster_size, q.cluster_size.numpy()), f"case {case} step {s}: cluster_size"
        assert np.array_equal(_unrows(r.z_q, z4.shape), z_q.numpy()), f"case {case} step {s}: z_q"
        assert abs(r.loss - float(loss)) <= 1e-5 * abs(float(loss)) + 1e-12
        assert np.max(np.abs(r.weight - q.embedding.weight.numpy())) <= 1e-6
        # continue from the reference's own state so that every step is checked bit-exactly
        w, cs = q.embedding.weight.numpy().copy(), q.cluster_size.numpy().copy()


def test_hard_forward_degenerate_inputs(oracle, ref):
    """duplicated codes (first index wins), rows exactly on a code, zero rows (norm clamps), duplicated rows, a code
    nobody picks (bins == 0 keeps the old row), an un-normalised codebook"""
    hard, _ = ref
    g = torch.Generator().manual_seed(77)
    k, d, n = 384, 64, 300
    weight = torch.randn(k, d, generator=g) * 0.8                    # raw codebook
    weight[5] = weight[300]
    weight[301] = weight[300]
    weight[383] = -weight[0]
    z = torch.randn(n, d, generator=g)
    z[0] = weight[300] * 2.5
    z[1] = 0.0
    z[2] = weight[301]
    z[10:20] = z[20:30]
    z[40] = -z[41]
    q = hard.NormEMAVectorQuantizer(n_embed=k, embedding_dim=d, beta=0.3, decay=0.9).train()
    q.embedding.weight.data.copy_(weight)
    w, cs = weight.numpy().copy(), np.zeros(k, np.float32)
    for s in range(3):
        with torch.no_grad():
            z_q, loss, idx = q(z.view(n, d, 1, 1))
        r = oracle.hard_forward(z.numpy(), w, cs, 0.3, 0.9)
        assert np.array_equal(r.idx, idx.numpy())
        assert np.array_equal(r.cluster_size, q.cluster_size.numpy())
        assert np.array_equal(r.z_q, z_q.view(n, d).numpy())
        assert abs(r.loss - float(loss)) <= 1e-5 * abs(float(loss))
        assert np.max(np.abs(r.weight - q.embedding.weight.numpy())) <= 1e-6
        if s == 0:
            picked = set(idx.tolist())                               # of three equal codes only the first can win
            assert 300 not in picked and 301 not in picked
        w, cs = q.embedding.weight.numpy().copy(), q.cluster_size.numpy().copy()


SOFT_SHAPES = [(1, 16, 30, 1), (64, 32, 100, 5), (255, 64, 700, 5), (256, 64, 2333, 5), (130, 128, 512, 3),
               (50, 256, 64, 8), (300, 48, 9, 8)]


@pytest.mark.parametrize("case", range(len(SOFT_SHAPES)))
def test_soft_core_random_shapes(oracle, ref, case):
    """the quantisation core of VectorQuantizer.specific_embedding (soft:195-214) with the reference's own ops"""
    n, d, kprime, kk = SOFT_SHAPES[case]
    _, soft = ref
    g = torch.Generator().manual_seed(12000 + case)
    x = torch.randn(n, d, generator=g) * 1.7
    table = torch.randn(kprime, d, generator=g)
    xn = F.normalize(x, p=2, dim=-1)
    cn = F.normalize(table, p=2, dim=-1)
    dist = soft.VectorQuantizer.get_distance(None, xn, cn) if hasattr(soft.VectorQuantizer, "get_distance") else None
    if dist is None:
        pytest.skip("reference has no get_distance")
    vals, idx = torch.topk(dist, kk, dim=-1, largest=False)
    wts = F.softmax(-vals, dim=-1)
    zq = (wts.unsqueeze(-1) * cn[idx]).sum(dim=1)
    vq = torch.mean((zq - x) ** 2)
    commit = 0.25 * torch.mean((zq - x) ** 2)
    r = oracle.soft_core(x.numpy(), table.numpy(), kk, 0.25)
    assert np.array_equal(r.idx, idx.numpy()), f"soft case {case}: top-k indices"
    assert np.array_equal(r.vals, vals.numpy()), f"soft case {case}: top-k distances"
    assert np.max(np.abs(r.weights - wts.numpy())) <= 1e-6
    assert np.max(np.abs(r.zq_raw - zq.numpy())) <= 2e-6
    assert abs(r.vq_loss - float(vq)) <= 2e-5 * float(vq)
    assert abs(r.commit_loss - float(commit)) <= 2e-5 * float(commit)
