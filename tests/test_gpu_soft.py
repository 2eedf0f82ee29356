"""GPU parity tests of the soft-one (top-k) quantizer against golden outputs of the unmodified
reference (forward values, indices, weights, losses, usage, and the autograd gradients)."""
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


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("name", golden_names("soft_specific"))
def test_specific_embedding_matches_reference(name, dev):
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    meta, a = load_golden(name)
    n_e, e_dim, kk = meta["n_e"], meta["e_dim"], meta["k"]
    sq = VectorQuantizer(n_e=n_e, e_dim=e_dim, beta=meta["beta"], entropy_loss_ratio=0.0, l2_norm=True,
                         show_usage=True, split=[e_dim, e_dim], k=kk).to(dev).train()
    assert sorted(sq.state_dict().keys()) == meta["state_dict_keys"]
    sq.codebook.weight.data.copy_(_t(a["codebook"], dev))
    for types, proj in (("text", sq.proj_text), ("graph", sq.proj_graph)):
        proj.weight.data.copy_(_t(a[f"{types}_proj_w"], dev))
        proj.bias.data.copy_(_t(a[f"{types}_proj_b"], dev))
        x_in = _t(a[f"{types}_x_in"], dev).requires_grad_(True)
        sq.zero_grad()
        z_q, (vq, commit, xn, zq2), usage = sq.specific_embedding(x_in, types=types)
        idx, wts = sq._last_specific
        # the GPU's nn.Linear (cuBLAS) is not bit-identical to the CPU one, so compare on the values the
        # quantizer actually saw: indices exact wherever the reference's own top-k margin is not fp32 noise
        ref_idx = a[f"{types}_idx"]
        same = (idx.cpu().numpy() == ref_idx)
        vals = a[f"{types}_vals"]
        gaps = np.diff(vals, axis=1).min(axis=1)
        assert same[gaps > 1e-5].all(), f"{name}/{types}: top-k indices differ outside near-ties"
        assert same.mean() > 0.995
        ok = same.all(axis=1)
        assert np.max(np.abs(wts.detach().cpu().numpy()[ok] - a[f"{types}_wts"][ok])) <= 2e-6
        assert np.max(np.abs(z_q.detach().cpu().numpy()[ok] - a[f"{types}_z_q"][ok])) <= 5e-6
        assert np.max(np.abs(xn.detach().cpu().numpy() - a[f"{types}_xn"])) <= 2e-6
        assert abs(float(vq) - float(a[f"{types}_vq"][0])) <= 2e-5 * float(a[f"{types}_vq"][0])
        assert abs(float(commit) - float(a[f"{types}_commit"][0])) <= 2e-5 * float(a[f"{types}_commit"][0])
        if ok.all():
            assert usage == float(a[f"{types}_usage"][0])
        # gradients (reference autograd through topk/softmax/gather/normalize)
        total = vq + commit + (z_q * _t(a[f"{types}_gout"], dev)).sum()
        total.backward()
        for got, key, tol in ((x_in.grad, "grad_x_in", 2e-5), (sq.codebook.weight.grad, "grad_codebook", 2e-5),
                              (proj.weight.grad, "grad_proj_w", 2e-4), (proj.bias.grad, "grad_proj_b", 2e-4)):
            ref = a[f"{types}_{key}"]
            err = np.max(np.abs(got.cpu().numpy() - ref))
            assert err <= tol * max(1.0, np.abs(ref).max()), f"{name}/{types}/{key}: max err {err}"
    # eval: zero losses, same quantised values
    sq.eval()
    with torch.no_grad():
        z_q, (vq, commit, xn, _), usage = sq.specific_embedding(_t(a["eval_x_in"], dev), types="text")
    assert float(vq) == 0.0 and float(commit) == 0.0
    assert np.max(np.abs(z_q.cpu().numpy() - a["eval_z_q"])) <= 5e-6


def test_soft_core_matches_oracle(oracle, dev):
    """The quantisation core alone (no nn.Linear in front): indices and distances bit-exact vs the oracle."""
    from medtok_b200 import _lib
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    # (top-1 is a legal constructor argument of the reference, k=1: it needs the soft workspace buffers too)
    for (n, d, k, kk) in ((4000, 64, 7000, 5), (1500, 128, 16384, 5), (999, 32, 333, 3), (600, 256, 1024, 8),
                          (777, 64, 500, 1),
                          # MedTok's training batches (train_MedTok.py:387): the small-batch route of small.cu
                          (1, 64, 7000, 5), (255, 64, 21000, 5), (256, 64, 7000, 5), (1024, 64, 21000, 5), (300, 128, 4000, 8)):
        g = torch.Generator().manual_seed(n + k)
        x = torch.randn(n, d, generator=g) * 1.7
        table = torch.randn(k, d, generator=g)
        r = oracle.soft_core(x.numpy(), table.numpy(), kk, 0.25, training=True)
        ws = _lib.Workspace()
        out, vq, commit, idx, wts = _SoftQuantizeFn.apply(x.to(dev), table.to(dev), 0, k, kk, 0.25, True, ws)
        cnt = _lib.debug_counters(ws.buf, n, d, k, kk)
        assert np.array_equal(idx.cpu().numpy(), r.idx), f"top-k indices differ (n={n}, d={d}, k={k})"
        assert np.max(np.abs(wts.cpu().numpy() - r.weights)) <= 1e-6
        assert np.max(np.abs(out.cpu().numpy() - r.z_q)) <= 2e-6
        assert abs(float(vq) - r.vq_loss) <= 1e-5 * r.vq_loss
        assert abs(float(commit) - r.commit_loss) <= 1e-5 * r.commit_loss
        assert cnt["full_scan_rows"] <= 0.02 * n, cnt


@pytest.mark.parametrize("mode", ["eval"])
def test_full_forward_matches_reference(mode, dev):
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    meta, a = load_golden(f"soft_forward_{mode}_ne1536_d64")
    sq = VectorQuantizer(n_e=meta["n_e"], e_dim=meta["e_dim"], beta=meta["beta"], entropy_loss_ratio=0.0,
                         l2_norm=True, show_usage=True, split=[meta["e_dim"]] * 2, k=meta["k"])
    sd = {k[4:]: torch.from_numpy(v) for k, v in a.items() if k.startswith("sd::")}
    sq.load_state_dict(sd, strict=True)                      # reference checkpoint loads unchanged
    sq = sq.to(dev).train(mode == "train")
    with torch.no_grad():
        out = sq(_t(a["z"], dev), _t(a["text_features"], dev), _t(a["graph_nodes"], dev), _t(a["mask"], dev),
                 _t(a["batch"], dev), _t(a["z_aug"], dev))
    for key in meta["out_keys"]:
        assert key in out
    for key in ("shared_text_tokens", "shared_graph_tokens", "text_tokens", "graph_tokens"):
        assert out[key].shape == (6, meta["k"]) and out[key].dtype == torch.int64
        assert out[key + "_weights"].shape == (6, meta["k"])
        assert torch.allclose(out[key + "_weights"].sum(-1), torch.ones(6, device=dev), atol=1e-6)
    for key in ("graph_feature", "text_feature", "shared_text_embedding", "shared_graph_embedding",
                "specific_embedding_text", "specific_embedding_graph", "specific_embedding_text_aug",
                "specific_embedding_graph_aug"):
        err = np.max(np.abs(out[key].cpu().numpy() - a["out::" + key]))
        assert err <= 2e-5, f"{key}: {err}"
    for key in ("shared_codebook_usage", "text_specific_usage", "graph_specific_usage"):
        assert isinstance(out[key], float) and out[key] == float(a["out::" + key][0])
    for key in ("shared_embed_loss", "text_specific_loss", "graph_specific_loss"):
        assert float(out[key][0]) == 0.0 and float(out[key][1]) == 0.0


def test_full_forward_train_mode_matches_reference(dev):
    """Train-mode forward against the golden of the unmodified reference: every key that does not pass through the
    cross-attention dropout (a different RNG stream on the GPU) -- the four specific quantisations with their loss
    tuples (vq, commit, normalised input, z_q) and the pass-through features -- is compared by value; the shared
    (dropout-dependent) entries are checked for structure."""
    from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer
    meta, a = load_golden("soft_forward_train_ne1536_d64")
    sq = VectorQuantizer(n_e=meta["n_e"], e_dim=meta["e_dim"], beta=meta["beta"], entropy_loss_ratio=0.0,
                         l2_norm=True, show_usage=True, split=[meta["e_dim"]] * 2, k=meta["k"])
    sd = {k[4:]: torch.from_numpy(v) for k, v in a.items() if k.startswith("sd::")}
    sq.load_state_dict(sd, strict=True)
    sq = sq.to(dev).train()
    out = sq(_t(a["z"], dev), _t(a["text_features"], dev), _t(a["graph_nodes"], dev), _t(a["mask"], dev),
             _t(a["batch"], dev), _t(a["z_aug"], dev))
    assert set(meta["out_keys"]) <= set(out.keys())
    for key in ("graph_feature", "text_feature", "specific_embedding_text", "specific_embedding_graph",
                "specific_embedding_text_aug", "specific_embedding_graph_aug"):
        err = np.max(np.abs(out[key].detach().cpu().numpy() - a["out::" + key]))
        assert err <= 2e-5, f"{key}: {err}"
    for key in ("text_specific_loss", "graph_specific_loss"):
        tup = out[key]
        assert len(tup) == 4
        for j in (0, 1):                                         # vq loss, commitment loss (train: device scalars)
            ref = float(a[f"out::{key}::{j}"])
            assert abs(float(tup[j]) - ref) <= 2e-5 * abs(ref) + 1e-9, f"{key}[{j}]: {float(tup[j])} vs {ref}"
        for j in (2, 3):                                         # normalised input, z_q
            err = np.max(np.abs(tup[j].detach().cpu().numpy() - a[f"out::{key}::{j}"]))
            assert err <= 2e-5, f"{key}[{j}]: {err}"
    for key in ("text_specific_usage", "graph_specific_usage"):
        assert isinstance(out[key], float)
    sh = out["shared_embed_loss"]
    assert len(sh) == 6 and sh[0].dim() == 0 and sh[1].dim() == 0 and float(sh[0]) > 0.0 and float(sh[1]) > 0.0
    for j in (2, 3, 4, 5):
        assert tuple(sh[j].shape) == a[f"out::shared_embed_loss::{j}"].shape
    assert out["shared_text_embedding"].shape == out["shared_graph_embedding"].shape == (6, meta["e_dim"])
    # the whole dict is differentiable in train mode (train_MedTok.py:213-240 sums these losses and calls backward)
    total = sh[0] + sh[1] + out["text_specific_loss"][0] + out["text_specific_loss"][1] + \
        out["graph_specific_loss"][0] + out["graph_specific_loss"][1]
    total.backward()
    assert sq.codebook.weight.grad is not None and torch.isfinite(sq.codebook.weight.grad).all()


@pytest.mark.single_route
@pytest.mark.parametrize("n,d,k,kk", [(60000, 128, 16384, 5), (50000, 64, 8192, 5), (20000, 256, 4096, 8),
                                      (30000, 32, 1000, 3), (40000, 64, 21000, 5)])
def test_single_pass_equals_two_pass_and_exact_scan(dev, n, d, k, kk):
    """The production route (one tensor pass, records in the L2 scratch) must give bit-identical top-k indices,
    distances, weights and outputs to the two-pass route and to the exact CUDA-core scan of every code."""
    from medtok_b200 import _lib
    from medtok_b200.vector_quantization_soft_one_new import _SoftQuantizeFn
    g = torch.Generator().manual_seed(n + d)
    x = (torch.randn(n, d, generator=g) * 0.7).to(dev)
    table = torch.randn(k, d, generator=g).to(dev)
    table[5] = table[4]                                          # duplicated codes: exact ties, lower index first
    res = []
    try:
        for route in (0, 1, 2):
            _lib.check(_lib.lib().mtvq_debug_set_soft_route(route), "route")
            ws = _lib.Workspace()
            with torch.no_grad():
                out, vq, commit, idx, wts = _SoftQuantizeFn.apply(x, table, 0, k, kk, 0.25, True, ws)
            cnt = _lib.debug_counters(ws.buf, n, d, k, kk)
            res.append((out, vq, idx, wts, cnt))
    finally:
        _lib.lib().mtvq_debug_set_soft_route(0)
    for r in res[1:]:
        assert torch.equal(res[0][2], r[2]), "top-k indices differ between routes"
        assert torch.equal(res[0][3], r[3]) and torch.equal(res[0][0], r[0])
        assert abs(float(res[0][1]) - float(r[1])) <= 1e-6 * float(r[1])
    assert res[0][4]["full_scan_rows"] <= 0.01 * n, res[0][4]
    assert res[2][4]["full_scan_rows"] == n
