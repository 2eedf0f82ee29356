"""Generate tests/golden/*.npz by running the UNMODIFIED reference quantizer modules on CPU.

Run in the authoring container only (needs /root/reference):   python -m oracle.make_golden
The GPU box never runs this; it consumes the committed fixtures.

Harness-side accommodations (none of them edits the reference):
  * MedTok/__init__.py is bypassed and two unused transformers names are shimmed (ref_loader).
  * soft quantizer: `codebook_used` is re-registered as a plain buffer, because the reference
    registers an nn.Parameter there and its in-place window writes raise on CPU in grad mode.
  * N*k per soft call is kept <= 300000 (the usage window raises otherwise).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def save(name, meta, **arrays):
    os.makedirs(OUT, exist_ok=True)
    arrays = {k: (np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.frombuffer(json.dumps(meta).encode(), np.uint8),
                        **arrays)
    print("wrote", name, {k: tuple(v.shape) for k, v in arrays.items()})


# ------------------------------------------------------------------------------------------
def hard_case(hard, name, z4, weight, beta=0.25, decay=0.99, training=True, steps=1, update=True,
              extra_meta=None):
    """Run NormEMAVectorQuantizer.forward `steps` times on the same input; record every step."""
    k, d = weight.shape
    q = hard.NormEMAVectorQuantizer(n_embed=k, embedding_dim=d, beta=beta, decay=decay)
    q.embedding.weight.data.copy_(weight)
    q.embedding.update = update
    q.train(training)
    rec = {"z": _np(z4), "weight0": _np(weight)}
    with torch.no_grad():
        for s in range(steps):
            z_q, loss, idx = q(z4)
            rec[f"z_q_{s}"] = _np(z_q)
            rec[f"loss_{s}"] = _np(loss).reshape(1)
            rec[f"idx_{s}"] = _np(idx)
            rec[f"cluster_size_{s}"] = _np(q.cluster_size)
            rec[f"weight_{s}"] = _np(q.embedding.weight)
    meta = {"kind": "hard", "beta": beta, "decay": decay, "training": training, "steps": steps,
            "update": update, "state_dict_keys": sorted(q.state_dict().keys())}
    meta.update(extra_meta or {})
    save(name, meta, **rec)


def gen_hard(hard):
    g = torch.Generator().manual_seed(1234)

    def randn(*s):
        return torch.randn(*s, generator=g)

    # 1. plain training forward, two consecutive steps (EMA state carries over)
    hard_case(hard, "hard_train_n512_d64_k1024", randn(512, 64, 1, 1), hard.l2norm(randn(1024, 64)), steps=2)
    # 2. eval forward (usage EMA only, codebook untouched)
    hard_case(hard, "hard_eval_n512_d64_k1024", randn(512, 64, 1, 1), hard.l2norm(randn(1024, 64)), training=False)
    # 3. ragged: b c h w with h,w > 1 (exercises the rearrange), N=1000, K=1000 (not /256, /3), D=32
    hard_case(hard, "hard_ragged_n1000_d32_k1000", randn(10, 32, 10, 10), hard.l2norm(randn(1000, 32)), steps=2)
    # 4. D=128, training, update disabled (embedding.update=False -> only cluster_size moves)
    hard_case(hard, "hard_noupdate_n384_d128_k1024", randn(384, 128, 1, 1), hard.l2norm(randn(1024, 128)),
              update=False)
    hard_case(hard, "hard_train_n384_d128_k1024", randn(384, 128, 1, 1), hard.l2norm(randn(1024, 128)))
    hard_case(hard, "hard_train_n300_d256_k512", randn(300, 256, 1, 1), hard.l2norm(randn(512, 256)))
    # 5. ties / dead codes / degenerate rows
    w = hard.l2norm(randn(2048, 64))
    w[100] = w[17]; w[500] = w[17]; w[2047] = w[1024]           # duplicated codes -> first index wins
    z = randn(256, 64)
    z[0] = w[17] * 3.0                                           # exactly on a duplicated code
    z[1] = w[2047] * 0.5
    z[2] = 0.0                                                   # zero row: norm clamps to eps
    z[3] = w[500]
    z[4:8] = z[8:12]                                             # duplicate inputs
    hard_case(hard, "hard_ties_n256_d64_k2048", z.reshape(256, 64, 1, 1), w, steps=2,
              extra_meta={"dup_codes": [17, 100, 500], "dup_codes2": [1024, 2047]})
    # 6. un-normalised codebook (as a codebook_init_path load could produce)
    hard_case(hard, "hard_rawcodebook_n256_d64_k768", randn(256, 64, 1, 1), randn(768, 64) * 0.7)
    # 7. clustered / skewed usage (Zipf over codes, SURVEY 8d "second distribution")
    w = hard.l2norm(randn(2048, 64))
    ranks = torch.arange(1, 2049, dtype=torch.float64)
    p = (1.0 / ranks ** 1.1); p = p / p.sum()
    j = torch.multinomial(p, 2048, replacement=True, generator=g)
    z = hard.l2norm(w[j] + 0.1 * randn(2048, 64))
    hard_case(hard, "hard_zipf_n2048_d64_k2048", z.reshape(2048, 64, 1, 1), w, steps=2)
    # 8. BASELINE config 1 shape (seed 0 inputs, seed 1 codebook)
    torch.manual_seed(0)
    z = torch.randn(4096, 64)
    torch.manual_seed(1)
    w = hard.l2norm(torch.randn(8192, 64))
    hard_case(hard, "hard_c1_n4096_d64_k8192", z.reshape(4096, 64, 1, 1), w)


# ------------------------------------------------------------------------------------------
def _fix_soft(sq):
    """harness accommodation: plain buffer instead of the Parameter-in-buffer (Appendix B-3)"""
    if "codebook_used" in sq._buffers:
        del sq._buffers["codebook_used"]
    sq.register_buffer("codebook_used", torch.zeros(300000))
    return sq


def gen_soft(soft):
    g = torch.Generator().manual_seed(4321)

    def randn(*s):
        return torch.randn(*s, generator=g)

    for name, n_e, e_dim, n, kk in (("soft_specific_ne3072_d64_n300", 3072, 64, 300, 5),
                                    ("soft_specific_ne1000_d32_n257", 1000, 32, 257, 5),
                                    ("soft_specific_ne1536_d128_n200", 1536, 128, 200, 3)):
        torch.manual_seed(7)
        sq = _fix_soft(soft.VectorQuantizer(n_e=n_e, e_dim=e_dim, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True,
                                            show_usage=True, split=[e_dim, e_dim], k=kk))
        sq.train()
        rec = {"codebook": _np(sq.codebook.weight)}
        for types, proj in (("text", sq.proj_text), ("graph", sq.proj_graph)):
            x_in = randn(n, e_dim).requires_grad_(True)
            gout = randn(n, e_dim)
            sq.zero_grad()
            z_q, (vq, commit, xn, zq2), usage = sq.specific_embedding(x_in, types=types)
            total = vq + commit + (z_q * gout).sum()
            total.backward()
            # indices / weights are internal to the reference: recompute them with ITS ops
            with torch.no_grad():
                x = proj(x_in)
                region = n_e // 3
                cn = F.normalize(sq.codebook.weight[:region] if types == "text" else sq.codebook.weight[-region:], p=2, dim=-1)
                dmat = sq.get_distance(F.normalize(x, p=2, dim=-1), cn)
                vals, idx = torch.topk(dmat, k=kk, largest=False)
                wts = torch.softmax(-vals, dim=1)
            rec.update({
                f"{types}_x_in": _np(x_in), f"{types}_gout": _np(gout),
                f"{types}_proj_w": _np(proj.weight), f"{types}_proj_b": _np(proj.bias),
                f"{types}_z_q": _np(z_q), f"{types}_vq": _np(vq).reshape(1), f"{types}_commit": _np(commit).reshape(1),
                f"{types}_xn": _np(xn), f"{types}_usage": np.array([usage], np.float64),
                f"{types}_idx": _np(idx), f"{types}_vals": _np(vals), f"{types}_wts": _np(wts),
                f"{types}_grad_x_in": _np(x_in.grad), f"{types}_grad_codebook": _np(sq.codebook.weight.grad),
                f"{types}_grad_proj_w": _np(proj.weight.grad), f"{types}_grad_proj_b": _np(proj.bias.grad),
            })
        # eval-mode specific embedding (losses are 0.0 CPU tensors, soft :211-212)
        sq.eval()
        with torch.no_grad():
            x_in = randn(n, e_dim)
            z_q, (vq, commit, xn, _), usage = sq.specific_embedding(x_in, types="text")
        rec.update({"eval_x_in": _np(x_in), "eval_z_q": _np(z_q), "eval_vq": _np(vq).reshape(1),
                    "eval_usage": np.array([usage], np.float64)})
        save(name, {"kind": "soft_specific", "n_e": n_e, "e_dim": e_dim, "k": kk, "beta": 0.25,
                    "state_dict_keys": sorted(sq.state_dict().keys())}, **rec)

    # full forward (dict of 14 keys), eval and train, with the cross-attention pre-step
    n_e, e_dim, bz, T = 1536, 64, 6, 9
    for mode in ("eval", "train"):
        torch.manual_seed(11)
        sq = _fix_soft(soft.VectorQuantizer(n_e=n_e, e_dim=e_dim, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True,
                                            show_usage=True, split=[e_dim, e_dim], k=5))
        sq.train(mode == "train")
        z = randn(bz, 2 * e_dim)
        z_aug = randn(bz, 2 * e_dim)
        text_features = randn(bz, T, e_dim)
        mask = torch.ones(bz, T, dtype=torch.long)
        for i in range(bz):
            mask[i, T - i:] = 0                                   # ragged text lengths
        nodes_per = [3, 5, 2, 7, 4, 6]
        batch = torch.cat([torch.full((c,), i, dtype=torch.long) for i, c in enumerate(nodes_per)])
        graph_nodes = randn(batch.shape[0], e_dim)
        sd = {k: _np(v) for k, v in sq.state_dict().items()}
        torch.manual_seed(99)                                    # dropout stream for the train case
        with torch.no_grad():
            out = sq(z, text_features, graph_nodes, mask, batch, z_aug)
        rec = {"z": _np(z), "z_aug": _np(z_aug), "text_features": _np(text_features), "mask": _np(mask),
               "batch": _np(batch), "graph_nodes": _np(graph_nodes)}
        for k_, v in sd.items():
            rec["sd::" + k_] = v
        for k_, v in out.items():
            if isinstance(v, tuple):
                for i, t in enumerate(v):
                    rec[f"out::{k_}::{i}"] = _np(t) if torch.is_tensor(t) else np.array([t])
            elif torch.is_tensor(v):
                rec["out::" + k_] = _np(v)
            else:
                rec["out::" + k_] = np.array([v], np.float64)
        save(f"soft_forward_{mode}_ne{n_e}_d{e_dim}", {"kind": "soft_forward", "mode": mode, "n_e": n_e,
                                                         "e_dim": e_dim, "k": 5, "beta": 0.25, "dropout_seed": 99,
                                                         "out_keys": list(out.keys())}, **rec)


# ------------------------------------------------------------------------------------------
def gen_kmeans(hard):
    import contextlib
    import io
    g = torch.Generator().manual_seed(77)
    samples = hard.l2norm(torch.randn(3000, 64, generator=g))
    k = 256
    torch.manual_seed(5)
    init = samples[torch.randperm(samples.shape[0])[:k]]          # what sample_vectors() will draw (:17-18)
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        means, bins = hard.kmeans(samples, k, 10, use_cosine_sim=True)
    save("kmeans_n3000_d64_k256", {"kind": "kmeans", "iters": 10}, samples=_np(samples), init=_np(init),
         means=_np(means), bins=_np(bins))


# ------------------------------------------------------------------------------------------
def gen_lookup(hard):
    """lookup-only VectorQuantizer (norm_ema_quantizer.py:222-268): eval (counts -> cluster_size) and train,
    with and without z, h*w > 1 included."""
    g = torch.Generator().manual_seed(31)
    for name, (b, h, w, d, k), training, with_z in (("lookup_eval_b64_d64_k512", (64, 1, 1, 64, 512), False, True),
                                                    ("lookup_train_b6_hw12_d32_k300", (6, 3, 4, 32, 300), True, True),
                                                    ("lookup_eval_noz_b5_hw4_d128_k64", (5, 2, 2, 128, 64), False, False)):
        q = hard.VectorQuantizer(n_embed=k, embedding_dim=d, beta=0.25)
        q.train(training)
        q.cluster_size.fill_(3.0)                                   # must be overwritten (decay is forced to 0)
        idx = torch.randint(0, k, (b * h * w,), generator=g)
        idx[: max(1, idx.numel() // 8)] = idx[0]                    # a crowded code
        z = torch.randn(b, d, h, w, generator=g) if with_z else None
        with torch.no_grad():
            z_q, loss, out_idx = q((b, h, w), idx, z)
        rec = {"weight": _np(q.embedding.weight), "idx": _np(idx), "z_q": _np(z_q), "loss": _np(loss).reshape(1),
               "cluster_size": _np(q.cluster_size)}
        if with_z:
            rec["z"] = _np(z)
        save(name, {"kind": "lookup", "bhw": [b, h, w], "beta": 0.25, "training": training, "with_z": with_z,
                    "state_dict_keys": sorted(q.state_dict().keys())}, **rec)


def gen_cvq(hard):
    """CVectorQuantiser (norm_ema_quantizer.py:272-403), anchor='closest': two training steps and one eval step."""
    g = torch.Generator().manual_seed(61)
    n, d, k = 600, 32, 96
    w0 = hard.l2norm(torch.randn(k, d, generator=g))
    zs = [torch.randn(n, d, 1, 1, generator=g) for _ in range(3)]
    q = hard.CVectorQuantiser(n_embed=k, embedding_dim=d, beta=0.25, kmeans_init=False, anchor='closest')
    q.embedding.weight.data.copy_(w0)
    rec = {"weight0": _np(w0)}
    with torch.no_grad():
        for s, mode in enumerate(("train", "train", "eval")):
            q.train(mode == "train")
            z_q, loss, idx = q(zs[s])
            rec.update({f"z_{s}": _np(zs[s]), f"z_q_{s}": _np(z_q), f"loss_{s}": _np(loss).reshape(1), f"idx_{s}": _np(idx),
                        f"cluster_size_{s}": _np(q.cluster_size), f"weight_{s}": _np(q.embedding.weight),
                        f"embed_prob_{s}": _np(q.embed_prob)})
    save("cvq_closest_n600_d32_k96", {"kind": "cvq", "beta": 0.25, "decay": 0.99, "modes": ["train", "train", "eval"],
                                      "state_dict_keys": sorted(q.state_dict().keys())}, **rec)


def gen_usage(soft):
    """codebook_usage (vector_quantization_soft_one_new.py:219-236): a sequence of calls on one window."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        q = soft.VectorQuantizer(n_e=3000, e_dim=16, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True, show_usage=True,
                                 split=[16, 16])
    q.register_buffer("codebook_used", torch.zeros(300000))       # plain buffer (harness accommodation i)
    g = torch.Generator().manual_seed(17)
    rec, usages = {}, []
    with torch.no_grad():
        for i, (rows, hi) in enumerate(((100, 1000), (7000, 3000), (59999, 40), (1, 3000), (60000, 2999), (33333, 1500))):
            idx = torch.randint(0, hi, (rows, 5), generator=g)
            usages.append(q.codebook_usage(idx, types=("shared", "text-specific", "graph-specific")[i % 3]))
            rec[f"idx_{i}"] = _np(idx)
    rec["window"] = _np(q.codebook_used)
    rec["usage"] = np.asarray(usages, np.float64)
    save("usage_window_ne3000", {"kind": "usage", "n_e": 3000, "calls": len(usages)}, **rec)


# ------------------------------------------------------------------------------------------
def _gloo_worker(rank, world, port, zs, w0, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import contextlib
    import io
    hard, _ = ref_loader.load()
    with contextlib.redirect_stdout(io.StringIO()):
        m = hard.NormEMAVectorQuantizer(n_embed=w0.shape[0], embedding_dim=w0.shape[1], beta=0.25)  # binds all_reduce (:155-157)
    m.embedding.weight.data.copy_(w0)
    m.train()
    with torch.no_grad():
        z_q, loss, idx = m(zs[rank])
    q.put((rank, _np(z_q), _np(loss).reshape(1), _np(idx), _np(m.cluster_size), _np(m.embedding.weight)))
    dist.barrier()
    dist.destroy_process_group()


def gen_gloo(hard):
    import torch.multiprocessing as mp
    g = torch.Generator().manual_seed(2024)
    world = 2
    zs = [torch.randn(320, 64, 1, 1, generator=g) for _ in range(world)]
    w0 = hard.l2norm(torch.randn(512, 64, generator=g))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, 29731, zs, w0, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get() for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join()
    rec = {"weight0": _np(w0)}
    for r, z_q, loss, idx, cs, w in res:
        rec.update({f"z_{r}": _np(zs[r]), f"z_q_{r}": z_q, f"loss_{r}": loss, f"idx_{r}": idx,
                    f"cluster_size_{r}": cs, f"weight_{r}": w})
    save("hard_gloo2_n320_d64_k512", {"kind": "hard_gloo", "world": world, "beta": 0.25, "decay": 0.99}, **rec)


def main():
    hard, soft = ref_loader.load()
    which = set(sys.argv[1:]) or {"hard", "soft", "kmeans", "gloo", "lookup", "usage", "cvq"}
    if "hard" in which:
        gen_hard(hard)
    if "soft" in which:
        gen_soft(soft)
    if "kmeans" in which:
        gen_kmeans(hard)
    if "gloo" in which:
        gen_gloo(hard)
    if "lookup" in which:
        gen_lookup(hard)
    if "usage" in which:
        gen_usage(soft)
    if "cvq" in which:
        gen_cvq(hard)


if __name__ == "__main__":
    main()
