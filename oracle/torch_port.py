"""ATen-level port of the reference's CPU path -- TEST/BENCH INFRASTRUCTURE ONLY.

The reference is pure PyTorch, so "its own CPU implementation" is a sequence of ATen ops.  This
file issues the same ops in the same order as NormEMAVectorQuantizer.forward
(norm_ema_quantizer.py:166-218) and the soft quantizer core
(vector_quantization_soft_one_new.py:120-125,195-214), with one accommodation from SURVEY.md
section 8(c)-(iii): rows are processed in chunks so that N >> 64K does not need >100 GiB for the
materialised N x K matrices; the per-code statistics are accumulated over chunks and the EMA
(:197-210) is applied once.  It is what bench.py times for `cpu_baseline` / `--impl reference`
on the GPU box (kind = "port"), and tests check it against the golden fixtures.
"""
import torch
import torch.nn.functional as F


@torch.no_grad()
def hard_forward(z, weight, cluster_size, beta, decay=0.99, training=True, update=True,
                 chunk_rows=4096):
    """z: the module's [b, c, h, w] input (or [N, D] rows, treated as [N, D, 1, 1]); weight [K, D]
    and cluster_size [K] are updated IN PLACE like the reference.
    Returns (z_q in the input's layout, loss, idx int64 [N])."""
    in_shape = z.shape
    if z.dim() == 2:
        z = z[:, :, None, None]
    k, dim = weight.shape
    z = z.permute(0, 2, 3, 1)                                               # :169 (a view, like rearrange)
    z = F.normalize(z, p=2, dim=-1)                                         # :170
    bhwc = z.shape
    zf = z.reshape(-1, dim)                                                 # :171
    n = zf.shape[0]
    idx = torch.empty(n, dtype=torch.int64)
    bins = torch.zeros(k, dtype=z.dtype)
    embed_sum = torch.zeros(dim, k, dtype=z.dtype)
    for lo in range(0, n, chunk_rows):
        zc = zf[lo:lo + chunk_rows]
        d = zc.pow(2).sum(dim=1, keepdim=True) + weight.pow(2).sum(dim=1) \
            - 2 * torch.einsum('bd,nd->bn', zc, weight)                     # :175-177
        ic = torch.argmin(d, dim=1)                                         # :179
        idx[lo:lo + chunk_rows] = ic
        enc = F.one_hot(ic, k).type(z.dtype)                                # :183
        bins += enc.sum(0)                                                  # :187 / :194
        if training and update:
            embed_sum += zc.t() @ enc                                       # :202
    z_q = F.embedding(idx, weight)                                          # :181 (pre-update rows)
    if (not training) or update:                                            # :185 / :191 are the only branches
        cluster_size.mul_(decay).add_(bins, alpha=1 - decay)                # :189 / :197
    if training and update:
        zero_mask = bins == 0                                               # :199
        bins = bins.masked_fill(zero_mask, 1.)                              # :200
        e = (embed_sum / bins.unsqueeze(0)).t()                             # :205
        e = F.normalize(e, p=2, dim=-1)                                     # :206
        e = torch.where(zero_mask[..., None], weight, e)                    # :208
        weight.mul_(decay).add_(e, alpha=1 - decay)                         # :137
        weight.copy_(F.normalize(weight, p=2, dim=-1))                      # :138
    loss = beta * F.mse_loss(z_q, zf)                                       # :212
    z_q = zf + (z_q - zf)                                                   # :214
    z_q = z_q.view(bhwc).permute(0, 3, 1, 2)                                # :217
    if len(in_shape) == 2:
        z_q = z_q.reshape(in_shape)
    return z_q, loss, idx


@torch.no_grad()
def soft_core(x, codebook_slice, k, beta, training=True, chunk_rows=4096):
    """One soft quantisation (soft :195-214 / :148-182) of projected rows x [N, D] against the raw
    rows of the active codebook slice.  Returns (z_q, vq_loss, commit_loss, idx, weights)."""
    xn = F.normalize(x, p=2, dim=-1)
    cn = F.normalize(codebook_slice, p=2, dim=-1)
    n = x.shape[0]
    idx = torch.empty(n, k, dtype=torch.int64)
    wts = torch.empty(n, k, dtype=x.dtype)
    zq = torch.empty_like(x)
    for lo in range(0, n, chunk_rows):
        xc = xn[lo:lo + chunk_rows]
        d = torch.sum(xc ** 2, dim=1, keepdim=True) + torch.sum(cn ** 2, dim=1) \
            - 2 * torch.einsum('bd,dn->bn', xc, torch.einsum('n d -> d n', cn))   # :121-123
        vals, ic = torch.topk(d, k=k, largest=False)                              # :203
        w = torch.softmax(-vals, dim=1)                                           # :204
        idx[lo:lo + chunk_rows] = ic
        wts[lo:lo + chunk_rows] = w
        zq[lo:lo + chunk_rows] = (w.unsqueeze(-1) * cn[ic]).sum(dim=1)            # :205
    if training:
        vq = torch.mean((zq - x) ** 2)                                            # :208
        commit = beta * torch.mean((zq - x) ** 2)                                 # :209
    else:
        vq = torch.tensor(0.0)
        commit = torch.tensor(0.0)
    out = x + (zq - x)                                                            # :214
    return out, vq, commit, idx, wts
