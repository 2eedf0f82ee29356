"""B200-native drop-in for MedTok/vector_quantization_soft_one_new.py (soft-one top-k quantizer with
text-specific, graph-specific and shared codebooks carved out of one table).

Same constructor, forward signature, output dict and state_dict keys as the reference
(`/root/reference/MedTok/vector_quantization_soft_one_new.py:90-271`).  The quantisation core
(normalise -> distances -> top-k -> softmax -> weighted sum -> losses -> straight-through, and its
backward) runs in libmedtok_vq.so; the cross-attention pre-step (:17-88,133-142) is encoder-side
PyTorch, outside the quantizer hot path, and is kept op-for-op so checkpoints load unchanged.

Deliberate differences from the reference, all additive:
  * the dict also carries `{shared_text,shared_graph,text,graph}_tokens` / `..._tokens_weights`
    ([N,k] int64 / fp32) that tokenizer.py:235-238 reads but the reference never returns (B-6);
  * `codebook_used` is a plain buffer (same state_dict key) instead of a Parameter-in-buffer (B-3);
  * a batch with N*k > 300000 keeps the last 300000 entries instead of raising (B-4);
  * eval-mode losses are 0.0 tensors on the module's device rather than on the CPU (B-9).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .norm_ema_quantizer import EmbeddingEMA


class CrossAttentionLayer(nn.Module):
    """reference :17-51"""

    def __init__(self, embed_dim, num_heads, dropout=0.1):
        super().__init__()
        self.multihead_attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout)
        self.layer_norm = nn.LayerNorm(embed_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, query, key, value, attn_mask=None):
        attn_output, _ = self.multihead_attn(query, key, value, attn_mask=attn_mask)
        return self.layer_norm(query + self.dropout(attn_output))


class CrossAttention(nn.Module):
    """reference :53-88 (note: every layer of the second direction attends to the ORIGINAL vector1)"""

    def __init__(self, embed_dim, num_heads, dropout=0.1, layers=2):
        super().__init__()
        self.model = nn.ModuleList([CrossAttentionLayer(embed_dim, num_heads, dropout) for _ in range(layers)])

    def forward(self, vector1, vector2, attn_mask=None):
        out1, out2 = vector1, vector2
        for layer in self.model:
            out1 = layer(out1, vector2, vector2, attn_mask)
        for layer in self.model:
            out2 = layer(out2, vector1, vector1, attn_mask)
        return out1, out2


class _SoftQuantizeFn(torch.autograd.Function):
    """x [N,D] (projected, un-normalised) and the raw codebook table -> the soft quantisation of
    reference :195-214 against rows [lo, hi) of the table."""

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, x, table, lo, hi, topk, beta, training, workspace):
        L = _lib.lib()
        x32 = x.detach().float().contiguous()
        codes = table.detach()[lo:hi]
        if codes.dtype != torch.float32 or not codes.is_contiguous():
            codes = codes.float().contiguous()
        n, d = x32.shape
        k = codes.shape[0]
        dev = x32.device
        ws = workspace.get(n, d, k, topk, dev)
        idx = torch.empty(n, topk, dtype=torch.int64, device=dev)
        vals = torch.empty(n, topk, dtype=torch.float32, device=dev)
        wts = torch.empty(n, topk, dtype=torch.float32, device=dev)
        zq = torch.empty_like(x32)
        out = torch.empty_like(x32)
        sse = torch.zeros(1, dtype=torch.float64, device=dev)
        _lib.check(L.mtvq_soft_forward(_lib.ptr(x32), n, d, _lib.ptr(codes), k, topk, _lib.ptr(idx), _lib.ptr(vals),
                                       _lib.ptr(wts), _lib.ptr(zq), _lib.ptr(out), _lib.ptr(sse), _lib.ptr(ws),
                                       ws.numel(), _lib.stream_ptr(dev)), "mtvq_soft_forward")
        if training:
            vq = (sse / float(max(n * d, 1))).float().reshape(())        # :208  mean((zq - x.detach())^2)
            commit = vq * beta                                           # :209  beta*mean((zq.detach() - x)^2)
        else:
            vq = torch.zeros((), dtype=torch.float32, device=dev)        # :211-212
            commit = torch.zeros((), dtype=torch.float32, device=dev)
        ctx.save_for_backward(x32, codes, idx, wts, zq)
        ctx.meta = (lo, hi, topk, beta, training, tuple(table.shape), table.dtype, x.dtype)
        ctx.mark_non_differentiable(idx, wts)
        return out.to(x.dtype), vq, commit, idx, wts

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, g_out, g_vq, g_commit, _gi, _gw):
        """One fused kernel (mtvq_soft_backward_fused): straight-through + commitment term + the path through the
        softmax weights and F.normalize for x, and the gathered rows' gradient for the raw table (atomic adds)."""
        x32, codes, idx, wts, zq = ctx.saved_tensors
        lo, hi, topk, beta, training, tshape, tdtype, xdtype = ctx.meta
        n, d = x32.shape
        k = codes.shape[0]
        dev = x32.device
        grad_x = torch.empty_like(x32)
        grad_table = torch.zeros(tshape, dtype=torch.float32, device=dev)
        gslice = grad_table[lo:hi]

        def dev_scalar(g):
            return None if g is None else g.detach().float().reshape(1).contiguous()

        go = None if g_out is None else g_out.detach().float().contiguous()
        gv, gc = dev_scalar(g_vq), dev_scalar(g_commit)
        scratch = torch.empty(k * d + ((k + 127) // 128) * 128, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().mtvq_soft_backward_fused(_lib.ptr(x32), n, d, _lib.ptr(codes), k, topk, _lib.ptr(idx),
                                                       _lib.ptr(wts), _lib.ptr(go), _lib.ptr(gv), _lib.ptr(gc), float(beta),
                                                       int(bool(training)), _lib.ptr(grad_x), _lib.ptr(gslice),
                                                       _lib.ptr(scratch), scratch.numel() * 4, _lib.stream_ptr(dev)),
                   "mtvq_soft_backward_fused")
        return grad_x.to(xdtype), grad_table.to(tdtype), None, None, None, None, None, None


class VectorQuantizer(nn.Module):
    """Drop-in for reference :90-271."""

    def __init__(self, n_e, e_dim, beta, entropy_loss_ratio, l2_norm, show_usage, split, kmeans=False, num_head=4, k=5):
        super().__init__()
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        self.entropy_loss_ratio = entropy_loss_ratio
        self.l2_norm = l2_norm
        self.show_usage = show_usage
        self.split = split
        self.k = k
        self.kmeans_init = kmeans
        self.initted = False
        if not l2_norm:
            raise NotImplementedError("l2_norm=False is a NameError in the reference too (:147-155)")
        _lib.check_embedding_dim(e_dim, "VectorQuantizer")
        if not 1 <= k <= 8:
            raise _lib.MedTokVQError(f"VectorQuantizer: top-k must be in [1, 8], got k={k}")

        self.cross_attn = CrossAttention(e_dim, num_head, dropout=0.1, layers=2)
        self.proj_text = nn.Linear(self.split[0], e_dim)
        self.proj_graph = nn.Linear(self.split[1], e_dim)
        if self.kmeans_init:
            print("using kmeans init")
            self.codebook = EmbeddingEMA(self.n_e, self.split[0])
        else:
            print("no kmeans init")
            self.codebook = nn.Embedding(self.n_e, self.e_dim)
        if self.show_usage:
            self.register_buffer("codebook_used", torch.zeros(300000))
        self._ws = _lib.Workspace()
        self._usage_next = None
        self._usage_scratch = None
        self._usage_last = 0.0

    # -- quantisation core ----------------------------------------------------------------------
    def _quantize(self, x, lo, hi):
        _lib.require_cuda(x, "input")
        return _SoftQuantizeFn.apply(x, self.codebook.weight, lo, hi, self.k, float(self.beta), self.training,
                                     self._ws)

    def get_distance(self, x, y):
        """reference :120-125 (kept for API compatibility; the forward never materialises it)"""
        return torch.sum(x ** 2, dim=1, keepdim=True) + torch.sum(y ** 2, dim=1) - 2 * x @ y.t()

    def _pool_shared(self, z_text, z_graph, text_mask, batch):
        """per-sample cross-attention pooling, reference :130-145 (encoder-side PyTorch)"""
        lengths = text_mask.sum(dim=-1).tolist()
        flat_t, flat_g = [], []
        for i in range(z_text.size(0)):
            zt = z_text[i, :int(lengths[i]), :]
            zg = z_graph[(batch == i).flatten()]
            at, ag = self.cross_attn(zt, zg)
            flat_t.append(at[0, :])
            flat_g.append(torch.mean(ag, dim=0))
        return torch.stack(flat_t, dim=0), torch.stack(flat_g, dim=0)

    def get_shared_info(self, z_text, z_graph, text_mask, batch):
        """reference :127-185: both modalities against the WHOLE table"""
        zt, zg = self._pool_shared(z_text, z_graph, text_mask, batch)
        out_t, vq_t, cm_t, idx_t, w_t = self._quantize(zt, 0, self.n_e)
        out_g, vq_g, cm_g, idx_g, w_g = self._quantize(zg, 0, self.n_e)
        usage = self.codebook_usage(torch.concat([idx_t, idx_g], dim=-1), types='shared')
        self._last_tokens = {'shared_text_tokens': idx_t, 'shared_text_tokens_weights': w_t,
                             'shared_graph_tokens': idx_g, 'shared_graph_tokens_weights': w_g}
        info = (vq_t + vq_g, cm_t + cm_g, F.normalize(zt, p=2, dim=-1), F.normalize(zg, p=2, dim=-1), out_t, out_g)
        return torch.concat([out_t, out_g], dim=-1), info, usage

    def specific_embedding(self, original_embedding, types='text'):
        """reference :187-217"""
        if types == 'text':
            original_embedding = self.proj_text(original_embedding)
        if types == 'graph':
            original_embedding = self.proj_graph(original_embedding)
        region = self.codebook.weight.shape[0] // 3
        lo, hi = (0, region) if types == 'text' else (self.n_e - region, self.n_e)
        z_q, vq, commit, idx, wts = self._quantize(original_embedding, lo, hi)
        usage = self.codebook_usage(idx, types=types + '-specific')
        self._last_specific = (idx, wts)
        return z_q, (vq, commit, F.normalize(original_embedding, p=2, dim=-1), z_q), usage

    @_lib.on_tensor_device
    def codebook_usage(self, min_encoding_indices, types='shared'):
        """reference :219-236: sliding float32 window of the last 300000 code ids, len(unique)/n_e (quirks kept: the
        zero padding is counted and slice-relative indices are divided by the full table size).  The shift and
        the distinct count are one pass of mtvq_codebook_usage (bitmap + popcount instead of clone + sort +
        unique); like the reference the result is a Python float, i.e. one device->host read of 4 bytes."""
        if not self.show_usage:
            return 0.0
        _lib.require_cuda(self.codebook_used, "codebook_used")
        with torch.no_grad():
            flat = min_encoding_indices.detach().reshape(-1)
            if flat.dtype != torch.int64 or not flat.is_contiguous():
                flat = flat.to(torch.int64).contiguous()
            if flat.numel() == 0:
                return self._usage_last
            window = self.codebook_used
            dev = window.device
            cap = window.shape[0]
            if self._usage_next is None or self._usage_next.shape != window.shape or self._usage_next.device != dev:
                self._usage_next = torch.empty_like(window)
            words = (self.n_e + 31) // 32 + 1
            if self._usage_scratch is None or self._usage_scratch.numel() < words + 1 or self._usage_scratch.device != dev:
                self._usage_scratch = torch.empty(words + 1, dtype=torch.int32, device=dev)
            distinct = self._usage_scratch[words:]
            _lib.check(_lib.lib().mtvq_codebook_usage(_lib.ptr(window), _lib.ptr(flat), flat.numel(), cap, self.n_e,
                                                      _lib.ptr(self._usage_next), _lib.ptr(self._usage_scratch),
                                                      4 * words, _lib.ptr(distinct), _lib.stream_ptr(dev)),
                       "mtvq_codebook_usage")
            # the registered buffer keeps its identity (DDP buffer broadcasts, hooks and anything else that cached the
            # tensor object keep seeing the live window): the shifted window is copied back in place (1.2 MB)
            window.copy_(self._usage_next)
            self._usage_last = int(distinct.item()) / self.n_e
            return self._usage_last

    def forward(self, z, text_features, graph_node_features, text_attention_mask, batch, z_aug=None):
        """reference :238-271"""
        shared_embedding, shared_embed_loss, shared_usage = self.get_shared_info(text_features, graph_node_features,
                                                                                text_attention_mask, batch)
        shared_text, shared_graph = torch.split(shared_embedding, split_size_or_sections=self.split, dim=-1)
        z_text, z_graph = torch.split(z, split_size_or_sections=self.split, dim=-1)
        spec_text, text_loss, text_usage = self.specific_embedding(z_text, types='text')
        text_tok, text_w = self._last_specific
        spec_graph, graph_loss, graph_usage = self.specific_embedding(z_graph, types='graph')
        graph_tok, graph_w = self._last_specific
        if z_aug is not None:
            za_text, za_graph = torch.split(z_aug, split_size_or_sections=self.split, dim=-1)
            spec_text_aug, _, _ = self.specific_embedding(za_text, types='text')
            spec_graph_aug, _, _ = self.specific_embedding(za_graph, types='graph')
        else:
            spec_text_aug = spec_graph_aug = None
        out = {
            'graph_feature': z_graph,
            'text_feature': z_text,
            'shared_text_embedding': shared_text,
            'shared_graph_embedding': shared_graph,
            'shared_embed_loss': shared_embed_loss,
            'shared_codebook_usage': shared_usage,
            'specific_embedding_text': spec_text,
            'text_specific_loss': text_loss,
            'text_specific_usage': text_usage,
            'specific_embedding_graph': spec_graph,
            'graph_specific_loss': graph_loss,
            'graph_specific_usage': graph_usage,
            'specific_embedding_text_aug': spec_text_aug,
            'specific_embedding_graph_aug': spec_graph_aug,
            # keys tokenizer.py:235-238 expects in eval mode (top-k code ids and their softmax weights)
            'text_tokens': text_tok, 'text_tokens_weights': text_w,
            'graph_tokens': graph_tok, 'graph_tokens_weights': graph_w,
        }
        out.update(self._last_tokens)
        return out
