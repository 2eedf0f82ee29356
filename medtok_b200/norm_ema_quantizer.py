"""B200-native drop-in for MedTok/norm_ema_quantizer.py (hard nearest-code quantizer, norm-EMA codebook).

Same public names, constructor signatures, forward contract and state_dict keys as the reference
(`/root/reference/MedTok/norm_ema_quantizer.py`), so `tokenizer.py` / `train_MedTok.py` /
`inference.py` can import this module instead.  All tensor work of the forward + EMA update runs in
libmedtok_vq.so (hand-written sm_100a CUDA, include/medtok_vq.h); this file is host-side glue:
layout handling, buffer ownership, the torch.distributed exchange and autograd wiring.
There is no CPU / PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import os
import sys

import torch
import torch.distributed as distributed
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .distributed import PeerExchange, all_reduce_stats


def _world_size() -> int:
    return distributed.get_world_size() if (distributed.is_available() and distributed.is_initialized()) else 1


def l2norm(t):
    """reference :8-9"""
    return F.normalize(t, p=2, dim=-1)


def ema_inplace(moving_avg, new, decay):
    """reference :11-12"""
    moving_avg.data.mul_(decay).add_(new, alpha=(1 - decay))


def norm_ema_inplace(moving_avg, new, decay):
    """reference :136-138"""
    moving_avg.data.mul_(decay).add_(new, alpha=(1 - decay))
    moving_avg.data.copy_(l2norm(moving_avg.data))


def _rows_f32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()          # cast policy: fp16/bf16 inputs (autocast) are widened, all maths is fp32
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# thin functional layer over the C ABI (used by the modules, k-means and the tests)
# ------------------------------------------------------------------------------------------------
@_lib.on_tensor_device
def hard_assign(z_rows: torch.Tensor, weight: torch.Tensor, norm_mode: int = _lib.NORM_CONTIGUOUS,
                workspace: _lib.Workspace | None = None, exact: bool = False):
    """idx[n] = first-index argmin_k ||zf_n - w_k||^2-style distance of reference :175-179.
    Returns (idx int64 [n], denom fp32 [n])."""
    _lib.require_cuda(z_rows, "z")
    z_rows, weight = _rows_f32(z_rows), _rows_f32(weight)
    n, d = z_rows.shape
    k = weight.shape[0]
    ws = (workspace or _lib.Workspace()).get(n, d, k, 0, z_rows.device)
    idx = torch.empty(n, dtype=torch.int64, device=z_rows.device)
    denom = torch.empty(n, dtype=torch.float32, device=z_rows.device)
    fn = _lib.lib().mtvq_hard_assign_exact if exact else _lib.lib().mtvq_hard_assign
    _lib.check(fn(_lib.ptr(z_rows), n, d, norm_mode, _lib.ptr(weight), k, _lib.ptr(idx), _lib.ptr(denom),
                  _lib.ptr(ws), ws.numel(), _lib.stream_ptr(z_rows.device)), "mtvq_hard_assign")
    return idx, denom


@_lib.on_tensor_device
def code_statistics(z_rows, denom, weight, idx, workspace: _lib.Workspace | None = None):
    """Packed fp64 statistics [bins(k) | embed_sum(k*d) | sse | rows] and z_q for given indices
    (reference :181-202 without the one-hot)."""
    z_rows, weight = _rows_f32(z_rows), _rows_f32(weight)
    n, d = z_rows.shape
    k = weight.shape[0]
    dev = z_rows.device
    ws = (workspace or _lib.Workspace()).get(n, d, k, 0, dev)
    stats = torch.zeros(_lib.stats_len(d, k), dtype=torch.float64, device=dev)
    z_q = torch.empty_like(z_rows)
    L = _lib.lib()
    _lib.check(L.mtvq_gather_ste(_lib.ptr(z_rows), _lib.ptr(denom), _lib.ptr(weight), _lib.ptr(idx), n, d, k,
                                 _lib.ptr(z_q), _lib.ptr(stats), _lib.stream_ptr(dev)), "mtvq_gather_ste")
    _lib.check(L.mtvq_code_sums(_lib.ptr(z_rows), _lib.ptr(denom), _lib.ptr(idx), n, d, k, _lib.ptr(stats),
                                _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "mtvq_code_sums")
    return z_q, stats


def sample_vectors(samples, num):
    """reference :14-22"""
    num_samples, device = samples.shape[0], samples.device
    if num_samples >= num:
        indices = torch.randperm(num_samples, device=device)[:num]
    else:
        indices = torch.randint(0, num_samples, (num,), device=device)
    return samples[indices]


@_lib.on_tensor_device
def kmeans(samples, num_clusters, num_iters=10, use_cosine_sim=False, means=None, verbose=True):
    """reference :24-57, cosine branch, on the CUDA kernels: `buckets = (samples @ means.t()).max(-1).indices`
    is the tensor-core assignment in inner-product mode (SCORE_DOT: exact fp32 re-scoring of near-ties by z.w alone,
    first index on ties) on the samples as given (NORM_NONE: no second normalisation); bincount + scatter_add are the
    packed statistics (fp64 sums of the samples); the centre update is mtvq_kmeans_update.  Returns (means, bins)."""
    if not use_cosine_sim:
        raise NotImplementedError("only the cosine-similarity branch is used by MedTok (:90)")
    _lib.require_cuda(samples, "samples")
    samples = _rows_f32(samples)
    if means is None:
        means = sample_vectors(samples, num_clusters)
    means = means.clone().contiguous()
    ws = _lib.Workspace()
    d = samples.shape[1]
    L = _lib.lib()
    bins = None
    for i in range(num_iters):
        idx, ones = hard_assign(samples, means, norm_mode=_lib.NORM_NONE | _lib.SCORE_DOT, workspace=ws)
        stats = torch.zeros(_lib.stats_len(d, num_clusters), dtype=torch.float64, device=samples.device)
        wsb = ws.get(samples.shape[0], d, num_clusters, 0, samples.device)
        _lib.check(L.mtvq_code_sums(_lib.ptr(samples), _lib.ptr(ones), _lib.ptr(idx), samples.shape[0], d, num_clusters,
                                    _lib.ptr(stats), _lib.ptr(wsb), wsb.numel(), _lib.stream_ptr(samples.device)),
                   "mtvq_code_sums")
        _lib.check(L.mtvq_kmeans_update(_lib.ptr(means), _lib.ptr(stats), num_clusters, d,
                                        _lib.stream_ptr(samples.device)), "mtvq_kmeans_update")
        bins = stats[:num_clusters]
        if verbose:
            print(f"Iteration {i + 1}/{num_iters} - Zero clusters: {int((bins == 0).sum().item())}")
    return means, bins.to(torch.int64)


# ------------------------------------------------------------------------------------------------
class EmbeddingEMA(nn.Module):
    """Codebook container, reference :60-134: `weight`, `cluster_size`, `embed_avg` parameters
    (requires_grad=False), `initted` buffer, `update` flag; identical state_dict keys."""

    def __init__(self, num_tokens, codebook_dim, decay=0.99, eps=1e-5, kmeans_init=True, codebook_init_path=''):
        super().__init__()
        self.num_tokens = num_tokens
        self.codebook_dim = codebook_dim
        self.decay = decay
        self.eps = eps
        if codebook_init_path == '':
            if not kmeans_init:
                weight = l2norm(torch.randn(num_tokens, codebook_dim))
            else:
                weight = torch.zeros(num_tokens, codebook_dim)
            self.register_buffer('initted', torch.Tensor([not kmeans_init]))
        else:
            print(f"load init codebook weight from {codebook_init_path}")
            weight = torch.load(codebook_init_path, map_location='cpu').clone()
            self.register_buffer('initted', torch.Tensor([True]))
        self.weight = nn.Parameter(weight, requires_grad=False)
        self.cluster_size = nn.Parameter(torch.zeros(num_tokens), requires_grad=False)
        self.embed_avg = nn.Parameter(weight.clone(), requires_grad=False)
        self.update = True
        self._initted_host = None       # host mirror of `initted`: avoids the reference's per-forward sync (:87)

    def _is_initted(self) -> bool:
        if self._initted_host is None:
            self._initted_host = bool(self.initted.item())
        return self._initted_host

    def _load_from_state_dict(self, *args, **kwargs):
        self._initted_host = None
        return super()._load_from_state_dict(*args, **kwargs)

    @torch.jit.ignore
    def init_embed_(self, data):
        """reference :85-93"""
        if self._is_initted():
            return
        print("Performing Kemans init for codebook")
        embed, cluster_size = kmeans(data, self.num_tokens, 10, use_cosine_sim=True)
        self.weight.data.copy_(embed)
        self.cluster_size.data.copy_(cluster_size)
        self.initted.data.copy_(torch.Tensor([True]))
        self._initted_host = True

    @torch.jit.ignore
    def init_embed_split(self, data, split):
        """reference :95-106 (unused by MedTok; kept for API parity): one k-means per half of the feature dimension"""
        if self._is_initted():
            return
        print("Performing Kemans init for codebook")
        embed1, cluster_size1 = kmeans(data[:, :split[0]].contiguous(), self.num_tokens, 10, use_cosine_sim=True)
        embed2, cluster_size2 = kmeans(data[:, split[0]:].contiguous(), self.num_tokens, 10, use_cosine_sim=True)
        self.weight.data.copy_(torch.cat([embed1, embed2], dim=-1))
        self.cluster_size.data.copy_((cluster_size1 + cluster_size2) / 2.)
        self.initted.data.copy_(torch.Tensor([True]))
        self._initted_host = True

    @torch.jit.ignore
    def init_embed_with_ind(self, data, inds):
        """reference :108-116 (unused by MedTok; `inds` is ignored there too)"""
        self.init_embed_(data)

    def forward(self, embed_id):
        return F.embedding(embed_id, self.weight)

    def cluster_size_ema_update(self, new_cluster_size):
        self.cluster_size.data.mul_(self.decay).add_(new_cluster_size, alpha=1 - self.decay)

    def embed_avg_ema_update(self, new_embed_avg):
        self.embed_avg.data.mul_(self.decay).add_(new_embed_avg, alpha=1 - self.decay)

    def weight_update(self, num_tokens):
        n = self.cluster_size.sum()
        smoothed_cluster_size = ((self.cluster_size + self.eps) / (n + num_tokens * self.eps) * n)
        self.weight.data.copy_(self.embed_avg / smoothed_cluster_size.unsqueeze(1))


# ------------------------------------------------------------------------------------------------
class _HardQuantizeFn(torch.autograd.Function):
    """Autograd shell around the forward: gradients reach only z (straight-through + commitment
    term, through l2norm), never the codebook (reference :212-214)."""

    @staticmethod
    def forward(ctx, z_rows_in, module, norm_mode):
        z_q, loss, idx = module._run(z_rows_in, norm_mode)
        ctx.save_for_backward(z_rows_in, z_q)
        ctx.beta = module.beta
        ctx.mark_non_differentiable(idx)
        return z_q, loss, idx

    @staticmethod
    def backward(ctx, g_zq, g_loss, _g_idx):
        """one fused row kernel (mtvq_hard_backward); a channel-strided forward input (h*w > 1) arrives here as the
        [N, D] rows autograd made contiguous, so the same kernel serves both layouts"""
        z_rows, z_q = ctx.saved_tensors
        z32 = _rows_f32(z_rows)
        n, d = z32.shape
        grad = torch.empty_like(z32)
        go = None if g_zq is None else g_zq.detach().float().contiguous()
        gl = None if g_loss is None else g_loss.detach().float().reshape(1).contiguous()
        with torch.cuda.device(z32.device):
            _lib.check(_lib.lib().mtvq_hard_backward(_lib.ptr(z32), _lib.ptr(z_q), _lib.ptr(go), _lib.ptr(gl), n, d,
                                                     float(ctx.beta), _lib.ptr(grad), _lib.stream_ptr(z32.device)),
                       "mtvq_hard_backward")
        return grad.to(z_rows.dtype), None, None


class NormEMAVectorQuantizer(nn.Module):
    """Drop-in for reference :140-218."""

    def __init__(self, n_embed, embedding_dim, beta, decay=0.99, eps=1e-5,
                 statistic_code_usage=True, kmeans_init=False, codebook_init_path=''):
        super().__init__()
        _lib.check_embedding_dim(embedding_dim, type(self).__name__)
        self.codebook_dim = embedding_dim
        self.num_tokens = n_embed
        self.beta = beta
        self.decay = decay
        self.embedding = EmbeddingEMA(self.num_tokens, self.codebook_dim, decay, eps, kmeans_init, codebook_init_path)
        self.statistic_code_usage = statistic_code_usage
        if statistic_code_usage:
            self.register_buffer('cluster_size', torch.zeros(n_embed))
        # like the reference (:155-159) the exchange is bound at construction time
        if distributed.is_available() and distributed.is_initialized():
            print("ddp is enable, so use ddp_reduce to sync the statistic_code_usage for each gpu!", file=sys.stderr)
            self.all_reduce_fn = distributed.all_reduce
        else:
            self.all_reduce_fn = nn.Identity()
        self.global_loss = False          # True: loss = beta * sum_ranks(sse) / (sum_ranks(rows) * D)
        self._ws = _lib.Workspace()
        self.last_stats = None            # packed fp64 statistics of the last forward (diagnostics)
        # exchange fused into the finalize kernel over NVLink peer memory (set up at the first multi-GPU forward);
        # False = unavailable / disabled -> one packed all-reduce + finalize
        self.fused_exchange = os.environ.get("MEDTOK_FUSED_EXCHANGE", "1") != "0"
        self._peers = None

    def reset_cluster_size(self, device):
        if self.statistic_code_usage:
            self.register_buffer('cluster_size', torch.zeros(self.num_tokens))
            self.cluster_size = self.cluster_size.to(device)

    def _peer_exchange(self, dev, d, k):
        """PeerExchange for the fused finalize, or None (-> all-reduce route).  Only for the stock binding
        (`all_reduce_fn is torch.distributed.all_reduce` on an NCCL group); set up once.  Every step that can fail on
        one rank only (allocation, eligibility, the mapping check) is followed by a MIN vote, so all ranks issue the
        same sequence of collectives and take the same route; the barrier comes last, after the final vote."""
        if self._peers is not None or not self.fused_exchange:
            return self._peers or None
        self.fused_exchange = False
        if self.all_reduce_fn is not distributed.all_reduce or distributed.get_backend() != "nccl":
            return None                  # (a property of how the job was launched: the same on every rank)

        def all_agree(ok_local: bool) -> bool:
            ok = torch.full((1,), 1.0 if ok_local else 0.0, device=dev)
            distributed.all_reduce(ok, op=distributed.ReduceOp.MIN)
            return float(ok) != 0.0

        def note(stage, e):
            print(f"medtok_b200: fused NVLink exchange unavailable ({stage}: {type(e).__name__}: {e}); using all_reduce",
                  file=sys.stderr)

        peers = None
        try:
            world = distributed.get_world_size()
            # one node only: the exchange reads peer memory over NVLink (LOCAL_WORLD_SIZE is set by torchrun)
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
            if local_world != world or world > torch.cuda.device_count():
                raise RuntimeError(f"multi-node world ({world} ranks, {local_world} on this node)")
            peers = PeerExchange(_lib.stats_len(d, k), dev, result_numel=k * d + k)
        except Exception as e:
            note("allocation", e)
            peers = None
        if not all_agree(peers is not None):
            return None
        try:
            peers.connect()                                   # collective; every rank is here
            peers.arm()
            ok = True
        except Exception as e:
            note("rendezvous", e)
            ok = False
        if not all_agree(ok):
            return None
        distributed.barrier()                                 # nobody signals before every flag array is zero
        self._peers = peers
        self.fused_exchange = True
        return peers

    # -- one forward on rows [N, D] -------------------------------------------------------------
    @_lib.on_tensor_device
    def _run(self, z_rows_in: torch.Tensor, norm_mode: int):
        L = _lib.lib()
        z_rows = _rows_f32(z_rows_in)
        n, d = z_rows.shape
        k = self.num_tokens
        dev = z_rows.device
        weight = self.embedding.weight.data
        _lib.require_cuda(weight, "embedding.weight")
        if weight.dtype != torch.float32 or not weight.is_contiguous():
            raise _lib.MedTokVQError("embedding.weight must be a contiguous fp32 tensor")
        has_cs = self.statistic_code_usage
        distributed_on = not isinstance(self.all_reduce_fn, nn.Identity)
        flags = (_lib.F_TRAINING if self.training else 0) | (_lib.F_UPDATE if self.embedding.update else 0)
        if distributed_on or not has_cs:
            flags |= _lib.F_SKIP_FINALIZE
        ws = self._ws.get(n, d, k, 0, dev)
        z_q = torch.empty_like(z_rows)
        idx = torch.empty(n, dtype=torch.int64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        upd_w = self.training and self.embedding.update
        upd_cs = has_cs and ((not self.training) or self.embedding.update)
        peers = self._peer_exchange(dev, d, k) if (distributed_on and (upd_w or upd_cs)) else None
        stats = peers.next_buffer() if peers is not None else \
            torch.empty(_lib.stats_len(d, k), dtype=torch.float64, device=dev)
        cs = self.cluster_size if has_cs else None
        # The two-pass route of the hard path has been validated on up to MEDTOK_TWO_PASS_MAX_WORLD ranks (default 4; see
        # DESIGN.md section 7): larger groups take the single-pass route unless a route has been set explicitly.
        pin_single = (distributed_on and L.mtvq_debug_get_hard_route() < 0 and not os.environ.get("MTVQ_HARD_ROUTE")
                      and _world_size() > int(os.environ.get("MEDTOK_TWO_PASS_MAX_WORLD", "4")))
        if pin_single:
            L.mtvq_debug_set_hard_route(1)
        try:
            _lib.check(L.mtvq_hard_forward(_lib.ptr(z_rows), n, d, norm_mode, _lib.ptr(weight), k, _lib.ptr(cs),
                                           float(self.beta), float(self.decay), flags, _lib.ptr(z_q), _lib.ptr(idx),
                                           _lib.ptr(loss), _lib.ptr(stats), _lib.ptr(ws), ws.numel(),
                                           _lib.stream_ptr(dev)), "mtvq_hard_forward")
        finally:
            if pin_single:
                L.mtvq_debug_set_hard_route(-1)
        self.last_route = 1 if pin_single else int(L.mtvq_hard_route(n, d, k))
        if peers is not None:
            # ONE kernel: signal / wait over NVLink, rank-ordered fp64 sum of the peers' statistics, EMA update
            _lib.check(L.mtvq_ema_finalize_peers(_lib.ptr(weight), _lib.ptr(cs), ctypes.c_void_p(peers.peer_stats_ptr()),
                                                 ctypes.c_void_p(peers.peer_flags_ptr()),
                                                 ctypes.c_void_p(peers.peer_results_ptr()), peers.world, peers.rank,
                                                 peers.epoch, k, d, float(self.beta), float(self.decay), int(upd_cs),
                                                 int(upd_w), _lib.ptr(loss) if self.global_loss else None,
                                                 _lib.stream_ptr(dev)), "mtvq_ema_finalize_peers")
        elif flags & _lib.F_SKIP_FINALIZE:
            if distributed_on and (upd_w or upd_cs):
                all_reduce_stats(stats, self.all_reduce_fn)   # ONE packed all-reduce (reference: :188 / :195 + :203)
            if upd_w or upd_cs:
                _lib.check(L.mtvq_ema_finalize(_lib.ptr(weight), _lib.ptr(cs), _lib.ptr(stats), k, d,
                                               float(self.beta), float(self.decay), int(upd_cs), int(upd_w),
                                               _lib.ptr(loss) if self.global_loss else None,
                                               _lib.stream_ptr(dev)), "mtvq_ema_finalize")
        self.last_stats = stats
        return z_q, loss.reshape(()), idx

    def forward(self, z):
        _lib.require_cuda(z, "z")
        b, c, h, w = z.shape
        if c != self.codebook_dim:
            raise _lib.MedTokVQError(f"expected {self.codebook_dim} channels, got {c}")
        if h * w == 1:
            rows = z.reshape(b, c)                                   # 'b c h w -> b h w c' is a no-op view
            norm_mode = _lib.NORM_CONTIGUOUS
        else:
            rows = z.permute(0, 2, 3, 1).reshape(-1, c)              # reference :169-171
            norm_mode = _lib.NORM_STRIDED                             # torch normalises the strided view there
        if not self.embedding._is_initted():
            self.embedding.init_embed_(l2norm(_rows_f32(rows)))      # reference :173 (k-means on first batch)
        if torch.is_grad_enabled() and z.requires_grad:
            z_q, loss, idx = _HardQuantizeFn.apply(rows, self, norm_mode)
        else:
            z_q, loss, idx = self._run(rows, norm_mode)
        if z_q.dtype != z.dtype:
            z_q = z_q.to(z.dtype)
        if h * w == 1:
            z_q = z_q.view(b, c, 1, 1)
        else:
            z_q = z_q.view(b, h, w, c).permute(0, 3, 1, 2)           # reference :217
        return z_q, loss, idx


# ------------------------------------------------------------------------------------------------
class VectorQuantizer(nn.Module):
    """Drop-in for the lookup-only quantizer, reference :222-268: code indices in, codebook rows out
    (`forward(bhw, encoding_indices, z=None) -> (z_q[b,c,h,w], loss, encoding_indices)`); in eval mode the
    per-code counts replace `cluster_size` (the reference's EMA there runs with decay forced to 0, :230-231)."""

    def __init__(self, n_embed, embedding_dim, beta, decay=0.99, eps=1e-5,
                 statistic_code_usage=True, kmeans_init=False, codebook_init_path=''):
        super().__init__()
        self.codebook_dim = embedding_dim
        self.num_tokens = n_embed
        self.beta = beta
        decay = 0.
        self.decay = decay
        self.embedding = EmbeddingEMA(self.num_tokens, self.codebook_dim, decay, eps, kmeans_init, codebook_init_path)
        self.statistic_code_usage = statistic_code_usage
        if statistic_code_usage:
            self.register_buffer('cluster_size', torch.zeros(n_embed))
        if distributed.is_available() and distributed.is_initialized():
            print("ddp is enable, so use ddp_reduce to sync the statistic_code_usage for each gpu!", file=sys.stderr)
            self.all_reduce_fn = distributed.all_reduce
        else:
            self.all_reduce_fn = nn.Identity()
        self._ws = _lib.Workspace()

    def reset_cluster_size(self, device):
        if self.statistic_code_usage:
            self.register_buffer('cluster_size', torch.zeros(self.num_tokens))
            self.cluster_size = self.cluster_size.to(device)

    @_lib.on_tensor_device
    def _run(self, idx: torch.Tensor, z_rows: torch.Tensor | None):
        weight = self.embedding.weight.data
        _lib.require_cuda(weight, "embedding.weight")
        _lib.require_cuda(idx, "encoding_indices")
        if weight.dtype != torch.float32 or not weight.is_contiguous():
            raise _lib.MedTokVQError("embedding.weight must be a contiguous fp32 tensor")
        dev = weight.device
        k, d = weight.shape
        flat = idx.reshape(-1)
        if flat.dtype != torch.int64 or not flat.is_contiguous():
            flat = flat.to(torch.int64).contiguous()
        n = flat.shape[0]
        z_q = torch.empty(n, d, dtype=torch.float32, device=dev)
        sse = torch.zeros(1, dtype=torch.float64, device=dev) if z_rows is not None else None
        # the reference only needs cluster_size when statistic_code_usage registered it (:236-238,257-261)
        want_counts = (not self.training) and self.statistic_code_usage
        counts = torch.empty(k, dtype=torch.float32, device=dev) if want_counts else None
        ws = self._ws.get(n, d, k, 0, dev) if want_counts else None
        _lib.check(_lib.lib().mtvq_lookup(_lib.ptr(z_rows), _lib.ptr(weight), _lib.ptr(flat), n, d, k, _lib.ptr(z_q),
                                          _lib.ptr(sse), _lib.ptr(counts), _lib.ptr(ws), 0 if ws is None else ws.numel(),
                                          _lib.stream_ptr(dev)), "mtvq_lookup")
        if want_counts:
            self.all_reduce_fn(counts)                                   # reference :259
            self.cluster_size.data.copy_(counts)                         # ema_inplace with decay 0 (:260)
        if z_rows is not None:
            loss = (sse * (float(self.beta) / float(max(n * d, 1)))).float().reshape(())
        else:
            loss = torch.zeros((), dtype=torch.float32, device=dev)      # :266
        return z_q, loss

    def forward(self, bhw, encoding_indices, z=None):
        b, h, w = int(bhw[0]), int(bhw[1]), int(bhw[2])
        z_rows = None
        if z is not None:
            _lib.require_cuda(z, "z")
            z_rows = _rows_f32(z.permute(0, 2, 3, 1)).reshape(-1, z.shape[1])      # :263 'b c h w -> b h w c'
        # the codebook parameter has requires_grad=False and the loss detaches z (:264): no output carries a gradient
        z_q, loss = self._run(encoding_indices, z_rows)
        z_q = z_q.view(b, h, w, -1).permute(0, 3, 1, 2)                             # :267
        return z_q, loss, encoding_indices


# ------------------------------------------------------------------------------------------------
class FeaturePool():
    """Feature history for the 'random' anchor (reference :405-460; host-side bookkeeping, behaviour kept):
    until the pool has filled once, batches are appended (a batch larger than the pool replaces it by a random sample
    of the batch and marks it full); afterwards a larger batch again replaces the pool by a random sample, a smaller one
    overwrites randomly chosen slots.  `query` returns the pool."""

    def __init__(self, pool_size, dim=64):
        self.pool_size = pool_size
        if pool_size > 0:
            self.nums_features = 0
            self.features = (2.0 * torch.rand((pool_size, dim)) - 1.0) / pool_size

    def _resample(self, batch):
        pick = torch.randint(0, batch.size(0), (int(self.pool_size),))
        self.features = batch[pick]

    def query(self, features):
        self.features = self.features.to(features.device)
        rows, cap = features.size(0), int(self.pool_size)
        filling = self.nums_features < cap
        if rows > cap:
            self._resample(features)
            if filling:
                self.nums_features = cap
        elif filling:
            end = self.nums_features + rows
            self.features[self.nums_features:end] = features
            self.nums_features = end
        else:
            slots = torch.randperm(cap)[:rows]
            self.features[slots] = features
        return self.features


class CVectorQuantiser(NormEMAVectorQuantizer):
    """Drop-in for reference :272-403 (norm-EMA quantizer + online re-initialisation of dead codes).

    The assignment, statistics, EMA and loss are exactly the NormEMAVectorQuantizer forward (the reference's
    `d = -|z|^2 - |e|^2 + 2 z.e`, argmax, is the exact negation of :175-179, ties included), run by the same kernels.  On
    top, in training mode: `embed_prob` tracks the LOCAL usage frequencies (:375-377) and every code is pulled towards an
    anchor sample with weight exp(-(embed_prob * K * 10) / (1 - decay) - 1e-3) (:379-393).  For anchor='closest' the anchor
    is the sample nearest to the code under the PRE-update codebook, found with the same tensor-core assignment kernel
    with the roles swapped (codes as rows, normalised samples as the codebook) and gathered with mtvq_lookup; the
    K x D blend itself is plain elementwise PyTorch (not on the hot path).  anchor='random' is the feature pool of
    :405-460, anchor='probrandom' (:387-390) a K x N softmax + multinomial in PyTorch (a random draw: not bit-comparable);
    contras_loss=True is a NameError in the reference (:396-402) and raises here too."""

    def __init__(self, n_embed, embedding_dim, beta, decay=0.99, eps=1e-5,
                 statistic_code_usage=True, kmeans_init=True, codebook_init_path='',
                 distance='l2', anchor='closest', first_batch=False, contras_loss=False):
        super().__init__(n_embed, embedding_dim, beta, decay, eps, statistic_code_usage, kmeans_init, codebook_init_path)
        if distance != 'l2':
            raise NotImplementedError("only distance='l2' (on l2-normalised rows 'cos' selects the same codes)")
        if contras_loss:
            raise NotImplementedError("contras_loss=True raises NameError in the reference (:396-402)")
        if anchor not in ('closest', 'random', 'probrandom'):
            raise ValueError(f"unknown anchor {anchor!r}")
        self.num_embed = n_embed
        self.n_e = n_embed
        self.embed_dim = embedding_dim
        self.distance = distance
        self.anchor = anchor
        self.first_batch = first_batch
        self.contras_loss = contras_loss
        self.init = False
        self.pool = FeaturePool(self.num_embed, self.embed_dim)
        self.register_buffer("embed_prob", torch.zeros(self.num_embed))
        self._ws_anchor = _lib.Workspace()

    def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
        reinit = self.training and self.anchor in ('closest', 'random', 'probrandom') and not self.init
        w_old = self.embedding.weight.data.clone() if (reinit and self.anchor in ('closest', 'probrandom')) else None
        z_q, loss, idx = super().forward(z)
        if self.training:
            with torch.no_grad():
                k, d = self.num_tokens, self.codebook_dim
                n = idx.numel()
                local_bins = torch.bincount(idx, minlength=k).to(torch.float32)          # :375 mean of the LOCAL one-hot
                self.embed_prob.mul_(self.decay).add_(local_bins / float(max(n, 1)), alpha=1 - self.decay)
                if reinit:
                    rows = z.permute(0, 2, 3, 1).reshape(-1, d)
                    zf = l2norm(_rows_f32(rows))                                         # z_flattened (:313-315)
                    if self.anchor == 'closest':
                        # sample with the largest d[:, k] per code (:381-383) == nearest sample to the code
                        near, _ = hard_assign(w_old, zf, workspace=self._ws_anchor)
                        feat = torch.empty(k, d, dtype=torch.float32, device=zf.device)
                        _lib.check(_lib.lib().mtvq_lookup(None, _lib.ptr(zf), _lib.ptr(near), k, d, n, _lib.ptr(feat), None,
                                                          None, None, 0, _lib.stream_ptr(zf.device)), "mtvq_lookup")
                    elif self.anchor == 'random':
                        feat = self.pool.query(zf)
                    else:
                        # 'probrandom' (:387-390): one sample per code drawn from softmax over the samples of
                        # d = -|z|^2 - |e|^2 + 2 z.e under the pre-update codebook.  A K x N softmax by definition
                        # (training batches are a few hundred rows); plain PyTorch, in row chunks of the codebook
                        w_src = self.embedding.weight.data if w_old is None else w_old
                        pick = torch.empty(k, dtype=torch.int64, device=zf.device)
                        zz = zf.pow(2).sum(1)
                        for lo in range(0, k, 4096):
                            wc = w_src[lo:lo + 4096]
                            dkn = -(wc.pow(2).sum(1, keepdim=True) + zz[None, :] - 2.0 * wc @ zf.t())
                            pick[lo:lo + 4096] = torch.multinomial(F.softmax(dkn, dim=1), num_samples=1).view(-1)
                        feat = zf[pick]
                    dec = torch.exp(-(self.embed_prob * self.num_embed * 10) / (1 - self.decay) - 1e-3).unsqueeze(1)
                    self.embedding.weight.data = self.embedding.weight.data * (1 - dec) + feat * dec
                    if self.first_batch:
                        self.init = True
        return z_q, loss, idx
