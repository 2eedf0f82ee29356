"""CPU oracle for MedTok's vector-quantizer hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  medtok_b200/ never does: the product path is CUDA-only and raises when its
extension is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the unmodified reference modules themselves, run on CPU in
the authoring container by oracle/make_golden.py and committed under tests/golden/
(tests/test_oracle_golden.py re-checks every fixture).

Two layers:
  * `C` -- ctypes binding of oracle/vq_oracle.c, the bit-defined ("canonical") arithmetic.
  * `hard_forward`, `soft_specific`, ... -- the reference's forward passes restated on top of it,
    chunk-free (the C code never materialises the N x K matrix), each citing the reference lines.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import build_oracle

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class _CLib:
    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle.build())
        L = self.lib
        L.vqo_l2norm_rows.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, _f32p, _f32p, ctypes.c_int]
        L.vqo_sumsq_rows.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, _f32p]
        L.vqo_hard_assign.argtypes = [_f32p, _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _i64p, _f32p]
        L.vqo_dist_matrix.argtypes = [_f32p, _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _f32p]
        L.vqo_soft_topk.argtypes = [_f32p, _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    _i64p, _f32p, _f32p, _f32p]
        L.vqo_stats.argtypes = [_f32p, _i64p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _f64p, _f64p]
        L.vqo_ema_cluster_size.argtypes = [_f32p, _f64p, ctypes.c_int, ctypes.c_double]
        L.vqo_ema_codebook.argtypes = [_f32p, _f64p, _f64p, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        L.vqo_gather_ste.argtypes = [_f32p, _f32p, _i64p, ctypes.c_int64, ctypes.c_int, _f32p]
        L.vqo_gather_ste.restype = ctypes.c_double
        for fn in (L.vqo_l2norm_rows, L.vqo_sumsq_rows, L.vqo_hard_assign, L.vqo_dist_matrix,
                   L.vqo_soft_topk, L.vqo_stats, L.vqo_ema_cluster_size, L.vqo_ema_codebook):
            fn.restype = None

    # -- thin typed wrappers ---------------------------------------------------------------
    def l2norm_rows(self, x, seq_mode=False):
        x = _f32(x)
        n, d = x.shape
        out = np.empty_like(x)
        den = np.empty(n, np.float32)
        self.lib.vqo_l2norm_rows(_p(x, _f32p), n, d, _p(out, _f32p), _p(den, _f32p), int(bool(seq_mode)))
        return out, den

    def sumsq_rows(self, x):
        x = _f32(x)
        n, d = x.shape
        out = np.empty(n, np.float32)
        self.lib.vqo_sumsq_rows(_p(x, _f32p), n, d, _p(out, _f32p))
        return out

    def hard_assign(self, zf, w):
        zf, w = _f32(zf), _f32(w)
        n, d = zf.shape
        k = w.shape[0]
        idx = np.empty(n, np.int64)
        dmin = np.empty(n, np.float32)
        self.lib.vqo_hard_assign(_p(zf, _f32p), _p(w, _f32p), n, d, k, _p(idx, _i64p), _p(dmin, _f32p))
        return idx, dmin

    def dist_matrix(self, zf, w):
        zf, w = _f32(zf), _f32(w)
        n, d = zf.shape
        k = w.shape[0]
        out = np.empty((n, k), np.float32)
        self.lib.vqo_dist_matrix(_p(zf, _f32p), _p(w, _f32p), n, d, k, _p(out, _f32p))
        return out

    def soft_topk(self, xn, cn, kk, want_zq=True):
        xn, cn = _f32(xn), _f32(cn)
        n, d = xn.shape
        k = cn.shape[0]
        idx = np.empty((n, kk), np.int64)
        val = np.empty((n, kk), np.float32)
        wgt = np.empty((n, kk), np.float32)
        zq = np.empty((n, d), np.float32) if want_zq else None
        self.lib.vqo_soft_topk(_p(xn, _f32p), _p(cn, _f32p), n, d, k, kk, _p(idx, _i64p), _p(val, _f32p),
                               _p(wgt, _f32p), _p(zq, _f32p))
        return idx, val, wgt, zq

    def stats(self, zf, idx, k):
        zf = _f32(zf)
        idx = np.ascontiguousarray(idx, np.int64)
        n, d = zf.shape
        bins = np.zeros(k, np.float64)
        esum = np.zeros((k, d), np.float64)
        self.lib.vqo_stats(_p(zf, _f32p), _p(idx, _i64p), n, d, k, _p(bins, _f64p), _p(esum, _f64p))
        return bins, esum

    def ema_cluster_size(self, cluster_size, bins, decay):
        cs = _f32(cluster_size).copy()
        bins = np.ascontiguousarray(bins, np.float64)
        self.lib.vqo_ema_cluster_size(_p(cs, _f32p), _p(bins, _f64p), cs.shape[0], decay)
        return cs

    def ema_codebook(self, weight, bins, embed_sum, decay):
        w = _f32(weight).copy()
        bins = np.ascontiguousarray(bins, np.float64)
        embed_sum = np.ascontiguousarray(embed_sum, np.float64)
        k, d = w.shape
        assert d <= 1024
        self.lib.vqo_ema_codebook(_p(w, _f32p), _p(bins, _f64p), _p(embed_sum, _f64p), k, d, decay)
        return w

    def gather_ste(self, zf, w, idx):
        zf, w = _f32(zf), _f32(w)
        idx = np.ascontiguousarray(idx, np.int64)
        n, d = zf.shape
        out = np.empty_like(zf)
        sse = self.lib.vqo_gather_ste(_p(zf, _f32p), _p(w, _f32p), _p(idx, _i64p), n, d, _p(out, _f32p))
        return out, float(sse)


_C = None


def C() -> _CLib:
    global _C
    if _C is None:
        _C = _CLib()
    return _C


# ------------------------------------------------------------------------------------------
# Hard path: NormEMAVectorQuantizer.forward   (norm_ema_quantizer.py:166-218)
# ------------------------------------------------------------------------------------------
@dataclass
class HardResult:
    z_q: np.ndarray          # [N, D]  straight-through value (== zf + (w[idx] - zf))
    loss: float              # beta * mean((z_q - zf)^2)
    idx: np.ndarray          # int64 [N]
    cluster_size: np.ndarray  # fp32 [K] after the EMA
    weight: np.ndarray       # fp32 [K, D] after the norm-EMA (unchanged in eval)
    zf: np.ndarray           # normalised input
    bins: np.ndarray         # float64 [K] (summed over ranks when partials are given)
    embed_sum: np.ndarray    # float64 [K, D]
    sse: float


def hard_local(z, weight, strided_norm=False):
    """Rank-local half of the forward: :170-181 (normalise, distances, argmin) + the raw
    statistics of :183-202.  z: [N, D] (the module's b c h w input flattened to rows);
    strided_norm=True when the module input had h*w > 1 (torch then reduces a strided view)."""
    c = C()
    zf, _ = c.l2norm_rows(z, seq_mode=strided_norm)           # :170
    idx, _ = c.hard_assign(zf, weight)                        # :175-179
    bins, esum = c.stats(zf, idx, weight.shape[0])            # :183,187/194,202
    zq, sse = c.gather_ste(zf, weight, idx)                   # :181,212,214
    return zf, idx, bins, esum, zq, sse


def hard_forward(z, weight, cluster_size, beta, decay=0.99, training=True, update=True,
                 partials=None, loss_rows=None, strided_norm=False) -> HardResult:
    """Whole forward.  `partials` = [(bins, embed_sum), ...] from other ranks, summed exactly
    like the reference's all_reduce (:188,195,203).  Loss is rank-local like the reference
    (:212) unless `loss_rows` (total row count) and extra SSE are folded in by the caller."""
    c = C()
    z = _f32(z)
    weight = _f32(weight)
    zf, idx, bins, esum, zq, sse = hard_local(z, weight, strided_norm)
    if partials:
        for b, e in partials:
            bins = bins + b
            esum = esum + e
    # :185-189 (eval) and :191-197 (train AND embedding.update) are the only two branches that
    # touch cluster_size: training with update=False leaves every buffer untouched.
    new_cs = _f32(cluster_size).copy()
    if (not training) or update:
        new_cs = c.ema_cluster_size(cluster_size, bins, decay)
    new_w = weight
    if training and update:
        new_w = c.ema_codebook(weight, bins, esum, decay)     # :199-210
    n, d = z.shape
    mse = np.float32(sse / float(n * d))
    loss = float(np.float32(beta) * mse)                      # :212
    return HardResult(zq, loss, idx, new_cs, new_w, zf, bins, esum, sse)


def cvq_forward(z, weight, cluster_size, embed_prob, beta, decay=0.99, training=True, init_done=False):
    """CVectorQuantiser.forward, norm_ema_quantizer.py:313-403, anchor='closest', distance='l2': the norm-EMA forward
    (d of :320-324 is the exact negation of :175-177, so argmax d == argmin with the same ties; :336-364 == :183-210),
    then in training mode embed_prob <- decay*embed_prob + (1-decay)*mean(one_hot) (:375-377) and, unless `init_done`,
    every code is blended with the sample that has the largest d in its column under the PRE-update codebook
    (:381-383,391-392).  Returns (HardResult, embed_prob, weight_after_reinit)."""
    z = _f32(z)
    weight = _f32(weight)
    r = hard_forward(z, weight, cluster_size, beta, decay, training=training, update=True)
    ep = _f32(embed_prob).copy()
    w_out = r.weight
    if training:
        n = z.shape[0]
        avg = (np.bincount(r.idx, minlength=weight.shape[0]).astype(np.float32) / np.float32(n)).astype(np.float32)
        ep = (ep * np.float32(decay) + avg * np.float32(1 - decay)).astype(np.float32)          # :377
        if not init_done:
            dm = -C().dist_matrix(r.zf, weight)                                                 # :320-324 (negated :175-177)
            near = dm.shape[0] - 1 - np.argmax(dm[::-1], axis=0)                                # last row among equal maxima
            feat = r.zf[near]
            dec = np.exp(-(ep * np.float32(weight.shape[0]) * np.float32(10)) / np.float32(1 - decay)
                         - np.float32(1e-3)).astype(np.float32)[:, None]
            w_out = (r.weight * (np.float32(1) - dec) + feat * dec).astype(np.float32)          # :392
    return r, ep, w_out


# ------------------------------------------------------------------------------------------
# Soft path: VectorQuantizer.specific_embedding / shared core
# (vector_quantization_soft_one_new.py:147-185, 187-217)
# ------------------------------------------------------------------------------------------
@dataclass
class SoftResult:
    z_q: np.ndarray        # [N, D] straight-through value x + (zq - x)
    zq_raw: np.ndarray     # [N, D] weighted sum of normalised codes
    idx: np.ndarray        # int64 [N, k] slice-relative
    vals: np.ndarray       # fp32 [N, k] ascending distances
    weights: np.ndarray    # fp32 [N, k]
    vq_loss: float
    commit_loss: float
    xn: np.ndarray


def soft_core(x, codebook_slice, k, beta, training=True) -> SoftResult:
    """x: [N, D] projected (un-normalised) inputs; codebook_slice: raw rows of the active slice
    (text = W[:R], graph = W[-R:], shared = W; soft :148,195-200)."""
    c = C()
    x = _f32(x)
    xn, _ = c.l2norm_rows(x)                                   # :150-151,196
    cn, _ = c.l2norm_rows(codebook_slice)                      # :148,198-200
    idx, vals, wgt, zq = c.soft_topk(xn, cn, k)                # :157-165,203-205
    if training:
        diff = zq - x
        mse = float(np.mean(diff.astype(np.float64) ** 2))
        vq, commit = mse, float(beta) * mse                    # :169-173,208-209
    else:
        vq = commit = 0.0                                      # :175-178,211-212
    out = x + (zq - x)                                         # :181-182,214
    return SoftResult(out, zq, idx, vals, wgt, vq, commit, xn)


def usage_window(window, new_idx, n_e):
    """soft :219-236 codebook_usage: slide a fixed float32 window, count distinct values
    (the 0.0 padding included) and divide by the FULL table size.  Returns (window, usage)."""
    flat = np.asarray(new_idx, np.float32).reshape(-1)
    cur = flat.shape[0]
    w = np.asarray(window, np.float32).copy()
    if cur >= w.shape[0]:
        # the reference raises here; defined behaviour: keep the last |w|
        w[:] = flat[-w.shape[0]:]
    else:
        w[:-cur] = w[cur:].copy()
        w[-cur:] = flat
    return w, len(np.unique(w)) / n_e


def lookup_forward(bhw, idx, weight, beta, z=None, training=False):
    """Lookup-only VectorQuantizer.forward, norm_ema_quantizer.py:250-268: z_q = weight[idx] viewed b h w c and
    returned b c h w (:253,267); eval: cluster_size <- bincount (EMA with decay forced to 0, :230,257-261);
    loss = beta * mean((z_q - z)^2) against z rearranged b h w c (:262-264), 0 without z (:266).
    Returns (z_q[b,c,h,w], loss, counts or None)."""
    b, h, w = bhw
    weight = _f32(weight)
    idx = np.asarray(idx, np.int64).reshape(-1)
    rows = weight[idx]                                          # F.embedding
    counts = None if training else np.bincount(idx, minlength=weight.shape[0]).astype(np.float32)
    if z is not None:
        zr = np.transpose(_f32(z), (0, 2, 3, 1)).reshape(-1, weight.shape[1])
        diff = rows - zr
        loss = float(beta) * float(np.mean((diff * diff).astype(np.float64)))
    else:
        loss = 0.0
    return np.ascontiguousarray(np.transpose(rows.reshape(b, h, w, -1), (0, 3, 1, 2))), loss, counts


# ------------------------------------------------------------------------------------------
# k-means init (norm_ema_quantizer.py:24-57, cosine branch) given the initial means
# ------------------------------------------------------------------------------------------
def kmeans_cosine(samples, means, num_iters=10):
    """samples @ means.t() -> argmax (first index) -> bincount -> scatter_add mean -> l2norm;
    empty clusters keep their previous mean.  Returns (means, bins of the last iteration)."""
    c = C()
    samples = _f32(samples)
    means = _f32(means).copy()
    k, d = means.shape
    bins = np.zeros(k, np.int64)
    for _ in range(num_iters):
        # argmax of the dot == argmin of (-dot); build it from the distance matrix kernel with
        # zero norms is not bit-identical, so do the dot explicitly (sequential fmaf chain).
        dots = _dot_chain(samples, means)                       # :35
        buckets = dots.argmax(axis=1)                           # :41 (first index on ties)
        bins = np.bincount(buckets, minlength=k)                # :42
        sums = np.zeros((k, d), np.float64)
        np.add.at(sums, buckets, samples.astype(np.float64))    # :46-47 (fp32 scatter_add in ref)
        new = (sums / np.maximum(bins, 1)[:, None]).astype(np.float32)   # :48
        new, _ = c.l2norm_rows(new)                             # :51
        means = np.where((bins == 0)[:, None], means, new)      # :53
    return means, bins


def _dot_chain(a, b):
    """[n,d] x [k,d] -> [n,k] with the sequential-fmaf chain (via the distance kernel:
    d = (zz + cc) - 2 dot is not invertible bit-exactly, so use float64 emulation)."""
    a64 = a.astype(np.float64)
    b64 = b.astype(np.float64)
    acc = np.zeros((a.shape[0], b.shape[0]), np.float32)
    for j in range(a.shape[1]):
        acc = (a64[:, j:j + 1] * b64[None, :, j] + acc.astype(np.float64)).astype(np.float32)
    return acc
