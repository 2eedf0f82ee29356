"""Host-side pieces of the multi-GPU path (one process per GPU, rows sharded, codebook replicated).

The only exchange step of the quantizer is the per-code statistics of the hard path
(reference: two fp32 all-reduces + a rank-local loss, norm_ema_quantizer.py:188,195,203,212).  Here every
rank fills ONE packed fp64 buffer [ bins(K) | embed_sum(K*D) | sse | rows ] (mtvq_stats_len) and a single
SUM all-reduce makes it global; the EMA step then runs identically on every rank, so replicas stay
bit-identical without a broadcast.
"""
from __future__ import annotations

import torch


def shard_rows(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """SURVEY 8e partitioning: rank r takes rows [r*ceil(N/W), min(N, (r+1)*ceil(N/W)))."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def stats_layout(d: int, k: int) -> dict:
    return {"bins": (0, k), "embed_sum": (k, k + k * d), "sse": k + k * d, "rows": k + k * d + 1,
            "len": k + k * d + 2}


def pack_stats(bins, embed_sum, sse, rows) -> torch.Tensor:
    """Build the packed buffer from its parts (used by tests and by hosts that compute parts separately)."""
    k, d = embed_sum.shape
    out = torch.empty(stats_layout(d, k)["len"], dtype=torch.float64, device=embed_sum.device)
    out[:k] = bins
    out[k:k + k * d] = embed_sum.reshape(-1)
    out[k + k * d] = sse
    out[k + k * d + 1] = rows
    return out


def unpack_stats(stats: torch.Tensor, d: int, k: int):
    lay = stats_layout(d, k)
    return (stats[:k], stats[k:k + k * d].view(k, d), stats[lay["sse"]], stats[lay["rows"]])


def all_reduce_stats(stats: torch.Tensor, all_reduce_fn) -> torch.Tensor:
    """One SUM all-reduce of the packed statistics (in place).  `all_reduce_fn` is what the module bound at
    construction: torch.distributed.all_reduce, or nn.Identity() when no process group exists."""
    if stats.dtype != torch.float64 or not stats.is_contiguous():
        raise ValueError("packed statistics must be a contiguous float64 tensor")
    all_reduce_fn(stats)
    return stats
