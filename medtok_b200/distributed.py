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


class PeerExchange:
    """Peer-mapped buffers for the exchange fused into the finalize kernel (mtvq_ema_finalize_peers): two packed
    statistics buffers per rank (alternating by step), one flag array and one result buffer (the slice owners push the
    updated rows into it), allocated as torch symmetric memory so that every rank holds device pointers into every
    peer's copy (NVLink P2P).  PyTorch only provides the mapping; the signalling, the waiting, the rank-ordered fp64
    sums of a rank's slice and the distribution of the results all happen inside the one kernel.

    Set-up is split so that a failure on ONE rank can never leave the others inside a collective: `__init__` only
    allocates (local, may raise), `connect()` is the collective rendezvous (every rank of the group must call it, which
    the caller guarantees by voting on the allocation first), `arm()` zeroes the flags; the caller votes again and then
    barriers before the first step (NormEMAVectorQuantizer._peer_exchange)."""

    MAX_WORLD = 64                      # flag slots per rank (mtvq_ema_finalize_peers rejects larger worlds)

    def __init__(self, numel: int, device, group=None, result_numel: int = 0):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > self.MAX_WORLD:
            raise RuntimeError(f"fused exchange supports at most {self.MAX_WORLD} ranks, got {self.world}")
        self.numel = numel
        self.device = device
        import warnings
        with warnings.catch_warnings():                            # newer torch enables groups lazily and deprecates this
            warnings.simplefilter("ignore")
            try:
                symm_mem.enable_symm_mem_for_group(self.group.group_name)
            except Exception:
                pass
        self.stats = [symm_mem.empty(numel, dtype=torch.float64, device=device) for _ in range(2)]
        # flags: [0, 64) "statistics ready" per peer, [64, 128) "slice final" per peer, [128] a local block counter
        self.flags = symm_mem.empty(3 * self.MAX_WORLD, dtype=torch.int32, device=device)
        # result buffer the slice owners push the updated rows / cluster sizes into (k*d + k floats)
        self.results = symm_mem.empty(max(result_numel, 1), dtype=torch.float32, device=device)
        self.table = None
        self.epoch = 0

    def connect(self):
        """Collective: exchange the peer mappings.  Raises if the handles do not describe our tensors."""
        import torch.distributed._symmetric_memory as symm_mem
        self._handles = [symm_mem.rendezvous(t, self.group) for t in (*self.stats, self.flags, self.results)]
        tables = []
        for t, h in zip((*self.stats, self.flags, self.results), self._handles):
            off = int(getattr(h, "offset", 0) or 0)
            ptrs = [int(p) + off for p in h.buffer_ptrs]
            if len(ptrs) != self.world or ptrs[self.rank] != t.data_ptr():
                raise RuntimeError("symmetric-memory handle does not describe the tensor it was made from")
            tables.append(ptrs)
        # device tables of peer pointers: [stats parity 0 | stats parity 1 | flags | results], `world` int64 each
        self.table = torch.tensor(tables, dtype=torch.int64, device=self.device)

    def arm(self):
        """Local: zero the flags and statistics (nobody signals before the caller's barrier that follows)."""
        self.flags.zero_()
        for t in self.stats:
            t.zero_()
        self.epoch = 0
        torch.cuda.synchronize(self.device)

    def next_buffer(self) -> torch.Tensor:
        """Statistics buffer of the coming step (call once per step, before mtvq_hard_forward)."""
        self.epoch += 1
        return self.stats[self.epoch & 1]

    def peer_stats_ptr(self) -> int:
        return self.table[self.epoch & 1].data_ptr()

    def peer_flags_ptr(self) -> int:
        return self.table[2].data_ptr()

    def peer_results_ptr(self) -> int:
        return self.table[3].data_ptr()
