"""Full-vocabulary tokenization sweep: the result side of `inference.py:97-134` (BASELINE config 4).

The reference loops over the code vocabulary in batches, keeps every batch's `(embeddings, tokens, weights)` in Python
lists, reorders them with `sorted(zip(x_indices, ...))` -- which sorts by the LAST batch's `code_indices` only
(`inference.py:119-121`) -- concatenates and writes `embeddings_all.npy [N, 4*D]`,
`tokens_all.npy [N, 4, k]` (int64) and `weights_all.npy [N, 4, k]` (fp32), the files `EHRModel_token.py:128` and
`MedTok_QA_inference.py:93` load.  Under DDP every rank writes the same three paths.

`TokenSweep` keeps the file formats and fixes the bookkeeping: rows are placed by their code index as they arrive
(preallocated pinned host arrays, asynchronous device->host copies on a side stream), ranks own disjoint sets of code
indices, and `finish()` merges the ranks' rows on rank 0 before the three files are written once.
"""
from __future__ import annotations

import os

import numpy as np
import torch


def compose_eval_outputs(out: dict):
    """`MedTok.forward` in eval mode, tokenizer.py:229-247: the quantizer's dict -> (embedding [B, 4*D],
    tokens [B, 4, k] int64, weights [B, 4, k] fp32) in the order text, graph, shared-text, shared-graph."""
    tokens = torch.cat((out['text_tokens'], out['graph_tokens'], out['shared_text_tokens'], out['shared_graph_tokens']), dim=-1)
    weights = torch.cat((out['text_tokens_weights'], out['graph_tokens_weights'], out['shared_text_tokens_weights'],
                         out['shared_graph_tokens_weights']), dim=-1)
    tokens = tokens.view(-1, 4, tokens.size(-1) // 4)
    weights = weights.view(-1, 4, weights.size(-1) // 4)
    emb = torch.cat((out['specific_embedding_text'], out['specific_embedding_graph'], out['shared_text_embedding'],
                     out['shared_graph_embedding']), dim=-1)
    return emb, tokens, weights


class TokenSweep:
    def __init__(self, n_codes: int, emb_dim: int, k: int = 5, pin: bool | None = None):
        pin = torch.cuda.is_available() if pin is None else pin
        self.n_codes, self.emb_dim, self.k = n_codes, emb_dim, k
        self.embeddings = torch.zeros(n_codes, emb_dim, dtype=torch.float32, pin_memory=pin)
        self.tokens = torch.zeros(n_codes, 4, k, dtype=torch.int64, pin_memory=pin)
        self.weights = torch.zeros(n_codes, 4, k, dtype=torch.float32, pin_memory=pin)
        self.seen = torch.zeros(n_codes, dtype=torch.bool)
        self._pending = []           # (code indices on the host, staged host copies, event)
        self._copy = None

    @torch.no_grad()
    def add(self, code_indices, embeddings, tokens, weights) -> None:
        """One batch: `code_indices` [B] (position of each row in the vocabulary, `x.code_indices` in the reference),
        results as returned by `compose_eval_outputs`.  Device tensors are copied out asynchronously."""
        ci = torch.as_tensor(code_indices).reshape(-1).to(torch.int64).cpu()
        if ci.numel() != embeddings.shape[0]:
            raise ValueError("code_indices and the batch disagree on the number of rows")
        if ci.numel() and (int(ci.min()) < 0 or int(ci.max()) >= self.n_codes):
            raise IndexError("code index outside the vocabulary")
        if embeddings.is_cuda:
            if self._copy is None:
                self._copy = torch.cuda.Stream(embeddings.device)
            ev = torch.cuda.Event()
            self._copy.wait_stream(torch.cuda.current_stream(embeddings.device))
            with torch.cuda.stream(self._copy):
                staged = [t.detach().to(dt).contiguous().to("cpu", non_blocking=True)
                          for t, dt in ((embeddings, torch.float32), (tokens, torch.int64), (weights, torch.float32))]
                ev.record(self._copy)
            for t in (embeddings, tokens, weights):
                t.record_stream(self._copy)
            self._pending.append((ci, staged, ev))
            if len(self._pending) > 4:
                self._drain(keep=2)
        else:
            self._place(ci, embeddings.detach().float(), tokens.detach().to(torch.int64), weights.detach().float())

    def _place(self, ci, e, t, w):
        self.embeddings[ci] = e.reshape(-1, self.emb_dim)
        self.tokens[ci] = t.reshape(-1, 4, self.k)
        self.weights[ci] = w.reshape(-1, 4, self.k)
        self.seen[ci] = True

    def _drain(self, keep: int = 0):
        while len(self._pending) > keep:
            ci, (e, t, w), ev = self._pending.pop(0)
            ev.synchronize()
            self._place(ci, e, t, w)

    def finish(self, group=None) -> bool:
        """Merge the ranks' rows on rank 0 (no-op without a process group).  Returns True on the rank that holds the
        complete arrays.  Every code index must have been added by exactly one rank."""
        self._drain()
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return True
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = self.seen.nonzero().flatten()
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine.numel()], dtype=torch.int64, device=dev), group=group)
        cap = int(max(int(c) for c in counts))
        # one padded message per array; the row payload is packed as bytes so that a single gather moves all of it
        row_bytes = 8 + 4 * self.emb_dim + 8 * 4 * self.k + 4 * 4 * self.k
        msg = torch.zeros(cap, row_bytes, dtype=torch.uint8)
        if mine.numel():
            parts = [mine.view(-1, 1).contiguous().view(torch.uint8).view(mine.numel(), 8),
                     self.embeddings[mine].contiguous().view(torch.uint8).view(mine.numel(), -1),
                     self.tokens[mine].contiguous().view(torch.uint8).view(mine.numel(), -1),
                     self.weights[mine].contiguous().view(torch.uint8).view(mine.numel(), -1)]
            msg[:mine.numel()] = torch.cat(parts, dim=1)
        msg = msg.to(dev)
        gathered = [torch.empty_like(msg) for _ in range(world)] if rank == 0 else None
        dist.gather(msg, gathered, dst=0, group=group)
        if rank != 0:
            return False
        for r in range(world):
            n_r = int(counts[r])
            if r == 0 or n_r == 0:
                continue
            m = gathered[r][:n_r].cpu()
            o = 0
            ci = m[:, o:o + 8].contiguous().view(torch.int64).view(-1); o += 8
            e = m[:, o:o + 4 * self.emb_dim].contiguous().view(torch.float32); o += 4 * self.emb_dim
            t = m[:, o:o + 32 * self.k].contiguous().view(torch.int64); o += 32 * self.k
            w = m[:, o:o + 16 * self.k].contiguous().view(torch.float32)
            if bool(self.seen[ci].any()):
                raise RuntimeError("a code index was tokenized by more than one rank")
            self._place(ci, e, t, w)
        return True

    def save(self, out_dir: str) -> None:
        """The three files of inference.py:132-134 (same names, dtypes and shapes)."""
        if not bool(self.seen.all()):
            missing = int((~self.seen).sum())
            raise RuntimeError(f"{missing} of {self.n_codes} codes were never tokenized")
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, "embeddings_all.npy"), self.embeddings.numpy())
        np.save(os.path.join(out_dir, "tokens_all.npy"), self.tokens.numpy())
        np.save(os.path.join(out_dir, "weights_all.npy"), self.weights.numpy())
