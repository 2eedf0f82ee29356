"""Streaming front-end for host-resident code embeddings (the tokenization sweep of inference.py:106-138:
embeddings come off the encoders / a loader on the host side batch by batch, code indices go back).

`quantize_stream` keeps the GPU busy with a 2-deep pipeline: while batch i is being quantised on the
compute stream, batch i+1 is uploaded from pinned host memory on a copy stream and the indices of batch
i-1 are downloaded.  All tensor work is the module's own forward (libmedtok_vq.so); this file only moves
bytes and orders streams.
"""
from __future__ import annotations

import torch


class HostStream:
    def __init__(self, module, rows: int, dim: int, device, depth: int = 2, return_zq: bool = False):
        self.module = module
        self.device = torch.device(device)
        self.depth = depth
        self.copy_in = torch.cuda.Stream(self.device)
        self.copy_out = torch.cuda.Stream(self.device)
        self.z_dev = [torch.empty(rows, dim, 1, 1, device=self.device) for _ in range(depth)]
        self.idx_host = [torch.empty(rows, dtype=torch.int64).pin_memory() for _ in range(depth)]
        self.loss_host = [torch.empty(1).pin_memory() for _ in range(depth)]
        # return_zq: the quantised rows come back too (the reference's whole return tuple); otherwise a consumer that
        # only needs tokens reconstructs them host-side as codebook[idx]
        self.zq_host = [torch.empty(rows, dim).pin_memory() for _ in range(depth)] if return_zq else None
        self.uploaded = [torch.cuda.Event() for _ in range(depth)]
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self.downloaded = [torch.cuda.Event() for _ in range(depth)]
        self.h2d_bytes = rows * dim * 4
        self.d2h_bytes = rows * 8 + 4 + (rows * dim * 4 if return_zq else 0)

    @torch.no_grad()
    def run(self, host_batches):
        """host_batches: iterable of pinned fp32 tensors [rows, dim].  Yields nothing; results of batch i
        are in idx_host[i % depth] / loss_host[i % depth] once `downloaded[i % depth]` has completed.
        Returns the number of batches processed."""
        main = torch.cuda.current_stream(self.device)
        n = 0
        for i, zb in enumerate(host_batches):
            s = i % self.depth
            with torch.cuda.stream(self.copy_in):
                self.copy_in.wait_event(self.computed[s])            # slot's previous batch has been consumed
                self.z_dev[s].view(zb.shape).copy_(zb, non_blocking=True)
                self.uploaded[s].record(self.copy_in)
            main.wait_event(self.uploaded[s])
            main.wait_event(self.downloaded[s])                      # previous results of this slot are out
            z_q, loss, idx = self.module(self.z_dev[s])
            self.computed[s].record(main)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(self.computed[s])
                self.idx_host[s].copy_(idx, non_blocking=True)
                self.loss_host[s].copy_(loss.reshape(1), non_blocking=True)
                if self.zq_host is not None:
                    self.zq_host[s].copy_(z_q.reshape(self.zq_host[s].shape), non_blocking=True)
                self.downloaded[s].record(self.copy_out)
            idx.record_stream(self.copy_out)
            loss.record_stream(self.copy_out)
            z_q.record_stream(self.copy_out)
            n += 1
        return n

    def finish(self):
        self.copy_out.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
