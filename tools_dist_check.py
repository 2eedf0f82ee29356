"""2+ rank NCCL check (run under torchrun on the GPU box): the row-sharded forward with one packed all-reduce
must give every rank the same codebook / cluster_size as a single-process forward on the concatenated batch."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from medtok_b200.distributed import shard_rows  # noqa: E402
from medtok_b200.norm_ema_quantizer import NormEMAVectorQuantizer  # noqa: E402

n_total, d, k = 200_000, 64, 4096
g = torch.Generator().manual_seed(123)
z = torch.randn(n_total, d, generator=g)
w = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=-1)
lo, hi = shard_rows(n_total, world, rank)

q = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()          # binds dist.all_reduce (process group is up)
q.embedding.weight.data.copy_(w.to(dev))
steps = 3                                                         # several steps: both statistics buffers, growing epochs
with torch.no_grad():
    for _ in range(steps):
        zq, loss, idx = q(z[lo:hi].view(-1, d, 1, 1).to(dev))
torch.cuda.synchronize()

ref = NormEMAVectorQuantizer(k, d, 0.25).to(dev).train()
ref.all_reduce_fn = nn.Identity()                                 # single-process semantics on the full batch
ref.embedding.weight.data.copy_(w.to(dev))
with torch.no_grad():
    for _ in range(steps):
        zq1, loss1, idx1 = ref(z.view(-1, d, 1, 1).to(dev))
torch.cuda.synchronize()

ok_idx = torch.equal(idx, idx1[lo:hi])
ok_cs = torch.equal(q.cluster_size, ref.cluster_size)
ok_w = torch.equal(q.embedding.weight, ref.embedding.weight)
maxdw = (q.embedding.weight - ref.embedding.weight).abs().max().item()
gathered = [torch.empty_like(q.embedding.weight) for _ in range(world)]
dist.all_gather(gathered, q.embedding.weight.data)
same_replicas = all(torch.equal(gathered[0], t) for t in gathered)
print(f"rank {rank}/{world}: route={'fused NVLink exchange in finalize' if q._peers is not None else 'all_reduce + finalize'} "
      f"idx shard equal={ok_idx} cluster_size equal={ok_cs} codebook bit-equal={ok_w} "
      f"(max diff {maxdw:.2e}) replicas identical={same_replicas} local loss={float(loss):.6f} full loss={float(loss1):.6f}")
assert ok_idx and ok_cs and same_replicas and maxdw <= 1.2e-7
dist.destroy_process_group()
