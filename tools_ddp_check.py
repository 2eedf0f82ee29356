"""2+ rank DDP drop-in check (run under torchrun on the GPU box; train_MedTok.py:185 wraps the tokenizer in
DistributedDataParallel(find_unused_parameters=True) and steps it under autocast + GradScaler): one training step of
a small model = Linear -> soft-one VectorQuantizer inside DDP; the gradients DDP all-reduced must be identical on every
rank, finite, and equal to the average of the per-rank gradients of the same module run without DDP."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.nn.parallel import DistributedDataParallel as DDP

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from medtok_b200.vector_quantization_soft_one_new import VectorQuantizer  # noqa: E402


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.pre = nn.Linear(128, 128)
        self.unused = nn.Linear(8, 8)                      # never touched: find_unused_parameters=True must cope
        self.quantize = VectorQuantizer(n_e=1536, e_dim=64, beta=0.25, entropy_loss_ratio=0.0, l2_norm=True,
                                        show_usage=True, split=[64, 64], k=5)

    def forward(self, z, text, nodes, mask, batch, z_aug):
        out = self.quantize(self.pre(z), text, nodes, mask, batch, self.pre(z_aug))
        return (out['shared_embed_loss'][0] + out['shared_embed_loss'][1] + out['text_specific_loss'][0]
                + out['text_specific_loss'][1] + out['graph_specific_loss'][0] + out['graph_specific_loss'][1])


def inputs(r):
    g = torch.Generator().manual_seed(100 + r)
    b, t, e = 64, 9, 64
    z = torch.randn(b, 2 * e, generator=g)
    mask = torch.ones(b, t, dtype=torch.int64)
    batch = torch.arange(b).repeat_interleave(3)
    return [x.to(dev) for x in (z, torch.randn(b, t, e, generator=g), torch.randn(batch.numel(), e, generator=g), mask, batch,
                                z + 0.05 * torch.randn(b, 2 * e, generator=g))]


torch.manual_seed(0)
model = Tiny().to(dev).train()
for m in model.modules():
    if isinstance(m, nn.Dropout):
        m.p = 0.0
    if isinstance(m, nn.MultiheadAttention):
        m.dropout = 0.0
# reference: the same module without DDP, gradients averaged by hand
model.zero_grad(set_to_none=True)
model(*inputs(rank)).backward()
manual = {}
for n, p in model.named_parameters():
    if p.grad is not None:
        g = p.grad.detach().clone()
        dist.all_reduce(g)
        manual[n] = g / world
model.zero_grad(set_to_none=True)

ddp = DDP(model, device_ids=[local], find_unused_parameters=True)
opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4)
scaler = torch.amp.GradScaler("cuda")
with torch.autocast("cuda", dtype=torch.float16):
    loss = ddp(*inputs(rank))
scaler.scale(loss).backward()
scaler.unscale_(opt)
ok_finite, ok_same, worst = True, True, 0.0
for n, p in model.named_parameters():
    if p.grad is None:
        continue
    ok_finite = ok_finite and bool(torch.isfinite(p.grad).all())
    gathered = [torch.empty_like(p.grad) for _ in range(world)]
    dist.all_gather(gathered, p.grad)
    ok_same = ok_same and all(torch.equal(gathered[0], t) for t in gathered)
    if n in manual and n != "quantize.codebook.weight":
        worst = max(worst, float((p.grad - manual[n]).norm() / manual[n].norm().clamp_min(1e-12)))
w0 = model.quantize.codebook.weight.detach().clone()
scaler.step(opt)
scaler.update()
stepped = not torch.equal(w0, model.quantize.codebook.weight)
print(f"rank {rank}/{world}: ddp step under autocast+GradScaler: loss={float(loss):.5f} grads finite={ok_finite} "
      f"identical across ranks={ok_same} rel err vs hand-averaged fp32 grads={worst:.4f} optimizer stepped={stepped}")
assert ok_finite and ok_same and stepped and worst <= 0.05
dist.destroy_process_group()
