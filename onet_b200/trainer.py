"""Training step of the reference loops (Train_Onet_on_simclutter_20250407.py:209-218, Train_Onet_on_zy3_20240606.py:
102-120) — zero_grad, forward, compute_loss, backward, Adam — as one object, single- or multi-GPU.

Data parallelism (new work; the reference is single-device): one process per GPU, the batch is sharded across
ranks, BatchNorm statistics stay per rank (== the reference run on that shard), and the only collective is the
gradient all-reduce.  All parameters live in one flat fp32 arena (`Onet.flatten_parameters`); the arena is cut into
buckets along U-Net block boundaries and each bucket's NCCL all-reduce is issued as soon as the backward pass has
finished that block, so communication overlaps the rest of backward.  Adam is one fused kernel over the arena.
"""
import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, ptr
from .model import invalidate_packed_weights


class OnetTrainer:
    def __init__(self, onet, lr=5e-6, betas=(0.9, 0.999), eps=1e-8, process_group=None, overlap=True):
        self.onet = onet
        self.lr, self.betas, self.eps = lr, betas, eps
        self.flat, self.grads = onet.flatten_parameters()
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step_count = 0
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.pg = process_group
        self.overlap = overlap
        self._handles = []
        self._bucket_of = self._make_buckets()

    # ------------------------------------------------------------------ buckets
    def _make_buckets(self):
        """block module -> (start, end) element range of the arena holding that block's parameters."""
        ar = self.onet._arena
        off = {id(p): (o, o + p.numel()) for p, o in zip(self.onet.parameters(), ar["offsets"])}
        buckets = {}
        for unet in self.onet._unets():
            for name, mod in unet.named_children():
                ps = [off[id(p)] for p in mod.parameters()]
                if ps:
                    buckets[(id(unet), name)] = (min(a for a, _ in ps), max(b for _, b in ps))
        return buckets

    def _after_block(self, unet, name):
        if self.world == 1 or not self.overlap:
            return
        a, b = self._bucket_of[(id(unet), name)]
        self._handles.append(dist.all_reduce(self.grads[a:b], group=self.pg, async_op=True))

    def broadcast_parameters(self, src=0):
        """Make every rank start from rank `src`'s weights and BatchNorm buffers."""
        if self.world == 1:
            return
        dist.broadcast(self.flat, src, group=self.pg)
        for b in self.onet.buffers():
            dist.broadcast(b, src, group=self.pg)
        invalidate_packed_weights()

    # ------------------------------------------------------------------ one step
    def step(self, X):
        """X: (B_local, C, H, W) fp32 on this rank's device.  Returns the (local) loss as a 0-dim device tensor."""
        onet = self.onet
        onet.train()
        self.grads.zero_()
        onet._after_block = self._after_block
        Lt, Vt, Ld, Vd, S = onet(X)
        St = S[:, 0, :, :].unsqueeze(dim=1)
        Sd = S[:, 1, :, :].unsqueeze(dim=1)
        loss = onet.compute_loss(Lt, St, Ld, Sd)
        loss.backward()
        onet._after_block = None
        if self.world > 1:
            if self.overlap:
                for h in self._handles:
                    h.wait()
                self._handles = []
            else:
                dist.all_reduce(self.grads, group=self.pg)
        self.step_count += 1
        call("onet_adam_step", ptr(self.flat), ptr(self.grads), ptr(self.m), ptr(self.v), self.flat.numel(), float(self.lr),
             float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_count, 1.0 / self.world,
             torch.cuda.current_stream(X.device).cuda_stream)
        invalidate_packed_weights()
        return loss.detach()
