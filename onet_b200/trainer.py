"""Training step of the reference loops (Train_Onet_on_simclutter_20250407.py:209-218, Train_Onet_on_zy3_20240606.py:
102-120) — zero_grad, forward, compute_loss, backward, Adam — as one object, single- or multi-GPU.

Data parallelism (new work; the reference is single-device): one process per GPU, the batch is sharded across
ranks, BatchNorm statistics stay per rank (== the reference run on that shard), and the only collective is the
gradient all-reduce.  All parameters live in one flat fp32 arena (`Onet.flatten_parameters`); the arena is cut into
buckets along U-Net block boundaries and each bucket's NCCL all-reduce is issued as soon as the backward pass has
finished that block, so communication overlaps the rest of backward.  Adam is one fused kernel over the arena.

The whole step (≈230 kernel launches, all shapes static) can be captured once in a CUDA graph and replayed
(`graph=True`): the step count and the hyper-parameters live in device memory (`onet_adam_step_dev`), the input is
copied into a static buffer, the loss is read from a static scalar.
"""
import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, ptr
from .model import invalidate_packed_weights


def plan_buckets(named_blocks, offsets):
    """{block name: (start, end)} element ranges of the flat arena, one bucket per U-Net block.
    named_blocks: [(name, [param, ...])]; offsets: {id(param): (start, end)}."""
    buckets = {}
    for name, params in named_blocks:
        ps = [offsets[id(p)] for p in params]
        if ps:
            buckets[name] = (min(a for a, _ in ps), max(b for _, b in ps))
    return buckets


def cosine_warm_restarts_lr(epoch, base_lr=1e-4, T_0=300, T_mult=2, eta_min=1e-6):
    """Learning rate of `torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0, T_mult, eta_min)` after `epoch` calls
    of scheduler.step() — the schedule of Train_Onet_on_zy3_20240606.py:88-89, 125 (lr 1e-4, T_0 300, T_mult 2, eta_min 1e-6).
    Use with `OnetTrainer.set_lr` once per epoch (the fused Adam reads the rate from device memory, also inside a graph)."""
    import math
    t, T_i = epoch, T_0
    if T_mult == 1:
        t = epoch % T_0
    else:
        while t >= T_i:
            t -= T_i
            T_i *= T_mult
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t / T_i)) / 2


class OnetTrainer:
    def __init__(self, onet, lr=5e-6, betas=(0.9, 0.999), eps=1e-8, process_group=None, overlap=True, graph=False):
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
        # device-resident optimizer state for the graph-capturable Adam
        self.use_graph = bool(graph)
        self._hyper = None
        self._step_dev = None
        self._graph = None
        self._static_x = None
        self._static_loss = None
        self.launches_per_step = None

    # ------------------------------------------------------------------ buckets
    def _make_buckets(self):
        """(unet id, block name) -> (start, end) element range of the arena holding that block's parameters."""
        ar = self.onet._arena
        off = {id(p): (o, o + p.numel()) for p, o in zip(self.onet.parameters(), ar["offsets"])}
        buckets = {}
        for unet in self.onet._unets():
            blocks = [(name, list(mod.parameters())) for name, mod in unet.named_children()]
            for name, rng in plan_buckets(blocks, off).items():
                buckets[(id(unet), name)] = rng
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

    def set_lr(self, lr):
        """Learning-rate schedules (Train_Onet_on_zy3_20240606.py uses cosine restarts): takes effect on the next step,
        also for an already captured graph."""
        self.lr = lr
        if self._hyper is not None:
            self._hyper[0] = float(lr)

    # ------------------------------------------------------------------ one step
    def _device_state(self, dev):
        if self._hyper is None:
            self._hyper = torch.tensor([self.lr, self.betas[0], self.betas[1], self.eps], dtype=torch.float32, device=dev)
            self._step_dev = torch.tensor([self.step_count], dtype=torch.int32, device=dev)

    def _step_body(self, X, adam_on_device):
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
        st = torch.cuda.current_stream(X.device).cuda_stream
        if adam_on_device:
            call("onet_adam_step_dev", ptr(self.flat), ptr(self.grads), ptr(self.m), ptr(self.v), self.flat.numel(),
                 ptr(self._hyper), ptr(self._step_dev), 1.0 / self.world, st, device=self.flat.device)
        else:
            call("onet_adam_step", ptr(self.flat), ptr(self.grads), ptr(self.m), ptr(self.v), self.flat.numel(), float(self.lr),
                 float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_count + 1, 1.0 / self.world, st, device=self.flat.device)
        invalidate_packed_weights()
        return loss.detach()

    def _capture(self, X):
        """Capture one step on a static input buffer.  The warm-up step that precedes the capture (lazy initialisation
        of kernels / NCCL outside the capture) is undone by restoring parameters, optimizer state and BN buffers."""
        dev = self.flat.device
        self._device_state(dev)
        self._static_x = torch.empty(X.shape, dtype=torch.float32, device=dev)
        self._static_x.copy_(X, non_blocking=True)
        saved = [t.clone() for t in (self.flat, self.m, self.v, self._step_dev)] + [b.clone() for b in self.onet.buffers()]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._step_body(self._static_x, adam_on_device=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for t, s in zip([self.flat, self.m, self.v, self._step_dev] + list(self.onet.buffers()), saved):
            t.copy_(s)
        invalidate_packed_weights()
        self.onet._last = None          # drop the warm-up step's autograd graph before capturing
        self._graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._step_body(self._static_x, adam_on_device=True)
        self.launches_per_step = _lib.launch_count() - l0     # kernels of this library inside one replay
        self.onet._last = None          # tensors of the captured forward belong to the graph's private pool

    def step(self, X):
        """X: (B_local, C, H, W) fp32, on this rank's device or in (pinned) host memory.  Returns the (local) loss as a
        0-dim device tensor."""
        dev = self.flat.device
        if self.use_graph and _lib.PROFILE is None:
            if self._graph is None or self._static_x.shape != X.shape:
                self._capture(X)
            self._static_x.copy_(X, non_blocking=True)
            self._graph.replay()
            # the replayed Adam kernel changed the weights through raw pointers: the packed operand copies the replay made
            # BEFORE its update are stale for any eval / eager forward that follows (the next replay repacks by itself)
            invalidate_packed_weights()
            self.step_count += 1
            return self._static_loss
        if X.device != dev:
            X = X.to(dev, non_blocking=True)
        if self._step_dev is not None:          # keep the device-side counter in step with eager steps
            loss = self._step_body(X, adam_on_device=True)
        else:
            loss = self._step_body(X, adam_on_device=False)
        self.step_count += 1
        return loss
