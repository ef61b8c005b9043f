"""Halo-tiled inference for large radar frames (BASELINE.json configs[4]: 2048 x 2048 frames across 8 GPUs).

In eval mode the twin U-Net is a fixed convolutional map: BatchNorm uses its running statistics, so an output pixel
depends only on the input inside its receptive field (radius 94 px through the four pooling levels).
A frame is therefore cut into core tiles whose origins are multiples of 16 (the pooling grids of tile and frame
coincide); every tile is evaluated on its core plus a halo of `halo >= 94` pixels (96 by default) taken from the frame,
CLAMPED at the frame border - there the convolution's own zero padding applies, exactly as in the whole-frame forward
(zero-padding the INPUT instead would be wrong: BatchNorm shifts turn a zero input into non-zero activations).  Only the
core of every result is kept.  Tiles are independent: they are grouped by padded shape, batched through
`Onet.forward`, and - with several ranks - dealt round-robin to the ranks with no collective on the data path; the label
mask is gathered at the end.  The result is bit-identical to the whole-frame forward (tests/test_infer_gpu.py).

Reference call sites this replaces: the whole-frame `onet(X)` + `predict_label(S)` of `test_simclutter`
(Train_Onet_on_simclutter_20250407.py:109-147) and `test_on_zy3_nail` (Train_Onet_on_zy3_20240606.py).
"""
from collections import OrderedDict, namedtuple

import torch

Tile = namedtuple("Tile", "y0 x0 y1 x1 py0 px0 py1 px1")   # core [y0,y1) x [x0,x1); padded region [py0,py1) x [px0,px1)

RECEPTIVE_RADIUS = 94
ALIGN = 16


def plan_tiles(H, W, tile=512, halo=96):
    """Core tiles of at most `tile` x `tile` pixels covering the frame exactly once, each with its clamped halo region.
    `tile` and `halo` must be multiples of 16 and H, W multiples of 16 (the network's own constraint)."""
    if tile % ALIGN or halo % ALIGN or tile <= 0:
        raise ValueError("tile and halo must be multiples of 16")
    if H % ALIGN or W % ALIGN:
        raise ValueError("frame height and width must be multiples of 16")
    if halo < RECEPTIVE_RADIUS and (H > tile or W > tile):
        raise ValueError(f"halo {halo} is smaller than the receptive-field radius {RECEPTIVE_RADIUS}: tiled results would differ")
    tiles = []
    for y0 in range(0, H, tile):
        for x0 in range(0, W, tile):
            y1, x1 = min(H, y0 + tile), min(W, x0 + tile)
            tiles.append(Tile(y0, x0, y1, x1, max(0, y0 - halo), max(0, x0 - halo), min(H, y1 + halo), min(W, x1 + halo)))
    return tiles


def group_by_shape(tiles):
    """{(padded height, padded width): [tile index, ...]} - tiles of one group are batched through the network."""
    groups = OrderedDict()
    for i, t in enumerate(tiles):
        groups.setdefault((t.py1 - t.py0, t.px1 - t.px0), []).append(i)
    return groups


def shard(indices, rank, world):
    """Round-robin share of a work list for one rank (no data-path collective: tiles are independent)."""
    return list(indices)[rank::world]


class TiledPredictor:
    """`forward_fn(batch (n,C,h,w) float32) -> (Vt (n,1,h,w), Vd (n,1,h,w))`; for the product path this is the eval-mode
    `onet_b200.Onet` (see `for_onet`).  `max_batch` bounds the number of tiles per network call.

    Two entry points:
      predict(frames)        -> (Vt, Vd, label int64): the full response maps on every rank (diagnostics, parity tests)
      predict_labels(frames) -> uint8 label masks: the throughput path.  Each rank computes the (frame, tile) pairs it owns;
                                there is no data-path collective (each pixel has one owner), frames are copied from (pinned)
                                host memory on a copy stream one frame ahead of the compute stream, and every finished mask
                                goes back as ONE BYTE per pixel on a third stream while the next frame computes."""

    def __init__(self, forward_fn, tile=512, halo=96, max_batch=4, device=None):
        self.forward_fn, self.tile, self.halo, self.max_batch = forward_fn, tile, halo, max_batch
        self.device = device           # CUDA device of the product path; None = host-side stand-in (CPU tests)
        self._cache = {}

    @staticmethod
    def for_onet(onet, tile=512, halo=96, max_batch=4):
        def fwd(x):
            onet.eval()
            with torch.no_grad():
                _, Vt, _, Vd, _ = onet(x)
            return Vt, Vd
        return TiledPredictor(fwd, tile, halo, max_batch, device=next(onet.parameters()).device)

    # ------------------------------------------------------------------ shared: tiles of ONE frame -> its response maps
    def _run_frame(self, frame, tiles, Vt, Vd):
        """frame (C,H,W) and maps Vt, Vd (1,H,W) on the same device; evaluates `tiles` and writes their cores."""
        by_shape = OrderedDict()
        for t in tiles:
            by_shape.setdefault((t.py1 - t.py0, t.px1 - t.px0), []).append(t)
        for items in by_shape.values():
            for k in range(0, len(items), self.max_batch):
                chunk = items[k:k + self.max_batch]
                if len(chunk) == 1 and chunk[0].py1 - chunk[0].py0 == frame.shape[1] and chunk[0].px1 - chunk[0].px0 == frame.shape[2]:
                    batch = frame.unsqueeze(0)                      # whole frame: no staging copy
                else:
                    batch = torch.stack([frame[:, t.py0:t.py1, t.px0:t.px1] for t in chunk]).contiguous()
                vt, vd = self.forward_fn(batch)
                for j, t in enumerate(chunk):
                    oy, ox = t.y0 - t.py0, t.x0 - t.px0
                    h, w = t.y1 - t.y0, t.x1 - t.x0
                    Vt[:, t.y0:t.y1, t.x0:t.x1] = vt[j, :, oy:oy + h, ox:ox + w]
                    Vd[:, t.y0:t.y1, t.x0:t.x1] = vd[j, :, oy:oy + h, ox:ox + w]

    def _owned(self, B, H, W, rank, world):
        """{frame index: [tiles this rank evaluates]}: (frame, tile) pairs dealt round-robin to the ranks."""
        tiles = plan_tiles(H, W, self.tile, self.halo)
        work = [(b, i) for b in range(B) for i in range(len(tiles))]
        per_frame = OrderedDict()
        for b, i in shard(work, rank, world):
            per_frame.setdefault(b, []).append(tiles[i])
        return per_frame

    def predict(self, frames, rank=0, world=1, process_group=None):
        """frames: (B,C,H,W) float32 in [0,1].  Returns (Vt, Vd, label) with Vt, Vd (B,1,H,W) float32 and label (B,H,W)
        int64 (1 iff Vd > Vt, reference predict_label :193-202).  With world > 1 every rank computes its share of the
        tiles and the maps are summed across ranks at the end so that EVERY rank holds the full fp32 maps (each pixel is
        written by exactly one rank) - the diagnostic entry point; `predict_labels` is the one that moves no maps."""
        B, C, H, W = frames.shape
        Vt = torch.zeros(B, 1, H, W, dtype=torch.float32, device=frames.device)
        Vd = torch.zeros_like(Vt)
        for b, tl in self._owned(B, H, W, rank, world).items():
            self._run_frame(frames[b], tl, Vt[b], Vd[b])
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(Vt, group=process_group)
            dist.all_reduce(Vd, group=process_group)
        if Vt.is_cuda:
            from ._lib import call, ptr
            label = torch.empty(B, H, W, dtype=torch.long, device=Vt.device)
            call("onet_predict_label", ptr(Vt), ptr(Vd), label.numel(), ptr(label), torch.cuda.current_stream(Vt.device).cuda_stream, device=Vt.device)
        else:           # host-side stand-in used by the CPU tests of the tiling / sharding logic
            label = (Vd > Vt).squeeze(1).long()
        return Vt, Vd, label

    # ------------------------------------------------------------------ throughput path: uint8 masks, no map exchange
    def _buffers(self, key, make):
        if key not in self._cache:
            self._cache[key] = make()
        return self._cache[key]

    def predict_labels(self, frames, rank=0, world=1, process_group=None, gather=False):
        """frames: (B,C,H,W) float32 in [0,1], in host memory (pinned for asynchronous copies) or on this rank's device.
        Returns the uint8 masks (B,H,W), 1 iff Vd > Vt (reference predict_label :193-202; `.long()` gives its dtype): in
        pinned host memory when the frames came from the host, on the device otherwise.  Pixels of (frame, tile) pairs owned
        by other ranks are 0 unless gather=True, which sums the uint8 masks across ranks (one byte per pixel; still no fp32
        map leaves a rank).  The returned tensor is a cached buffer, overwritten by the next call with the same shape."""
        B, C, H, W = frames.shape
        owned = self._owned(B, H, W, rank, world)
        if self.device is None:            # host-side stand-in (CPU tests of the tiling / sharding logic)
            lab = torch.zeros(B, H, W, dtype=torch.uint8)
            for b, tl in owned.items():
                Vt, Vd = torch.zeros(1, H, W), torch.zeros(1, H, W)
                self._run_frame(frames[b], tl, Vt, Vd)
                lab[b] = (Vd > Vt)[0].to(torch.uint8)
            if gather and world > 1:
                import torch.distributed as dist
                dist.all_reduce(lab, group=process_group)
            return lab
        from ._lib import call, ptr
        dev = self.device
        host_in = not frames.is_cuda
        main = torch.cuda.current_stream(dev)
        s_in, s_out = self._buffers(("streams", dev), lambda: (torch.cuda.Stream(dev), torch.cuda.Stream(dev)))
        with torch.cuda.device(dev):
            maps = self._buffers(("maps", H, W), lambda: torch.zeros(2, 1, H, W, dtype=torch.float32, device=dev))
            lab_dev = self._buffers(("lab", B, H, W, rank, world), lambda: torch.zeros(B, H, W, dtype=torch.uint8, device=dev))
            whole = all(len(tl) == 1 and (tl[0].y1 - tl[0].y0, tl[0].x1 - tl[0].x0) == (H, W) for tl in owned.values())
            if host_in:
                xbuf = self._buffers(("xbuf", C, H, W), lambda: torch.empty(2, C, H, W, dtype=torch.float32, device=dev))
                out = self._buffers(("out", B, H, W, rank, world), lambda: torch.zeros(B, H, W, dtype=torch.uint8).pin_memory())
                ev_in = [torch.cuda.Event(), torch.cuda.Event()]
                ev_done = [torch.cuda.Event(), torch.cuda.Event()]
                s_in.wait_stream(main)
                s_out.wait_stream(main)
            for k, (b, tl) in enumerate(owned.items()):
                slot = k & 1
                if host_in:
                    if k >= 2:
                        s_in.wait_event(ev_done[slot])           # the compute of frame k-2 has finished reading xbuf[slot]
                    with torch.cuda.stream(s_in):
                        xbuf[slot].copy_(frames[b], non_blocking=True)
                        ev_in[slot].record(s_in)
                    main.wait_event(ev_in[slot])
                    frame = xbuf[slot]
                else:
                    frame = frames[b]
                if not whole:
                    maps.zero_()                                   # pixels this rank does not own must read as label 0
                self._run_frame(frame, tl, maps[0], maps[1])
                call("onet_predict_label_u8", ptr(maps[0]), ptr(maps[1]), H * W, ptr(lab_dev[b]), main.cuda_stream)
                if host_in:
                    ev_done[slot].record(main)
                    if not (gather and world > 1):
                        s_out.wait_event(ev_done[slot])
                        with torch.cuda.stream(s_out):
                            out[b].copy_(lab_dev[b], non_blocking=True)      # 1 byte per pixel, overlaps the next frame
            if gather and world > 1:
                import torch.distributed as dist
                for b in range(B):
                    if b not in owned:
                        lab_dev[b].zero_()
                dist.all_reduce(lab_dev, group=process_group)
                if host_in:
                    out.copy_(lab_dev, non_blocking=True)
                    main.synchronize()
            elif host_in:
                s_out.synchronize()                                # the caller reads the masks now
        return out if host_in else lab_dev
