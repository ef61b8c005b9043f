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
    `onet_b200.Onet` (see `for_onet`).  `max_batch` bounds the number of tiles per network call."""

    def __init__(self, forward_fn, tile=512, halo=96, max_batch=4):
        self.forward_fn, self.tile, self.halo, self.max_batch = forward_fn, tile, halo, max_batch

    @staticmethod
    def for_onet(onet, tile=512, halo=96, max_batch=4):
        def fwd(x):
            onet.eval()
            with torch.no_grad():
                _, Vt, _, Vd, _ = onet(x)
            return Vt, Vd
        return TiledPredictor(fwd, tile, halo, max_batch)

    def predict(self, frames, rank=0, world=1, process_group=None):
        """frames: (B,C,H,W) float32 in [0,1].  Returns (Vt, Vd, label) with Vt, Vd (B,1,H,W) float32 and label (B,H,W)
        int64 (1 iff Vd > Vt, reference predict_label :193-202).  With world > 1 every rank computes its share of the
        tiles and the three maps are summed across ranks at the end (each pixel is written by exactly one rank)."""
        B, C, H, W = frames.shape
        tiles = plan_tiles(H, W, self.tile, self.halo)
        Vt = torch.zeros(B, 1, H, W, dtype=torch.float32, device=frames.device)
        Vd = torch.zeros_like(Vt)
        work = [(b, i) for b in range(B) for i in range(len(tiles))]          # (frame, tile) pairs, dealt to the ranks
        mine = shard(work, rank, world)
        by_shape = OrderedDict()
        for b, i in mine:
            t = tiles[i]
            by_shape.setdefault((t.py1 - t.py0, t.px1 - t.px0), []).append((b, t))
        for items in by_shape.values():
            for k in range(0, len(items), self.max_batch):
                chunk = items[k:k + self.max_batch]
                batch = torch.stack([frames[b, :, t.py0:t.py1, t.px0:t.px1] for b, t in chunk]).contiguous()
                vt, vd = self.forward_fn(batch)
                for j, (b, t) in enumerate(chunk):
                    oy, ox = t.y0 - t.py0, t.x0 - t.px0
                    h, w = t.y1 - t.y0, t.x1 - t.x0
                    Vt[b, :, t.y0:t.y1, t.x0:t.x1] = vt[j, :, oy:oy + h, ox:ox + w]
                    Vd[b, :, t.y0:t.y1, t.x0:t.x1] = vd[j, :, oy:oy + h, ox:ox + w]
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(Vt, group=process_group)
            dist.all_reduce(Vd, group=process_group)
        if Vt.is_cuda:
            from ._lib import call, ptr
            label = torch.empty(B, H, W, dtype=torch.long, device=Vt.device)
            call("onet_predict_label", ptr(Vt), ptr(Vd), label.numel(), ptr(label), torch.cuda.current_stream(Vt.device).cuda_stream)
        else:           # host-side stand-in used by the CPU tests of the tiling / sharding logic
            label = (Vd > Vt).squeeze(1).long()
        return Vt, Vd, label
