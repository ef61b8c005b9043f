"""Tiling / sharding logic of the halo-tiled inference path on CPU (tiles are independent, no data-path
collective).  The network is replaced by a small torch conv stack whose receptive field fits inside the halo and whose
biases make "zero-padded input" differ from "convolution padding" - the case the clamped halo must get right."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onet_b200.infer import TiledPredictor, group_by_shape, plan_tiles, shard


def test_plan_tiles_covers_frame_once_and_is_aligned():
    for H, W, tile, halo in ((2048, 2048, 512, 96), (256, 320, 64, 96), (224, 224, 512, 96), (48, 400, 128, 112)):
        tiles = plan_tiles(H, W, tile, halo)
        cover = torch.zeros(H, W, dtype=torch.int32)
        for t in tiles:
            cover[t.y0:t.y1, t.x0:t.x1] += 1
            assert t.y0 % 16 == 0 and t.x0 % 16 == 0 and t.py0 % 16 == 0 and t.px0 % 16 == 0
            assert (t.py1 - t.py0) % 16 == 0 and (t.px1 - t.px0) % 16 == 0
            assert t.py0 == max(0, t.y0 - halo) and t.py1 == min(H, t.y1 + halo)      # halo clamped at the frame border
            assert t.px0 == max(0, t.x0 - halo) and t.px1 == min(W, t.x1 + halo)
        assert int(cover.min()) == 1 and int(cover.max()) == 1
    assert len(group_by_shape(plan_tiles(2048, 2048, 512, 96))) == 4       # corner / two edge kinds / interior
    with pytest.raises(ValueError):
        plan_tiles(2048, 2048, 512, 64)        # halo below the receptive-field radius
    with pytest.raises(ValueError):
        plan_tiles(200, 200, 64, 96)           # not a multiple of 16
    assert sorted(shard(range(10), 0, 3) + shard(range(10), 1, 3) + shard(range(10), 2, 3)) == list(range(10))


def _standin():
    torch.manual_seed(3)
    convs = [torch.nn.Conv2d(1 if i == 0 else 4, 4, 3, padding=1) for i in range(5)]
    pool = torch.nn.MaxPool2d(2)
    up = torch.nn.Upsample(scale_factor=2, mode="nearest")

    def fwd(x):
        with torch.no_grad():
            a = torch.relu(convs[0](x))
            b = torch.relu(convs[1](pool(a)))
            c = torch.relu(convs[2](pool(b)))
            d = torch.relu(convs[3](up(c))) + b
            e = convs[4](up(d)) + a
            return e[:, 0:1], e[:, 1:2]
    return fwd


def test_tiled_equals_whole_frame_single_rank():
    fwd = _standin()
    x = torch.rand(2, 1, 256, 320)
    vt, vd = fwd(x)
    Vt, Vd, lab = TiledPredictor(fwd, tile=64, halo=96, max_batch=3).predict(x)
    # the CPU conv library picks shape-dependent accumulation orders: equal up to fp32 rounding (a halo / clamping mistake
    # shows up as O(0.1))
    assert torch.allclose(Vt, vt, atol=2e-5, rtol=0) and torch.allclose(Vd, vd, atol=2e-5, rtol=0)
    assert float((lab != (vd > vt).squeeze(1).long()).float().mean()) < 1e-3
    lab8 = TiledPredictor(fwd, tile=64, halo=96, max_batch=3).predict_labels(x)
    assert lab8.dtype == torch.uint8 and torch.equal(lab8.long(), lab)


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        fwd = _standin()
        torch.manual_seed(11)
        x = torch.rand(1, 1, 192, 256)
        vt, vd = fwd(x)
        Vt, Vd, lab = TiledPredictor(fwd, tile=64, halo=96, max_batch=2).predict(x, rank=rank, world=world)
        assert torch.allclose(Vt, vt, atol=2e-5, rtol=0) and torch.allclose(Vd, vd, atol=2e-5, rtol=0)
        assert float((lab != (vd > vt).squeeze(1).long()).float().mean()) < 1e-3
        # throughput entry point: uint8 masks, each rank only its own tiles; gather=True sums the one-byte masks
        pred = TiledPredictor(fwd, tile=64, halo=96, max_batch=2)
        mine = pred.predict_labels(x, rank=rank, world=world)
        full = pred.predict_labels(x, rank=rank, world=world, gather=True)
        assert mine.dtype == torch.uint8 and full.dtype == torch.uint8
        assert float((full.long() != lab).float().mean()) < 1e-3
        assert int(mine.sum()) < int(full.sum()) or int(full.sum()) == 0
        assert bool(((mine == full) | (mine == 0)).all())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
        raise


@pytest.mark.timeout(300)
def test_tiles_sharded_over_two_gloo_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
