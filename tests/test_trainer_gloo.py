"""Host-side logic of the data-parallel trainer on CPU: two gloo ranks, world_size 2.

The CUDA step itself needs a GPU; what is covered here is everything around it: the flat parameter arena, the
per-block gradient buckets (every parameter in exactly one bucket, issued in backward order), the overlapped
all-reduce producing the cross-rank SUM that the Adam kernel then scales by 1/world, and the initial broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        import onet_b200
        from onet_b200.trainer import OnetTrainer
        from onet_b200.model import _DEC, _ENC

        torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT weights
        net = onet_b200.Onet(1, True, True)
        tr = OnetTrainer(net, lr=1e-3)
        assert tr.world == world
        flat, grads = tr.flat, tr.grads

        # arena: every parameter is a view into `flat`, its gradient a view into `grads`
        total = 0
        for p, o in zip(net.parameters(), net._arena["offsets"]):
            assert p.data_ptr() == flat.data_ptr() + 4 * o and p.grad.data_ptr() == grads.data_ptr() + 4 * o
            total += p.numel()
        assert total == 31036416

        # buckets: disjoint, and together they cover every parameter
        cover = torch.zeros(flat.numel(), dtype=torch.int8)
        for a, b in tr._bucket_of.values():
            cover[a:b] += 1
        assert int(cover.max()) == 1
        for p, o in zip(net.parameters(), net._arena["offsets"]):
            assert bool((cover[o:o + p.numel()] == 1).all())

        # broadcast: all ranks end with rank 0's parameters and BatchNorm buffers
        ref = flat.clone()
        tr.broadcast_parameters(0)
        gathered = [torch.zeros(8) for _ in range(world)]
        dist.all_gather(gathered, flat[:8].clone())
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        if rank == 0:
            assert torch.equal(ref, flat)

        # overlapped all-reduce in backward block order: the result is the SUM over ranks in every parameter slot
        grads.copy_(torch.arange(grads.numel(), dtype=torch.float32).remainder_(97.0).mul_(rank + 1))
        expect = torch.arange(grads.numel(), dtype=torch.float32).remainder_(97.0).mul_(sum(r + 1 for r in range(world)))
        order = [n for n, _, _ in reversed(_DEC)] + ["down4"] + [n for n, _, _ in reversed(_ENC[:3])] + ["inc"]
        assert sorted(order) == sorted(n for n, _ in net.topu.named_children())
        for name in order:
            tr._after_block(net.topu, name)
        assert len(tr._handles) == len(order)
        for h in tr._handles:
            h.wait()
        tr._handles = []
        assert bool((grads[cover.bool()] == expect[cover.bool()]).all())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
        raise e


@pytest.mark.timeout(600)
def test_dp_trainer_host_logic_two_gloo_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
