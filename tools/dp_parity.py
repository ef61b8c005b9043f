"""Data-parallel parity on real GPUs (run under torchrun, N >= 2 ranks):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py

Every rank takes shard r of one global batch and runs OnetTrainer.step (bucketed NCCL all-reduce overlapped with
backward, then Adam with grad_scale 1/world).  Rank 0 then replays every shard on its own GPU through the single-GPU
engine with identical initial weights, averages the per-shard gradients ("rank r == reference run on
shard r, grads averaged") and compares with the all-reduced gradient arena and the updated parameters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import onet_b200  # noqa: E402
from onet_b200.data import rayleigh_target_frames  # noqa: E402
from onet_b200.trainer import OnetTrainer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, H, W = 4, 64, 64
    x_all = rayleigh_target_frames(B * world, 1, H, W, seed=77)
    for mode in ("fp32", "bf16"):
        torch.manual_seed(1234 + rank)            # different init per rank: broadcast_parameters must fix it
        net = onet_b200.Onet(1, True, True, mode=mode).to(dev)
        tr = OnetTrainer(net, lr=1e-3)
        tr.broadcast_parameters(0)
        w0 = tr.flat.clone()
        loss = tr.step(x_all[rank * B:(rank + 1) * B].to(dev))
        torch.cuda.synchronize()
        g_dp = tr.grads.clone() / world             # arena holds the SUM; Adam applies 1/world
        w1 = tr.flat.clone()
        losses = [torch.zeros((), device=dev) for _ in range(world)]
        dist.all_gather(losses, loss.float())
        if rank == 0:
            ref = onet_b200.Onet(1, True, True, mode=mode).to(dev)
            rt = OnetTrainer(ref, lr=1e-3)
            g_sum = torch.zeros_like(rt.grads)
            ref_losses = []
            for r in range(world):
                rt.flat.copy_(w0)
                onet_b200.invalidate_packed_weights()
                ref.train()
                rt.grads.zero_()
                Lt, Vt, Ld, Vd, S = ref(x_all[r * B:(r + 1) * B].to(dev))
                l = ref.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
                l.backward()
                g_sum += rt.grads
                ref_losses.append(float(l))
            g_ref = g_sum / world
            rel = float((g_dp - g_ref).norm() / g_ref.norm())
            # expected parameter after one Adam step from the all-reduced gradient scaled by 1/world
            # (m = (1-b1) g, v = (1-b2) g^2, bias-corrected => lr * g / (|g| + eps))
            upd = w0 - 1e-3 * g_dp / (g_dp.abs() + 1e-8)
            prel = float((w1 - upd).norm() / (w1 - w0).norm().clamp_min(1e-30))
            lrel = max(abs(float(a) - b) / abs(b) for a, b in zip(losses, ref_losses))
            print(f"dp_parity[{mode}] world={world}: grad rel-L2 {rel:.3e}, Adam update rel {prel:.3e}, per-rank loss rel {lrel:.3e}",
                  flush=True)
            # the gradient at random init is ill-conditioned (tests/test_model_gpu.py): the order of the fp32 atomics alone
            # moves it by ~1e-4..1e-3 (fp32 mode) / ~1e-2 (bf16 operands) run to run; a plumbing error would be O(1)
            tol = 2e-3 if mode == "fp32" else 5e-2
            assert rel < tol and lrel < 1e-5 and prel < 1e-4, (rel, prel, lrel)
        dist.barrier()
    if rank == 0:
        print("dp_parity OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
