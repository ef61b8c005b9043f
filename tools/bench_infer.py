"""Inference throughput on large radar frames (BASELINE.json configs[4]):

    python tools/bench_infer.py [--frames 4] [--size 2048] [--tile 1024] [--steps 5]                    # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_infer.py     # N GPUs

Eval-mode Onet (running BatchNorm statistics), halo-tiled (halo 96 px), (frame, tile) pairs dealt round-robin to the ranks,
label mask = argmax of the 2-way softmax as uint8 (`TiledPredictor.predict_labels`: no data-path collective).  The driver-run
figure is the `infer` sub-record of bench.py; this tool is for other frame counts / tile sizes.  Prints ONE JSON line: whole-job Mpix/s (frame pixels, halo overhead not counted) with the
frames resident in HBM, the same with the frames copied from pinned host memory and the mask copied back every step
(`e2e`), and the CPU oracle port timed on a bounded sample (one 512 x 512 frame)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--max-batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import onet_b200
    from onet_b200.data import rayleigh_target_frames
    from onet_b200.infer import TiledPredictor

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1981)
    net = onet_b200.Onet(1, True, True, mode="bf16").to(dev)
    for b in net.buffers():
        if world > 1:
            dist.broadcast(b, 0)
    if world > 1:
        for p in net.parameters():
            dist.broadcast(p.data, 0)
        onet_b200.invalidate_packed_weights()
    pred = TiledPredictor.for_onet(net, tile=args.tile, halo=96, max_batch=args.max_batch)
    S = args.size
    host = rayleigh_target_frames(args.frames, 1, S, S, seed=7, n_targets=20).pin_memory()
    frames = host.to(dev)

    def run(resident):
        # uint8 masks; each rank computes the (frame, tile) pairs it owns, nothing is exchanged between ranks.  Host frames:
        # copied in on a copy stream one frame ahead, masks copied back (1 byte per pixel) on a third stream, host-synchronised
        return pred.predict_labels(frames if resident else host, rank=rank, world=world)

    def timed(resident):
        for _ in range(args.warmup):
            run(resident)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run(resident)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / args.steps

    ms_dev = timed(True)
    ms_e2e = timed(False)
    if rank == 0:
        mpix = args.frames * S * S / 1e6
        line = dict(metric="onet_infer_mpix_per_sec", value=mpix / (ms_dev * 1e-3), unit="Mpix/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_dev, higher_is_better=True, scaling="strong", dtype="bf16", data="synthetic",
                    config=dict(workload=f"Onet eval-mode inference, {args.frames} frames of 1x{S}x{S} Rayleigh clutter + targets, "
                                         f"halo-tiled (tile {args.tile}, halo 96), tiles dealt to {world} rank(s)",
                                frames=args.frames, tile=args.tile, halo=96),
                    e2e=dict(value=mpix / (ms_e2e * 1e-3), unit="Mpix/s", ms_per_step=ms_e2e, h2d_bytes_per_step=args.frames * S * S * 4 // world,
                             d2h_bytes_per_step=args.frames * S * S // world))
        if not args.no_cpu_baseline:      # the CPU leg lives in bench.py (the one place outside tests/ that executes the oracle)
            import bench
            threads = os.cpu_count() or 1
            mpix_cpu, dt = bench.cpu_infer_baseline(threads)
            line["cpu_baseline"] = dict(value=mpix_cpu, unit="Mpix/s", cores=threads, kind="port",
                                        sample=f"one 1x512x512 frame, eval-mode oracle forward ({dt:.2f} s)")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
