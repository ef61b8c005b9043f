"""Timeline evidence for the gradient all-reduce overlap (SURVEY.md section 5 "Distributed communication backend"):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/overlap_timeline.py [--batch 64]

Every rank runs the data-parallel training step of bench.py (OnetTrainer, bucketed all-reduce issued per finished U-Net block);
rank 0 records three steps with the CUPTI kernel tracer (torch.profiler, nsys is not in this image) and prints, per NCCL
all-reduce kernel of the LAST step: start, duration, which of this library's kernels ran on the other streams while it was in
flight, and how much of it was EXPOSED (no compute kernel running at the same time).  Also the step's critical numbers: span of
the step, sum of the NCCL kernel times, sum of the exposed parts.  The output is committed under profiles/."""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--graph", action="store_true", help="replay the captured step graph instead of launching from the host")
    args = ap.parse_args()
    import onet_b200
    from onet_b200.data import k_clutter_frames
    from onet_b200.trainer import OnetTrainer
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1981)
    net = onet_b200.Onet(1, True, True, mode="bf16").to(dev)
    tr = OnetTrainer(net, lr=5e-6, graph=args.graph)
    tr.broadcast_parameters(0)
    x = k_clutter_frames(args.batch, 1, 256, 256, seed=11 + rank, n_targets=8).to(dev)
    for _ in range(4):
        tr.step(x)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    acts = [ProfilerActivity.CUDA] if rank == 0 else []
    if rank == 0:
        with profile(activities=acts) as prof:
            for _ in range(3):
                tr.step(x)
            torch.cuda.synchronize()
        path = os.path.join(tempfile.gettempdir(), "onet_trace.json")
        prof.export_chrome_trace(path)
    else:
        for _ in range(3):
            tr.step(x)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel" and "dur" in e]
        ev.sort(key=lambda e: e["ts"])
        adam = [i for i, e in enumerate(ev) if "adam_dev_kernel" in e["name"] or "adam_kernel" in e["name"]]
        lo = ev[adam[-2]]["ts"] + ev[adam[-2]]["dur"] if len(adam) >= 2 else ev[0]["ts"]
        hi = ev[adam[-1]]["ts"] + ev[adam[-1]]["dur"]
        step = [e for e in ev if lo <= e["ts"] < hi]
        nccl = [e for e in step if "nccl" in e["name"].lower()]
        comp = [e for e in step if "nccl" not in e["name"].lower()]
        print(f"# one data-parallel training step, world {world}, batch {args.batch}/GPU, rank 0, {'graph replay' if args.graph else 'eager launches'}")
        print(f"step span {(hi - lo) / 1e3:.3f} ms, {len(comp)} compute kernels ({sum(e['dur'] for e in comp) / 1e3:.3f} ms summed), "
              f"{len(nccl)} NCCL kernels ({sum(e['dur'] for e in nccl) / 1e3:.3f} ms summed)")
        print("| # | NCCL kernel | start (ms into step) | duration (ms) | exposed (ms) | compute kernels running meanwhile |")
        print("|---|---|---|---|---|---|")
        tot_exposed = 0.0
        for i, n in enumerate(nccl):
            a, b = n["ts"], n["ts"] + n["dur"]
            cover = sorted((max(a, c["ts"]), min(b, c["ts"] + c["dur"])) for c in comp if c["ts"] < b and c["ts"] + c["dur"] > a)
            covered, cur = 0.0, a
            for s, e in cover:
                if e > cur:
                    covered += e - max(s, cur)
                    cur = max(cur, e)
            exposed = n["dur"] - covered
            tot_exposed += exposed
            names = {}
            for c in comp:
                if c["ts"] < b and c["ts"] + c["dur"] > a:
                    k = c["name"].split("(")[0].replace("void onet::", "").replace("onet::", "")[:40]
                    names[k] = names.get(k, 0) + 1
            top = ", ".join(f"{k} x{v}" for k, v in sorted(names.items(), key=lambda kv: -kv[1])[:4])
            print(f"| {i} | {n['name'][:48]} | {(a - lo) / 1e3:.3f} | {n['dur'] / 1e3:.3f} | {exposed / 1e3:.3f} | {top} |")
        print(f"\nexposed communication: {tot_exposed / 1e3:.3f} ms of a {(hi - lo) / 1e3:.3f} ms step "
              f"({100 * tot_exposed / (hi - lo):.1f} %); NCCL busy {sum(e['dur'] for e in nccl) / 1e3:.3f} ms")
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
