#!/usr/bin/env python
"""Benchmark of the Onet hot path (BASELINE.json metric: Onet train imgs/s, fwd + bwd + JSD + Adam).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm's CPU path on the host cores

A "step" is one pass of the training hot path over one batch of synthetic frames:
zero_grad -> forward (twin U-Net, 2B images) -> JSD loss -> backward -> gradient all-reduce (N > 1) -> Adam.
Workload at N = 1 is BASELINE.json configs[1]: batch 64 of 1x256x256 K-distributed sea-clutter frames, bf16
operands with fp32 accumulation.  N > 1 keeps 64 frames per GPU (weak scaling; N = 8 is configs[2]'s global 512).

`value`  : images/s with the batch already resident in HBM.
`e2e`    : the same step through the public module API with the batch copied from pinned host memory every step and
           the loss read back to the host every step (the loop of Train_Onet_on_simclutter_20250407.py:209-219).
`roofline`: dominant kernel family, algorithmic FLOPs / CUDA-event time measured inside this process.
`cpu_baseline`: the CPU oracle port (same ATen CPU kernels the reference dispatches to) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "onet_train_images_per_sec"
UNIT = "images/s"
H = W = 256
CIN = 1
LR = 5e-6          # Train_Onet_on_simclutter_20250407.py:181
EXTRA_MODES = [("tf32", 64), ("fp32", 16)]      # (mode, frames per GPU) of the `modes` sub-record


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(bf16_burst=d.get("bf16_tflops"), bf16_sustained=d.get("bf16_tflops_sustained"), hbm=d.get("hbm_gbs"),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        pw.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                    power_w_median=pw[len(pw) // 2] if pw else None, power_w_max=pw[-1] if pw else None)


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (restatement of the reference algorithm on the same ATen CPU kernels)
# --------------------------------------------------------------------------------------------------
def cpu_steps(batch, steps, warmup, threads):
    import torch
    from oracle import onet_oracle as orc
    from onet_b200.data import k_clutter_frames
    torch.set_num_threads(threads)
    st = orc.init_state(CIN, seed=1981)
    leaves = [k for k, v in st.items() if v.dtype.is_floating_point and "running" not in k]
    params = [st[k].requires_grad_(True) for k in leaves]
    opt = torch.optim.Adam(params, lr=LR, betas=(0.9, 0.999), eps=1e-8)
    x = k_clutter_frames(batch, CIN, H, W, seed=7)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        Lt, Vt, Ld, Vd, S = orc.onet_forward(st, x, training=True)
        loss = orc.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
        loss.backward()
        opt.step()
        float(loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def cpu_infer_baseline(threads):
    """Inference leg of the CPU baseline (tools/bench_infer.py): eval-mode oracle forward of one 1x512x512 frame.
    Returns (Mpix/s, seconds)."""
    import torch
    from oracle import onet_oracle as orc
    from onet_b200.data import rayleigh_target_frames
    torch.set_num_threads(threads)
    st = orc.init_state(1, seed=1981)
    xs = rayleigh_target_frames(1, 1, 512, 512, seed=9)
    with torch.no_grad():
        orc.onet_forward(st, xs[:, :, :64, :64], training=False)
        t0 = time.perf_counter()
        orc.onet_forward(st, xs, training=False)
        dt = time.perf_counter() - t0
    return 0.262144 / dt, dt


def _config(batch, world, mode="bf16", workload="simclutter"):
    """The `config` object of the JSON line; the reference arm reports the SAME object (it times a bounded sample of it)."""
    prec = {"bf16": "bf16 operands, fp32 accumulate", "fp32": "FP32 verification mode (CUDA cores)",
            "tf32": "tf32 tensor-core operands, fp32 storage and accumulate"}[mode]
    return dict(workload=f"Onet(in_chns={CIN}, shared twin) fwd+bwd+JSD+Adam, batch {batch}/GPU of {CIN}x{H}x{W} "
                         + ("K-distributed clutter frames (BASELINE configs[1])" if workload == "simclutter"
                            else "synthetic multispectral patches (BASELINE configs[3] shape)") + "; " + prec,
                per_gpu_batch=batch, global_batch=batch * world, image=[CIN, H, W], parallelism=f"dp{world}",
                l2="working set (~19 GB of activations per step) is far larger than the 126 MB L2; inputs rotate over 4 batches")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 8          # BASELINE.json configs[0]: batch 8 of 1x256x256 frames on the CPU
    ips, sec = cpu_steps(batch, args.steps, args.warmup, threads)
    sample = (f"each step = batch {batch} (BASELINE configs[0]) of the workload's {CIN}x{H}x{W} K-clutter frames, "
              f"{args.warmup} warm-up + {args.steps} timed steps, oracle port on ATen CPU kernels; images/s is per image, the "
              f"GPU arm runs {args.batch} frames per GPU and step")
    line = dict(impl="reference", metric=METRIC, value=ips, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", config=_config(args.batch, max(1, args.gpus), args.mode, args.workload),
                cpu_baseline=dict(value=ips, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=ips, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    _emit(line)


# --------------------------------------------------------------------------------------------------
# per-kernel accounting for the roofline entry
# --------------------------------------------------------------------------------------------------
def _family(name, a):
    """(family, algorithmic FLOPs, shape string) of one C-ABI call from its integer arguments."""
    if name == "onet_conv3x3_fwd":
        fl = 2.0 * 9 * a[3] * a[4] * a[5] * a[6] * a[8]
        return ("conv3x3 fwd/dgrad (tcgen05 implicit GEMM)" if a[16] == 1 else "conv_first fwd (CUDA cores)"), fl, \
            f"{a[6]}->{a[8]} @{a[4]}x{a[5]} N={a[3]}"
    if name == "onet_conv3x3_dgrad_bnred":      # dgrad with the previous layer's BatchNorm-backward reduce in its epilogue
        return "conv3x3 fwd/dgrad (tcgen05 implicit GEMM)", 2.0 * 9 * a[3] * a[4] * a[5] * a[6] * a[8], \
            f"{a[6]}->{a[8]} @{a[4]}x{a[5]} N={a[3]} +bnred"
    if name == "onet_conv3x3_wgrad":
        fl = 2.0 * 9 * a[6] * a[7] * a[8] * a[9] * a[10]
        return ("conv3x3 wgrad (tcgen05 split-K)" if a[13] == 1 else "conv_first wgrad (CUDA cores)"), fl, \
            f"{a[9]}->{a[10]} @{a[7]}x{a[8]} N={a[6]}"
    if name == "onet_convT2x2_fwd":
        return ("up-conv fwd/dgrad (tcgen05)" if a[16] == 1 else "convT_simt"), 2.0 * 4 * a[3] * a[4] * a[5] * a[6] * a[9], \
            f"{a[6]}->{a[9]} @{a[4]}x{a[5]} N={a[3]}"
    if name == "onet_convT2x2_dgrad":
        return ("up-conv fwd/dgrad (tcgen05)" if a[15] == 1 else "convT_simt"), 2.0 * 4 * a[3] * a[4] * a[5] * a[6] * a[8], \
            f"{a[6]}->{a[8]} @{a[4]}x{a[5]} N={a[3]}"
    if name == "onet_convT2x2_wgrad":
        return ("up-conv wgrad (tcgen05)" if a[16] == 1 else "convT_simt"), 2.0 * 4 * a[6] * a[7] * a[8] * a[9] * a[10], \
            f"{a[9]}->{a[10]} @{a[7]}x{a[8]} N={a[6]}"
    return name.replace("onet_", ""), 0.0, ""


def profile_steps(trainer, x, nsteps, record=True, elem_bytes=2.0):
    """Per-call CUDA-event timing of `nsteps` more steps.  Every rank must run the steps (they contain the gradient
    all-reduce); only ranks with record=True keep the events.  Returns (by family, by kernel variant, by variant+shape)."""
    import torch
    from onet_b200 import _lib
    _lib.PROFILE = [] if record else None
    for _ in range(nsteps):
        trainer.step(x)
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    if not record:
        return {}, {}, {}
    detail = os.environ.get("ONET_BENCH_DETAIL")
    if detail:      # per-call dump of the LAST profiled step: name, kernel variant, integer args, ms, TFLOP/s
        per = len(prof) // nsteps
        with open(detail, "w") as f:
            for name, a, e0, e1, kern in prof[-per:]:
                fam_name, fl, shape = _family(name, a)
                ms = e0.elapsed_time(e1)
                ints = [v for v in a if isinstance(v, int) and not isinstance(v, bool) and abs(v) < (1 << 40)]
                f.write(f"{name}\t{kern}\t{fam_name}\t{ms:.4f}\t{(fl / (ms * 1e-3) / 1e12) if fl else 0:.1f}\t{ints}\n")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from roofline_table import work          # algorithmic HBM bytes of every call
    fam, kern_t, shape_t = {}, {}, {}
    for name, a, e0, e1, kern in prof:
        f, fl, shape = _family(name, a)
        ms = e0.elapsed_time(e1)
        ints = [v for v in a if isinstance(v, int) and not isinstance(v, bool) and abs(v) < (1 << 40)]
        desc, _, nbytes = work(name, ints, elem_bytes)
        if fl == 0.0:            # calls that launch several kernels (BatchNorm backward = reduce + apply + parameter gradient)
            kern = name.replace("onet_", "")
            shape = desc
        for table, key in ((fam, f), (kern_t, kern), (shape_t, (kern, shape))):
            d = table.setdefault(key, dict(ms=0.0, flops=0.0, calls=0, bytes=0.0))
            d["ms"] += ms
            d["flops"] += fl
            d["bytes"] += nbytes
            d["calls"] += 1
    for table in (fam, kern_t, shape_t):
        for d in table.values():
            d["ms"] /= nsteps
            d["flops"] /= nsteps
            d["bytes"] /= nsteps
            d["calls"] //= nsteps
    return fam, kern_t, shape_t


# --------------------------------------------------------------------------------------------------
# sub-records of the same JSON line: the other BASELINE configurations, measured in the same driver run
# --------------------------------------------------------------------------------------------------
def _timed_region(fn, steps, dev, world):
    """CUDA-event time (ms, max over ranks) of `steps` calls of fn(i), barrier + synchronize on both sides."""
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def infer_record(dev, rank, world, steps=4, warmup=2, frames_per_gpu=2, size=2048, cpu_leg=True):
    """BASELINE configs[4]: eval-mode inference on 1 x 2048 x 2048 radar frames, `frames_per_gpu` whole frames per rank and
    step (halo tiling only splits a frame when there are fewer frames than ranks; tests/test_infer_gpu.py covers it), masks
    as uint8.  `value`: frames resident in HBM; `e2e`: frames copied from pinned host memory on a copy stream one frame ahead
    of the compute, every mask copied back (1 byte per pixel) while the next frame computes, host-synchronised per step."""
    import torch
    import onet_b200
    from onet_b200 import _lib, synth
    from onet_b200.evaluate import normalize_per_frame
    from onet_b200.infer import TiledPredictor
    torch.manual_seed(1981)
    net = onet_b200.Onet(1, True, True, mode="bf16").to(dev)
    pred = TiledPredictor.for_onet(net, tile=size, halo=96, max_batch=1)
    f, _ = synth.get_rayleigh_frames(frames_per_gpu, snr=2, img_sz=(size, size), seed=7 + rank, device=dev)
    frames = normalize_per_frame(f.unsqueeze(1))
    host = frames.cpu().pin_memory()
    for _ in range(warmup):
        pred.predict_labels(frames)
        pred.predict_labels(host)
    l0 = _lib.launch_count()
    ms_dev = _timed_region(lambda i: pred.predict_labels(frames), steps, dev, world)
    launches = _lib.launch_count() - l0
    ms_e2e = _timed_region(lambda i: pred.predict_labels(host), steps, dev, world)
    mpix = world * frames_per_gpu * size * size / 1e6
    pk = _peaks()
    roof_mpix = pk["bf16_sustained"] * 1e12 / 2.935e6 / 1e6          # 2.935 MFLOP per pixel
    v = mpix * steps / (ms_dev * 1e-3)
    rec = dict(metric="onet_infer_mpix_per_sec", value=v, unit="Mpix/s", n_gpus=world, steps=steps, warmup=warmup,
               ms_per_step=ms_dev / steps, higher_is_better=True, scaling="weak", dtype="bf16", data="synthetic",
               config=dict(workload=f"Onet eval-mode inference (BASELINE configs[4]), {frames_per_gpu} frames of 1x{size}x{size} "
                                    f"Rayleigh clutter + targets per GPU and step, whole frames per rank, uint8 masks",
                           frames_per_gpu=frames_per_gpu, size=size),
               e2e=dict(value=mpix * steps / (ms_e2e * 1e-3), unit="Mpix/s", ms_per_step=ms_e2e / steps,
                        h2d_bytes_per_step=frames_per_gpu * size * size * 4, d2h_bytes_per_step=frames_per_gpu * size * size),
               gpu_launches=int(launches),
               roofline=dict(bound="tensor", achieved=v / world * 2.935e6 / 1e6, peak=pk["bf16_sustained"], unit="TFLOP/s",
                             frac=v / world / roof_mpix, roofline_mpix_per_gpu=roof_mpix,
                             peak_source=pk["src"] + ", sustained bf16 figure (whole forward pass, 2.935 MFLOP per pixel)"))
    if cpu_leg and rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        mpix_cpu, dt = cpu_infer_baseline(threads)
        rec["cpu_baseline"] = dict(value=mpix_cpu, unit="Mpix/s", cores=threads, kind="port",
                                   sample=f"one 1x512x512 frame, eval-mode oracle forward ({dt:.2f} s)")
    del pred, net, frames, f
    torch.cuda.empty_cache()
    return rec


def train_record(dev, rank, world, cin, hw, batch, mode, steps, warmup, graph, workload):
    """One more training configuration through the same trainer: (value, e2e) images/s, weak scaling."""
    import torch
    import onet_b200
    from onet_b200.data import k_clutter_frames
    from onet_b200.trainer import OnetTrainer
    torch.manual_seed(1981)
    net = onet_b200.Onet(cin, True, True, mode=mode).to(dev)
    tr = OnetTrainer(net, lr=LR, graph=graph)
    tr.broadcast_parameters(0)
    host = [k_clutter_frames(batch, cin, hw, hw, seed=4000 + 97 * rank + i, n_targets=8).pin_memory() for i in range(2)]
    res = [h.to(dev) for h in host]
    for i in range(warmup):
        tr.step(res[i % 2])
    ms_dev = _timed_region(lambda i: tr.step(res[i % 2]), steps, dev, world)
    loss = []
    tr.step(host[0]).item()
    ms_e2e = _timed_region(lambda i: loss.append(tr.step(host[i % 2]).item()), steps, dev, world)
    imgs = world * batch * steps
    rec = dict(metric=METRIC, value=imgs / (ms_dev * 1e-3), unit=UNIT, n_gpus=world, steps=steps, warmup=warmup,
               ms_per_step=ms_dev / steps, higher_is_better=True, scaling="weak", dtype={"bf16": "bf16", "fp32": "f32", "tf32": "tf32"}[mode],
               data="synthetic", config=dict(workload=workload, per_gpu_batch=batch, global_batch=batch * world, image=[cin, hw, hw],
                                             parallelism=f"dp{world}", graph=bool(graph)),
               e2e=dict(value=imgs / (ms_e2e * 1e-3), unit=UNIT, ms_per_step=ms_e2e / steps, h2d_bytes_per_step=batch * cin * hw * hw * 4,
                        d2h_bytes_per_step=4),
               peak_memory_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30, loss_last=loss[-1] if loss else None)
    del tr, net, res
    torch.cuda.empty_cache()
    return rec


_REAL_STDOUT = None


def _claim_stdout():
    """Keep fd 1 for the ONE JSON line: libraries that print to stdout (NCCL prints its version there) go to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="simclutter", choices=["simclutter", "zy3"],
                    help="simclutter = BASELINE configs[1] (1x256x256, the headline); zy3 = configs[3] shape (3x224x224 patches)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying one CUDA graph per step")
    ap.add_argument("--no-profile", action="store_true", help="skip the eager per-kernel timing pass (no `roofline` objects)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the sub-records (inference configs[4], ZY-3 shape configs[3], verification modes) of the JSON line")
    args = ap.parse_args()
    global H, W, CIN
    if args.workload == "zy3":
        H = W = 224
        CIN = 3
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import onet_b200
    from onet_b200 import _lib
    from onet_b200.data import k_clutter_frames
    from onet_b200.trainer import OnetTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path for --impl ours)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=4))

    B = args.batch
    torch.manual_seed(1981)
    net = onet_b200.Onet(CIN, True, True, mode=args.mode).to(dev)
    trainer = OnetTrainer(net, lr=LR, graph=not args.no_graph)
    trainer.broadcast_parameters(0)

    POOL = 4
    if args.workload == "zy3":     # synthetic multispectral patches: smooth texture + noise in [0,1] (values do not affect throughput)
        host = [k_clutter_frames(B, CIN, H, W, seed=1981 + 97 * rank + i, n_targets=40, nu=2.0).pin_memory() for i in range(POOL)]
    else:
        host = [k_clutter_frames(B, CIN, H, W, seed=1981 + 97 * rank + i, n_targets=8).pin_memory() for i in range(POOL)]
    resident = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        trainer.step(resident[i % POOL])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms_dev = timed(lambda i: trainer.step(resident[i % POOL]), args.steps)
    launches = _lib.launch_count() - l0
    if trainer.use_graph:       # replays do not pass through the library's counter: kernels captured per step x steps
        launches = trainer.launches_per_step * args.steps

    losses = []

    def e2e_step(i):      # pinned host batch -> device inside the step, loss read back on the host every step
        losses.append(trainer.step(host[i % POOL]).item())
    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    peak_mem_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30

    # per-kernel pass (eager, events around every call).  It allocates a second set of activations next to the captured graph's
    # private pool: skipped when that would not fit (e.g. 256 frames per GPU) - the line then carries no per-kernel roofline
    free_b, total_b = torch.cuda.mem_get_info(dev)
    if args.no_profile or (trainer.use_graph and peak_mem_gb * 2 ** 30 * 1.2 > free_b):
        fam, kern_t, shape_t = {}, {}, {}
    else:
        fam, kern_t, shape_t = profile_steps(trainer, resident[0], 2, record=(rank == 0), elem_bytes=2.0 if args.mode == "bf16" else 4.0)
    if world > 1:
        dist.barrier()

    extra = {}
    if not args.no_extra and args.workload == "simclutter" and args.mode == "bf16":
        # the other BASELINE configurations in the same driver run (every rank takes part: they contain collectives)
        extra["infer"] = infer_record(dev, rank, world, steps=4, warmup=2, cpu_leg=not args.no_cpu_baseline)
        H, W, CIN = 224, 224, 3
        extra["zy3"] = train_record(dev, rank, world, 3, 224, B, "bf16", min(args.steps, 10), 3, not args.no_graph,
                                    _config(B, world, "bf16", "zy3")["workload"])
        H, W, CIN = 256, 256, 1
        # the modes that meet the north_star tolerances literally, beside the bf16 throughput mode (same step, same shape)
        extra["modes"] = {}
        for mode, mb in EXTRA_MODES:
            extra["modes"][mode] = train_record(dev, rank, world, 1, 256, mb, mode, 2, 1, False, _config(mb, world, mode)["workload"])

    if rank == 0:
        pk = _peaks()
        value = world * B * args.steps / (ms_dev / 1e3)
        e2e_v = world * B * args.steps / (ms_e2e / 1e3)
        total_ms = sum(d["ms"] for d in fam.values()) or 1.0
        # dominant kernel = the kernel variant (one __global__ function) with the most time in the step; its launches
        # have different layer shapes, so achieved = sum of algorithmic FLOPs / sum of CUDA-event times = the
        # per-launch average of both.
        gemm = {k: d for k, d in kern_t.items() if d["flops"] > 0}
        top = max(gemm, key=lambda k: gemm[k]["ms"]) if gemm else None
        roof = None
        if top:
            d = gemm[top]
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            shapes = {k[1]: v for k, v in shape_t.items() if k[0] == top}
            # the per-kernel times come from an eager pass with CUDA events around every launch: each kernel runs alone,
            # at burst clocks (their sum is below the replayed step) - so the denominator is the BURST figure
            roof = dict(bound="tensor", kernel=top, achieved=ach, peak=pk["bf16_burst"], unit="TFLOP/s",
                        frac=ach / pk["bf16_burst"], frac_of_sustained_peak=ach / pk["bf16_sustained"],
                        traffic=None,
                        peak_source=pk["src"] + ", burst bf16 figure (kernel timed alone between CUDA events; sustained figure "
                                                f"{pk['bf16_sustained']}, which the whole-step `step_tflops` is held against)",
                        share_of_step=d["ms"] / total_ms, launches_per_step=d["calls"],
                        flops_per_launch=d["flops"] / max(d["calls"], 1), ms_per_launch=d["ms"] / max(d["calls"], 1),
                        shapes={sh: dict(calls=v["calls"], ms_per_launch=round(v["ms"] / max(v["calls"], 1), 4),
                                         tflops=round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)) for sh, v in
                                sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])})
            # DRAM bytes per launch from the committed `ncu --set full` captures (profiles/): the capture of this kernel
            caps = []
            for fn in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
                if "ncu_traffic" in fn and fn.endswith(".json"):
                    caps = json.load(open(os.path.join(ROOT, "profiles", fn)))      # the newest round's file wins
                    cap_file = fn
            def _norm(k):        # "conv3x3_halo2_px_kernel<256, 0, OpBf16>" (ncu) and "conv3x3_halo2_px_kernel<256>" (library) -> same key
                base, _, targs = k.replace(" ", "").partition("<")
                first = targs.split(",")[0].rstrip(">")
                return base.replace("/tf32", ""), (first if first.isdigit() else ""), ("Tf32" in k or "/tf32" in k)
            for c in caps:
                if _norm(c["kernel"]) == _norm(top):
                    roof["traffic"] = c["dram_bytes"]
                    roof["traffic_capture"] = dict(file="profiles/" + cap_file, capture=c["capture"],
                                                   algorithmic_bytes=c["algorithmic_bytes"], duration_us=c["duration_us"],
                                                   tcgen05_pct_of_peak=c.get("tcgen05_pct_of_peak"),
                                                   note="bytes of ONE launch of the captured layer shape "
                                                        "(N_H_W_Cin_Cout in `capture`), cold cache")
                    break
            roof["ncu_captures"] = caps
            famd = fam.get("conv3x3 fwd/dgrad (tcgen05 implicit GEMM)")
            if famd:
                roof["conv_fwd_dgrad_family"] = dict(tflops=famd["flops"] / (famd["ms"] * 1e-3) / 1e12, ms_per_step=famd["ms"],
                                                     launches_per_step=famd["calls"], share_of_step=famd["ms"] / total_ms)
        # largest HBM-bound call family of the step (BatchNorm backward = reduce pass + apply pass + parameter gradients)
        hbm = {k: d for k, d in kern_t.items() if d["flops"] == 0 and d["bytes"] > 0}
        roof_hbm = None
        if hbm:
            topb = max(hbm, key=lambda k: hbm[k]["ms"])
            d = hbm[topb]
            achb = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            shapes = {k[1]: v for k, v in shape_t.items() if k[0] == topb}
            roof_hbm = dict(bound="hbm", kernel=topb, achieved=achb, peak=pk["hbm"], unit="GB/s", frac=achb / pk["hbm"], traffic=None,
                            peak_source=pk["src"] + ", copy bandwidth", share_of_step=d["ms"] / total_ms, launches_per_step=d["calls"],
                            bytes_per_launch=d["bytes"] / max(d["calls"], 1), ms_per_launch=d["ms"] / max(d["calls"], 1),
                            note="algorithmic bytes: every operand read once per pass and the result written once "
                                 "(tools/roofline_table.py::work)",
                            shapes={sh: dict(calls=v["calls"], ms_per_launch=round(v["ms"] / max(v["calls"], 1), 4),
                                             gbs=round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 0)) for sh, v in
                                    sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])})
            for c in (roof or {}).get("ncu_captures", []):
                if c.get("family") == topb:
                    roof_hbm["traffic"] = c["dram_bytes"]
                    roof_hbm["traffic_capture"] = dict(capture=c["capture"], algorithmic_bytes=c["algorithmic_bytes"],
                                                       duration_us=c["duration_us"], kernel=c["kernel"])
                    break
        per_kernel = {k: dict(ms_per_step=round(d["ms"], 3), share=round(d["ms"] / total_ms, 4), calls=d["calls"],
                              tflops=(round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1) if d["flops"] else None))
                      for k, d in sorted(kern_t.items(), key=lambda kv: -kv[1]["ms"])}
        kernels = {k: dict(ms_per_step=round(d["ms"], 3), share=round(d["ms"] / total_ms, 4), calls=d["calls"],
                           tflops=(round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1) if d["flops"] else None))
                   for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        train_flops = sum(d["flops"] for d in fam.values())
        if not fam:          # no per-kernel pass: the algorithmic figure of SURVEY.md section 8d (per image, twin network)
            train_flops = (576.90e9 if args.workload == "simclutter" else 442.15e9) * B
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={"bf16": "bf16", "fp32": "f32", "tf32": "tf32"}[args.mode], data="synthetic",
                    config=_config(B, world, args.mode, args.workload),
                    e2e=dict(value=e2e_v, unit=UNIT, ms_per_step=ms_e2e / args.steps, h2d_bytes_per_step=B * CIN * H * W * 4,
                             d2h_bytes_per_step=4),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, roofline_hbm=roof_hbm, kernels=kernels,
                    per_kernel=per_kernel,
                    step_tflops=train_flops / (ms_dev / args.steps * 1e-3) / 1e12,
                    step_frac_of_sustained_peak=train_flops / (ms_dev / args.steps * 1e-3) / 1e12 / pk["bf16_sustained"],
                    peak_memory_gb=peak_mem_gb, loss_last=losses[-1] if losses else None)
        line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ips, sec = cpu_steps(8, 2, 1, threads)
            line["cpu_baseline"] = dict(value=ips, unit=UNIT, cores=threads, kind="port",
                                        sample=f"batch 8 (BASELINE configs[0]) of the same {CIN}x{H}x{W} frames, 1 warm-up + 2 timed "
                                               f"steps ({sec:.1f} s/step), oracle port on ATen CPU kernels")
        _emit(line)
    if world > 1:
        # a captured graph holds NCCL work; tear down in order and do not rely on interpreter-exit destructors
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
