"""Run ONE kernel family of the Onet path repeatedly (for ncu captures and quick timing).

    python tools/profile_layer.py KIND N H W Cin Cout [iters]
KIND: fwd (conv + BN partial statistics) | dgrad (conv, no statistics) | wgrad | fwd_tf32 | dgrad_tf32 | wgrad_tf32 |
      convT | convT_dgrad | convT_wgrad |
      bnapply | bnapply_pool | bnbwd | bnbwd_pool | bnbwd_pool_g2   (the BN kinds use C = Cout; Cin is ignored)
Prints the CUDA-event time per launch and the achieved TFLOP/s (convolutions) or GB/s of algorithmic bytes (BN kinds)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
import gpu_util as U  # noqa: E402


def make_runner(kind, n, h, w, cin, cout):
    """Returns (run, flops, bytes): a closure launching the kernel(s) and the algorithmic work of one call."""
    call, ptr, st = U.call, U.ptr, U.stream
    bf = torch.bfloat16
    dev = "cuda"
    if kind in ("first_fwd", "first_wgrad"):          # first layer of the U-Net (Cin = 1 or 3): CUDA-core direct kernels
        x = torch.rand(n, h, w, cin, device=dev).to(bf)
        wt = torch.randn(cout, cin, 3, 3, device=dev) * 0.3
        wf, _ = U.pack_conv(wt, U.BF16)
        gy = torch.randn(n, h, w, cout, device=dev).to(bf)
        byts = 2.0 * n * h * w * (cin + cout)
        if kind == "first_fwd":
            return (lambda: U.conv3x3(x, wf, cout, U.BF16, U.ENGINE_SIMT, group_images=max(n // 2, 1), stats=True)), 0.0, byts
        return (lambda: U.conv3x3_wgrad(gy, x, U.BF16, U.ENGINE_SIMT)), 0.0, byts
    if kind in ("fwd_tf32", "dgrad_tf32", "wgrad_tf32"):        # tcgen05 kind::tf32 kernels: fp32 storage (C-ABI: ONET_F32 + ONET_ENGINE_TC)
        x = torch.randn(n, h, w, cin, device=dev)
        wt = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
        wf, wd = U.pack_conv(wt, U.F32)
        gy = torch.randn(n, h, w, cout, device=dev)
        flops = 2.0 * 9 * n * h * w * cin * cout
        byts = 4.0 * n * h * w * (cin + cout)
        if kind == "fwd_tf32":
            return (lambda: U.conv3x3(x, wf, cout, U.F32, U.ENGINE_TC, group_images=max(n // 2, 1), stats=True)), flops, byts
        if kind == "dgrad_tf32":
            return (lambda: U.conv3x3(x, wf, cout, U.F32, U.ENGINE_TC, stats=False)), flops, byts
        return (lambda: U.conv3x3_wgrad(gy, x, U.F32, U.ENGINE_TC)), flops, byts
    if kind in ("fwd", "dgrad", "wgrad"):
        x = torch.randn(n, h, w, cin, device=dev).to(bf)
        wt = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
        wf, wd = U.pack_conv(wt, U.BF16)
        gy = torch.randn(n, h, w, cout, device=dev).to(bf)
        flops = 2.0 * 9 * n * h * w * cin * cout
        byts = 2.0 * n * h * w * (cin + cout)
        if kind == "fwd":
            return (lambda: U.conv3x3(x, wf, cout, U.BF16, U.ENGINE_TC, group_images=max(n // 2, 1), stats=True)), flops, byts
        if kind == "dgrad":
            return (lambda: U.conv3x3(x, wf, cout, U.BF16, U.ENGINE_TC, stats=False)), flops, byts
        return (lambda: U.conv3x3_wgrad(gy, x, U.BF16, U.ENGINE_TC)), flops, byts
    if kind in ("convT", "convT_dgrad", "convT_wgrad"):
        co = cin // 2
        x = torch.randn(n, h, w, cin, device=dev).to(bf)
        wt = torch.randn(cin, co, 2, 2, device=dev) * 0.05
        bias = torch.randn(co, device=dev)
        wf, wd = U.pack_convT(wt, U.BF16)
        cat = torch.zeros(n, 2 * h, 2 * w, 2 * co, device=dev, dtype=bf)
        dx = torch.empty(n, h, w, cin, device=dev, dtype=bf)
        dw = torch.zeros(cin, co, 2, 2, device=dev)
        db = torch.zeros(co, device=dev)
        flops = 2.0 * 4 * n * h * w * cin * co
        byts = 2.0 * n * h * w * (cin + 4 * co)
        if kind == "convT":
            return (lambda: call("onet_convT2x2_fwd", ptr(x), cin, 0, n, h, w, cin, ptr(wf), ptr(bias), co, ptr(cat, co), 2 * co,
                                 0, 0, 0, U.BF16, U.ENGINE_TC, st())), flops, byts
        if kind == "convT_dgrad":
            return (lambda: call("onet_convT2x2_dgrad", ptr(cat, co), 2 * co, 0, n, h, w, cin, ptr(wd), co, ptr(dx), cin, 0,
                                 0, 0, U.BF16, U.ENGINE_TC, st())), flops, byts
        return (lambda: call("onet_convT2x2_wgrad", ptr(x), cin, 0, ptr(cat, co), 2 * co, 0, n, h, w, cin, co, ptr(dw), ptr(db),
                             0, 0, U.BF16, U.ENGINE_TC, st())), flops, byts
    # ---- BatchNorm kinds
    c = cout
    g = max(n // 2, 1)
    y = (torch.randn(n, h, w, c, device=dev) * 1.5 + 0.3).to(bf)
    aff = torch.empty(4, 2, c, device=dev)
    aff[0].fill_(0.3); aff[1].fill_(0.66); aff[2].fill_(0.66); aff[3].fill_(-0.2)
    pool = kind.endswith("_pool") or kind.endswith("_pool_g2")
    elems = float(n) * h * w * c
    if kind.startswith("bnapply"):
        out = torch.empty(n, h, w, 2 * c, device=dev, dtype=bf) if pool else torch.empty(n, h, w, c, device=dev, dtype=bf)
        ldo = out.shape[-1]
        pl = torch.empty(n, h // 2, w // 2, c, device=dev, dtype=bf) if pool else None
        byts = elems * (2 + 2) + (elems / 4 * 2 if pool else 0)
        parg = torch.empty(n, h // 2, w // 2, c // 8, device=dev, dtype=torch.int16) if pool else None
        return (lambda: call("onet_bn_relu_apply", ptr(y), n, h, w, c, ptr(aff[2]), ptr(aff[3]), g, ptr(out), ldo, 0, ptr(pl),
                             ptr(parg), U.BF16, st())), 0.0, byts
    g1 = torch.randn(n, h, w, 2 * c if pool else c, device=dev).to(bf)
    ld1 = g1.shape[-1]
    g2 = torch.randn(n, h, w, c, device=dev).to(bf) if kind.endswith("_g2") else None
    gp = torch.randn(n, h // 2, w // 2, c, device=dev).to(bf) if pool else None
    sums = torch.zeros(2, 2, c, dtype=torch.float64, device=dev)
    dy = torch.empty(n, h, w, c, device=dev, dtype=bf)
    dgam = torch.zeros(c, device=dev)
    dbet = torch.zeros(c, device=dev)
    count = float(g * h * w)
    parg = torch.randint(0, 1 << 15, (n, h // 2, w // 2, c // 8), device=dev, dtype=torch.int16) if pool else None
    # algorithmic bytes: y + g1 (+ g2) read once each, pooled gradient + argmax bytes once, dy written once
    byts = elems * (2 + 2 + 2) + (elems * 2 if g2 is not None else 0) + (elems / 4 * 2 if pool else 0)

    def run():
        sums.zero_()
        call("onet_bn_relu_bwd", ptr(y), n, h, w, c, ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]), g, ptr(g1), ld1, 0,
             ptr(g2), c, 0, ptr(gp), ptr(parg), ptr(sums), count, ptr(dy), ptr(dgam), ptr(dbet), ptr(dgam), ptr(dbet), U.BF16, st())
    return run, 0.0, byts


def time_kind(kind, n, h, w, cin, cout, iters, flush):
    run, flops, byts = make_runner(kind, n, h, w, cin, cout)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    rate = f"{flops / (ms * 1e-3) / 1e12:.1f} TFLOP/s" if flops else f"{byts / (ms * 1e-3) / 1e9:.0f} GB/s (algorithmic)"
    print(f"{kind} N={n} {h}x{w} {cin}->{cout}: {ms:.3f} ms  {rate}", flush=True)
    return ms


def main():
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if sys.argv[1] == "sweep":      # every distinct shape of one training step at per-GPU batch B (argv[2], default 64)
        B2 = 2 * (int(sys.argv[2]) if len(sys.argv) > 2 else 64)
        chans = [64, 128, 256, 512, 1024]
        for k in range(5):
            s = 256 >> k
            c = chans[k]
            shapes = [(c, c)] + ([(c // 2, c)] if k > 0 else []) + ([(2 * c, c)] if k < 4 else [])
            for ci, co in shapes:
                for kind in ("fwd", "dgrad", "wgrad"):
                    time_kind(kind, B2, s, s, ci, co, 3, flush)
                if ci != co:
                    time_kind("dgrad", B2, s, s, co, ci, 3, flush)
            for kind in ("bnapply", "bnbwd") + (("bnapply_pool", "bnbwd_pool") if k < 4 else ()):
                time_kind(kind, B2, s, s, c, c, 3, flush)
            if k < 4:
                for kind in ("convT", "convT_dgrad", "convT_wgrad"):
                    time_kind(kind, B2, s // 2, s // 2, 2 * c, c, 3, flush)
        time_kind("bnbwd_pool_g2", B2, 256, 256, 64, 64, 3, flush)
        time_kind("first_fwd", B2, 256, 256, 1, 64, 3, flush)
        time_kind("first_wgrad", B2, 256, 256, 1, 64, 3, flush)
        return
    if sys.argv[1] == "sweep_bw":   # only the bandwidth-bound kinds
        B2 = 2 * (int(sys.argv[2]) if len(sys.argv) > 2 else 64)
        chans = [64, 128, 256, 512, 1024]
        for k in range(5):
            s = 256 >> k
            c = chans[k]
            for kind in ("bnapply", "bnbwd") + (("bnapply_pool", "bnbwd_pool") if k < 4 else ()):
                time_kind(kind, B2, s, s, c, c, 3, flush)
            if k < 4:
                for kind in ("convT", "convT_dgrad", "convT_wgrad"):
                    time_kind(kind, B2, s // 2, s // 2, 2 * c, c, 3, flush)
        time_kind("bnbwd_pool_g2", B2, 256, 256, 64, 64, 3, flush)
        time_kind("first_fwd", B2, 256, 256, 1, 64, 3, flush)
        time_kind("first_wgrad", B2, 256, 256, 1, 64, 3, flush)
        return
    kind = sys.argv[1]
    n, h, w, cin, cout = (int(v) for v in sys.argv[2:7])
    iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
    time_kind(kind, n, h, w, cin, cout, iters, flush)


if __name__ == "__main__":
    main()
