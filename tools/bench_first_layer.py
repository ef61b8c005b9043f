"""Times the first-convolution calls of a 1-channel network (closed-form statistics, conv + BN + ReLU, single-pass backward) at the
training shape (128 twin images of 256 x 256) and at the inference shape (4 twin frames of 2048 x 2048):

    python tools/bench_first_layer.py            # one line per call: ms, GB/s of algorithmic traffic
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    import gpu_util as U
    call, ptr = U.call, U.ptr
    dt, tdt = U.BF16, torch.bfloat16
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for n, h, w in ((128, 256, 256), (4, 2048, 2048)):
        g = n // 2
        torch.manual_seed(3)
        x = torch.rand(n, h, w, 1, device="cuda").to(tdt)
        wt = (torch.randn(64, 1, 3, 3, device="cuda") / 3).to(tdt).float()
        wf, _ = U.pack_conv(wt, dt)
        gn = torch.randn(n, h, w, 64, device="cuda").to(tdt)
        aff = torch.randn(4, 2, 64, device="cuda")
        aff[1] = aff[1].abs() + 0.5
        gram = torch.zeros(2, 90, dtype=torch.float64, device="cuda")
        st = torch.zeros(2, 2, 64, dtype=torch.float64, device="cuda")
        act = torch.empty(n, h, w, 64, dtype=tdt, device="cuda")
        acc_a = torch.zeros(2, 64, 9, device="cuda")
        sums = torch.zeros(2, 2, 64, dtype=torch.float64, device="cuda")
        dw = torch.zeros(64, 1, 3, 3, device="cuda")
        dg, db = torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda")
        px = n * h * w
        calls = {
            "first_conv_stats": (lambda: call("onet_first_conv_stats", ptr(x), n, h, w, 1, ptr(wf), ptr(gram), ptr(st[0]), ptr(st[1]), g, dt, U.stream()), 2 * px),
            "first_conv_bn_relu": (lambda: call("onet_first_conv_bn_relu", ptr(x), n, h, w, 1, ptr(wf), ptr(aff[2]), ptr(aff[3]), g, ptr(act), 0, dt, U.stream()), 130 * px),
            "first_conv_bwd": (lambda: call("onet_first_conv_bwd", ptr(x), n, h, w, 1, ptr(wf), ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]), g, ptr(gn),
                                            ptr(gram), ptr(acc_a), ptr(sums), float(g * h * w), ptr(dw), ptr(dg), ptr(db), ptr(dg), ptr(db), dt, U.stream()), 130 * px),
        }
        for name, (fn, byts) in calls.items():
            for _ in range(2):
                fn()
            ms = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            m = sorted(ms)[len(ms) // 2]
            print(f"{name} N={n} {h}x{w}: {m:.3f} ms  {byts / m / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
