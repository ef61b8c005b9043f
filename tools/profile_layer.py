"""Run ONE tensor-core layer of the Onet path repeatedly (for ncu captures and quick timing).

    python tools/profile_layer.py fwd   N H W Cin Cout [iters]
    python tools/profile_layer.py wgrad N H W Cin Cout [iters]
Prints the CUDA-event time per launch and the achieved TFLOP/s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
import gpu_util as U  # noqa: E402


def main():
    kind = sys.argv[1]
    n, h, w, cin, cout = (int(v) for v in sys.argv[2:7])
    iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
    torch.manual_seed(0)
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    wf, wd = U.pack_conv(wt, U.BF16)
    gy = torch.randn(n, h, w, cout, device="cuda").to(torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flops = 2.0 * 9 * n * h * w * cin * cout

    def run():
        if kind == "fwd":
            U.conv3x3(x, wf, cout, U.BF16, U.ENGINE_TC, group_images=n // 2, stats=True)
        else:
            U.conv3x3_wgrad(gy, x, U.BF16, U.ENGINE_TC)

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
    print(f"{kind} N={n} {h}x{w} {cin}->{cout}: {ms:.3f} ms  {flops / (ms * 1e-3) / 1e12:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
