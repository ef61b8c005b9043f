"""Does a weight-gradient kernel really run NEXT TO a BatchNorm-backward kernel?  Times the pair serially on one stream
and forked onto two streams (both host launch orders), per layer shape of the training step.

    python tools/overlap_probe.py [N=128]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from profile_layer import make_runner  # noqa: E402


def timed(fn, flush, iters=5):
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream(priority=-1)
    for s, c, tensor_kind, bn_kind in ((256, 64, "wgrad", "bnbwd"), (128, 128, "wgrad", "bnbwd"), (64, 256, "wgrad", "bnbwd"),
                                       (32, 512, "wgrad", "bnbwd"), (256, 64, "wgrad", "bnbwd_pool_g2"), (128, 128, "wgrad", "bnbwd_pool"),
                                       (256, 64, "dgrad", "bnbwd"), (64, 256, "dgrad", "bnbwd"), (256, 64, "wgrad", "bnapply")):
        tk, _, _ = make_runner(tensor_kind, n, s, s, c, c)
        bn, _, _ = make_runner(bn_kind, n, s, s, c, c)
        for _ in range(2):
            tk(); bn()
        torch.cuda.synchronize()

        def forked(first_tensor):
            side.wait_stream(main_s)
            if first_tensor:
                with torch.cuda.stream(side):
                    tk()
                bn()
            else:
                bn()
                with torch.cuda.stream(side):
                    tk()
            main_s.wait_stream(side)

        t_t, t_b = timed(tk, flush), timed(bn, flush)
        t_ser = timed(lambda: (tk(), bn()), flush)
        t_f1 = timed(lambda: forked(True), flush)
        t_f2 = timed(lambda: forked(False), flush)
        print(f"{tensor_kind}+{bn_kind} {s}x{s} C={c}: tensor {t_t:.3f}  bn {t_b:.3f}  serial {t_ser:.3f}  forked(tensor first) {t_f1:.3f}  "
              f"forked(bn first) {t_f2:.3f}  ms", flush=True)


if __name__ == "__main__":
    main()
