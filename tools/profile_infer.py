"""Per-call CUDA-event times of one eval-mode inference pass over a 1 x S x S frame (BASELINE configs[4]):

    python tools/profile_infer.py [S=2048]

Prints every C-ABI call of the pass (name, kernel variant, ms, integer arguments) and the totals per call name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    import onet_b200
    from onet_b200 import _lib, synth
    from onet_b200.evaluate import normalize_per_frame
    from onet_b200.infer import TiledPredictor
    dev = torch.device("cuda", 0)
    torch.manual_seed(1981)
    net = onet_b200.Onet(1, True, True, mode="bf16").to(dev)
    pred = TiledPredictor.for_onet(net, tile=size, halo=96, max_batch=1)
    f, _ = synth.get_rayleigh_frames(1, snr=2, img_sz=(size, size), seed=7, device=dev)
    frames = normalize_per_frame(f.unsqueeze(1))
    for _ in range(2):
        pred.predict_labels(frames)
    torch.cuda.synchronize()
    _lib.PROFILE = []
    pred.predict_labels(frames)
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    tot = {}
    total = 0.0
    for name, a, e0, e1, kern in prof:
        ms = e0.elapsed_time(e1)
        ints = [v for v in a if isinstance(v, int) and not isinstance(v, bool) and abs(v) < (1 << 40)]
        print(f"{name}\t{kern}\t{ms:.4f}\t{ints}")
        d = tot.setdefault(name, [0.0, 0])
        d[0] += ms
        d[1] += 1
        total += ms
    print(f"# total {total:.3f} ms for {size * size / 1e6:.2f} Mpix = {size * size / 1e6 / (total * 1e-3):.1f} Mpix/s (serial sum of the calls)")
    for name, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"# {name:34s} {ms:8.3f} ms  {n:3d} calls")


if __name__ == "__main__":
    main()
