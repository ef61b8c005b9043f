"""Two small training steps in the given mode, for compute-sanitizer runs:

    compute-sanitizer --tool memcheck python tools/sanitize_step.py bf16 2 256 256
    compute-sanitizer --tool racecheck python tools/sanitize_step.py bf16 2 64 64

Arguments: mode (fp32 | bf16 | tf32), frames B, height, width.  B=2 at 256 x 256 reaches every kernel variant of the benchmark
step: CTA pairs (cta_group::2) for Cin >= 128, the weight-resident kernel (>= 4 tiles per SM with Cin = 64), the CTA-pair weight
gradient (Cout >= 256), the fused BatchNorm-backward reduce (C >= 256), the pooled BatchNorm backward and the two-stream backward."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import onet_b200  # noqa: E402
from onet_b200 import _lib  # noqa: E402
from onet_b200.data import rayleigh_target_frames  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 32, 32)
torch.manual_seed(3)
net = onet_b200.Onet(1, True, True, mode=mode).cuda()
x = rayleigh_target_frames(B, 1, H, W, seed=5).cuda()
if os.environ.get("ONET_TRACE_CALLS"):          # synchronise after every C-ABI call and name the first one that faults
    _orig = _lib.call

    def traced(name, *args, **kw):
        _orig(name, *args, **kw)
        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            ints = [a for a in args if isinstance(a, int) and abs(a) < (1 << 40)]
            print(f"FAULT after {name} ({_lib.lib().onet_last_kernel().decode()}) ints={ints}: {str(e)[:120]}", flush=True)
            os._exit(3)
    import onet_b200.model as _m
    _lib.call = traced
    _m.call = traced
    os.environ["ONET_NO_WGRAD_OVERLAP"] = "1"
if os.environ.get("ONET_TRACE_JOINS"):          # two-stream schedule kept; synchronise at every join and list the calls since the last one
    import onet_b200.model as _m
    _orig2, _recent = _lib.call, []

    def logged(name, *args, **kw):
        _recent.append((name, [a for a in args if isinstance(a, int) and abs(a) < (1 << 40)]))
        _orig2(name, *args, **kw)
    _lib.call = logged
    _m.call = logged
    _join = _m._Engine._join_side

    def join(self):
        _join(self)
        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"FAULT at a join: {str(e)[:100]}; calls since the previous join:", flush=True)
            for c in _recent:
                print("   ", c, flush=True)
            os._exit(3)
        del _recent[:]
    _m._Engine._join_side = join
for rep in range(2):
    net.train()
    net.zero_grad()
    Lt, Vt, Ld, Vd, S = net(x)
    loss = net.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
    loss.backward()
    torch.cuda.synchronize()
    g = torch.cat([p.grad.flatten() for p in net.parameters()])
    print(f"rep {rep}: mode {mode} B={B} {H}x{W} loss {loss.item():.8f} grad norm {float(g.double().norm()):.10e} "
          f"kernels launched {_lib.launch_count()}", flush=True)
