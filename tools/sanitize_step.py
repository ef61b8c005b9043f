"""One small training step (B=2, 32x32) in the given mode, for compute-sanitizer runs:
    compute-sanitizer --tool initcheck python tools/sanitize_step.py fp32"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import onet_b200
from onet_b200.data import rayleigh_target_frames

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
torch.manual_seed(3)
net = onet_b200.Onet(1, True, True, mode=mode).cuda()
x = rayleigh_target_frames(2, 1, 32, 32, seed=5).cuda()
for rep in range(2):
    net.train()
    net.zero_grad()
    Lt, Vt, Ld, Vd, S = net(x)
    loss = net.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
    loss.backward()
    torch.cuda.synchronize()
    g = torch.cat([p.grad.flatten() for p in net.parameters()])
    print(f"rep {rep}: loss {loss.item():.8f} grad norm {float(g.double().norm()):.10e}", flush=True)
