"""CPU emulation: which bf16 roundings hurt the gradient? (scratch experiment)"""
import sys, itertools
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
from oracle import onet_oracle as orc
torch.set_num_threads(8)

R = lambda t: t.bfloat16().float()
class RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rg):
        ctx.rg = rg
        return R(x)
    @staticmethod
    def backward(ctx, g):
        return (R(g) if ctx.rg else g), None

def run(st, x, q_op, q_y, q_a, q_g):
    st = {k: v.clone() for k, v in st.items()}
    leaves = [k for k, v in st.items() if v.dtype.is_floating_point and 'running' not in k]
    for k in leaves: st[k].requires_grad_(True)
    def bn(prefix, t):
        mean = t.mean(dim=(0,2,3)); var = t.var(dim=(0,2,3), unbiased=False)
        inv = torch.rsqrt(var + 1e-5)
        return (t - mean[None,:,None,None]) * (inv*st[prefix+'.weight'])[None,:,None,None] + st[prefix+'.bias'][None,:,None,None]
    def dc(block, t):
        p = orc._dc_prefix(block)
        for ci, bi in ((0,1),(3,4)):
            w = st[f'{p}.{ci}.weight']
            if q_op: w = RoundSTE.apply(w, False)
            t = F.conv2d(t, w, None, padding=1)
            if q_y: t = RoundSTE.apply(t, q_g)
            t = torch.relu(bn(f'{p}.{bi}', t))
            if q_a: t = RoundSTE.apply(t, q_g)
        return t
    def unet(t):
        if q_a: t = R(t)
        x1 = dc('inc', t); skips=[x1]; h=x1
        for name,_,_ in orc.ENCODER:
            h = F.max_pool2d(h,2); h = dc(name,h); skips.append(h)
        y = skips[-1]
        for i,(name,_,_) in enumerate(orc.DECODER):
            w = st[f'{name}.up.weight']
            if q_op: w = RoundSTE.apply(w, False)
            u = F.conv_transpose2d(y, w, st[f'{name}.up.bias'], stride=2)
            if q_a: u = RoundSTE.apply(u, q_g)
            y = dc(name, torch.cat([skips[3-i], u], 1))
        return x1, y
    Lt, Ht = unet(x); Ld, Hd = unet(torch.clip(1-x,0,1))
    Vt = (Lt*Ht).sum(1, keepdim=True); Vd = (Ld*Hd).sum(1, keepdim=True)
    S = torch.softmax(torch.cat([Vt,Vd],1),1)
    loss = orc.compute_loss(Lt, S[:,0:1], Ld, S[:,1:2]); loss.backward()
    return loss.item(), {k: st[k].grad.clone() for k in leaves}, Vt.detach(), S.detach(), Lt.detach(), Ld.detach()

B,H = int(sys.argv[1]), int(sys.argv[2])
st = orc.perturb_bn_affine(orc.init_state(1, seed=31), seed=131)
x = orc.rayleigh_frames(B,1,H,H,seed=31)
l0,g0,V0,S0,Lt0,Ld0 = run(st,x,False,False,False,False)
a0 = Lt0.sum(1); b0 = Ld0.sum(1)
print('a mean', a0.mean().item(), 'b mean', b0.mean().item(), '|a-b| mean', (a0-b0).abs().mean().item())
import numpy as np
for name,cfg in [('ops only',(True,False,True,False)),('fwd storage all',(True,True,True,False)),('fwd+grad storage',(True,True,True,True))]:
    l,g,V,S,Lt,Ld = run(st,x,*cfg)
    errs = [float((g[k]-g0[k]).norm()/g0[k].norm()) for k in g0]
    print(name, 'loss rel %.2e'%(abs(l-l0)/abs(l0)), 'Vt %.2e'%float((V-V0).norm()/V0.norm()), 'S %.2e'%float((S-S0).norm()/S0.norm()),
          'Lt %.2e'%float((Lt-Lt0).norm()/Lt0.norm()), 'grad median %.2e max %.2e'%(np.median(errs), max(errs)))

print('--- conditioning: perturb x by relative eps (fp32 everywhere)')
for eps in (1e-6, 1e-4, 1e-3):
    xp = x * (1 + eps * torch.randn_like(x))
    l,g,V,S,Lt,Ld = run(st,xp,False,False,False,False)
    errs = {k: float((g[k]-g0[k]).norm()/g0[k].norm()) for k in g0}
    ks = sorted(errs, key=errs.get)
    print('eps', eps, 'Vt %.2e'%float((V-V0).norm()/V0.norm()), 'S %.2e'%float((S-S0).norm()/S0.norm()), 'grad median %.2e max %.2e'%(np.median(list(errs.values())), max(errs.values())), 'best', ks[0], '%.1e'%errs[ks[0]], 'worst', ks[-1], '%.1e'%errs[ks[-1]])
