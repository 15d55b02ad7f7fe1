import torch
from neosr_b200 import ops
from neosr_b200.archs import build_network
from oracle.unet import synth_unet, unet_forward
nf=64; skip=False
p,b=synth_unet(num_feat=nf,seed=31)
net=build_network({"type":"unet","num_feat":nf,"skip_connection":skip}); net.load_state_dict({**p,**b}); net=net.cuda().train()
gen=torch.Generator().manual_seed(32)
xs=[torch.rand(2,3,32,48,generator=gen) for _ in range(3)]; ts=[torch.randn(2,1,32,48,generator=gen) for _ in range(3)]; x,t=xs[0],ts[0]
res={}
for dt in (torch.float32, torch.float64):
    pr={k:v.clone().to(dt) for k,v in p.items()}; bo={k:v.clone().to(dt) for k,v in b.items()}
    xo=x.clone().to(dt).requires_grad_(True)
    yo=unet_forward(pr,bo,xo,True,skip)
    gx,=torch.autograd.grad(((yo-t.to(dt))**2).mean(),xo)
    res[dt]=gx.double()
ops.DEFAULT_ENGINE="simt"
y,S=net.engine_forward(x.cuda(),save=True)
dy=(2.0/y.numel())*(y-t.cuda())
dx=net.engine_backward(S,dy,param_grads=False).cpu().double()
def r2(a,b): return float((a-b).norm()/b.norm())
print("gpu-vs-cpu32",r2(dx,res[torch.float32]),"gpu-vs-cpu64",r2(dx,res[torch.float64]),"cpu32-vs-cpu64",r2(res[torch.float32],res[torch.float64]))
