import torch
from neosr_b200 import ops
def rel(a,b): return float((a-b).abs().max()/b.abs().max())
g=torch.Generator().manual_seed(0)
for (cin,cout,H,W) in [(64,64,32,48),(128,64,32,48),(256,128,16,24),(512,256,8,12),(1024,512,4,6),(512,256,8,12),(256,128,16,24)]:
    w=(torch.randn(cout,cin,3,3,generator=g)*0.05).cuda()
    pw=ops.PackedWeight(w).refresh()
    dy=torch.randn(2,H,W,cout,generator=g).cuda(); aux=torch.randn(2,H,W,cin,generator=g).cuda()
    x=torch.randn(2,H,W,cin,generator=g).cuda()
    a=ops.conv_fprop(dy,pw,None,dgrad=True,engine="simt"); b=ops.conv_fprop(dy,pw,None,dgrad=True,engine="auto")
    f1=ops.conv_fprop(x,pw,None,engine="simt",act="lrelu",act_slope=0.2); f2=ops.conv_fprop(x,pw,None,engine="auto",act="lrelu",act_slope=0.2)
    a2,p2=ops.conv_fprop(dy,pw,None,dgrad=True,engine="simt",actgrad="lrelu",actgrad_slope=0.2,aux=aux,want_pre=True)
    b2,q2=ops.conv_fprop(dy,pw,None,dgrad=True,engine="auto",actgrad="lrelu",actgrad_slope=0.2,aux=aux,want_pre=True)
    print((cin,cout,H,W),"fprop",rel(f2,f1),"dgrad",rel(b,a),"dgrad+actgrad",rel(b2,a2),"pre",rel(q2,p2),flush=True)
