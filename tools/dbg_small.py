import torch, time, os
from neosr_b200 import ops
def bench(B,H,W,cin,cout,n=20):
    x=torch.randn(B,H,W,cin,device="cuda"); dy=torch.randn(B,H,W,cout,device="cuda")
    dw=torch.empty(cout,cin,3,3,device="cuda")
    for _ in range(3): ops.conv_wgrad(x,dy,dw,None,3,3)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): ops.conv_wgrad(x,dy,dw,None,3,3)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n
    print(f"wgrad B{B} {H}x{W} {cin}->{cout}: {ms*1e3:8.1f} us  {2*B*H*W*cin*cout*9/ms/1e9:7.1f} TF",flush=True)
for s in [(16,64,64,192,64),(16,64,64,160,32),(16,64,64,64,32),(32,64,64,180,180),(8,128,128,64,256),(16,256,256,64,64),(4,32,32,512,512)]:
    bench(*s)
