import torch, os
from neosr_b200 import ops
def t(fn,n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3
for (B,H,W,cin,cout) in [(32,256,256,3,64),(32,256,256,64,3)]:
    x=torch.randn(B,H,W,cin,device="cuda"); dy=torch.randn(B,H,W,cout,device="cuda")
    w=torch.randn(cout,cin,3,3,device="cuda")*0.1; b=torch.zeros(cout,device="cuda")
    pw=ops.PackedWeight(w).refresh(); dw=torch.empty_like(w); db=torch.empty_like(b)
    print(f"{cin}->{cout} fprop {t(lambda: ops.conv_fprop(x,pw,b)):8.1f} us  dgrad {t(lambda: ops.conv_fprop(dy,pw,None,dgrad=True)):8.1f} us  wgrad {t(lambda: ops.conv_wgrad(x,dy,dw,db,3,3)):8.1f} us  (NSR_NARROW_GEMM={os.environ.get('NSR_NARROW_GEMM','1')})")
