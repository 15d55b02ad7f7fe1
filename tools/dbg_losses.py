import sys, torch
sys.path.insert(0, ".")
from neosr_b200.losses import build_loss
from oracle import losses as OL
from oracle.make_golden_otf import loss_inputs
from oracle.ref_otf import structured_gt

def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

gt = structured_gt(22, 2, 96, 80)
x = (gt + 0.1 * torch.randn(gt.shape, generator=torch.Generator().manual_seed(2))).clamp(-0.05, 1.05)
cases = [("far", x, gt, dict(saturation=1.2, brightness=0.9, loss_weight=0.5)), ("far_s", x, gt, dict(saturation=1.2)),
         ("far_b", x, gt, dict(brightness=0.9)), ("far_w", x, gt, dict(loss_weight=0.5))]
xn, gn = loss_inputs()["near"]
cases += [("near", xn, gn, dict()), ("near_nb", xn, gn, dict(blur=False))]
for tag, xx, gg, kw in cases:
    mod = build_loss({"type": "consistency_loss", **kw}).cuda()
    okw = dict(kw); lw = okw.pop("loss_weight", 1.0)
    xo = xx.clone().requires_grad_(True)
    vo = OL.consistency_loss(xo, gg, lw, **okw)
    go, = torch.autograd.grad(vo, xo)
    v, g = mod.value_and_grad(xx.cuda().contiguous(), gg.cuda(), True, None)
    d = (g.cpu() - go).abs()
    print(tag, kw, "value", float(v), float(vo), "grad rel", rel(g, go), "argmax", divmod(int(d.argmax()), 1), "gmax", float(go.abs().max()))
