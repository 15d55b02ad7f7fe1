"""The literal drop-in: the reference's UNMODIFIED `train.py` (from the staged baseline/_ref copy) trains for a few
iterations on a real paired dataset on disk with the B200 hook installed the way INTEGRATION.md section 1 describes
(one two-line file in a scanned folder).  Everything outside the hot path - option parser, dataset, sampler,
DataLoader workers, CUDAPrefetcher, loggers, checkpoint cadence - is the reference's own code; `model_type = "image"`,
`[network_g]`, `[network_d]`, the `*_opt` losses and `adan_sf` resolve to this repo's classes and run on the kernels.
"""
import os
import re
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"

TOML = """
name = "dropin_{tag}"
model_type = "image"
scale = 4
manual_seed = 1024
{extra_top}
[datasets.train]
type = "paired"
dataroot_gt = '{gt}'
dataroot_lq = '{lq}'
patch_size = 32
batch_size = 4
num_worker_per_gpu = 2

[path]

[network_g]
type = "{net_g}"
{net_g_extra}
{net_d}
[train]
ema = 0.999
{train_extra}
[train.optim_g]
type = "adan_sf"
lr = 1e-3
betas = [ 0.98, 0.92, 0.987 ]
weight_decay = 0.02
schedule_free = true
warmup_steps = 1600
{optim_d}
[train.pixel_opt]
type = "L1Loss"
loss_weight = 1.0

[train.perceptual_opt]
type = "vgg_perceptual_loss"
loss_weight = 0.5
criterion = "chc"
allow_random_init = true
{gan}
[logger]
total_iter = {iters}
save_checkpoint_freq = {iters}
print_freq = 2
use_tb_logger = false
"""


def _dataset(root: Path, n: int = 12) -> tuple[Path, Path]:
    """A few structured 8-bit RGB pairs (GT 160x160, LQ 40x40) written as PNG, the format paired_dataset.py reads."""
    import cv2
    from neosr_b200.data.synthetic import structured_gt
    gt_dir, lq_dir = root / "gt", root / "lq"
    gt_dir.mkdir(parents=True), lq_dir.mkdir(parents=True)
    gt = structured_gt(7, n, 160, 160)
    lq = torch.nn.functional.interpolate(gt, scale_factor=0.25, mode="bicubic", antialias=True).clamp(0, 1)
    for i in range(n):
        for d, t in ((gt_dir, gt), (lq_dir, lq)):
            img = (t[i].permute(1, 2, 0).numpy() * 255.0).round().astype(np.uint8)[:, :, ::-1]
            assert cv2.imwrite(str(d / f"img{i:03d}.png"), img)
    return gt_dir, lq_dir


def _run_train(tmp_path: Path, tag: str, iters: int, **fmt) -> tuple[str, Path]:
    if not (REF / "train.py").exists():
        pytest.skip("baseline/_ref is not staged (run __graft_entry__.build() where /root/reference is mounted)")
    work = tmp_path / "neosr"
    shutil.copytree(REF, work)
    # the maintainer-side change: ONE file in a folder the reference scans (models/__init__.py:13-22)
    (work / "neosr" / "models" / "zz_b200.py").write_text("import neosr_b200\nneosr_b200.install_into_neosr()\n")
    gt_dir, lq_dir = _dataset(tmp_path / "data")
    base = dict(extra_top="", net_g_extra="", net_d="", train_extra="", optim_d="", gan="")
    base.update(fmt)
    (work / "opt.toml").write_text(TOML.format(tag=tag, gt=gt_dir, lq=lq_dir, iters=iters, **base))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT), str(work / "_stubs")]), NSR_PLUGIN_VERBOSE="1")
    out = subprocess.run([sys.executable, "train.py", "-opt", "opt.toml"], cwd=work, env=env, capture_output=True,
                         text=True, timeout=1500)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-4000:]
    return text, work / "experiments" / f"dropin_{tag}"


def test_unmodified_train_py_runs_the_image_model_on_the_kernels(tmp_path):
    iters = 6
    text, exp = _run_train(tmp_path, "swinir", iters, net_g="swinir_small", net_g_extra="drop_path_rate = 0.0")
    assert "neosr_b200.install_into_neosr:" in text
    # the reference's MessageLogger printed every iteration with the loss keys of image.closure (image.py:476-586)
    # (print_freq = 2: with print_freq = 1 the reference's own ETA arithmetic divides by zero at iteration 1, logger.py:100)
    its = [int(m) for m in re.findall(r"iter:\s*(\d+)", text.replace(",", ""))]
    assert its and max(its) == iters, its
    vals = [float(v) for v in re.findall(r"l_g_total: ([0-9.eE+-]+)", text)]
    assert len(vals) >= iters // 2 and all(np.isfinite(vals)) and all(v > 0 for v in vals)
    assert "l_g_pix" in text and "l_g_percep" in text
    # checkpoints in the reference's format and key names (base.py:311-335, 433-442): EMA weights + training state
    ck = torch.load(exp / "models" / f"net_g_{iters}.pth", map_location="cpu", weights_only=True)
    assert "params" in ck
    sys.path.insert(0, str(REF))
    try:
        from oracle import ref_shim
        ref_shim.activate()
        ref_net = ref_shim.build_network({"type": "swinir_small"})
    finally:
        sys.path.pop(0)
    missing = ref_net.load_state_dict(ck["params"], strict=True)  # the reference's own module loads it
    assert not missing.missing_keys and not missing.unexpected_keys
    st = torch.load(exp / "training_states" / f"{iters}.state", map_location="cpu", weights_only=False)
    assert st["iter"] == iters and len(st["optimizers"]) == 1
    s0 = next(iter(st["optimizers"][0]["state"].values()))
    assert {"exp_avg", "exp_avg_sq", "exp_avg_diff", "z", "neg_pre_grad"} <= set(s0)


def test_unmodified_train_py_gan_step_with_scheduler_and_warmup(tmp_path):
    """GAN configuration (unet discriminator, bce) + `[train.scheduler]` + `warmup_iter`: the options the reference's
    train loop forwards to `update_learning_rate` (train.py:274, base.py:174-252)."""
    iters = 4
    text, exp = _run_train(
        tmp_path, "gan", iters, net_g="compact",
        net_d='[network_d]\ntype = "unet"\n',
        train_extra='warmup_iter = 3\n[train.scheduler]\ntype = "MultiStepLR"\nmilestones = [ 2 ]\ngamma = 0.5\n',
        optim_d='[train.optim_d]\ntype = "adan_sf"\nlr = 5e-4\nbetas = [ 0.98, 0.92, 0.995 ]\nweight_decay = 0.02\n'
                'schedule_free = true\n',
        gan='[train.gan_opt]\ntype = "gan_loss"\ngan_type = "bce"\nloss_weight = 0.3\n')
    for key in ("l_g_gan", "l_d_real", "l_d_fake", "out_d_real", "out_d_fake"):
        assert key in text, key
    lrs = [float(v) for v in re.findall(r"lr: ([0-9.eE+-]+)", text)]
    assert len(lrs) >= iters // 2, text[-2000:]
    # expected: exactly what base.update_learning_rate does with torch's scheduler object (base.py:229-252): the
    # scheduler steps every iteration, and for it < warmup_iter the lr is overwritten with initial_lr * it / warmup_iter
    q = torch.nn.Parameter(torch.zeros(1))
    o = torch.optim.SGD([q], lr=1e-3)
    sch = torch.optim.lr_scheduler.MultiStepLR(o, milestones=[2], gamma=0.5)
    want = []
    for it in range(1, iters + 1):
        o.step()
        sch.step()
        if it < 3:
            o.param_groups[0]["lr"] = o.param_groups[0]["initial_lr"] / 3 * it
        if it % 2 == 0:
            want.append(o.param_groups[0]["lr"])
    for a, b in zip(lrs, want):
        assert abs(a - b) <= 6e-3 * b, (lrs, want)  # the log prints 3 significant digits
    assert (exp / "models" / f"net_d_{iters}.pth").exists()
