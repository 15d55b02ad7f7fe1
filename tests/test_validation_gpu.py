"""Validation path (image.py:664-925): `test()` runs the EMA weights through the same engine in
forward-only mode — whole image and the reference's 16-px-overlap tiling — and `validation()` walks a loader, computes
metrics and keeps the best."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(tmp_path, tile=-1):
    from neosr_b200.models import build_model
    opt = {"name": "valtest", "model_type": "image", "scale": 2, "is_train": True, "dist": False, "rank": 0, "world_size": 1,
           "network_g": {"type": "compact", "num_feat": 16, "num_conv": 2, "upscale": 2},
           "datasets": {"train": {"patch_size": 24}},
           "train": {"ema": 0.999, "optim_g": {"type": "adan_sf", "lr": 1e-3, "betas": (0.98, 0.92, 0.987), "weight_decay": 0.02,
                                               "schedule_free": True, "warmup_steps": 100},
                     "pixel_opt": {"type": "L1Loss", "loss_weight": 1.0}},
           "val": {"tile": tile, "save_img": True, "metrics": {"psnr": {"type": "calculate_psnr", "crop_border": 4, "better": "higher"}}},
           "path": {"visualization": str(tmp_path / "vis")}, "cuda_graph": False}
    return build_model(opt)


class _Loader(list):
    class dataset:  # noqa: N801
        opt = {"name": "synthetic_val", "type": "paired"}


def test_test_uses_ema_weights_and_matches_oracle(tmp_path):
    from neosr_b200.data.synthetic import structured_gt
    from oracle.compact import compact_forward
    m = _model(tmp_path)
    g = torch.Generator().manual_seed(0)
    for it in range(3):  # make EMA != weights
        gt = structured_gt(it, 2, 48, 48)
        m.feed_data({"lq": torch.nn.functional.avg_pool2d(gt, 2), "gt": gt})
        m.optimize_parameters(it)
    lq = torch.rand(1, 3, 40, 56, generator=g)
    m.feed_data({"lq": lq})
    m.test()
    ema_p = {k: v.detach().cpu() for k, v in m.net_g_ema.module.named_parameters()}
    ref = compact_forward(ema_p, lq, num_conv=2, upscale=2)
    assert m.output.shape == (1, 3, 80, 112)
    assert float((m.output.cpu() - ref).abs().max()) < 1e-4
    w_p = {k: v.detach().cpu() for k, v in m.net_g.named_parameters()}
    assert any(float((w_p[k] - ema_p[k]).abs().max()) > 0 for k in w_p)  # EMA really differs from the live weights
    assert m.net_g.training  # test() restores train mode (image.py:680-682)


def test_tiled_inference_equals_whole_image(tmp_path):
    """2 convs => receptive radius 4 < the 16-px overlap: the partitioned result must equal the one-shot result."""
    whole, tiled = _model(tmp_path, -1), _model(tmp_path, 32)
    tiled.net_g.load_state_dict(whole.net_g.state_dict())
    tiled.net_g_ema.load_state_dict(whole.net_g_ema.state_dict())
    lq = torch.rand(1, 3, 75, 100, generator=torch.Generator().manual_seed(1))
    for m in (whole, tiled):
        m.feed_data({"lq": lq})
        m.test()
    assert tiled.output.shape == whole.output.shape == (1, 3, 150, 200)
    assert float((tiled.output - whole.output).abs().max()) < 1e-5


def test_validation_loop_metrics_and_images(tmp_path):
    from neosr_b200.data.synthetic import structured_gt
    m = _model(tmp_path)
    items = _Loader()
    for i in range(3):
        gt = structured_gt(10 + i, 1, 48, 64)
        items.append({"lq": torch.nn.functional.avg_pool2d(gt, 2), "gt": gt, "lq_path": [f"/data/val/img{i}.png"]})
    m.validation(items, 100, None, True)
    assert 5.0 < m.metric_results["psnr"] < 60.0
    assert m.best_metric_results["synthetic_val"]["psnr"]["iter"] == 100
    assert (tmp_path / "vis" / "img0" / "img0_100.png").exists() or (tmp_path / "vis" / "img0" / "img0_100.png.npy").exists()
    assert m.is_train
    gt = structured_gt(1, 2, 48, 48)  # training continues afterwards
    m.feed_data({"lq": torch.nn.functional.avg_pool2d(gt, 2), "gt": gt})
    m.optimize_parameters(101)
