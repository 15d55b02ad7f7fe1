"""Host-side drop-in surface on CPU: registries, state_dict key/shape compatibility with the
reference, optimizer state layout, option handling."""
import pytest
import torch

from oracle import ref_shim
from oracle.swinir import swinir_medium_config, swinir_param_shapes


def test_registries_and_factories():
    from neosr_b200 import ARCH_REGISTRY, LOSS_REGISTRY, MODEL_REGISTRY
    for n in ("swinir_small", "swinir_medium", "swinir_large", "VGGFeatureExtractor"):
        assert n in ARCH_REGISTRY
    import neosr_b200.losses  # noqa: F401
    import neosr_b200.models  # noqa: F401
    for n in ("L1Loss", "chc_loss", "vgg_perceptual_loss", "gan_loss"):
        assert n in LOSS_REGISTRY
    assert "image" in MODEL_REGISTRY
    with pytest.raises(KeyError):
        ARCH_REGISTRY.get("nope")
    with pytest.raises(AssertionError):
        ARCH_REGISTRY.register(ARCH_REGISTRY.get("swinir_small"))


def test_swinir_medium_state_dict_matches_oracle_shapes():
    from neosr_b200.archs import build_network
    net = build_network({"type": "swinir_medium", "drop_path_rate": 0.0, "upscale": 4})
    assert {k: tuple(v.shape) for k, v in net.named_parameters()} == swinir_param_shapes(swinir_medium_config(4))
    assert [k for k, _ in net.named_parameters()] == list(swinir_param_shapes(swinir_medium_config(4)))


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="live reference not mounted")
def test_state_dict_interchanges_with_reference():
    """Key-for-key, buffers included (relative_position_index, attn_mask): load ours into the
    reference module and back (strict=True both ways)."""
    from neosr_b200.archs import build_network
    ours = build_network({"type": "swinir_medium", "drop_path_rate": 0.0, "upscale": 4})
    ref = ref_shim.build_network({"type": "swinir_medium", "drop_path_rate": 0.0})
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o) == list(sd_r)
    for k in sd_o:
        assert tuple(sd_o[k].shape) == tuple(sd_r[k].shape), k
        if "relative_position_index" in k or "attn_mask" in k:
            assert torch.equal(sd_o[k].float(), sd_r[k].float()), k
    ref.load_state_dict(sd_o, strict=True)
    ours.load_state_dict(sd_r, strict=True)


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="live reference not mounted")
def test_vgg_extractor_state_dict_keys_match_reference():
    from neosr_b200.archs.vgg_arch import VGGFeatureExtractor
    from oracle import losses as OL
    from oracle.swinir import synth_params
    ours = VGGFeatureExtractor(["conv1_2", "conv5_4"], allow_random_init=True)
    ref = ref_shim.build_vgg_perceptual(synth_params(OL.vgg19_conv_shapes(), seed=5)).vgg
    assert set(ours.state_dict()) == set(ref.state_dict())


def test_adan_sf_param_group_and_state_layout():
    from neosr_b200.optimizers import adan_sf
    p = [torch.nn.Parameter(torch.zeros(4, 3))]
    opt = adan_sf(p, lr=1e-3, betas=(0.98, 0.92, 0.987), weight_decay=0.02, schedule_free=True, warmup_steps=10)
    keys = set(opt.param_groups[0])
    assert {"lr", "betas", "eps", "r", "weight_decay", "max_grad_norm", "warmup_steps", "train_mode", "weight_sum",
            "lr_max", "weight_lr_power", "schedule_free"} <= keys
    with pytest.raises(ValueError):
        adan_sf(p, lr=-1.0)
    with pytest.raises(ValueError):
        adan_sf(p, betas=(1.0, 0.9, 0.9))
    sd = opt.state_dict()
    opt2 = adan_sf(p, lr=5e-4)
    opt2.load_state_dict(sd)
    assert opt2.param_groups[0]["lr"] == 1e-3


def test_net_opt_reads_scale_from_argv(tmp_path, monkeypatch):
    from neosr_b200.archs import arch_util
    toml = tmp_path / "o.toml"
    toml.write_text('name="x"\nscale = 2\n[datasets.train]\ntype="paired"\n')
    monkeypatch.setattr("sys.argv", ["train.py", "-opt", str(toml)])
    monkeypatch.setitem(arch_util._default, "scale", None)
    assert arch_util.net_opt() == (2, True)
    arch_util.set_default_scale(4)
    assert arch_util.net_opt()[0] == 4
