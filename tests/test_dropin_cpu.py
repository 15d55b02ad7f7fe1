"""Drop-in boundary on CPU: with the reference importable (the /root/reference mount here, baseline/_ref on the GPU box),
`install_into_neosr()` must leave neosr's OWN registries resolving the hot-path names to the B200 classes, whatever
order the reference's directory scans run in, and must not disturb names this repo does not build."""
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no reference tree (/root/reference or baseline/_ref)")


@pytest.fixture(autouse=True)
def _restore_reference_registries():
    """The override is process-wide; the other CPU tests use the reference's registries as the ORACLE, so put every
    entry back after each test here."""
    ref_shim.activate()
    import neosr_b200.plugin as plugin
    plugin._force_reference_scans()
    from neosr.utils.registry import ARCH_REGISTRY, LOSS_REGISTRY, MODEL_REGISTRY
    saved = [(r, dict(r._obj_map)) for r in (ARCH_REGISTRY, LOSS_REGISTRY, MODEL_REGISTRY)]
    yield
    for r, m in saved:
        r._obj_map.clear()
        r._obj_map.update(m)


def test_install_into_neosr_overrides_the_reference_registries():
    ref_shim.activate()
    import neosr_b200
    done = neosr_b200.install_into_neosr()
    from neosr.utils.registry import ARCH_REGISTRY, LOSS_REGISTRY, MODEL_REGISTRY
    assert set(done) == {"arch", "loss", "model"}
    for name in ("swinir_small", "swinir_medium", "swinir_large", "hat_s", "hat_m", "hat_l", "esrgan", "compact",
                 "realplksr", "unet"):
        assert ARCH_REGISTRY.get(name).__module__.startswith("neosr_b200.archs."), name
    for name in ("L1Loss", "chc_loss", "vgg_perceptual_loss", "gan_loss", "mssim_loss", "consistency_loss"):
        assert LOSS_REGISTRY.get(name).__module__.startswith("neosr_b200.losses."), name
    for name in ("image", "otf"):
        assert MODEL_REGISTRY.get(name).__module__.startswith("neosr_b200.models."), name
    # names outside the built hot path keep pointing at the reference
    assert ARCH_REGISTRY.get("span").__module__.startswith("neosr.archs.")
    # idempotent, and the reference's lazy arch scan (archs/__init__.py:17-27) no longer trips the duplicate assert
    neosr_b200.install_into_neosr()
    from neosr.archs import build_network
    net = build_network({"type": "span"})
    assert type(net).__module__.startswith("neosr.archs.")


def test_b200_arch_refuses_cpu_even_through_the_reference_factory():
    """No CPU fallback behind the plugin surface: the module builds (parameters only) and its forward raises."""
    import torch
    ref_shim.activate()
    import neosr_b200
    neosr_b200.install_into_neosr()
    from neosr.archs import build_network
    net = build_network({"type": "compact"})
    assert type(net).__module__.startswith("neosr_b200.archs.")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            net(torch.rand(1, 3, 8, 8))
