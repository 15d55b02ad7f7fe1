"""`CUDAPrefetcher` (the reference's neosr/data/prefetch_dataloader.py:69-113 interface): batches arrive on the device
unchanged and in order, the epoch ends with None, `reset()` restarts it, unpinned sources are staged through pinned memory."""
import pytest
import torch


def _loader(n, pinned):
    g = torch.Generator().manual_seed(3)
    out = []
    for i in range(n):
        lq, gt = torch.rand(2, 3, 8, 8, generator=g), torch.rand(2, 3, 32, 32, generator=g)
        if pinned:
            lq, gt = lq.pin_memory(), gt.pin_memory()
        out.append({"lq": lq, "gt": gt, "lq_path": f"img{i}.png"})
    return out


def test_prefetcher_needs_a_cuda_device():
    from neosr_b200.data import CUDAPrefetcher
    with pytest.raises(RuntimeError):
        CUDAPrefetcher(_loader(1, False), {}, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [True, False])
def test_prefetcher_order_epoch_end_and_reset(pinned):
    from neosr_b200.data import CUDAPrefetcher
    data = _loader(5, pinned)
    pre = CUDAPrefetcher(data, {})
    for epoch in range(2):
        for i in range(5):
            b = pre.next()
            assert b["lq"].is_cuda and b["gt"].is_cuda and b["lq_path"] == f"img{i}.png"
            # consume on the compute stream right away, as the model does
            s = (b["lq"].sum() + b["gt"].sum()).item()
            assert abs(s - float(data[i]["lq"].sum() + data[i]["gt"].sum())) < 1e-2
            assert torch.equal(b["gt"].cpu(), data[i]["gt"])
        assert pre.next() is None
        pre.reset()
