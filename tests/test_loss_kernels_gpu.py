"""Dense loss kernels (first two terms) against the oracle's restatement of the reference's loss functions
(pinned to the unmodified reference by tests/golden/loss_golden.pt): values to 1e-5, gradients to 1e-4 of their largest
entry, on the same fp32 inputs."""
import pytest
import torch
import torch.nn.functional as F

from oracle import prn_loss_oracle as LO
from planerecnet_b200 import losses as PL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,ld", [(2 * 3728, 16), (777, 2)])
def test_focal_cate_loss(cuda_lib, n, ld):
    g = torch.Generator().manual_seed(n)
    logits = torch.randn(n, ld, generator=g) * 3
    labels = torch.randint(0, 3, (n,), generator=g)             # 2 = background
    labels[:5] = torch.tensor([0, 1, 2, 0, 2])
    num_ins = 37
    x = logits[:, :2].clone().requires_grad_(True)
    oh = torch.zeros(n, 2)
    pos = torch.nonzero(labels != 2).squeeze(1)
    oh[pos, labels[pos]] = 1
    ref = LO.sigmoid_focal_sum(x, oh, 0.25, 2.0) / (num_ins + 1)
    ref.backward()
    xd = logits.cuda().requires_grad_(True)
    got = PL.focal_cate_loss(xd, labels.cuda(), num_ins)
    got.backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
    gd = xd.grad.cpu()
    assert float((gd[:, :2] - x.grad).abs().max()) <= 1e-4 * float(x.grad.abs().max())
    assert float(gd[:, 2:].abs().max()) == 0.0 if ld > 2 else True


@pytest.mark.parametrize("B,h,w", [(2, 240, 320), (3, 17, 23)])
def test_depth_rmselog_loss(cuda_lib, B, h, w):
    g = torch.Generator().manual_seed(B)
    depth = (torch.rand(B, 1, h, w, generator=g) * 4 + 0.5)
    depth[0, 0, h - 1, w - 1] = 1e-12                            # its corner up-pixel falls below the log clamp: no gradient there
    gt = 0.5 + 4 * torch.rand(B, 1, 2 * h, 2 * w, generator=g)
    gt[0, 0, :5, :7] = 0.0                                       # invalid ground truth
    d = depth.clone().requires_grad_(True)
    up = F.interpolate(d, scale_factor=2, mode="bilinear", align_corners=False)
    ref = 5.0 * LO.rmse_log_mean(up, gt, gt > 1 / 1000)
    ref.backward()
    dd = depth.cuda().requires_grad_(True)
    got = PL.depth_rmselog_loss(dd, gt.cuda())
    (2.0 * got).backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float((dd.grad.cpu() / 2.0 - d.grad).abs().max()) <= 1e-4 * float(d.grad.abs().max())
