"""oracle/prn_loss_oracle.py (restatement of PlaneRecNetLoss) against the golden values produced by the
unmodified reference (tests/golden/make_loss_golden.py -> tests/golden/loss_golden.pt) on the seeded cases of
tests/loss_cases.py: the five loss terms (1e-5 relative; the reference's NaN for a degenerate plane is reproduced), the
norms of the gradients w.r.t. every prediction tensor, and the target assignment exactly."""
import math
import os

import numpy as np
import pytest
import torch

import loss_cases as LC
from oracle import prn_loss_oracle as LO

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.pt"))


@pytest.mark.parametrize("name", list(LC.CASES))
def test_joint_loss_matches_reference(name):
    ref = GOLD[name]
    mask, cate, kern, depth, gts, gt_depth = LC.synth(**LC.CASES[name])
    leaves = [mask] + cate + kern + [depth]
    for t in leaves:
        t.requires_grad_(True)
    np.random.seed(0)                       # the plane loss samples triplets with numpy's global RNG (vnl.py:48-53)
    taps = {}
    out = LO.loss_forward(mask, cate, kern, depth, gts, gt_depth, taps=taps)
    assert list(out) == ["ins", "cat", "dpt", "pln", "lav"]
    for k, v in ref["losses"].items():
        got = float(out[k].detach().sum())
        if math.isnan(v):
            assert math.isnan(got), (k, got)
        else:
            assert abs(got - v) <= 1e-5 * max(1.0, abs(v)), (k, got, v)
    # target assignment: identical cells, order and masks
    for b, tg in enumerate(taps["targets"]):
        assert [list(t[3]) for t in tg] == ref["grid_orders"][b]
        assert [(t[1] != LO.CFG["num_classes"]).nonzero().tolist() for t in tg] == ref["cate_pos"][b]
        assert [int(t[0].sum()) for t in tg] == ref["ins_label_sums"][b]
    # gradients of the total (autograd through the restatement)
    if not any(math.isnan(v) for v in ref["losses"].values()):
        sum(v.sum() for v in out.values()).backward()
        got = [0.0 if t.grad is None else float(t.grad.double().norm()) for t in leaves]
        for a, b in zip(got, ref["grad_norms"]):
            assert abs(a - b) <= 1e-4 * max(b, 1e-6), (got, ref["grad_norms"])


def test_elementary_losses_closed_forms():
    p = torch.tensor([[1.0, 0.0, 1.0, 0.0]])
    assert torch.allclose(LO.dice_loss(p, p.bool()), torch.tensor([1 - 4 / 4.002]), atol=1e-6)
    z = torch.zeros(3, 2)
    assert torch.allclose(LO.sigmoid_focal_sum(z, torch.zeros(3, 2), 0.25, 2.0), torch.tensor(6 * 0.75 * 0.25 * math.log(2.0)))
    d = torch.full((1, 1, 4, 4), 2.0)
    assert float(LO.rmse_log_mean(d, d * math.e, torch.ones_like(d))) == pytest.approx(1.0, rel=1e-6)
    ramp = torch.arange(6.0).view(1, 1, 1, 6).expand(1, 1, 5, 6).contiguous()
    g = LO.gradient_map(ramp)
    assert torch.allclose(g[0, 0, 2, 1:5], torch.ones(4))          # |d/dx| = 1 -> ((1+2+1)*2/8)^2 = 1 in the interior
