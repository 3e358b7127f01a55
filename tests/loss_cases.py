"""Seeded synthetic predictions + ground truth for the joint-loss oracle (3-10 axis-aligned plane masks,
xyxy boxes, class 0, unit normal + offset, ScanNet intrinsics, gt depth in [0.5, 4.5] m).  Shared by
tests/golden/make_loss_golden.py (reference side) and tests/test_loss_oracle_cpu.py (oracle side) so that the fixture
only has to store the reference's outputs."""
import torch

GRIDS = (40, 36, 24, 16)


def synth(seed, B=2, n_lo=3, n_hi=8, tiny=False):
    g = torch.Generator().manual_seed(seed)
    mask = torch.randn(B, 128, 120, 160, generator=g) * 0.5
    cate = [torch.randn(B, 2, S, S, generator=g) - 2 for S in GRIDS]
    kern = [torch.randn(B, 128, S, S, generator=g) * 0.1 for S in GRIDS]
    depth = torch.rand(B, 1, 240, 320, generator=g) * 4 + 0.5
    gts = []
    for b in range(B):
        n = int(torch.randint(n_lo, n_hi, (1,), generator=g))
        masks = torch.zeros(n, 480, 640, dtype=torch.uint8)
        boxes = torch.zeros(n, 4, dtype=torch.float64)
        for i in range(n):
            w = int(torch.randint(30, 420, (1,), generator=g))
            h = int(torch.randint(30, 360, (1,), generator=g))
            x0 = int(torch.randint(0, 640 - w, (1,), generator=g))
            y0 = int(torch.randint(0, 480 - h, (1,), generator=g))
            if tiny and i == 0:                                   # a plane of a few pixels (below the first scale range's mask)
                w, h = 6, 5
            masks[i, y0:y0 + h, x0:x0 + w] = 1
            boxes[i] = torch.tensor([x0, y0, x0 + w, y0 + h], dtype=torch.float64)
        nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64), dim=1)
        planes = torch.cat([nrm, torch.rand(n, 1, generator=g, dtype=torch.float64) + 1], 1)
        K = torch.tensor([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], dtype=torch.float64)
        gts.append(dict(masks=masks, boxes=boxes, classes=torch.zeros(n, dtype=torch.int64), plane_paras=planes, k_matrix=K))
    gt_depth = 0.5 + 4 * torch.rand(B, 1, 480, 640, generator=g)
    return mask, cate, kern, depth, gts, gt_depth


CASES = {"loss_seed0": dict(seed=0), "loss_seed1_b1": dict(seed=1, B=1), "loss_seed2_many_tiny": dict(seed=2, n_lo=8, n_hi=11, tiny=True)}
