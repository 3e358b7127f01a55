"""First thing to run on a GPU in round 2: the assembled planerecnet_b200.losses.PlaneRecNetLoss (device-side targets,
kernel-backed dice / lava / focal / depth terms, host-side plane term) against the pinned CPU oracle on the golden cases,
plus whether the device-side target assignment equals the CPU one.  (Its device steps are parity-green individually and
the assembly reproduces the reference on the CPU with emulated device steps; round 1 ran out of GPU budget before this
script could be run.)  Usage: python tools/loss_module_check.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import loss_cases as LC  # noqa: E402
from oracle import prn_loss_oracle as LO  # noqa: E402
from planerecnet_b200 import losses as PL, targets as T  # noqa: E402
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402


def main():
    set_cfg("PlaneRecNet_101_config")
    crit = PL.PlaneRecNetLoss(cfg)
    for name, kw in LC.CASES.items():
        mask, cate, kern, depth, gts, gt_depth = LC.synth(**kw)
        leaves = [mask] + cate + kern + [depth]
        for t in leaves:
            t.requires_grad_(True)
        np.random.seed(0)
        ref = LO.loss_forward(mask, cate, kern, depth, gts, gt_depth)
        tot = sum(v.sum() for v in ref.values())
        if torch.isfinite(tot):
            tot.backward()
        dl = [t.detach().cuda().requires_grad_(True) for t in leaves]
        gts_d = [{k: v.cuda() for k, v in g.items()} for g in gts]
        same = all(a[3] == b[3] for g, gd in zip(gts, gts_d)
                   for a, b in zip(T.assign_targets(g, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"]),
                                   T.assign_targets(gd, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"])))
        np.random.seed(0)
        out = crit(None, dl[0], dl[1:5], dl[5:9], dl[9], gts_d, gt_depth.cuda())
        print(name, "device targets == CPU targets:", same)
        for k in ref:
            print(f"   {k}: ours {float(out[k].sum()):.6f}  oracle {float(ref[k].sum()):.6f}")
        if torch.isfinite(tot):
            sum(v.sum() for v in out.values()).backward()
            for t, d in zip(leaves, dl):
                if t.grad is not None and d.grad is not None:
                    a, b = d.grad.cpu().double().flatten(), t.grad.double().flatten()
                    print(f"   grad {tuple(t.shape)}: cos {float((a @ b) / (a.norm() * b.norm() + 1e-30)):.5f} "
                          f"rel-L2 {float((a - b).norm() / (b.norm() + 1e-30)):.3e}")


if __name__ == "__main__":
    main()
