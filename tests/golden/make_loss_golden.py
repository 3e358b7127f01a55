"""Golden values of the joint loss from the UNMODIFIED reference: imports PlaneRecNetLoss from
/root/reference in the build container (runtime patches of SURVEY §8c: CUDA-less import, `.cuda()` as identity), runs it
on the seeded synthetic cases of tests/loss_cases.py with numpy's RNG seeded, and stores the five loss terms, the
gradient norms of the total w.r.t. every prediction and the target assignment.  Usage: python tests/golden/make_loss_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    sys.path.insert(0, "/root/reference")
    sys.path.insert(1, os.path.join(ROOT, "tests"))
    torch.cuda.current_device = lambda: 0
    torch.Tensor.cuda = lambda self, *a, **k: self
    import planerecnet  # noqa: F401  (reference module: sets up cfg)
    from data.config import cfg, set_cfg
    from utils import timer
    timer.disable_all()
    from models.functions.losses import PlaneRecNetLoss
    import loss_cases as LC
    set_cfg("PlaneRecNet_101_config")
    assert cfg.use_lava_loss and cfg.use_plane_loss
    crit = PlaneRecNetLoss()
    out = {}
    for name, kw in LC.CASES.items():
        mask, cate, kern, depth, gts, gt_depth = LC.synth(**kw)
        leaves = [mask] + cate + kern + [depth]
        for t in leaves:
            t.requires_grad_(True)
        np.random.seed(0)
        losses = crit(None, mask, cate, kern, depth, gts, gt_depth)
        total = sum(v.sum() for v in losses.values())
        total.backward()
        tg = [crit.prepare_ground_truth(g, mask_feat_size=mask.shape[-2:]) for g in gts]
        out[name] = {"losses": {k: float(v.detach().sum()) for k, v in losses.items()},
                     "grad_norms": [0.0 if t.grad is None else float(t.grad.double().norm()) for t in leaves],
                     "grid_orders": [[list(map(int, o)) for o in t[3]] for t in tg],
                     "cate_pos": [[(c != cfg.num_classes).nonzero().tolist() for c in t[1]] for t in tg],
                     "ins_label_sums": [[int(m.sum()) for m in t[0]] for t in tg]}
        print(name, out[name]["losses"], [len(o) for o in out[name]["grid_orders"][0]])
    torch.save(out, os.path.join(HERE, "loss_golden.pt"))


if __name__ == "__main__":
    main()
