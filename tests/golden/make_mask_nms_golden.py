"""Golden vectors for the greedy mask-NMS (nms.py:53-80): runs the UNMODIFIED reference function, imported from
/root/reference in the build container, on seeded blob masks and stores inputs + its keep vector.
Usage (build container only): python tests/golden/make_mask_nms_golden.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from models.functions.nms import mask_nms  # noqa: E402  (pure torch, no CUDA needed)


def main():
    g = torch.Generator().manual_seed(11)
    cases = []
    h, w = 16, 20
    yy, xx = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    for n in (1, 7, 60):
        masks = torch.zeros(n, h, w, dtype=torch.bool)
        for i in range(n):
            cy, cx = float(torch.rand(1, generator=g)) * h, float(torch.rand(1, generator=g)) * w
            rad = 1.5 + 5.0 * float(torch.rand(1, generator=g))
            masks[i] = ((yy - cy) ** 2 + (xx - cx) ** 2) < rad * rad
        if n > 3:
            masks[2] = masks[1]          # duplicates: IoU 1
            masks[3] = False             # empty mask together with ...
            masks[n - 1] = False         # ... another empty one: union 0 branch
        labels = torch.randint(0, 2, (n,), generator=g)
        scores = torch.sort(torch.rand(n, generator=g), descending=True).values
        sums = masks.sum((1, 2)).float()
        for thr in (0.1, 0.5):
            keep = mask_nms(labels, masks, sums, scores, nms_thr=thr)
            cases.append(dict(masks=masks, labels=labels, scores=scores, sums=sums, thr=thr, keep=keep.bool()))
    torch.save(cases, os.path.join(HERE, "mask_nms.pt"))
    print("wrote", len(cases), "cases; kept", [int(c["keep"].sum()) for c in cases])


if __name__ == "__main__":
    main()
