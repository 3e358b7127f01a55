"""Synthetic weights/inputs for benchmarks and tests (no datasets or checkpoints are available offline)."""
import torch


def perturb_(net, seed=1234):
    """Deterministically move a freshly constructed model away from its degenerate default init
    (BN running stats 0/1, zero DCN offsets, unit GN affine) so that every branch of the arithmetic is
    exercised.  Works on the reference model and on ours alike (same state_dict keys and order)."""
    g = torch.Generator().manual_seed(seed)
    sd = net.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            if not v.is_floating_point():
                continue
            is_bn = ("bn" in k.split(".")[-2] or ".downsample.1." in k or
                     (k.startswith("depth_decoder") and k.split(".")[-2] in ("2", "3") and v.dim() == 1 and "latlayer" not in k))
            if k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            elif k.endswith("running_var"):
                v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
            elif "offset_conv.weight" in k:
                v.copy_(torch.randn(v.shape, generator=g) * 0.03)
            elif "offset_conv.bias" in k:
                v.copy_(torch.randn(v.shape, generator=g) * 0.5)
            elif "modulator_conv.weight" in k:
                v.copy_(torch.randn(v.shape, generator=g) * 0.02)
            elif "modulator_conv.bias" in k:
                v.copy_(torch.randn(v.shape, generator=g) * 0.3)
            elif v.dim() == 1 and k.endswith(".weight"):      # BN / GN scale
                v.copy_(torch.rand(v.shape, generator=g) * 0.4 + 0.8)
            elif v.dim() == 1 and k.endswith(".bias") and ("tower" in k or "convs_all_levels" in k or "conv_pred.1" in k or is_bn):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    return net


def make_input(B, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, H, W, generator=g, device="cpu")


def make_gt(B, H=480, W=640, seed=0, n_lo=3, n_hi=10):
    """Synthetic ground truth of SURVEY §8d config 4: per image 3-10 axis-aligned plane masks uint8 [n,H,W], float64 xyxy boxes,
    class 0, unit normal + offset, ScanNet intrinsics; gt depth in [0.5, 4.5] m.  Returns (gt_instances, gt_depths) on the CPU
    in the layout train.py's collate hands to PlaneRecNetLoss (losses.py:53-72)."""
    g = torch.Generator().manual_seed(1000 + seed)
    gts = []
    for _ in range(B):
        n = int(torch.randint(n_lo, n_hi + 1, (1,), generator=g, device="cpu"))
        masks = torch.zeros(n, H, W, dtype=torch.uint8, device="cpu")
        boxes = torch.zeros(n, 4, dtype=torch.float64, device="cpu")
        for i in range(n):
            w = int(torch.randint(W // 20, (2 * W) // 3, (1,), generator=g, device="cpu"))
            h = int(torch.randint(H // 16, (3 * H) // 4, (1,), generator=g, device="cpu"))
            x0 = int(torch.randint(0, W - w, (1,), generator=g, device="cpu"))
            y0 = int(torch.randint(0, H - h, (1,), generator=g, device="cpu"))
            masks[i, y0:y0 + h, x0:x0 + w] = 1
            boxes[i] = torch.tensor([x0, y0, x0 + w, y0 + h], dtype=torch.float64, device="cpu")
        nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64, device="cpu"), dim=1)
        planes = torch.cat([nrm, torch.rand(n, 1, generator=g, dtype=torch.float64, device="cpu") + 1], 1)
        K = torch.tensor([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], dtype=torch.float64, device="cpu")
        gts.append(dict(masks=masks, boxes=boxes, classes=torch.zeros(n, dtype=torch.int64, device="cpu"), plane_paras=planes, k_matrix=K))
    gt_depth = 0.5 + 4 * torch.rand(B, 1, H, W, generator=g, device="cpu")
    return gts, gt_depth


def make_cotangents(outs, seed=1, device=None):
    """Fixed seeded cotangents for the 10 training outputs (mask_pred, [cate]x4, [kernel]x4, depth_pred): SURVEY §8d config 4 (i).
    Generated on the CPU so that every implementation sees the same values."""
    g = torch.Generator().manual_seed(seed)
    m, cs, ks, d = outs

    def mk(t):
        v = torch.randn(t.shape, generator=g, device="cpu") / t[0].numel() ** 0.5
        return v.to(device if device is not None else t.device)

    return (mk(m), [mk(c) for c in cs], [mk(k) for k in ks], mk(d))
