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
    return torch.randn(B, 3, H, W, generator=g)
