"""Local fwd+bwd check of single bottlenecks (batch-statistics BatchNorm) against autograd through the CPU oracle,
fed with identical 16-bit-rounded inputs and cotangents."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_ours, perturb_, rel_l2  # noqa: E402
from oracle import prn_oracle as O  # noqa: E402


def check_block(netc, sd, preset, s, b, B=2, H=16, W=20, prec="bf16", verbose=True):
    from planerecnet_b200.train_engine import TrainEngine
    eng = TrainEngine(prec)
    blk = netc.backbone.layers[s][b]
    prefix = f"backbone.layers.{s}.{b}"
    cin = blk.conv1.in_channels
    stride = 2 if (b == 0 and s > 0) else 1
    g = torch.Generator().manual_seed(100 * s + b)
    x = torch.randn(B, cin, H, W, generator=g).relu().to(eng.tdt).float()
    o = O.Oracle(sd, preset, bn_train=True)
    for k, v in o.sd.items():
        if k.startswith(prefix) and v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = o._bottleneck(xr, prefix, stride, o.flags[s][b], b == 0)
    # quadratic loss: the cotangent vanishes where the ReLU output does, so that sign flips of near-zero outputs
    # (16-bit rounding) do not dominate the parameter sums the way they do under random cotangents
    wgt = 1.0 / ref[0].numel() ** 0.5
    (0.5 * wgt * ref * ref).sum().backward()

    eng.reset()
    tin = eng.to_nhwc(x.cuda())
    out = eng.bottleneck_t(tin, blk)
    eng._set_grad(out, eng.to_nhwc(eng.to_nchw(out, ref.shape[1]) * wgt))
    grads = eng.backward_keep_inputs(tin)
    dx = eng.to_nchw(grads["dx"], cin).cpu()
    res = {"out": rel_l2(eng.to_nchw(out, ref.shape[1]).cpu(), ref.detach()), "dx": rel_l2(dx, xr.grad)}
    gmax = max(float(o.sd[prefix + "." + n].grad.norm()) for n, _ in blk.named_parameters() if o.sd[prefix + "." + n].grad is not None)
    for n, p in blk.named_parameters():
        gr = o.sd[prefix + "." + n].grad
        if gr is None or float(gr.norm()) < 1e-5 * gmax:
            continue
        res[n] = rel_l2(grads["params"][id(p)].cpu(), gr)
    if verbose:
        print(f"{prefix} [{prec}] " + " ".join(f"{k}={v:.2e}" for k, v in res.items()), flush=True)
    return res


def main():
    preset = "PlaneRecNet_50_config"
    torch.manual_seed(0)
    net = build_ours(preset)
    perturb_(net)
    net.train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    netc = net.cuda()
    for prec in ("bf16", "f16"):
        for (s, b) in ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (3, 0), (3, 2)):
            check_block(netc, sd, preset, s, b, prec=prec)


if __name__ == "__main__":
    main()
