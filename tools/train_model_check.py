"""Whole-model training step on the GPU against autograd through the CPU oracle (batch-statistics BatchNorm):
prints per-parameter-family gradient agreement.  Usage: python tools/train_model_check.py [preset] [B H W] [stage]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import build_ours, perturb_, rel_l2  # noqa: E402
from oracle import prn_oracle as O  # noqa: E402


def reference_grads(net, preset, x, cots, stage="full", bn_train=True):
    o = O.Oracle(net.state_dict(), preset, bn_train=bn_train)
    for k, v in o.sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    if stage == "backbone":
        outs = o.backbone(x)
        flat = list(outs)
    else:
        mask, cate, kern, depth = o.forward_dense(x)
        flat = [mask] + list(cate) + list(kern) + [depth]
    loss = sum((0.5 * c * a * a).sum() for a, c in zip(flat, cots))     # cots = per-output weights of a quadratic loss
    loss.backward()
    return {k: v.grad for k, v in o.sd.items() if v.is_floating_point() and v.grad is not None}, [f.detach() for f in flat]


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "PlaneRecNet_50_config"
    B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 128, 160)
    stage = sys.argv[5] if len(sys.argv) > 5 else "full"
    prec = sys.argv[6] if len(sys.argv) > 6 else "bf16"
    bn_mode = sys.argv[7] if len(sys.argv) > 7 else "train"
    print(f"== {preset} {B}x{H}x{W} stage={stage} precision={prec} bn={bn_mode}")
    torch.manual_seed(0)
    net = build_ours(preset)
    perturb_(net)
    if len(sys.argv) > 8 and sys.argv[8] == "cond":
        # residual-dominant regime (like a trained ResNet): train-mode BatchNorm on a random init is otherwise chaotic
        # (1e-3 input noise -> 100 % change at C5 in the fp32 oracle itself), which makes element-wise parity meaningless
        with torch.no_grad():
            for k, v in net.state_dict().items():
                if k.endswith("bn3.weight"):
                    v.mul_(0.1)
    net.train()
    if bn_mode == "frozen":     # running statistics, affine parameters still trainable
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 3, H, W, generator=g)
    netc = net.cuda()
    if prec != "bf16":
        from planerecnet_b200.train_engine import TrainEngine
        netc._train_engine = TrainEngine(prec)
    eng = netc.train_engine
    t0 = time.time()
    if stage == "backbone":
        eng.reset()
        outs16 = eng.backbone_t(x.cuda(), netc.backbone)
        outs = [eng.to_nchw(o, c) for o, c in zip(outs16, netc.backbone.channels)]
        cots = [1.0 / o[0].numel() ** 0.5 for o in outs]
        for o16, o, c in zip(outs16, outs, cots):
            eng._set_grad(o16, eng.to_nhwc(o * c, c_pad=o16.shape[-1]))
        grads = eng.backward()
        ours = {n: grads[id(p)].cpu() for n, p in netc.named_parameters() if id(p) in grads}
    else:
        outs_t = netc(x.cuda())
        outs = [outs_t[0]] + list(outs_t[1]) + list(outs_t[2]) + [outs_t[3]]
        cots = [1.0 / o[0].numel() ** 0.5 for o in outs]
        loss = sum((0.5 * c * a * a).sum() for a, c in zip(outs, cots))
        loss.backward()
        ours = {n: p.grad.float().cpu() for n, p in netc.named_parameters() if p.grad is not None}
    torch.cuda.synchronize()
    print(f"ours: {time.time() - t0:.1f}s, {len(ours)} parameter gradients, launches {eng.launches}", flush=True)
    net_cpu_sd = {k: v.detach().cpu() for k, v in netc.state_dict().items()}

    class _N:   # oracle takes a state_dict holder; BN running stats were already updated by our step (unused in bn_train mode)
        def state_dict(self):
            return net_cpu_sd

    t0 = time.time()
    ref, ref_outs = reference_grads(_N(), preset, x, cots, stage, bn_train=(bn_mode == 'train'))
    print(f"oracle fwd+bwd: {time.time() - t0:.1f}s", flush=True)
    for i, (a, b) in enumerate(zip(outs, ref_outs)):
        print(f"  out[{i}] {tuple(b.shape)} rel_l2 {rel_l2(a.float().cpu(), b):.3e}")
    fam = {}
    worst = []
    gmax = max(float(v.norm()) for v in ref.values())
    for k, gr in ref.items():
        if float(gr.norm()) < 1e-5 * gmax:      # exactly-zero gradients (conv bias in front of a batch-stat BatchNorm)
            continue
        if k not in ours:
            if float(gr.abs().max()) > 0:
                print("  MISSING gradient for", k)
            continue
        a, b = ours[k].double().flatten(), gr.double().flatten()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        rl = float((a - b).norm() / (b.norm() + 1e-30))
        f = ".".join(k.split(".")[:2])
        fam.setdefault(f, []).append((cos, rl, k, float(b.norm())))
        worst.append((cos, rl, k, float(b.norm())))
    for f, v in sorted(fam.items()):
        print(f"  {f:28s} n={len(v):3d}  cos min {min(c for c, *_ in v):.4f} mean {sum(c for c, *_ in v) / len(v):.4f}   "
              f"rel_l2 max {max(r for _, r, *_ in v):.3f} mean {sum(r for _, r, *_ in v) / len(v):.3f}")
    worst.sort()
    print("  worst 12:")
    for cos, rl, k, nb in worst[:12]:
        print(f"    cos {cos:.4f} rel_l2 {rl:.3f} |ref| {nb:.3e}  {k}")
    a = torch.cat([ours[k].double().flatten() for k in ref if k in ours])
    b = torch.cat([ref[k].double().flatten() for k in ref if k in ours])
    print(f"  ALL: cos {float((a @ b) / (a.norm() * b.norm())):.5f} rel_l2 {float((a - b).norm() / b.norm()):.4f}")


if __name__ == "__main__":
    main()
