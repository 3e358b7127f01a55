"""Block-by-block comparison of the training-mode backbone forward (GPU) with the CPU oracle (bn_train)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_ours, perturb_, rel_l2  # noqa: E402
from oracle import prn_oracle as O  # noqa: E402
from planerecnet_b200.models.dcn import DeformableConv2d  # noqa: E402


def main():
    preset = "PlaneRecNet_50_config"
    B, H, W = 2, 128, 160
    torch.manual_seed(0)
    net = build_ours(preset)
    perturb_(net)
    net.train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    o = O.Oracle(sd, preset, bn_train=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 3, H, W, generator=g)
    netc = net.cuda()
    eng = netc.train_engine
    eng.reset()
    bb = netc.backbone

    def nchw(t, c):
        return eng.to_nchw(t, c).float().cpu()

    with torch.no_grad():
        # --- stem
        xr = F.relu(o._bn(o._conv(x, "backbone.conv1", 2, 3), "backbone.bn1", 1e-5))
        xr = F.max_pool2d(xr, 3, 2, 1)
        # ours: replicate backbone_t piecewise
        import ctypes as C
        from planerecnet_b200 import ops, _lib as L
        xc = x.cuda()
        a = eng._empty(B, H // 2, W // 2, 192)
        eng._call(eng.lib.prn_stem_im2col, C.c_void_p(xc.data_ptr()), C.c_void_p(a.data_ptr()), B, H, W, eng.dt, eng._st())
        wk = bb.conv1.weight.detach().float().permute(0, 2, 3, 1).reshape(64, 147)
        wk = F.pad(wk, (0, 192 - 147)).to(eng.tdt).contiguous().cuda()
        stats = eng._zeros(64, 2)
        y = eng._empty(B, H // 2, W // 2, 64)
        ops.conv2d(a, wk, batch=B, h_in=H // 2, w_in=W // 2, ksize=1, out16=y, stats=stats, stats_cg=0, dtype=eng.dt)
        y._prn_stats = stats
        yref = o._conv(x, "backbone.conv1", 2, 3)
        print("stem conv  ", rel_l2(nchw(y, 64), yref))
        s_ref = torch.stack([yref.sum((0, 2, 3)), (yref * yref).sum((0, 2, 3))], 1)
        print("stem stats ", rel_l2(stats.cpu(), s_ref))
        t = eng.maxpool_t(eng.t_bn(y, bb.bn1, True))
        print("stem out   ", rel_l2(nchw(t, 64), xr))
        for s, layer in enumerate(bb.layers):
            for b, blk in enumerate(layer):
                prefix = f"backbone.layers.{s}.{b}"
                stride = 2 if (b == 0 and s > 0) else 1
                # feed OUR block with the oracle's input to isolate per-block error
                tin = eng.to_nhwc(xr.cuda())
                a1 = eng.conv_bn(tin, blk.conv1, blk.bn1, True)
                r1 = F.relu(o._bn(o._conv(xr, prefix + ".conv1"), prefix + ".bn1", 1e-5))
                msg = f"{prefix}: a1 {rel_l2(nchw(a1, blk.conv1.out_channels), r1):.2e}"
                if isinstance(blk.conv2, DeformableConv2d):
                    y2 = eng.dcn_t(a1, blk.conv2)
                    r2 = o._dcn(r1, prefix + ".conv2", stride)
                    msg += f" dcn {rel_l2(nchw(y2, r2.shape[1]), r2):.2e}"
                    # same DCN fed with the oracle's a1
                    y2b = eng.dcn_t(eng.to_nhwc(r1.cuda()), blk.conv2)
                    msg += f" dcn(ref in) {rel_l2(nchw(y2b, r2.shape[1]), r2):.2e}"
                    a2 = eng.t_bn(y2, blk.bn2, True)
                else:
                    a2 = eng.conv_bn(a1, blk.conv2, blk.bn2, True)
                    r2 = o._conv(r1, prefix + ".conv2", stride, 1)
                rr2 = F.relu(o._bn(r2, prefix + ".bn2", 1e-5))
                msg += f" a2 {rel_l2(nchw(a2, rr2.shape[1]), rr2):.2e}"
                out = eng.bottleneck_t(tin, blk)
                ref = o._bottleneck(xr, prefix, stride, o.flags[s][b], b == 0)
                msg += f" out {rel_l2(nchw(out, ref.shape[1]), ref):.2e}"
                print(msg, flush=True)
                xr = ref


if __name__ == "__main__":
    main()
