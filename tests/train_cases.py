"""Shared parity cases for the training-step kernels (through the C ABI) against torch autograd in fp32 on the
CPU, on identical 16-bit-rounded operands.  Used by tests/test_train_ops_gpu.py and tools/gpu_selftest.py.

Tolerances (stated per check): weight gradients are fp32 sums of exact 16-bit products -> 2e-3 of the largest
reference entry; 16-bit outputs carry one extra rounding (2^-8 bf16, 2^-11 f16) of the result."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from planerecnet_b200 import _lib as L  # noqa: E402
from planerecnet_b200 import ops  # noqa: E402

DEV = "cuda"

WGRAD_CASES = {
    "wg_1x1_c64_n64": dict(B=1, H=8, W=16, C=64, N=64, k=1),
    "wg_1x1_c256_n512": dict(B=2, H=30, W=40, C=256, N=512, k=1),
    "wg_3x3_c64_n64": dict(B=1, H=16, W=16, C=64, N=64, k=3, pad=1),
    "wg_3x3_c256_n256_ragged": dict(B=2, H=15, W=20, C=256, N=256, k=3, pad=1),
    "wg_3x3_c128_n128_big": dict(B=2, H=60, W=80, C=128, N=128, k=3, pad=1),
    "wg_3x3_s2_c128_n128": dict(B=2, H=30, W=40, C=128, N=128, k=3, pad=1, stride=2),
    "wg_1x1_s2_c256_n512": dict(B=2, H=30, W=40, C=256, N=512, k=1, stride=2),
    "wg_3x3_reflect_c256_n128": dict(B=2, H=15, W=20, C=256, N=128, k=3, pad=1, reflect=True),
    "wg_3x3_up2_reflect_concat": dict(B=2, H=15, W=20, C=128, C1=128, N=128, k=3, pad=1, reflect=True, up=2),
    "wg_3x3_n2": dict(B=2, H=24, W=24, C=256, N=2, k=3, pad=1),
    "wg_3x3_n27": dict(B=2, H=15, W=20, C=128, N=27, k=3, pad=1),
    "wg_1x1_k3776_n256": dict(B=1, H=30, W=40, C=3776, N=256, k=1),
    "wg_1x1_c192_n64_stem": dict(B=1, H=48, W=64, C=192, N=64, k=1),
    "wg_3x3_c320_n256": dict(B=2, H=16, W=16, C=320, N=256, k=3, pad=1),
    "wg_f16_3x3_c128": dict(B=2, H=15, W=20, C=128, N=128, k=3, pad=1, dtype="f16"),
}


def _dt(c):
    dt = L.PRN_F16 if c.get("dtype") == "f16" else L.PRN_BF16
    return dt, ops.torch_dtype(dt)


def run_wgrad_case(name, seed=0, flags=0):
    c = dict(WGRAD_CASES[name])
    g = torch.Generator().manual_seed(seed)
    B, H, W, Cc, N, k = c["B"], c["H"], c["W"], c["C"], c["N"], c["k"]
    C1 = c.get("C1", 0)
    stride, pad, up = c.get("stride", 1), c.get("pad", 0), c.get("up", 1)
    dt, tdt = _dt(c)
    x0 = torch.randn(B, H, W, Cc, generator=g).to(tdt)
    x1 = torch.randn(B, H, W, C1, generator=g).to(tdt) if C1 else None
    ctot = Cc + C1
    Ho = (H * up + 2 * pad - k) // stride + 1
    Wo = (W * up + 2 * pad - k) // stride + 1
    M = B * Ho * Wo
    ld_dy = ops.round_up(N, 16) if N < 64 else N
    dy = torch.zeros(M, ld_dy, dtype=tdt)
    dy[:, :N] = (torch.randn(M, N, generator=g) * 0.5).to(tdt)

    # reference: autograd of the fp32 convolution on the rounded operands
    xin = torch.cat([x0] + ([x1] if C1 else []), dim=-1).float().permute(0, 3, 1, 2)
    w = torch.zeros(N, ctot, k, k, requires_grad=True)
    xi = xin
    if up == 2:
        xi = F.interpolate(xi, scale_factor=2, mode="nearest")
    if c.get("reflect"):
        xi = F.pad(xi, (pad, pad, pad, pad), mode="reflect")
        y = F.conv2d(xi, w, None, stride, 0)
    else:
        y = F.conv2d(xi, w, None, stride, pad)
    y.backward(dy[:, :N].float().reshape(B, Ho, Wo, N).permute(0, 3, 1, 2))
    ref = w.grad

    dw = torch.zeros(ops.round_up(N, 4), k * k * ctot, dtype=torch.float32, device=DEV)
    ops.conv2d_wgrad(x0.to(DEV), dy.to(DEV), dw, batch=B, h_in=H, w_in=W, n=N, ksize=k, stride=stride, pad=pad,
                     pad_mode=L.PAD_REFLECT if c.get("reflect") else L.PAD_ZERO, upsample=up,
                     src1=x1.to(DEV) if C1 else None, dtype=dt, flags=flags)
    torch.cuda.synchronize()
    splits = [(Cc, Cc)] + ([(C1, C1)] if C1 else [])
    got = ops.unpack_wgrad(dw, (N, ctot, k, k), splits).cpu()
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    assert torch.isfinite(got).all(), f"{name}: non-finite weight gradient"
    assert err <= 2e-3 * scale, f"{name}: wgrad max err {err:.4g} > {2e-3 * scale:.4g} (scale {scale:.3g})"
    return {"M": M, "N": N, "K": ctot * k * k, "err": err, "scale": scale}


DGRAD_CASES = {
    "dg_1x1_c256_n64": dict(B=2, H=15, W=20, C=256, N=64, k=1),
    "dg_3x3_c128_n256": dict(B=2, H=15, W=20, C=128, N=256, k=3, pad=1),
    "dg_3x3_c64_n64_acc": dict(B=1, H=16, W=16, C=64, N=64, k=3, pad=1, accumulate=True),
}


def run_dgrad_case(name, seed=0):
    """Input gradient = prn_conv2d_fwd over dY with flipped/transposed weights (stride 1)."""
    c = dict(DGRAD_CASES[name])
    g = torch.Generator().manual_seed(seed)
    B, H, W, Cc, N, k = c["B"], c["H"], c["W"], c["C"], c["N"], c["k"]
    pad = c.get("pad", 0)
    dt, tdt = _dt(c)
    w = (torch.randn(N, Cc, k, k, generator=g) / (N * k * k) ** 0.5).to(tdt)
    dy = torch.randn(B, H, W, N, generator=g).to(tdt)
    prev = torch.randn(B, H, W, Cc, generator=g).to(tdt) if c.get("accumulate") else None
    x = torch.zeros(B, Cc, H, W, requires_grad=True)
    F.conv2d(x, w.float(), None, 1, pad).backward(dy.float().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1)
    if prev is not None:
        ref = ref + prev.float()
    wp = ops.pack_dgrad_weight(w.float(), dtype=dt).to(DEV)
    out = torch.empty(B, H, W, wp.shape[0], dtype=tdt, device=DEV)
    ops.conv2d(dy.to(DEV), wp, batch=B, h_in=H, w_in=W, ksize=k, stride=1, pad=k - 1 - pad,
               residual=prev.to(DEV) if prev is not None else None, out16=out, dtype=dt)
    torch.cuda.synchronize()
    got = out.float().cpu()[..., :Cc]
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    eps16 = 2.0 ** -8 if dt == L.PRN_BF16 else 2.0 ** -11
    assert err <= (eps16 + 2e-3) * scale, f"{name}: dgrad max err {err:.4g} (scale {scale:.3g})"
    return {"err": err, "scale": scale}


def run_bn_case(seed=0, rows=2 * 15 * 20, Cc=320, residual=True, relu=True, dtype="bf16"):
    """Train-mode BatchNorm forward (statistics from prn_conv2d_fwd-style sums) and backward vs autograd."""
    g = torch.Generator().manual_seed(seed)
    dt, tdt = _dt({"dtype": dtype})
    x = (torch.randn(rows, Cc, generator=g) * 2 + 0.5).to(tdt)
    res = torch.randn(rows, Cc, generator=g).to(tdt) if residual else None
    gamma = torch.rand(Cc, generator=g) + 0.5
    beta = torch.randn(Cc, generator=g) * 0.1
    dz = torch.randn(rows, Cc, generator=g).to(tdt)
    eps, mom = 1e-5, 0.1
    rm0, rv0 = torch.randn(Cc, generator=g) * 0.1, torch.rand(Cc, generator=g) + 0.5

    xr = x.float().clone().requires_grad_(True)
    rr = res.float().clone().requires_grad_(True) if residual else None
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    y = F.batch_norm(xr.t().reshape(1, Cc, rows), rm, rv, gr, br, True, mom, eps).reshape(Cc, rows).t()
    if residual:
        y = y + rr
    if relu:
        y = y.relu()
    y.backward(dz.float())

    xd = x.to(DEV)
    stats = torch.stack([xd.float().sum(0), (xd.float() ** 2).sum(0)], 1).contiguous()     # [C,2] as the conv epilogue leaves them
    mi = torch.empty(Cc, 2, device=DEV)
    rmd, rvd = rm0.to(DEV), rv0.to(DEV)
    ops.bn_finalize(stats, mi, rmd, rvd, rows, eps, mom)
    out = torch.empty(rows, Cc, dtype=tdt, device=DEV)
    gd, bd = gamma.to(DEV), beta.to(DEV)
    ops.bn_apply(xd, out, mi, gd, bd, res.to(DEV) if residual else None, relu, dt)
    sums = torch.zeros(Cc, 2, device=DEV)
    dzd = dz.to(DEV)
    ops.chan_reduce(dzd, out if relu else None, xd, mi, sums, dt)
    dx = torch.empty(rows, Cc, dtype=tdt, device=DEV)
    gout = torch.empty(rows, Cc, dtype=tdt, device=DEV) if residual else None
    ops.bn_bwd_apply(dzd, out if relu else None, xd, mi, gd, sums, dx, gout, dt)
    torch.cuda.synchronize()
    eps16 = 2.0 ** -8 if dt == L.PRN_BF16 else 2.0 ** -11

    def chk(what, got, ref, tol):
        scale = ref.abs().max().item() + 1e-6
        err = (got.float().cpu() - ref).abs().max().item()
        assert err <= tol * scale, f"bn {what}: max err {err:.4g} > {tol * scale:.4g}"
        return err / scale

    r = {}
    r["out"] = chk("forward", out, y.detach(), eps16 + 2e-3)
    r["rm"] = chk("running_mean", rmd, rm, 1e-4)
    r["rv"] = chk("running_var", rvd, rv, 1e-4)
    r["dx"] = chk("dx", dx, xr.grad, eps16 + 3e-3)
    r["dgamma"] = chk("dgamma", sums[:, 1], gr.grad, 2e-3)
    r["dbeta"] = chk("dbeta", sums[:, 0], br.grad, 2e-3)
    if residual:
        r["dres"] = chk("dres", gout, rr.grad, eps16 + 1e-4)
    return r


def run_maxpool_bwd_case(seed=0, B=2, H=30, W=40, Cc=64):
    g = torch.Generator().manual_seed(seed)
    tdt = torch.bfloat16
    # few distinct values -> many ties inside windows: the first maximum must take the gradient
    x = torch.randint(0, 4, (B, H, W, Cc), generator=g).to(tdt)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dout = torch.randn(B, Ho, Wo, Cc, generator=g).to(tdt)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.max_pool2d(xr, 3, 2, 1).backward(dout.float().permute(0, 3, 1, 2))
    ref = xr.grad.permute(0, 2, 3, 1)
    din = torch.empty(B, H, W, Cc, dtype=tdt, device=DEV)
    ops.maxpool_bwd(x.to(DEV), dout.to(DEV), din, L.PRN_BF16)
    torch.cuda.synchronize()
    err = (din.float().cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2.0 ** -7 * scale, f"maxpool bwd: max err {err:.4g} (scale {scale:.3g})"
    return {"err": err}


def run_add_cases(seed=0):
    g = torch.Generator().manual_seed(seed)
    tdt = torch.bfloat16
    dst = torch.randn(2, 8, 12, 64, generator=g).to(tdt)
    src = torch.randn(2, 4, 6, 64, generator=g).to(tdt)
    ref = dst.float().clone()
    ref[:, ::2, ::2] += src.float()
    d = dst.to(DEV)
    ops.add_strided(d, src.to(DEV), 2, L.PRN_BF16)
    a32 = torch.randn(2, 8, 12, 64, generator=g)
    o = torch.empty_like(d)
    ops.add_f32(a32.to(DEV), dst.to(DEV), o, L.PRN_BF16)
    o2 = torch.empty_like(d)
    ops.add16(dst.to(DEV), d, o2, L.PRN_BF16)
    rb = torch.empty_like(d)
    ops.relu_bwd(dst.to(DEV), d, rb, L.PRN_BF16)
    torch.cuda.synchronize()
    assert (d.float().cpu() - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()
    r2 = a32 + dst.float()
    assert (o.float().cpu() - r2).abs().max().item() <= 2.0 ** -7 * r2.abs().max().item()
    r3 = dst.float() + d.float().cpu()
    assert (o2.float().cpu() - r3).abs().max().item() <= 2.0 ** -7 * r3.abs().max().item()
    r4 = dst.float() * (d.float().cpu() > 0)
    assert (rb.float().cpu() - r4).abs().max().item() == 0.0
    return {}


DCN_CASES = {
    "dcn_train_s1_c128": dict(B=2, H=15, W=20, C=128, N=128, stride=1),
    "dcn_train_s2_c256": dict(B=2, H=30, W=40, C=256, N=256, stride=2),
}


def run_dcn_case(name, seed=0):
    """prn_dcn_im2col + 1x1 contraction, and prn_dcn_col2im_bwd, against autograd through torchvision's
    deform_conv2d (fp32 CPU) with the clamp / 2*sigmoid pre-activations of models/dcn.py:53-57."""
    from torchvision.ops import deform_conv2d
    c = dict(DCN_CASES[name])
    g = torch.Generator().manual_seed(seed)
    B, H, W, Cc, N, stride = c["B"], c["H"], c["W"], c["C"], c["N"], c["stride"]
    tdt, dt = torch.bfloat16, L.PRN_BF16
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    M = B * Ho * Wo
    bound = max(H, W) / 4.0
    x = torch.randn(B, H, W, Cc, generator=g).to(tdt)
    w = (torch.randn(N, Cc, 3, 3, generator=g) / (Cc * 9) ** 0.5).to(tdt)
    pre = torch.randn(M, 27, generator=g)                    # pre-activation output of the offset/modulator conv
    pre[:, :18] *= 3.0                                       # some offsets beyond the clamp
    pre[0, :18] += 40.0
    dy = torch.randn(M, N, generator=g).to(tdt)

    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    pr = pre.clone().requires_grad_(True)
    off = pr[:, :18].clamp(-bound, bound).reshape(B, Ho, Wo, 18).permute(0, 3, 1, 2)
    msk = (2 * torch.sigmoid(pr[:, 18:])).reshape(B, Ho, Wo, 9).permute(0, 3, 1, 2)
    wr = w.float().clone().requires_grad_(True)
    y = deform_conv2d(xr, off, wr, None, stride=stride, padding=1, mask=msk)
    y.backward(dy.float().reshape(B, Ho, Wo, N).permute(0, 3, 1, 2))

    offmask = torch.zeros(M, 32)
    offmask[:, :18] = pre[:, :18].clamp(-bound, bound)
    offmask[:, 18:27] = 2 * torch.sigmoid(pre[:, 18:])
    xd, omd = x.to(DEV), offmask.to(DEV)
    col = torch.empty(M, 9 * Cc, dtype=tdt, device=DEV)
    ops.dcn_im2col(xd, omd, col, stride, 1, dt)
    wp = ops.pack_conv_weight(w.float(), [(Cc, Cc)], N, dt).to(DEV)               # [N, 9*C] with k = (tap, c)
    out = torch.empty(M, N, dtype=tdt, device=DEV)
    ops.conv2d(col, wp, batch=1, h_in=M, w_in=1, ksize=1, c0=9 * Cc, out16=out, dtype=dt)
    # backward: dcol = dY . W (1x1 contraction with the transposed weights), then the scatter
    wt = wp.t().contiguous()                                                        # [9*C, N]
    dcol = torch.empty(M, 9 * Cc, dtype=tdt, device=DEV)
    dyd = dy.to(DEV)
    ops.conv2d(dyd, wt, batch=1, h_in=M, w_in=1, ksize=1, out16=dcol, dtype=dt)
    dw = torch.zeros(N, 9 * Cc, device=DEV)
    ops.conv2d_wgrad(col, dyd, dw, batch=1, h_in=M, w_in=1, n=N, ksize=1, dtype=dt)
    dx32 = torch.zeros(B, H, W, Cc, device=DEV)
    dpre = torch.empty(M, 64, dtype=tdt, device=DEV)
    ops.dcn_col2im_bwd(xd, omd, dcol, dx32, dpre, stride, 1, bound, dt)
    torch.cuda.synchronize()

    def chk(what, got, ref, tol):
        scale = ref.abs().max().item() + 1e-6
        err = (got.float().cpu() - ref).abs().max().item()
        assert err <= tol * scale, f"{name} {what}: max err {err:.4g} > {tol * scale:.4g}"
        return err / scale

    r = {}
    r["fwd"] = chk("forward", out, y.detach().permute(0, 2, 3, 1).reshape(M, N), 2.0 ** -7 + 4e-3)
    r["dw"] = chk("dw", ops.unpack_wgrad(dw, (N, Cc, 3, 3), [(Cc, Cc)]), wr.grad, 1e-2)
    r["dx"] = chk("dx", dx32, xr.grad.permute(0, 2, 3, 1), 1e-2)
    r["dpre"] = chk("dpre", dpre[:, :27], pr.grad, 2e-2)
    assert float(dpre[:, 27:].float().abs().max()) == 0.0
    return r


def _chk(what, got, ref, tol):
    scale = ref.abs().max().item() + 1e-6
    err = (got.float().cpu() - ref).abs().max().item()
    assert err <= tol * scale, f"{what}: max err {err:.4g} > {tol * scale:.4g} (scale {scale:.3g})"
    return err / scale


def run_gn_case(seed=0, B=3, HW=15 * 20, Cc=256, G=32):
    """GroupNorm(32)+ReLU backward against autograd of F.group_norm."""
    g = torch.Generator().manual_seed(seed)
    tdt, dt = torch.bfloat16, L.PRN_BF16
    cg = Cc // G
    x = (torch.randn(B, HW, Cc, generator=g) * 1.5 + 0.3).to(tdt)
    gamma = torch.rand(Cc, generator=g) + 0.5
    beta = torch.randn(Cc, generator=g) * 0.2
    dz = torch.randn(B, HW, Cc, generator=g).to(tdt)
    eps = 1e-5
    xr = x.float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr.permute(0, 2, 1), G, gr, br, eps).permute(0, 2, 1).relu()
    y.backward(dz.float())
    xd = x.to(DEV)
    xf = xd.float().reshape(B, HW, G, cg)
    stats = torch.stack([xf.sum((1, 3)), (xf * xf).sum((1, 3))], -1).reshape(-1).contiguous()
    out = torch.empty(B, HW, Cc, dtype=tdt, device=DEV)
    gd, bd = gamma.to(DEV), beta.to(DEV)
    import ctypes as C
    L.check(L.lib().prn_groupnorm_apply(C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(stats.data_ptr()),
                                        C.c_void_p(gd.data_ptr()), C.c_void_p(bd.data_ptr()), B, HW, Cc, cg, C.c_float(eps), 1, dt,
                                        L.current_stream()))
    sums_bc = torch.zeros(B, Cc, 2, device=DEV)
    dgb = torch.zeros(Cc, 2, device=DEV)
    dzd = dz.to(DEV)
    ops.gn_bwd_reduce(dzd, out, xd, stats, sums_bc, dgb, cg, eps, dt)
    dx = torch.empty(B, HW, Cc, dtype=tdt, device=DEV)
    ops.gn_bwd_apply(dzd, out, xd, stats, gd, sums_bc, dx, cg, eps, dt)
    torch.cuda.synchronize()
    r = {"out": _chk("gn forward", out, y.detach(), 2.0 ** -8 + 2e-3)}
    r["dx"] = _chk("gn dx", dx, xr.grad, 2.0 ** -8 + 3e-3)
    r["dgamma"] = _chk("gn dgamma", dgb[:, 1], gr.grad, 3e-3)
    r["dbeta"] = _chk("gn dbeta", dgb[:, 0], br.grad, 3e-3)
    return r


def run_resample_bwd_cases(seed=0):
    """2x2 mean, bilinear x2 and arbitrary bilinear resize backward against autograd of F.interpolate."""
    g = torch.Generator().manual_seed(seed)
    tdt, dt = torch.bfloat16, L.PRN_BF16
    r = {}
    B, H, W, Cc = 2, 12, 16, 64
    # 2x2 mean
    dout = torch.randn(B, H // 2, W // 2, Cc, generator=g).to(tdt)
    xr = torch.zeros(B, Cc, H, W, requires_grad=True)
    F.interpolate(xr, scale_factor=0.5, mode="bilinear", align_corners=False).backward(dout.float().permute(0, 3, 1, 2))
    din = torch.empty(B, H, W, Cc, dtype=tdt, device=DEV)
    ops.avgpool2_bwd(dout.to(DEV), din, False, dt)
    prev = torch.randn(B, H, W, Cc, generator=g).to(tdt)
    din2 = prev.to(DEV)
    ops.avgpool2_bwd(dout.to(DEV), din2, True, dt)
    torch.cuda.synchronize()
    ref = xr.grad.permute(0, 2, 3, 1)
    r["avg"] = _chk("avgpool2 bwd", din, ref, 2.0 ** -8)
    r["avg_acc"] = _chk("avgpool2 bwd acc", din2, ref + prev.float(), 2.0 ** -7)
    # bilinear x2
    dout = torch.randn(B, 2 * H, 2 * W, Cc, generator=g).to(tdt)
    xr = torch.zeros(B, Cc, H, W, requires_grad=True)
    F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False).backward(dout.float().permute(0, 3, 1, 2))
    din = torch.empty(B, H, W, Cc, dtype=tdt, device=DEV)
    ops.upsample2x_bwd(dout.to(DEV), din, dt)
    torch.cuda.synchronize()
    r["up2"] = _chk("upsample2x bwd", din, xr.grad.permute(0, 2, 3, 1), 2.0 ** -7)
    # arbitrary resize (down 60x80 -> 40x40 style and up 15x20 -> 16x16 style), dout carries 2 extra coord channels
    for (h, w, S) in ((30, 40, 20), (15, 20, 16), (12, 16, 24)):
        dout = torch.randn(B, S, S, Cc + 64, generator=g).to(tdt)
        xr = torch.zeros(B, Cc, h, w, requires_grad=True)
        F.interpolate(xr, size=(S, S), mode="bilinear", align_corners=False).backward(dout[..., :Cc].float().permute(0, 3, 1, 2))
        din32 = torch.zeros(B, h, w, Cc, device=DEV)
        ops.resize_bilinear_bwd(dout.to(DEV), din32, dt)
        torch.cuda.synchronize()
        r[f"resize_{h}x{w}_{S}"] = _chk("resize bwd", din32, xr.grad.permute(0, 2, 3, 1), 1e-4)
    return r


def run_reflect_dgrad_case(seed=0, up=1, B=2, H=9, W=12, Cc=64, N=128, accumulate=False):
    """Input gradient of [nearest x2 ->] ReflectionPad2d(1) -> conv3x3: contraction with zero padding 2, then the fold."""
    g = torch.Generator().manual_seed(seed)
    tdt, dt = torch.bfloat16, L.PRN_BF16
    w = (torch.randn(N, Cc, 3, 3, generator=g) / (N * 9) ** 0.5).to(tdt)
    He, We = H * up, W * up
    dy = torch.randn(B, He, We, N, generator=g).to(tdt)
    prev = torch.randn(B, H, W, Cc, generator=g).to(tdt) if accumulate else None
    xr = torch.zeros(B, Cc, H, W, requires_grad=True)
    xi = F.interpolate(xr, scale_factor=2, mode="nearest") if up == 2 else xr
    F.conv2d(F.pad(xi, (1, 1, 1, 1), mode="reflect"), w.float()).backward(dy.float().permute(0, 3, 1, 2))
    ref = xr.grad.permute(0, 2, 3, 1)
    if accumulate:
        ref = ref + prev.float()
    wp = ops.pack_dgrad_weight(w.float(), dtype=dt).to(DEV)
    dpad = torch.empty(B, He + 2, We + 2, wp.shape[0], dtype=tdt, device=DEV)
    ops.conv2d(dy.to(DEV), wp, batch=B, h_in=He, w_in=We, ksize=3, stride=1, pad=2, out16=dpad, dtype=dt)
    din = prev.to(DEV) if accumulate else torch.empty(B, H, W, Cc, dtype=tdt, device=DEV)
    ops.reflect_fold(dpad, din, up, accumulate, dt)
    torch.cuda.synchronize()
    return {"err": _chk(f"reflect dgrad up{up}", din, ref, 2.0 ** -6)}


def run_softplus_case(seed=0, rows=1000):
    g = torch.Generator().manual_seed(seed)
    pre = (torch.randn(rows, generator=g) * 3).requires_grad_(True)
    out = F.softplus(pre)
    dout = torch.randn(rows, generator=g)
    out.backward(dout)
    dpre = torch.empty(rows, 64, dtype=torch.bfloat16, device=DEV)
    ops.softplus_bwd_pad(dout.to(DEV), out.detach().to(DEV), dpre, L.PRN_BF16)
    torch.cuda.synchronize()
    assert float(dpre[:, 1:].float().abs().max()) == 0.0
    return {"err": _chk("softplus bwd", dpre[:, 0], pre.grad, 2.0 ** -8 + 1e-4)}


# ------------------------------------------------------------------------------------------ whole blocks / whole model
def _build(preset, cond=False, seed=0):
    from helpers import build_ours, perturb_
    torch.manual_seed(seed)
    net = build_ours(preset)
    perturb_(net)
    if cond:
        # residual-dominant regime (like a trained ResNet).  Train-mode BatchNorm on a random init is chaotic: 1e-3
        # input noise changes C5 by ~100 % in the fp32 oracle itself, which makes element-wise parity meaningless.
        with torch.no_grad():
            for k, v in net.state_dict().items():
                if k.endswith("bn3.weight"):
                    v.mul_(0.1)
    return net


def run_block_check(s, b, prec="bf16", preset="PlaneRecNet_50_config", B=2, H=16, W=20, net=None, sd=None):
    """One bottleneck, batch-statistics BatchNorm, forward + backward, against autograd through the CPU oracle on
    identical 16-bit-rounded inputs.  Loss = quadratic (its cotangent vanishes where the ReLU output does; random
    cotangents make the per-channel sums cancellation-dominated and sign flips of near-zero outputs dominate them).
    Returns {name: rel-L2}."""
    from helpers import rel_l2
    from oracle import prn_oracle as O
    from planerecnet_b200.train_engine import TrainEngine
    if net is None:
        net = _build(preset).train()
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        net = net.cuda()
    eng = TrainEngine(prec)
    blk = net.backbone.layers[s][b]
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
    wgt = 1.0 / ref[0].numel() ** 0.5
    (0.5 * wgt * ref * ref).sum().backward()
    eng.reset()
    tin = eng.to_nhwc(x.cuda())
    out = eng.bottleneck_t(tin, blk)
    eng._set_grad(out, eng.to_nhwc(eng.to_nchw(out, ref.shape[1]) * wgt))
    grads = eng.backward_keep_inputs(tin)
    res = {"out": rel_l2(eng.to_nchw(out, ref.shape[1]).cpu(), ref.detach()),
           "dx": rel_l2(eng.to_nchw(grads["dx"], cin).cpu(), xr.grad)}
    refg = {n: o.sd[prefix + "." + n].grad for n, _ in blk.named_parameters()}
    gmax = max(float(v.norm()) for v in refg.values() if v is not None)
    for n, p in blk.named_parameters():
        gr = refg[n]
        if gr is None or float(gr.norm()) < 1e-5 * gmax:      # conv bias in front of a batch-stat BatchNorm: exactly 0
            continue
        res[n] = rel_l2(grads["params"][id(p)].cpu(), gr)
    return res


def run_model_check(preset="PlaneRecNet_50_config", B=2, H=128, W=160, prec="bf16", bn_mode="train", cond=True):
    """Whole training step through net(x) + loss.backward() against autograd through the CPU oracle.
    Returns dict(outs=[rel-L2 per output], all_cos, all_rel, fam={family: (min cos, mean cos)}, missing=[...])."""
    from helpers import rel_l2
    from oracle import prn_oracle as O
    from planerecnet_b200.train_engine import TrainEngine
    net = _build(preset, cond=cond).train()
    if bn_mode == "frozen":       # running statistics, affine parameters still trainable
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 3, H, W, generator=g)
    netc = net.cuda()
    netc._train_engine = TrainEngine(prec)
    outs_t = netc(x.cuda())
    outs = [outs_t[0]] + list(outs_t[1]) + list(outs_t[2]) + [outs_t[3]]
    wts = [1.0 / o[0].numel() ** 0.5 for o in outs]
    sum((0.5 * c * a * a).sum() for a, c in zip(outs, wts)).backward()
    torch.cuda.synchronize()
    ours = {n: p.grad.float().cpu() for n, p in netc.named_parameters() if p.grad is not None}
    o = O.Oracle(sd0, preset, bn_train=(bn_mode == "train"))
    for k, v in o.sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    mask, cate, kern, depth = o.forward_dense(x)
    flat = [mask] + list(cate) + list(kern) + [depth]
    sum((0.5 * c * a * a).sum() for a, c in zip(flat, wts)).backward()
    ref = {k: v.grad for k, v in o.sd.items() if v.is_floating_point() and v.grad is not None}
    r = {"outs": [rel_l2(a.detach().float().cpu(), b.detach()) for a, b in zip(outs, flat)], "missing": [], "fam": {},
         "launches": netc.train_engine.launches, "state": netc.state_dict(), "sd0": sd0}
    gmax = max(float(v.norm()) for v in ref.values())
    fam, av, bv = {}, [], []
    for k, gr in ref.items():
        if float(gr.norm()) < 1e-5 * gmax:
            continue
        if k not in ours:
            r["missing"].append(k)
            continue
        a, b = ours[k].double().flatten(), gr.double().flatten()
        fam.setdefault(".".join(k.split(".")[:2]), []).append(float((a @ b) / (a.norm() * b.norm() + 1e-30)))
        av.append(a)
        bv.append(b)
    a, b = torch.cat(av), torch.cat(bv)
    r["all_cos"] = float((a @ b) / (a.norm() * b.norm()))
    r["all_rel"] = float((a - b).norm() / b.norm())
    r["fam"] = {f: (min(v), sum(v) / len(v)) for f, v in fam.items()}
    return r


def run_pack_cases():
    """prn_pack_conv_weight / prn_pack_dgrad_weight == the torch reference packers, bit for bit."""
    g = torch.Generator().manual_seed(0)
    for (cout, cin, k, splits, dt) in ((64, 64, 1, [(64, 64)], L.PRN_BF16), (256, 258, 3, [(258, 320)], L.PRN_BF16),
                                       (128, 384, 3, [(256, 256), (128, 128)], L.PRN_BF16), (2, 256, 3, [(256, 256)], L.PRN_F16),
                                       (256, 3728, 1, [(3728, 3776)], L.PRN_BF16), (27, 128, 3, [(128, 128)], L.PRN_BF16)):
        w = torch.randn(cout, cin, k, k, generator=g)
        n_pad = ops.round_up(cout, 16)
        ref = ops.pack_conv_weight(w, splits, n_pad, dt)
        got = ops.pack_conv_weight_dev(w.cuda(), splits, n_pad, dt)
        assert torch.equal(got.cpu().view(torch.int16), ref.view(torch.int16)), ("fwd pack", cout, cin, k)
        lo = 0
        for real, pad in splits:
            cp = ops.round_up(cout, 64)
            refd = ops.pack_dgrad_weight(w[:, lo:lo + real], cout_pad=cp, n_pad=pad, dtype=dt)
            gotd = ops.pack_dgrad_weight_dev(w.cuda(), lo, lo + real, pad, cp, dt)
            assert torch.equal(gotd.cpu().view(torch.int16), refd.view(torch.int16)), ("dgrad pack", cout, cin, k, lo)
            lo += real
    return {}
