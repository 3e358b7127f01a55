"""Shared conv parity cases: the sm_100a implicit-GEMM kernel (through the C ABI) against plain fp32
torch CPU operators on the same 16-bit-rounded inputs.  Used by tests/test_conv_gpu.py and
tools/gpu_selftest.py."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from planerecnet_b200 import _lib as L  # noqa: E402
from planerecnet_b200 import ops  # noqa: E402
from oracle import prn_oracle as O  # noqa: E402

# name -> kwargs
CASES = {
    "1x1_c64_n64_m128": dict(B=1, H=8, W=16, C=64, N=64, k=1),
    "1x1_c128_n128": dict(B=1, H=16, W=16, C=128, N=128, k=1),
    "1x1_c256_n512_multitile": dict(B=2, H=30, W=40, C=256, N=512, k=1, bias=True, act="relu"),
    "3x3_c64_n64": dict(B=1, H=16, W=16, C=64, N=64, k=3, pad=1),
    "3x3_c256_n256_ragged": dict(B=2, H=15, W=20, C=256, N=256, k=3, pad=1, bias=True, residual=True, act="relu"),
    "bisect_nores": dict(B=2, H=15, W=20, C=256, N=256, k=3, pad=1, bias=True, act="relu"),
    "bisect_even_res": dict(B=2, H=16, W=16, C=256, N=256, k=3, pad=1, residual=True),
    "bisect_c128_res": dict(B=2, H=15, W=20, C=128, N=128, k=3, pad=1, residual=True),
    "bisect_1x1_k2304": dict(B=2, H=15, W=20, C=2304, N=256, k=1),
    "3x3_s2_c128_n128": dict(B=2, H=30, W=40, C=128, N=128, k=3, pad=1, stride=2, bias=True),
    "1x1_s2_c256_n512": dict(B=2, H=30, W=40, C=256, N=512, k=1, stride=2),
    "3x3_reflect_c256_n128": dict(B=2, H=15, W=20, C=256, N=128, k=3, pad=1, reflect=True, bias=True, act="relu"),
    "3x3_up2_reflect_concat": dict(B=2, H=15, W=20, C=128, C1=128, N=128, k=3, pad=1, reflect=True, up=2, bias=True, act="relu"),
    "3x3_n2_f32out": dict(B=2, H=24, W=24, C=256, N=2, k=3, pad=1, bias=True, out32=True),
    "3x3_n1_softplus": dict(B=1, H=24, W=32, C=64, N=1, k=3, pad=1, reflect=True, bias=True, out32=True, act="softplus"),
    "3x3_gnstats_cg8": dict(B=3, H=24, W=24, C=256, N=256, k=3, pad=1, stats_cg=8),
    "3x3_gnstats_cg4_ragged": dict(B=3, H=15, W=20, C=256, N=128, k=3, pad=1, stats_cg=4),
    "1x1_bnstats": dict(B=2, H=15, W=20, C=128, N=64, k=1, stats_cg=0, bnstats=True),
    "3x3_big_c256_n256": dict(B=2, H=60, W=80, C=256, N=256, k=3, pad=1, bias=True, act="relu"),
    "1x1_k3776_n256": dict(B=1, H=30, W=40, C=3776, N=256, k=1, bias=True),
    "grouped_sigmoid_avg4": dict(B=2, H=40, W=30, C=128, N=400, k=1, grouped=True, act="sigmoid_avg4", out32=True),
    "dcn_s1_c128": dict(B=2, H=15, W=20, C=128, N=128, k=3, pad=1, bias=True, dcn=True),
    "dcn_s2_c256": dict(B=2, H=30, W=40, C=256, N=256, k=3, pad=1, stride=2, bias=True, dcn=True),
    "offmask_conv": dict(B=2, H=15, W=20, C=128, N=27, k=3, pad=1, bias=True, act="offmask", out32=True),
    "f16_3x3_c128": dict(B=2, H=15, W=20, C=128, N=128, k=3, pad=1, bias=True, act="relu", dtype="f16"),
    # ---- TMA halo-tile kernel (3x3 / stride 1 / pad 1): patch geometry, padding modes, concat, sub-pixel deconv
    "3x3_halo_odd_sizes": dict(B=3, H=21, W=13, C=64, N=64, k=3, pad=1, bias=True),
    "3x3_halo_120x160_c64_n64": dict(B=1, H=120, W=160, C=64, N=64, k=3, pad=1, bias=True, act="relu", dtype="f16"),
    "3x3_halo_concat_c64_c128": dict(B=2, H=30, W=40, C=64, C1=128, N=128, k=3, pad=1, bias=True, act="relu"),
    "3x3_halo_n512": dict(B=2, H=15, W=20, C=128, N=512, k=3, pad=1, bias=True),
    "3x3_reflect_big": dict(B=2, H=33, W=41, C=128, N=128, k=3, pad=1, reflect=True, bias=True),
    "3x3_reflect_concat": dict(B=2, H=16, W=24, C=128, C1=128, N=128, k=3, pad=1, reflect=True, bias=True, act="relu"),
    "3x3_clamp_c128": dict(B=2, H=15, W=20, C=128, N=128, k=3, pad=1, clamp=True, bias=True),
    "3x3_subpixel_up2_reflect_n64": dict(B=2, H=15, W=20, C=128, C1=128, N=64, k=3, pad=1, reflect=True, up=2, bias=True, act="relu",
                                         subpixel=True),
    "3x3_subpixel_up2_reflect_n128_f16": dict(B=2, H=30, W=24, C=256, N=128, k=3, pad=1, reflect=True, up=2, bias=True, act="relu",
                                              subpixel=True, dtype="f16"),
    "3x3_halo_pad2_full_correlation": dict(B=2, H=15, W=20, C=128, N=128, k=3, pad=2, bias=True),
    "3x3_halo_kernelpred_rows": dict(B=2, H=24, W=24, C=256, N=128, k=3, pad=1, bias=True, out32=True),
    "1x1_residual_relu_n1024": dict(B=2, H=30, W=40, C=256, N=1024, k=1, bias=True, residual=True, act="relu"),
    "1x1_c64_n256_big": dict(B=2, H=60, W=80, C=64, N=256, k=1, bias=True, act="relu", dtype="f16"),
}

ACTS = {None: L.ACT_NONE, "relu": L.ACT_RELU, "sigmoid": L.ACT_SIGMOID, "softplus": L.ACT_SOFTPLUS,
        "offmask": L.ACT_DCN_OFFMASK, "sigmoid_avg4": L.ACT_SIGMOID_AVG4}


def run_case(name, seed=0, verbose=False):
    """Returns (max_abs_err, ref_scale, details).  Raises on mismatch."""
    c = dict(CASES[name])
    g = torch.Generator().manual_seed(seed)
    B, H, W, Cc, N, k = c["B"], c["H"], c["W"], c["C"], c["N"], c["k"]
    C1 = c.get("C1", 0)
    stride, pad, up = c.get("stride", 1), c.get("pad", 0), c.get("up", 1)
    dt = L.PRN_F16 if c.get("dtype") == "f16" else L.PRN_BF16
    tdt = ops.torch_dtype(dt)
    dev = "cuda"

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(tdt)

    x0 = rnd(B, H, W, Cc)
    x1 = rnd(B, H, W, C1) if C1 else None
    ctot = Cc + C1
    grouped = c.get("grouped", False)
    if grouped:
        w_all = rnd(B * N, ctot, 1, 1, scale=0.2)          # per-image weights
    else:
        w_all = rnd(N, ctot, k, k, scale=1.0 / (ctot * k * k) ** 0.5 * 2)
    n_pad = ops.round_up(N, 16)
    bias = torch.randn(N, generator=g) if c.get("bias") else None
    Ho = (H * up + 2 * pad - k) // stride + 1
    Wo = (W * up + 2 * pad - k) // stride + 1
    M = B * Ho * Wo
    res = rnd(M, n_pad) if c.get("residual") else None
    offmask = None
    if c.get("dcn"):
        off = torch.randn(B, 18, Ho, Wo, generator=g) * 2.0
        off[:, :, 0, 0] += 30.0   # push some samples far out of the image
        off[:, :, -1, -1] -= 30.0
        msk = torch.rand(B, 9, Ho, Wo, generator=g) * 2
        offmask = torch.zeros(M, 32)
        offmask[:, :18] = off.permute(0, 2, 3, 1).reshape(M, 18)
        offmask[:, 18:27] = msk.permute(0, 2, 3, 1).reshape(M, 9)

    # ---------------- reference (fp32 CPU on the rounded operands)
    xin = torch.cat([x0] + ([x1] if C1 else []), dim=-1).float().permute(0, 3, 1, 2)
    wf = w_all.float()
    if grouped:
        ref = torch.stack([F.conv2d(xin[b:b + 1], wf[b * N:(b + 1) * N])[0] for b in range(B)])
    elif c.get("dcn"):
        ref = O.deform_conv2d(xin, off, wf, bias, stride, pad, msk)
    else:
        xi = xin
        if up == 2:
            xi = F.interpolate(xi, scale_factor=2, mode="nearest")
        if c.get("reflect"):
            xi = F.pad(xi, (pad, pad, pad, pad), mode="reflect")
            ref = F.conv2d(xi, wf, bias, stride, 0)
        elif c.get("clamp"):
            xi = F.pad(xi, (pad, pad, pad, pad), mode="replicate")
            ref = F.conv2d(xi, wf, bias, stride, 0)
        else:
            ref = F.conv2d(xi, wf, bias, stride, pad)
    ref = ref.permute(0, 2, 3, 1).reshape(M, N)          # [M, N]
    if res is not None:
        ref = ref + res.float()[:, :N]
    pre_act = ref.clone()
    act = c.get("act")
    if act == "relu":
        ref = ref.relu()
    elif act == "softplus":
        ref = F.softplus(ref)
    elif act == "offmask":
        ref = torch.cat([ref[:, :18].clamp(-5.0, 5.0), 2 * torch.sigmoid(ref[:, 18:27])], 1)
    elif act == "sigmoid_avg4":
        ref = torch.sigmoid(ref).reshape(M // 4, 4, N).mean(1)

    # ---------------- device
    subpixel = c.get("subpixel", False)
    if grouped:
        wp = w_all.reshape(B * N, ctot).contiguous().to(dev)
    elif subpixel:
        # Upsample(x2) -> ReflectionPad2d(1) -> conv3x3 evaluated at the low resolution: phase-combined weights, replicate
        # padding, pixel-shuffled store (the reference above is the plain upsample + reflect + conv)
        splits = [(Cc, Cc)] + ([(C1, C1)] if C1 else [])
        wp = ops.pack_conv_weight(ops.subpixel_weights(w_all.float()), splits, 4 * N, dt).to(dev)
    else:
        splits = [(Cc, Cc)] + ([(C1, C1)] if C1 else [])
        wp = ops.pack_conv_weight(w_all, splits, n_pad, dt).to(dev)
    m_out = M // 4 if act == "sigmoid_avg4" else M
    out16 = torch.full((m_out, n_pad), float("nan"), dtype=tdt, device=dev)
    out32 = torch.full((m_out, n_pad), float("nan"), dtype=torch.float32, device=dev) if c.get("out32") else None
    stats = None
    if "stats_cg" in c:
        nstat = (B * (n_pad // c["stats_cg"])) if c["stats_cg"] else n_pad
        stats = torch.zeros(nstat, 2, device=dev)
    if subpixel:
        assert N % 32 == 0 and up == 2 and c.get("reflect")
        ops.conv2d(x0.to(dev), wp, batch=B, h_in=H, w_in=W, ksize=3, stride=1, pad=1, pad_mode=L.PAD_CLAMP, upsample=1,
                   src1=x1.to(dev) if C1 else None, bias=bias.repeat(4).to(dev) if bias is not None else None, act=ACTS[act],
                   out16=out16, ld_out16=n_pad, n_pad=4 * N, dtype=dt, shuffle_n=N)
    else:
      ops.conv2d(x0.to(dev), wp, batch=B, h_in=H, w_in=W, ksize=k, stride=stride, pad=pad,
               pad_mode=L.PAD_REFLECT if c.get("reflect") else (L.PAD_CLAMP if c.get("clamp") else L.PAD_ZERO), upsample=up,
               src1=x1.to(dev) if C1 else None, bias=ops.pad_vec(bias, n_pad).to(dev) if bias is not None else None,
               residual=res.to(dev) if res is not None else None, act=ACTS[act], act_param=5.0,
               out16=out16, out32=out32, stats=stats, stats_cg=c.get("stats_cg", 0),
               dcn_offmask=offmask.to(dev) if offmask is not None else None,
               n_pad=n_pad if not grouped else N, w_group_rows=N if grouped else 0, dtype=dt)
    torch.cuda.synchronize()
    ncols = 27 if act == "offmask" else N
    got16 = out16.float().cpu()[:, :ncols]
    scale = ref.abs().max().item() + 1e-6
    # 16-bit output: one rounding of the result (2^-8 bf16 / 2^-11 f16 relative) + fp32 accumulation noise
    eps16 = (2.0 ** -8 if dt == L.PRN_BF16 else 2.0 ** -11)
    err16 = (got16 - ref).abs().max().item()
    tol16 = eps16 * scale + 2e-3 * scale
    details = {"M": M, "N": N, "K": ctot * k * k, "err16": err16, "tol16": tol16, "scale": scale}
    assert torch.isfinite(got16).all(), f"{name}: non-finite values in 16-bit output"
    assert err16 <= tol16, f"{name}: 16-bit output max err {err16:.4g} > tol {tol16:.4g} (scale {scale:.3g})"
    if out32 is not None:
        got32 = out32.cpu()[:, :ncols]
        err32 = (got32 - ref).abs().max().item()
        details["err32"] = err32
        assert err32 <= 2e-3 * scale, f"{name}: fp32 output max err {err32:.4g} (scale {scale:.3g})"
    if stats is not None:
        st = stats.cpu()
        if c["stats_cg"]:
            cg = c["stats_cg"]
            pa = pre_act.reshape(B, Ho * Wo, N // cg, cg)
            s1 = pa.sum((1, 3)).reshape(-1)
            s2 = (pa * pa).sum((1, 3)).reshape(-1)
        else:
            s1, s2 = pre_act.sum(0), (pre_act * pre_act).sum(0)
        e1 = (st[:len(s1), 0] - s1).abs().max().item() / (s1.abs().max().item() + 1e-6)
        e2 = (st[:len(s2), 1] - s2).abs().max().item() / (s2.abs().max().item() + 1e-6)
        details.update(stat_err1=e1, stat_err2=e2)
        assert e1 < 5e-3 and e2 < 5e-3, f"{name}: statistics mismatch {e1:.3g} {e2:.3g}"
    return details
