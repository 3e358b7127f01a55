"""Parity of the HBM-bound passes (called through the C ABI) against fp32 torch CPU operators on the same
16-bit-rounded inputs.  Tolerance: one 16-bit rounding of the result (2^-8 bf16 / 2^-11 f16, relative to
the tensor's max) plus fp32 noise."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from planerecnet_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["bf16", "f16"])
def eng(request, cuda_lib):
    from planerecnet_b200.engine import Engine
    return Engine(request.param)


def _tol(eng, ref):
    return (2.0 ** -8 if eng.dt == L.PRN_BF16 else 2.0 ** -11) * float(ref.abs().max()) + 1e-5


def _rand_nhwc(eng, B, H, W, Cc, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(B, H, W, Cc, generator=g).to(eng.tdt)
    return t.cuda(), t.float().permute(0, 3, 1, 2)


def _to_nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2)


def test_layout_roundtrip(eng):
    x = torch.randn(2, 70, 9, 11, generator=torch.Generator().manual_seed(1))
    t = eng.to_nhwc(x.cuda())
    assert t.shape == (2, 9, 11, 128)
    assert torch.equal(t[..., 70:].float().cpu(), torch.zeros(2, 9, 11, 58))
    back = eng.to_nchw(t, 70).cpu()
    assert (back - x.to(eng.tdt).float()).abs().max() == 0


def test_maxpool_avgpool(eng):
    t, ref = _rand_nhwc(eng, 2, 14, 18, 64)
    got = _to_nchw(eng.maxpool(t))
    assert torch.equal(got, F.max_pool2d(ref, 3, 2, 1))
    got = _to_nchw(eng.avgpool2(t))
    exp = F.interpolate(ref, scale_factor=0.5, mode="bilinear", align_corners=False, recompute_scale_factor=False)
    assert (got - exp).abs().max() <= _tol(eng, exp)


def test_stem_im2col_matches_unfold(eng):
    x = torch.randn(2, 3, 20, 28, generator=torch.Generator().manual_seed(2))
    a = eng._empty(2, 10, 14, 192)
    eng._call(eng.lib.prn_stem_im2col, C.c_void_p(x.cuda().data_ptr()), C.c_void_p(a.data_ptr()), 2, 20, 28, eng.dt, eng._st())
    cols = F.unfold(x, 7, padding=3, stride=2)                       # [B, 3*49, L] ordered (c, ky, kx)
    cols = cols.reshape(2, 3, 49, -1).permute(0, 3, 2, 1).reshape(2, 10, 14, 147)   # -> (ky,kx,c)
    got = a.float().cpu()
    assert (got[..., :147] - cols.to(eng.tdt).float()).abs().max() == 0
    assert got[..., 147:].abs().max() == 0


@pytest.mark.parametrize("S", [40, 36, 24, 16, 7])
def test_resize_with_coord(eng, S):
    t, ref = _rand_nhwc(eng, 2, 15, 20, 64, seed=S)
    out = eng._empty(2, S, S, 128)
    eng._call(eng.lib.prn_resize_bilinear, C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), 2, 15, 20, 64, S, S, 128, 1,
              eng.dt, eng._st())
    xr, yr = torch.linspace(-1, 1, 20), torch.linspace(-1, 1, 15)
    yy, xx = torch.meshgrid(yr, xr, indexing="ij")
    full = torch.cat([ref, xx.expand(2, 1, 15, 20), yy.expand(2, 1, 15, 20)], 1)
    exp = F.interpolate(full, size=S, mode="bilinear", align_corners=False)
    got = _to_nchw(out)
    assert (got[:, :66] - exp).abs().max() <= _tol(eng, exp)
    assert got[:, 66:].abs().max() == 0


def test_append_coord(eng):
    t, ref = _rand_nhwc(eng, 2, 15, 20, 64)
    out = eng._empty(2, 15, 20, 128)
    eng._call(eng.lib.prn_append_coord, C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), 2, 15, 20, 64, 128, eng.dt, eng._st())
    got = _to_nchw(out)
    assert torch.equal(got[:, :64], ref)
    xr, yr = torch.linspace(-1, 1, 20), torch.linspace(-1, 1, 15)
    assert (got[0, 64, 3] - xr).abs().max() <= _tol(eng, xr) and (got[1, 65, :, 5] - yr).abs().max() <= _tol(eng, yr)


def test_upsample2x_and_accumulate(eng):
    t, ref = _rand_nhwc(eng, 2, 7, 9, 64)
    exp = F.interpolate(ref, scale_factor=2, mode="bilinear", align_corners=False)
    got = _to_nchw(eng.upsample2x(t))
    assert (got - exp).abs().max() <= _tol(eng, exp)
    acc, accref = _rand_nhwc(eng, 2, 14, 18, 64, seed=5)
    eng.upsample2x(t, into=acc)
    assert (_to_nchw(acc) - (accref + exp)).abs().max() <= _tol(eng, accref + exp)


def test_groupnorm_apply_from_sums(eng):
    t, ref = _rand_nhwc(eng, 2, 9, 11, 128)
    gn = torch.nn.GroupNorm(32, 128)
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5)
        gn.bias.normal_(0, 0.2)
    cg = 4
    grp = ref.reshape(2, 32, cg, -1)
    stats = torch.stack([grp.sum((2, 3)), (grp * grp).sum((2, 3))], -1).reshape(-1).cuda()
    got = _to_nchw(eng.gn_relu(t, stats, gn))
    exp = F.relu(gn(ref)).detach()
    assert (got - exp).abs().max() <= _tol(eng, exp) + 2e-3


def test_mul_and_ppa_gather(eng):
    a, ar = _rand_nhwc(eng, 2, 8, 12, 128, seed=1)
    b, br = _rand_nhwc(eng, 2, 8, 12, 128, seed=2)
    out = eng._empty(2, 8, 12, 128)
    eng._call(eng.lib.prn_mul, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
              C.c_int64(a.numel()), eng.dt, eng._st())
    assert (_to_nchw(out) - ar * br).abs().max() <= _tol(eng, ar * br)
    q = eng._empty(2, 2 * 3 * 4, 128)
    eng._call(eng.lib.prn_ppa_gather, C.c_void_p(a.data_ptr()), C.c_void_p(q.data_ptr()), 2, 8, 12, 128, eng.dt, eng._st())
    got = q.float().cpu().reshape(2, 2, 3, 2, 2, 128)        # (b, by, bx, dy, dx, c)
    exp = ar.permute(0, 2, 3, 1).reshape(2, 2, 4, 3, 4, 128)[:, :, 1:3, :, 1:3].permute(0, 1, 3, 2, 4, 5)
    assert torch.equal(got, exp)
    # mean of those 4 pixels == bilinear x0.25 (align_corners=False)
    down = F.interpolate(ar, scale_factor=0.25, mode="bilinear", align_corners=False, recompute_scale_factor=False)
    assert (got.mean((3, 4)).permute(0, 3, 1, 2) - down).abs().max() < 1e-5


def test_bad_arguments_are_rejected(eng):
    t, _ = _rand_nhwc(eng, 1, 4, 4, 64)
    with pytest.raises(L.PrnError):
        eng._call(eng.lib.prn_avgpool2x2, C.c_void_p(t.data_ptr()), C.c_void_p(t.data_ptr()), 1, 3, 4, 64, eng.dt, eng._st())
    with pytest.raises(L.PrnError):
        eng._call(eng.lib.prn_mul, None, None, None, C.c_int64(8), eng.dt, eng._st())


def test_conv3x3_to1_reflect_softplus(eng):
    t, ref = _rand_nhwc(eng, 2, 21, 35, 64, seed=9)            # sizes that are not multiples of the 16x16 tile
    g = torch.Generator().manual_seed(4)
    w = torch.randn(1, 64, 3, 3, generator=g) * 0.1
    w9c = w[0].permute(1, 2, 0).reshape(9, 64).contiguous().cuda()
    out = torch.empty(2, 21, 35, 1, device="cuda")
    eng._call(eng.lib.prn_conv3x3_to1_reflect, C.c_void_p(t.data_ptr()), C.c_void_p(w9c.data_ptr()), C.c_float(0.3),
              C.c_void_p(out.data_ptr()), 2, 21, 35, 64, 1, eng.dt, eng._st())
    exp = F.softplus(F.conv2d(F.pad(ref, (1, 1, 1, 1), mode="reflect"), w, torch.tensor([0.3])))
    assert (out.cpu().permute(0, 3, 1, 2) - exp).abs().max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("u8,hw", [(True, (96, 128)), (False, (96, 128)), (True, (75, 101))])
def test_stem_rows_from_camera_frames_equal_transform_then_im2col(cuda_lib, u8, hw):
    """FastBaseTransform (data/augmentations.py:496-530: (x - MEANS) / STD per BGR channel, BGR -> RGB) and
    pad_even_divided (models/functions/funcs.py:204-210) folded into the stem's im2col give BIT-IDENTICAL rows to running the
    transform in torch and feeding prn_stem_im2col (same fp32 arithmetic before the one 16-bit rounding)."""
    import torch
    from planerecnet_b200.config import MEANS, STD
    from planerecnet_b200.engine import Engine
    eng = Engine("f16")
    Hi, Wi = hw
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (2, Hi, Wi, 3), generator=g, dtype=torch.uint8)
    fr = frames.cuda() if u8 else frames.float().cuda()
    rows, Hp, Wp = eng.stem_rows_from_images(fr)
    # the reference's path: pad_even_divided on the raw image, then FastBaseTransform
    padded = torch.zeros(2, Hp, Wp, 3)
    padded[:, :Hi, :Wi] = frames.float()
    img = padded.permute(0, 3, 1, 2).contiguous()
    mean = torch.tensor(MEANS).float()[None, :, None, None]
    std = torch.tensor(STD).float()[None, :, None, None]
    x = ((img - mean) / std)[:, (2, 1, 0), :, :].contiguous().cuda()
    ref = eng._empty(2, Hp // 2, Wp // 2, 192)
    import ctypes as C
    eng._call(eng.lib.prn_stem_im2col, C.c_void_p(x.data_ptr()), C.c_void_p(ref.data_ptr()), 2, Hp, Wp, eng.dt, eng._st())
    torch.cuda.synchronize()
    assert (Hp, Wp) == ((Hi + 31) // 32 * 32, (Wi + 31) // 32 * 32)
    assert torch.equal(rows.view(torch.int16), ref.view(torch.int16))


@pytest.mark.gpu
def test_forward_frames_equals_forward_of_transformed_batch(cuda_lib):
    """net.forward_frames(frames) == net(FastBaseTransform()(frames)): same detections and depth (dense forward inputs are
    bit-identical, so only the GroupNorm atomics' run-to-run noise remains)."""
    import torch
    import helpers as H
    from planerecnet_b200.config import MEANS, STD
    net = H.perturb_(H.build_ours("PlaneRecNet_50_config")).eval().cuda()
    g = torch.Generator().manual_seed(1)
    frames = torch.randint(0, 256, (2, 128, 160, 3), generator=g, dtype=torch.uint8).cuda()
    mean = torch.tensor(MEANS, device="cuda").float()[None, :, None, None]
    std = torch.tensor(STD, device="cuda").float()[None, :, None, None]
    x = ((frames.float().permute(0, 3, 1, 2).contiguous() - mean) / std)[:, (2, 1, 0), :, :].contiguous()
    with torch.no_grad():
        ref = net(x)
        ref = [{k: (None if v is None else v.clone()) for k, v in r.items()} for r in ref]
        got = net.forward_frames(frames)
    for r, o in zip(got, ref):
        assert list(r.keys()) == list(o.keys())
        assert H.rel_l2(r["pred_depth"], o["pred_depth"]) < 2e-3
        n_r = 0 if r["pred_scores"] is None else len(r["pred_scores"])
        n_o = 0 if o["pred_scores"] is None else len(o["pred_scores"])
        assert abs(n_r - n_o) <= max(2, n_o // 5), (n_r, n_o)
