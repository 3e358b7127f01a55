"""Dense loss kernels (first two terms) against the oracle's restatement of the reference's loss functions
(pinned to the unmodified reference by tests/golden/loss_golden.pt): values to 1e-5, gradients to 1e-4 of their largest
entry, on the same fp32 inputs."""
import pytest
import torch
import torch.nn.functional as F

from oracle import prn_loss_oracle as LO
from planerecnet_b200 import losses as PL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,ld", [(2 * 3728, 16), (777, 2)])
def test_focal_cate_loss(cuda_lib, n, ld):
    g = torch.Generator().manual_seed(n)
    logits = torch.randn(n, ld, generator=g) * 3
    labels = torch.randint(0, 3, (n,), generator=g)             # 2 = background
    labels[:5] = torch.tensor([0, 1, 2, 0, 2])
    num_ins = 37
    x = logits[:, :2].clone().requires_grad_(True)
    oh = torch.zeros(n, 2)
    pos = torch.nonzero(labels != 2).squeeze(1)
    oh[pos, labels[pos]] = 1
    ref = LO.sigmoid_focal_sum(x, oh, 0.25, 2.0) / (num_ins + 1)
    ref.backward()
    xd = logits.cuda().requires_grad_(True)
    got = PL.focal_cate_loss(xd, labels.cuda(), num_ins)
    got.backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
    gd = xd.grad.cpu()
    assert float((gd[:, :2] - x.grad).abs().max()) <= 1e-4 * float(x.grad.abs().max())
    assert float(gd[:, 2:].abs().max()) == 0.0 if ld > 2 else True


@pytest.mark.parametrize("B,h,w", [(2, 240, 320), (3, 17, 23)])
def test_depth_rmselog_loss(cuda_lib, B, h, w):
    g = torch.Generator().manual_seed(B)
    depth = (torch.rand(B, 1, h, w, generator=g) * 4 + 0.5)
    depth[0, 0, h - 1, w - 1] = 1e-12                            # its corner up-pixel falls below the log clamp: no gradient there
    gt = 0.5 + 4 * torch.rand(B, 1, 2 * h, 2 * w, generator=g)
    gt[0, 0, :5, :7] = 0.0                                       # invalid ground truth
    d = depth.clone().requires_grad_(True)
    up = F.interpolate(d, scale_factor=2, mode="bilinear", align_corners=False)
    ref = 5.0 * LO.rmse_log_mean(up, gt, gt > 1 / 1000)
    ref.backward()
    dd = depth.cuda().requires_grad_(True)
    got = PL.depth_rmselog_loss(dd, gt.cuda())
    (2.0 * got).backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float((dd.grad.cpu() / 2.0 - d.grad).abs().max()) <= 1e-4 * float(d.grad.abs().max())


def test_dice_lava_row_kernels_match_emulation(cuda_lib):
    """prn_lava_weights / prn_dice_lava_rows / prn_dice_lava_bwd against their torch emulation on identical fp32 inputs."""
    import loss_cases as LC
    from loss_emulation import EmuBackend
    emu, be = EmuBackend(), PL.CudaBackend()
    g = torch.Generator().manual_seed(0)
    gt = 0.5 + 4 * torch.rand(2, 1, 96, 128, generator=g)
    gt[:, :, 20:60, 30:90] += torch.linspace(0, 3, 60)[None, None, None, :]          # a slope and steps: non-trivial gradients
    gw_r, gs_r = emu.lava_weights(gt, 24, 32, 1 / 1000)
    gw, gs = be.lava_weights(gt.cuda(), 24, 32, 1 / 1000)
    assert float(gs_r.min()) > 0
    assert float((gs.cpu() - gs_r).abs().max()) <= 1e-4 * float(gs_r.max())
    assert float((gw.cpu() - gw_r).abs().max()) <= 1e-4 * float(gw_r.max())
    B, n, P = 2, 16, 24 * 32
    seg = torch.rand(B * n, P, generator=g)
    tgt = (torch.rand(B * n, P, generator=g) < 0.3).to(torch.uint8)
    st_r = emu.row_stats(seg, tgt, gw_r, n)
    st = be.row_stats(seg.cuda(), tgt.cuda(), gw, n)
    assert float((st.cpu() - st_r).abs().max() / st_r.abs().max()) <= 1e-5
    coef = torch.randn(B * n, 3, generator=g)
    dx_r = emu.row_bwd(seg, tgt, gw_r, coef, n)
    dx = be.row_bwd(seg.cuda(), tgt.cuda(), gw, coef.cuda(), n)
    assert float((dx.float().cpu() - dx_r).abs().max()) <= (2.0 ** -8 + 1e-4) * float(dx_r.abs().max())      # one 16-bit rounding


@pytest.mark.parametrize("name", ["loss_seed0", "loss_seed2_many_tiny"])
def test_ins_lava_losses_match_oracle(cuda_lib, name):
    """The dice and lava terms end to end on the GPU (f16 operands of the three grouped contractions, fp32 sums) against
    the oracle: values within 1e-2, gradients w.r.t. the mask features and every level's kernels by cosine / rel-L2."""
    import numpy as np
    import loss_cases as LC
    from helpers import rel_l2
    from planerecnet_b200 import targets as T
    mask, cate, kern, depth, gts, gt_depth = LC.synth(**LC.CASES[name])
    leaves = [mask] + kern
    for t in leaves:
        t.requires_grad_(True)
    np.random.seed(0)
    ref = LO.loss_forward(mask, cate, kern, depth, gts, gt_depth)
    (ref["ins"] + ref["lav"].sum()).backward()
    dm, dk = mask.detach().cuda().requires_grad_(True), [k.detach().cuda().requires_grad_(True) for k in kern]
    # the device-side assignment must equal the (CPU) assignment the oracle uses, bit for bit; the losses consume it
    targets = [T.assign_targets(g, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"]) for g in gts]
    gts_d = [{k: v.cuda() for k, v in g.items()} for g in gts]
    targets_dev = [T.assign_targets(g, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"]) for g in gts_d]
    same = all(td[3] == tc[3] and torch.equal(td[0].cpu(), tc[0]) for a, b in zip(targets_dev, targets) for td, tc in zip(a, b))
    assert same, f"{name}: device-side target assignment differs from the CPU one"
    l_ins, l_lav = PL.ins_lava_losses(dm, dk, targets_dev, gt_depth.cuda())
    (l_ins + l_lav).backward()
    assert abs(float(l_ins) - float(ref["ins"])) <= 1e-2 * abs(float(ref["ins"]))
    assert abs(float(l_lav) - float(ref["lav"].sum())) <= 1e-2 * abs(float(ref["lav"].sum()))
    for got, t in zip([dm] + dk, leaves):
        if t.grad is None or float(t.grad.abs().max()) == 0.0:
            assert got.grad is None or float(got.grad.abs().max()) == 0.0
            continue
        a, b = got.grad.cpu().double().flatten(), t.grad.double().flatten()
        cos = float((a @ b) / (a.norm() * b.norm()))
        print(f"{name}: grad {tuple(t.shape)} cos {cos:.5f} rel-L2 {rel_l2(got.grad.cpu(), t.grad):.3e}")
        # the dice gradient is cancellation-dominated: 16-bit rounding of the operands is amplified
        assert cos >= 0.995 and rel_l2(got.grad.cpu(), t.grad) <= 0.1, (cos, rel_l2(got.grad.cpu(), t.grad))


@pytest.mark.parametrize("name", ["loss_seed0", "loss_seed1_b1", "loss_seed2_many_tiny"])
def test_device_target_assignment_is_bit_exact(cuda_lib, name):
    """Integer / index work must be bit-exact: the SOLO target assignment computed on CUDA tensors equals the CPU one and the
    unmodified reference's own assignment stored in tests/golden/loss_golden.pt (grid cells, positive cells, mask sums)."""
    import os
    import loss_cases as LC
    from planerecnet_b200 import targets as T
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.pt"))[name]
    _, _, _, _, gts, _ = LC.synth(**LC.CASES[name])
    for b, g in enumerate(gts):
        cpu = T.assign_targets(g, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"])
        dev = T.assign_targets({k: v.cuda() for k, v in g.items()}, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"])
        for lvl, (tc, td) in enumerate(zip(cpu, dev)):
            assert td[3] == tc[3], (name, b, lvl, "grid order differs between CUDA and CPU")
            assert torch.equal(td[0].cpu(), tc[0]) and torch.equal(td[1].cpu(), tc[1]) and torch.equal(td[2].cpu(), tc[2])
            assert [int(o) for o in td[3]] == gold["grid_orders"][b][lvl], (name, b, lvl, "differs from the reference's assignment")
            assert (td[1].cpu() != 2).nonzero().tolist() == gold["cate_pos"][b][lvl]
        assert [int(t[0].sum()) for t in dev] == gold["ins_label_sums"][b]


@pytest.mark.parametrize("name", ["loss_seed0", "loss_seed1_b1", "loss_seed2_many_tiny"])
def test_assembled_loss_module_matches_reference_golden_on_gpu(cuda_lib, name):
    """planerecnet_b200.losses.PlaneRecNetLoss end to end on the GPU (device-side targets, kernel-backed dice / lava / focal /
    depth terms, host-side plane term) against the UNMODIFIED reference's values in tests/golden/loss_golden.pt:
    loss terms within 1e-2 relative (f16 operands in the mask contractions; fp32 elsewhere: 1e-4), gradient norms within 2e-2."""
    import os
    import numpy as np
    import loss_cases as LC
    from planerecnet_b200.config import cfg, set_cfg
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.pt"))[name]
    set_cfg("PlaneRecNet_101_config")
    crit = PL.PlaneRecNetLoss(cfg)
    mask, cate, kern, depth, gts, gt_depth = LC.synth(**LC.CASES[name])
    leaves = [t.cuda().requires_grad_(True) for t in [mask] + cate + kern + [depth]]
    gts_d = [{k: v.cuda() for k, v in g.items()} for g in gts]
    np.random.seed(0)
    out = crit(None, leaves[0], leaves[1:5], leaves[5:9], leaves[9], gts_d, gt_depth.cuda())
    assert set(out.keys()) == {"ins", "cat", "dpt", "pln", "lav"}
    tol = {"ins": 1e-2, "lav": 1e-2, "cat": 1e-4, "dpt": 1e-4, "pln": 2e-3}
    for k, ref in gold["losses"].items():
        got = float(out[k].sum())
        if ref != ref:                      # the reference's NaN for a degenerate plane
            assert got != got, (k, got)
            continue
        assert abs(got - ref) <= tol[k] * abs(ref) + 1e-6, (name, k, got, ref)
    total = sum(v.sum() for v in out.values())
    total.backward()                        # like train.py:349-350 (the reference backpropagates a non-finite total as well)
    for t, ref in zip(leaves, gold["grad_norms"]):
        got = 0.0 if t.grad is None else float(t.grad.double().norm())
        if ref != ref:
            assert got != got, (name, tuple(t.shape), got)
        else:
            assert abs(got - ref) <= 2e-2 * abs(ref) + 1e-7, (name, tuple(t.shape), got, ref)


def test_background_preparation_and_lookahead_give_the_same_losses(cuda_lib):
    """crit.prepare(gts) (worker thread + side stream + C sampler) followed by crit(...) returns what crit(...) alone returns:
    same targets, the same numpy-stream triplets, same loss terms and gradients; and a preparation made for the NEXT batch
    while the current one is still in use (the look-ahead of bench.py's e2e loop) is picked up by the next call."""
    import numpy as np
    import loss_cases as LC
    from planerecnet_b200.config import cfg, set_cfg
    set_cfg("PlaneRecNet_101_config")
    names = list(LC.CASES)[:2]
    data = []
    for name in names:
        mask, cate, kern, depth, gts, gt_depth = LC.synth(**LC.CASES[name])
        data.append(([t.cuda() for t in [mask] + cate + kern + [depth]], [{k: v.cuda() for k, v in g.items()} for g in gts],
                     gt_depth.cuda()))

    def run(mode):
        crit = PL.PlaneRecNetLoss(cfg)
        np.random.seed(0)
        res = []
        if mode == "lookahead":
            crit.prepare(data[0][1])
        for i, (tens, gts, gtd) in enumerate(data):
            leaves = [t.clone().requires_grad_(True) for t in tens]
            if mode == "prepare":
                crit.prepare(gts)
            out = crit(None, leaves[0], leaves[1:5], leaves[5:9], leaves[9], gts, gtd)
            if mode == "lookahead" and i + 1 < len(data):
                crit.prepare(data[i + 1][1])                  # next batch's preparation overlaps this backward
            torch.nansum(torch.stack([v.sum() for v in out.values()])).backward()
            res.append(({k: v.detach().double().sum().cpu() for k, v in out.items()},
                        [None if t.grad is None else t.grad.double().norm().cpu() for t in leaves]))
        tail = np.random.randint(0, 1 << 30, size=4)          # numpy's global stream ends at the same position
        return res, tail

    ref, tail_ref = run("inline")
    for mode in ("prepare", "lookahead"):
        got, tail = run(mode)
        assert np.array_equal(tail, tail_ref), mode
        for (lo_r, gr_r), (lo_g, gr_g) in zip(ref, got):
            for k in lo_r:
                a, b = float(lo_g[k]), float(lo_r[k])
                assert (a != a and b != b) or abs(a - b) <= 1e-5 * abs(b) + 1e-7, (mode, k, a, b)
            for a, b in zip(gr_g, gr_r):
                if b is None:
                    assert a is None
                    continue
                a, b = float(a), float(b)
                assert (a != a and b != b) or abs(a - b) <= 1e-4 * abs(b) + 1e-9, (mode, a, b)


@pytest.mark.parametrize("name", ["loss_seed0", "loss_seed1_b1", "loss_seed2_many_tiny"])
def test_plane_term_triplet_kernels_match_the_tensor_formulation(cuda_lib, name):
    """prn_vnl_triplets_fwd / _bwd (per-triplet geometry, selection mask, 1 - |cos|, analytic depth gradient) against the tensor
    formulation of the same term (itself pinned to the per-plane mirror of vnl.py on the CPU and to the reference's golden values):
    same numpy-stream triplets, per-image losses within 1e-5, depth gradient within 1e-4 (fp32 reductions in another order; a
    triplet that sits on a selection threshold to the last bit may flip)."""
    import numpy as np
    import torch.nn.functional as F
    import loss_cases as LC
    _, _, _, depth, gts, gt_depth = LC.synth(**LC.CASES[name])
    gts_d = [{k: v.cuda() for k, v in g.items()} for g in gts]
    res = {}
    for kernels in (False, True):
        d = depth.cuda().clone().requires_grad_(True)
        up = F.interpolate(d, scale_factor=2, mode="bilinear", align_corners=False)
        np.random.seed(0)
        out = PL._PlaneNormalBatched((480, 640), kernels=kernels)(up, gts_d, gt_depth.cuda())
        torch.nansum(out).backward()
        res[kernels] = (out.detach().double().cpu(), d.grad.double().cpu())
    (lo_t, g_t), (lo_k, g_k) = res[False], res[True]
    assert torch.equal(torch.isnan(lo_t), torch.isnan(lo_k)), (lo_t, lo_k)
    ok = ~torch.isnan(lo_t)
    if not bool(ok.any()):                      # every image degenerate (tiny planes): the reference's NaN in both, nothing to compare
        assert name == "loss_seed2_many_tiny"
        return
    assert torch.allclose(lo_k[ok], lo_t[ok], rtol=1e-5, atol=1e-9), (lo_t, lo_k)
    fin = torch.isfinite(g_t) & torch.isfinite(g_k)
    assert bool((torch.isfinite(g_t) == torch.isfinite(g_k)).all())
    rel = float((g_k[fin] - g_t[fin]).norm() / g_t[fin].norm())
    assert rel <= 1e-4 and float(g_t[fin].norm()) > 0, rel
