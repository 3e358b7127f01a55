"""Whole-path parity on the GPU: our sm_100a forward (through the module API and the C ABI) against
(a) the CPU oracle on the same seeded weights/inputs and (b) the golden fixtures generated from the
unmodified reference (tests/golden/make_golden.py).

Tolerances (rel-L2 per output tensor, eval-mode BatchNorm), from BASELINE.json north_star:
  f16 path  (default; 10-bit mantissa operands, fp32 accumulate):  3e-3  (north_star: 1e-2 for 16-bit, 1e-3 for fp32)
  bf16 path (8-bit mantissa):                                       2e-2  (measured 0.8-1.5e-2; reported in DESIGN.md)
Bookkeeping (indices / counts) is compared exactly where the dense inputs are identical
(tests/test_postprocess_gpu.py)."""
import os

import pytest
import torch

import helpers as H
from oracle import prn_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {"f16": 3e-3, "bf16": 2e-2}


def _run(preset, B, Hh, Ww, prec, seed=0):
    net = H.perturb_(H.build_ours(preset, 0)).eval()
    x = H.make_input(B, Hh, Ww, seed)
    orc = O.Oracle(net.state_dict(), preset)
    with torch.no_grad():
        ref = orc.forward_dense(x)
    net = net.cuda().set_precision(prec)
    with torch.no_grad():
        st = net.engine.forward_dense(net, x.cuda())
    torch.cuda.synchronize()
    return net, x, orc, ref, st


@pytest.mark.parametrize("preset,prec", [("PlaneRecNet_50_config", "f16"), ("PlaneRecNet_50_config", "bf16"),
                                         ("PlaneRecNet_101_config", "f16")])
def test_dense_forward_matches_oracle(cuda_lib, preset, prec):
    net, x, orc, (omask, ocate, okern, odepth), st = _run(preset, 2, 192, 256, prec)
    mask, cate, kern, depth = st["outputs"]
    tol = TOL[prec]
    errs = {"mask": H.rel_l2(mask.cpu(), omask), "depth": H.rel_l2(depth.cpu(), odepth)}
    for i in range(4):
        errs[f"cate{i}"] = H.rel_l2(cate[i].cpu(), ocate[i])
        errs[f"kern{i}"] = H.rel_l2(kern[i].cpu(), okern[i])
        c = st["cs"][i][..., :orc.taps["cs"][i].shape[1]].float().permute(0, 3, 1, 2).cpu()
        errs[f"C{i + 2}"] = H.rel_l2(c, orc.taps["cs"][i])
        p = st["ps"][i].float().permute(0, 3, 1, 2).cpu()
        errs[f"P{i + 2}"] = H.rel_l2(p, orc.taps["ps"][i])
    errs["attn"] = H.rel_l2(st["attn"].float().permute(0, 3, 1, 2).cpu(), orc.taps["ppa_attn"])
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"{preset}/{prec}: rel-L2 above {tol}: {bad} (all: {errs})"
    # contract: NCHW fp32 contiguous outputs of the reference's training branch
    assert mask.shape == omask.shape and mask.dtype == torch.float32 and mask.is_contiguous()
    assert [tuple(c.shape) for c in cate] == [(2, 2, s, s) for s in (40, 36, 24, 16)]
    assert [tuple(k.shape) for k in kern] == [(2, 128, s, s) for s in (40, 36, 24, 16)]
    assert depth.shape == (2, 1, 96, 128)


def test_dense_forward_matches_reference_golden(cuda_lib):
    """Directly against samples of the unmodified reference's tensors (no oracle in between)."""
    from golden.make_golden import sample_idx
    g = torch.load(os.path.join(GOLD, "r50_b2_192x256.pt"))
    net = H.perturb_(H.build_ours(g["preset"], g["seed"])).eval().cuda()
    x = H.make_input(g["B"], g["H"], g["W"], g["seed"]).cuda()
    with torch.no_grad():
        mask, cate, kern, depth = net.forward_dense(x)
    for name, t in [("mask", mask), ("depth", depth)] + [(f"kern{i}", kern[i]) for i in range(4)]:
        gs = g["stages"][name]
        flat = t.float().cpu().flatten()
        got = flat[sample_idx(flat.numel())]
        err = float((got - gs["samples"]).norm() / gs["samples"].norm())
        assert err <= TOL["f16"], f"{name}: {err:.3g}"
    for i in range(4):
        assert H.rel_l2(cate[i].cpu(), g["cate_full"][i]) <= TOL["f16"]
    assert H.rel_l2(depth.cpu()[:, :, ::4, ::4], g["depth_ds4"]) <= TOL["f16"]


def test_graph_replay_equals_eager_and_tracks_weight_updates(cuda_lib):
    net, x, orc, ref, st = _run("PlaneRecNet_50_config", 1, 128, 160, "f16")
    eng = net.engine
    xc = x.cuda()
    # GroupNorm sums are accumulated with fp32 atomics (order varies run to run), so two runs agree up to
    # isolated 16-bit rounding flips, not bit for bit
    same = lambda p, q: H.rel_l2(p, q) < 1e-3
    with torch.no_grad():
        eager = [st["outputs"][0].clone(), st["outputs"][3].clone()]
        g1 = eng.forward_dense_graph(net, xc)
        a = [g1["outputs"][0].clone(), g1["outputs"][3].clone()]
        x2 = H.make_input(1, 128, 160, seed=3).cuda()
        g2 = eng.forward_dense_graph(net, x2)
        b = [g2["outputs"][0].clone(), g2["outputs"][3].clone()]
        e2 = eng.forward_dense(net, x2)["outputs"]
    assert same(a[0], eager[0]) and same(a[1], eager[1]), "graph replay differs from the eager launch sequence"
    assert same(b[0], e2[0]) and same(b[1], e2[3]), "graph replay ignored the new input"
    assert not same(b[1], a[1])
    # an in-place weight update must invalidate the packed weights and the captured graph
    with torch.no_grad():
        net.depth_decoder.depth_pred[1].bias.add_(1.0)
        g3 = eng.forward_dense_graph(net, xc)["outputs"][3].clone()
        e3 = eng.forward_dense(net, xc)["outputs"][3]
    assert same(g3, e3) and not same(g3, a[1])


def test_module_level_entry_points(cuda_lib):
    """Sub-module forwards keep the reference's NCHW fp32 contracts (backbone / fpn / heads / decoder)."""
    net, x, orc, (omask, ocate, okern, odepth), st = _run("PlaneRecNet_50_config", 1, 128, 160, "f16")
    with torch.no_grad():
        cs = net.backbone(x.cuda())
        ps = net.fpn([cs[i] for i in net.fpn_indices])
        cate, kern = net.inst_head(net.split_feats(ps))
        mask = net.mask_head(ps)
        depth = net.depth_decoder([cs[i] for i in net.depth_decoder_indices], mask, kern)
    assert isinstance(cs, tuple) and [c.shape[1] for c in cs] == [256, 512, 1024, 2048]
    for i in range(4):
        assert H.rel_l2(cs[i].cpu(), orc.taps["cs"][i]) <= TOL["f16"]
        assert H.rel_l2(ps[i].cpu(), orc.taps["ps"][i]) <= TOL["f16"] * 1.5
        assert H.rel_l2(kern[i].cpu(), okern[i]) <= TOL["f16"] * 2
    assert H.rel_l2(mask.cpu(), omask) <= TOL["f16"] * 2
    assert H.rel_l2(depth.cpu(), odepth) <= TOL["f16"] * 2


def test_dcn_module_matches_oracle(cuda_lib):
    from planerecnet_b200.models.dcn import DeformableConv2d
    torch.manual_seed(3)
    for stride in (1, 2):
        m = DeformableConv2d(128, 128, stride=stride, bias=True)
        with torch.no_grad():
            m.offset_conv.weight.normal_(0, 0.03)
            m.offset_conv.bias.normal_(0, 0.5)
            m.modulator_conv.weight.normal_(0, 0.02)
        x = torch.randn(2, 128, 30, 40)
        sd = {"d." + k: v for k, v in m.state_dict().items()}
        o = O.Oracle.__new__(O.Oracle)
        o.sd, o.dtype = {k: v.float() for k, v in sd.items()}, torch.float32
        with torch.no_grad():
            ref = o._dcn(x, "d", stride)
            got = m.cuda()(x.cuda()).cpu()
        assert H.rel_l2(got, ref) <= 3e-3, (stride, H.rel_l2(got, ref))


def test_eval_forward_detections_close_to_oracle(cuda_lib):
    net, x, orc, ref, st = _run("PlaneRecNet_50_config", 2, 192, 256, "f16")
    with torch.no_grad():
        ores = orc.forward_eval(x)
        res = net(x.cuda())
    assert len(res) == 2
    for r, o in zip(res, ores):
        assert list(r.keys()) == ["pred_masks", "pred_boxes", "pred_classes", "pred_scores", "pred_depth"]
        assert H.rel_l2(r["pred_depth"].cpu(), o["pred_depth"]) <= TOL["f16"]
        n_r = 0 if r["pred_scores"] is None else len(r["pred_scores"])
        n_o = 0 if o["pred_scores"] is None else len(o["pred_scores"])
        assert abs(n_r - n_o) <= max(2, n_o // 10), (n_r, n_o)     # candidates sitting on a threshold may flip
        if n_r and n_o:
            assert r["pred_masks"].dtype == torch.bool and r["pred_masks"].shape[1:] == (192, 256)
            assert r["pred_boxes"].shape == (n_r, 4) and r["pred_classes"].dtype == torch.int64
            assert abs(float(r["pred_scores"][0]) - float(o["pred_scores"][0])) < 5e-3
            m0, o0 = r["pred_masks"][0].cpu(), o["pred_masks"][0]
            assert (m0 & o0).sum() / (m0 | o0).sum() > 0.98


def test_infer_pipelined_equals_sequential_calls():
    """net.infer_pipelined(batches) (two graph slots, forward of batch k+1 overlapping the bookkeeping of batch k)
    must deliver what net(x) delivers, batch by batch, in order.  The dense forward is not bit-reproducible run to run
    (GroupNorm sums use fp32 atomics), so dense values are compared to 16-bit-rounding tolerance and the detections
    loosely where they sit on hard thresholds."""
    net = H.perturb_(H.build_ours("PlaneRecNet_50_config")).eval().cuda()
    xs = [H.make_input(2, 128, 160, seed=s) * (0.4 + 0.4 * s) for s in (0, 1, 2, 3, 4)]     # clearly different batches
    with torch.no_grad():
        ref = [net(x.cuda()) for x in xs]
        ref = [[{k: (None if v is None else v.clone()) for k, v in r.items()} for r in batch] for batch in ref]
    got = list(net.infer_pipelined(x.pin_memory() for x in xs))
    assert len(got) == len(ref)
    n_det = 0
    for bi, (gb, rb) in enumerate(zip(got, ref)):
        assert len(gb) == len(rb)
        for g, r in zip(gb, rb):
            assert list(g.keys()) == list(r.keys())
            assert H.rel_l2(g["pred_depth"], r["pred_depth"]) < 2e-3, bi
            if r["pred_scores"] is None:
                assert g["pred_scores"] is None or g["pred_scores"].numel() <= 2
                continue
            # detections sit on hard thresholds (score_thr / update_thr) and near-tied scores: run-to-run noise of the
            # dense forward may add / drop a borderline one or swap neighbours -> compare counts loosely, the best ones tightly
            if g["pred_scores"] is None:
                assert r["pred_scores"].numel() <= 2
                continue
            ng, nr = g["pred_scores"].numel(), r["pred_scores"].numel()
            assert abs(ng - nr) <= max(2, 0.3 * max(ng, nr)), (bi, ng, nr)     # the low-score tail of a random-init model is fragile
            k = min(3, ng, nr)
            gs, rs = g["pred_scores"].sort(descending=True).values[:k], r["pred_scores"].sort(descending=True).values[:k]
            assert torch.allclose(gs, rs, rtol=2e-2, atol=1e-3), (gs, rs)
            n_det += r["pred_scores"].numel()
    assert n_det > 0
    # a different batch order must change the answers accordingly (the slots really carry different batches)
    assert H.rel_l2(got[0][0]["pred_depth"], got[1][0]["pred_depth"]) > 3e-3
