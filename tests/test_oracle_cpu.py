"""CPU tests of the oracle (oracle/prn_oracle.py) against the golden fixtures generated from the
unmodified reference by tests/golden/make_golden.py, plus the third-party deform_conv2d semantics."""
import os

import pytest
import torch

import helpers as H
from oracle import prn_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# oracle and reference run the same fp32 CPU operators; the residue is reduction-order noise
TOL = 2e-4


def _gold(name):
    return torch.load(os.path.join(GOLD, name + ".pt"))


def _check_stage(gs, t, label):
    t = t.detach().float().contiguous()
    assert list(t.shape) == gs["shape"], label
    from golden.make_golden import sample_idx
    idx = sample_idx(t.numel())
    got = t.flatten()[idx]
    err = float((got - gs["samples"]).norm() / (gs["samples"].norm() + 1e-30))
    assert err <= TOL, f"{label}: sampled rel-L2 {err:.3g}"
    assert abs(float(t.double().norm()) - gs["l2"]) <= TOL * gs["l2"] + 1e-6, f"{label}: norm differs"


@pytest.mark.parametrize("name", ["r50_b2_192x256", "r101_b1_192x256"])
def test_oracle_matches_reference_golden(name):
    g = _gold(name)
    net = H.perturb_(H.build_ours(g["preset"], g["seed"]))
    sd = net.state_dict()
    assert len(sd) == g["n_state"]
    l2 = float(sum(v.double().norm() ** 2 for v in sd.values() if v.is_floating_point()) ** 0.5)
    assert abs(l2 - g["state_l2"]) < 1e-9 * g["state_l2"], "seeded construction must reproduce the reference's weights"
    x = H.make_input(g["B"], g["H"], g["W"], g["seed"])
    orc = O.Oracle(sd, g["preset"])
    with torch.no_grad():
        mask, cate, kern, depth = orc.forward_dense(x)
        res = orc.forward_eval(x)
    st = g["stages"]
    for i in range(4):
        _check_stage(st[f"C{i + 2}"], orc.taps["cs"][i], f"C{i + 2}")
        _check_stage(st[f"P{i + 2}"], orc.taps["ps"][i], f"P{i + 2}")
        _check_stage(st[f"cate{i}"], cate[i], f"cate{i}")
        _check_stage(st[f"kern{i}"], kern[i], f"kern{i}")
        assert H.rel_l2(cate[i], g["cate_full"][i]) <= TOL
    _check_stage(st["mask"], mask, "mask")
    _check_stage(st["attn"], orc.taps["ppa_attn"], "attn")
    _check_stage(st["depth"], depth, "depth")
    assert H.rel_l2(depth[:, :, ::4, ::4], g["depth_ds4"]) <= TOL
    # bookkeeping: same detections, same classes, same mask areas
    for r, gr in zip(res, g["results"]):
        if gr["pred_scores"] is None:
            assert r["pred_scores"] is None
            continue
        assert r["pred_scores"].shape == gr["pred_scores"].shape
        assert torch.equal(r["pred_classes"], gr["pred_classes"])
        assert torch.allclose(r["pred_scores"], gr["pred_scores"], atol=2e-3)
        assert torch.equal(r["pred_boxes"], gr["pred_boxes"])


def test_oracle_full_resolution_golden():
    g = _gold("r50_b1_480x640")
    net = H.perturb_(H.build_ours(g["preset"], g["seed"]))
    x = H.make_input(g["B"], g["H"], g["W"], g["seed"])
    orc = O.Oracle(net.state_dict(), g["preset"])
    with torch.no_grad():
        mask, cate, kern, depth = orc.forward_dense(x)
    assert list(depth.shape) == [1, 1, 240, 320] and list(mask.shape) == [1, 128, 120, 160]
    _check_stage(g["stages"]["depth"], depth, "depth")
    _check_stage(g["stages"]["mask"], mask, "mask")
    for i in range(4):
        assert H.rel_l2(cate[i], g["cate_full"][i]) <= TOL


@pytest.mark.parametrize("stride", [1, 2])
def test_deform_conv2d_matches_torchvision(stride):
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    x = torch.randn(2, 16, 15, 20, dtype=torch.float64)
    w = torch.randn(8, 16, 3, 3, dtype=torch.float64)
    b = torch.randn(8, dtype=torch.float64)
    Ho, Wo = (15 + 2 - 3) // stride + 1, (20 + 2 - 3) // stride + 1
    off = torch.randn(2, 18, Ho, Wo, dtype=torch.float64) * 3
    off[:, :, 0, 0] += 30      # far outside -> zero contribution
    off[:, :, 1, 1] = -0.5     # straddles the border
    m = torch.rand(2, 9, Ho, Wo, dtype=torch.float64) * 2
    ref = tv.ops.deform_conv2d(x, off, w, b, stride=stride, padding=1, mask=m)
    got = O.deform_conv2d(x, off, w, b, stride, 1, m)
    assert (ref - got).abs().max() < 1e-12


def test_deform_conv2d_zero_offsets_is_plain_conv():
    torch.manual_seed(1)
    x = torch.randn(1, 8, 9, 11, dtype=torch.float64)
    w = torch.randn(4, 8, 3, 3, dtype=torch.float64)
    got = O.deform_conv2d(x, torch.zeros(1, 18, 9, 11, dtype=torch.float64), w, None, 1, 1,
                          torch.ones(1, 9, 9, 11, dtype=torch.float64))
    assert (got - torch.nn.functional.conv2d(x, w, None, 1, 1)).abs().max() < 1e-12


def test_point_nms_keeps_up_left_maxima():
    heat = torch.tensor([[[[0.1, 0.9, 0.2], [0.8, 0.3, 0.95], [0.5, 0.5, 0.1]]]])
    out = O.point_nms(heat)
    # a cell survives iff it equals the max over itself and its up / left / up-left neighbours
    exp = torch.tensor([[[[0.1, 0.9, 0.0], [0.8, 0.0, 0.95], [0.0, 0.0, 0.0]]]])
    assert torch.equal(out, exp)


def test_matrix_nms_decays_duplicates_only():
    m = torch.zeros(3, 4, 4, dtype=torch.bool)
    m[0, :2] = True
    m[1, :2] = True          # duplicate of 0
    m[2, 2:] = True          # disjoint
    s = torch.tensor([0.9, 0.8, 0.7])
    out = O.matrix_nms(torch.zeros(3, dtype=torch.long), m, m.flatten(1).sum(1).float(), s)
    assert torch.isclose(out[0], s[0]) and torch.isclose(out[2], s[2])
    assert torch.isclose(out[1], s[1] * torch.exp(torch.tensor(-2.0)))
    assert O.matrix_nms(torch.zeros(0, dtype=torch.long), m[:0], torch.zeros(0), torch.zeros(0)) == []


def test_inference_early_outs_return_none_fields():
    seg = torch.zeros(1, 128, 8, 8)
    depth = torch.ones(1, 1, 16, 16)
    r = O.inference_single(seg, torch.zeros(3728, 2), torch.zeros(3728, 128), depth, (32, 32))
    assert list(r.keys()) == ["pred_masks", "pred_boxes", "pred_classes", "pred_scores", "pred_depth"]
    assert r["pred_scores"] is None and r["pred_depth"].shape == (1, 1, 32, 32)
    # candidates above the score threshold but with empty masks (area <= stride) are dropped too
    cate = torch.zeros(3728, 2)
    cate[5, 0] = 0.9
    r = O.inference_single(seg - 10.0, cate, torch.ones(3728, 128), depth, (32, 32))
    assert r["pred_scores"] is None


def test_mask_nms_matches_reference_golden():
    """oracle.mask_nms (nms.py:53-80 restated) against the keep vectors of the unmodified reference function
    (tests/golden/mask_nms.pt, generated by tests/golden/make_mask_nms_golden.py)."""
    import os
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mask_nms.pt"))
    assert len(cases) == 6
    for c in cases:
        keep = O.mask_nms(c["labels"], c["masks"], c["sums"], c["scores"], nms_thr=c["thr"])
        assert torch.equal(keep, c["keep"])
    assert O.mask_nms(torch.zeros(0), torch.zeros(0, 4, 4), torch.zeros(0), torch.zeros(0)) == []
