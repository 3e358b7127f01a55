"""planerecnet_b200.targets (device-resident ground-truth assignment without the cv2 round trip) against cv2 itself and
against the oracle's restatement of losses.py:200-275 (which is pinned to the unmodified reference by
tests/golden/loss_golden.pt) — exact equality."""
import cv2
import numpy as np
import pytest
import torch

import loss_cases as LC
from oracle import prn_loss_oracle as LO
from planerecnet_b200 import targets as T


@pytest.mark.parametrize("shape", [(3, 480, 640), (1, 64, 96), (7, 128, 160)])
def test_quarter_masks_equals_cv2_bilinear(shape):
    n, H, W = shape
    rng = np.random.default_rng(n)
    m = (rng.random((H, W, n)) < 0.5).astype(np.uint8)
    m[:H // 2, :W // 3] = 1                                   # solid regions and edges, not only noise
    ref = cv2.resize(m, (int(W * 0.25 + 0.5), int(H * 0.25 + 0.5)), interpolation=cv2.INTER_LINEAR)
    ref = ref[..., None] if ref.ndim == 2 else ref
    got = T.quarter_masks(torch.from_numpy(m).permute(2, 0, 1))
    assert got.dtype == torch.uint8 and np.array_equal(got.numpy(), ref.transpose(2, 0, 1))
    with pytest.raises(ValueError):
        T.quarter_masks(torch.zeros(1, 30, 40))


@pytest.mark.parametrize("name", list(LC.CASES))
def test_assignment_equals_oracle(name):
    _, _, _, _, gts, _ = LC.synth(**LC.CASES[name])
    for gt in gts:
        ref = LO.assign_targets(gt, (120, 160))
        got = T.assign_targets(gt, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"], LO.CFG["num_classes"], LO.CFG["sigma"])
        assert len(got) == len(ref) == 4
        for (gi, gc, gd, go), (ri, rc, rd, ro) in zip(got, ref):
            assert go == ro
            assert torch.equal(gc, rc) and torch.equal(gd, rd)
            assert gi.shape == ri.shape and torch.equal(gi, ri)


def test_assignment_skips_empty_masks_and_out_of_range_boxes():
    masks = torch.zeros(3, 480, 640, dtype=torch.uint8)
    masks[1, 100:300, 200:500] = 1
    gt = dict(masks=masks, classes=torch.zeros(3, dtype=torch.int64),
              boxes=torch.tensor([[10, 10, 200, 200], [200, 100, 500, 300], [0, 0, 0.5, 0.5]], dtype=torch.float64))
    got = T.assign_targets(gt, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"])
    ref = LO.assign_targets(gt, (120, 160))
    for (gi, gc, gd, go), (ri, rc, rd, ro) in zip(got, ref):
        assert go == ro and torch.equal(gc, rc) and torch.equal(gd, rd) and torch.equal(gi, ri)
    assert sum(len(t[3]) for t in got) > 0


@pytest.mark.parametrize("name", ["loss_seed0", "loss_seed2_many_tiny"])
def test_ins_lava_orchestration_equals_oracle_on_cpu(name):
    """planerecnet_b200.losses.ins_lava_losses with the torch emulation of its five device steps == the oracle's 'ins' and
    'lav' terms and their gradients w.r.t. the mask features and every level's kernels (fp32, 1e-4)."""
    import numpy as np
    from loss_emulation import EmuBackend
    from planerecnet_b200 import losses as PL
    mask, cate, kern, depth, gts, gt_depth = LC.synth(**LC.CASES[name])
    leaves = [mask] + kern
    for t in leaves:
        t.requires_grad_(True)
    np.random.seed(0)
    ref = LO.loss_forward(mask, cate, kern, depth, gts, gt_depth)
    (ref["ins"] + 2.0 * ref["lav"].sum()).backward()
    ref_g = [None if t.grad is None else t.grad.clone() for t in leaves]
    for t in leaves:
        t.grad = None
    targets = [T.assign_targets(g, (120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"]) for g in gts]
    l_ins, l_lav = PL.ins_lava_losses(mask, kern, targets, gt_depth, backend=EmuBackend())
    assert abs(float(l_ins) - float(ref["ins"])) <= 1e-5 * abs(float(ref["ins"]))
    assert abs(float(l_lav) - float(ref["lav"].sum())) <= 1e-5 * abs(float(ref["lav"].sum()))
    (l_ins + 2.0 * l_lav).backward()
    for t, rg in zip(leaves, ref_g):
        if rg is None:
            assert t.grad is None or float(t.grad.abs().max()) == 0.0
        else:
            assert float((t.grad - rg).abs().max()) <= 1e-4 * float(rg.abs().max())


@pytest.mark.parametrize("name", list(LC.CASES))
def test_assembled_loss_module_equals_reference_golden_on_cpu(name):
    """planerecnet_b200.losses.PlaneRecNetLoss (target assignment + five terms, reference signature) with the torch
    emulation of the device steps == the UNMODIFIED reference's loss values (tests/golden/loss_golden.pt) and gradient norms."""
    import math
    import os
    import numpy as np
    from loss_emulation import EmuBackend
    from planerecnet_b200 import losses as PL
    from planerecnet_b200.config import cfg, set_cfg
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.pt"))[name]
    set_cfg("PlaneRecNet_101_config")
    crit = PL.PlaneRecNetLoss(cfg, backend=EmuBackend())
    mask, cate, kern, depth, gts, gt_depth = LC.synth(**LC.CASES[name])
    leaves = [mask] + cate + kern + [depth]
    for t in leaves:
        t.requires_grad_(True)
    np.random.seed(0)
    out = crit(None, mask, cate, kern, depth, gts, gt_depth)
    assert list(out) == ["ins", "cat", "dpt", "pln", "lav"]
    for k, v in gold["losses"].items():
        got = float(out[k].detach().sum())
        assert (math.isnan(got) if math.isnan(v) else abs(got - v) <= 2e-5 * max(1.0, abs(v))), (k, got, v)
    if not any(math.isnan(v) for v in gold["losses"].values()):
        sum(v.sum() for v in out.values()).backward()
        got = [0.0 if t.grad is None else float(t.grad.double().norm()) for t in leaves]
        for a, b in zip(got, gold["grad_norms"]):
            assert abs(a - b) <= 2e-4 * max(b, 1e-6), (got, gold["grad_norms"])


@pytest.mark.parametrize("name", list(LC.CASES))
def test_batched_assignment_equals_per_image_assignment(name):
    """targets.assign_targets_batch (one host round trip per batch) == [assign_targets(gt) for gt in batch], exactly."""
    _, _, _, _, gts, _ = LC.synth(**LC.CASES[name])
    args = ((120, 160), LO.CFG["grids"], LO.CFG["scale_ranges"], LO.CFG["num_classes"], LO.CFG["sigma"])
    ref = [T.assign_targets(g, *args) for g in gts]
    got = T.assign_targets_batch(gts, *args)
    assert len(got) == len(ref)
    for rb, gb in zip(ref, got):
        assert len(rb) == len(gb) == 4
        for (ri, rc, rd, ro), (gi, gc, gd, go) in zip(rb, gb):
            assert go == ro and torch.equal(gc, rc) and torch.equal(gd, rd)
            assert gi.shape == ri.shape and gi.dtype == ri.dtype and torch.equal(gi, ri)


def test_unconsumed_preparation_is_drained_before_the_next_one():
    """PlaneRecNetLoss.prepare(A) followed by prepare(B) without a forward in between: A's sampler thread has finished (and has
    advanced numpy's global stream by exactly its draws) before B's starts, so B's loss equals the one of a run that drew A's
    triplets and then B's, in order."""
    import numpy as np
    from loss_emulation import EmuBackend
    from planerecnet_b200 import losses as PL
    from planerecnet_b200.config import cfg, set_cfg
    set_cfg("PlaneRecNet_101_config")
    names = list(LC.CASES)
    a = LC.synth(**LC.CASES[names[0]])
    b = LC.synth(**LC.CASES[names[1]])

    def loss_b(crit):
        mask, cate, kern, depth, gts, gt_depth = b
        return crit(None, mask, cate, kern, depth, gts, gt_depth)["pln"].detach().double()

    np.random.seed(3)
    crit = PL.PlaneRecNetLoss(cfg, backend=EmuBackend())
    crit.prepare(a[4])                     # never consumed
    crit.prepare(b[4])
    got = loss_b(crit)
    np.random.seed(3)
    ref_crit = PL.PlaneRecNetLoss(cfg, backend=EmuBackend())
    pa = ref_crit.vnl_batched.prepare(a[4])
    if pa["thread"] is not None:
        pa["thread"].join()
    ref = loss_b(ref_crit)
    assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(ref)), (got, ref)


@pytest.mark.parametrize("name", list(LC.CASES))
def test_dense_batch_layout_matches_the_per_level_lists(name):
    """TargetsBatch.dense (what losses._InsLava gathers from in one shot) against the per-(image, level) results it replaces: row r
    of image b is the r-th positive cell in level order, gidx = level offset + cell, tgt row = that cell's target mask, valid marks
    exactly the n_b real rows, cate = the level-concatenated label maps, padding rows are zero."""
    _, _, _, _, gts, _ = LC.synth(**LC.CASES[name])
    grids = LO.CFG["grids"]
    tb = T.assign_targets_batch(gts, (120, 160), grids, LO.CFG["scale_ranges"], LO.CFG["num_classes"], LO.CFG["sigma"])
    d = tb.dense
    offs = [sum(S * S for S in grids[:l]) for l in range(len(grids))]
    assert d["level_off"] == offs and d["n"] % 16 == 0 and d["tgt"].shape == (len(gts), d["n"], 120 * 160)
    for b, per_level in enumerate(tb):
        rows_idx, rows_tgt = [], []
        for l, (ins, cate, ind, order) in enumerate(per_level):
            rows_idx += [offs[l] + c for c in order]
            rows_tgt.append(ins.reshape(len(order), 120 * 160))
            assert torch.equal(d["cate"][b, offs[l]:offs[l] + grids[l] ** 2].view(grids[l], grids[l]), cate)
        n_b = len(rows_idx)
        assert d["n_b"][b] == n_b
        assert d["gidx"][b, :n_b].tolist() == rows_idx
        assert bool(d["valid"][b, :n_b].all()) and not bool(d["valid"][b, n_b:].any())
        if n_b:
            assert torch.equal(d["tgt"][b, :n_b], torch.cat(rows_tgt))
        assert int(d["tgt"][b, n_b:].sum()) == 0
