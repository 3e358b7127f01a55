"""Bookkeeping parity (planerecnet.py:189-289, nms.py) on IDENTICAL dense inputs: the batched, sync-light
device implementation must select the same candidates, in the same order, with the same labels, as the oracle's
per-image restatement — exact on indices / labels / counts / boxes, 1e-5 on scores."""
import ctypes as C

import pytest
import torch

import helpers as H
from oracle import prn_oracle as O
from planerecnet_b200 import _lib as L
from planerecnet_b200 import ops, postprocess as PP

pytestmark = pytest.mark.gpu
GRIDS = [40, 36, 24, 16]
TOTAL = sum(g * g for g in GRIDS)


def _synthetic(B, h, w, seed, n_cand):
    """Category scores with ~n_cand candidates per image and a per-row table of blob-like sigmoid masks."""
    g = torch.Generator().manual_seed(seed)
    scores = torch.zeros(B, TOTAL, 2)
    P = h * w
    seg_table = torch.zeros(B, TOTAL, P)
    yy, xx = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    for b in range(B):
        rows = torch.randperm(TOTAL, generator=g)[:n_cand]
        for r in rows.tolist():
            c = int(torch.randint(0, 2, (1,), generator=g))
            scores[b, r, c] = 0.12 + 0.8 * float(torch.rand(1, generator=g))
            if float(torch.rand(1, generator=g)) < 0.15:
                scores[b, r, 1 - c] = 0.11 + 0.5 * float(torch.rand(1, generator=g))
            cy, cx = float(torch.rand(1, generator=g)) * h, float(torch.rand(1, generator=g)) * w
            rad = 1.0 + 6.0 * float(torch.rand(1, generator=g))       # some blobs are too small for their stride
            d2 = ((yy - cy) ** 2 + (xx - cx) ** 2) / (rad * rad)
            seg_table[b, r] = torch.sigmoid(4.0 - 4.0 * d2).flatten() * (0.6 + 0.4 * float(torch.rand(1, generator=g)))
    return scores, seg_table


@pytest.mark.parametrize("seed,n_cand", [(0, 40), (1, 150), (2, 3)])
def test_selection_and_matrix_nms_match_oracle_exactly(cuda_lib, seed, n_cand):
    B, h, w = 2, 16, 20          # P = 320 = 5 x 64 (Gram contraction needs a multiple of 64)
    P = h * w
    scores, seg_table = _synthetic(B, h, w, seed, n_cand)
    strides_all = torch.tensor([s for gsz, s in zip(GRIDS, O.INSTANCE_STRIDES) for _ in range(gsz * gsz)], dtype=torch.float32)
    # ---- oracle, per image
    exp = []
    for b in range(B):
        cate = scores[b]
        inds = cate > O.INFER["score_thr"]
        cs = cate[inds]
        nz = inds.nonzero(as_tuple=False)
        seg = seg_table[b, nz[:, 0]].reshape(-1, h, w)
        det = O.bookkeeping(seg, cs, nz[:, 1], strides_all[nz[:, 0]]) if len(cs) else None
        exp.append(None if det is None else (nz[det[3], 0] * 2 + nz[det[3], 1], det[1], det[2], det[0]))
    # ---- device
    from planerecnet_b200.engine import Engine
    eng = Engine("f16")
    dev = "cuda"
    tab = seg_table.to(dev)

    def seg_fn(rows, valid, n):
        seg32 = tab.gather(1, rows[:, :, None].expand(-1, -1, P)).reshape(B * n, P).contiguous()
        m16 = torch.empty(B * n, P, dtype=eng.tdt, device=dev)
        area = torch.empty(B * n, device=dev)
        ssum = torch.empty(B * n, device=dev)
        eng._call(eng.lib.prn_mask_stats, C.c_void_p(seg32.data_ptr()), C.c_void_p(m16.data_ptr()), C.c_void_p(area.data_ptr()),
                  C.c_void_p(ssum.data_ptr()), B * n, P, C.c_float(O.INFER["mask_thr"]), eng.dt, eng._st())
        mk = (seg32 > O.INFER["mask_thr"]).float().view(B, n, P)
        assert torch.equal(area.view(B, n), mk.sum(-1))          # the kernel's areas are exact

        def gram_fn(order, n1):
            msel = m16.view(B, n, P).gather(1, order[:, :, None].expand(-1, -1, P))
            inter = torch.empty(B, n1, n1, device=dev)
            ops.conv2d(msel.view(B, n1, 1, P), msel.view(B * n1, P), batch=B, h_in=n1, w_in=1, ksize=1, out32=inter,
                       ld_out32=n1, n_pad=n1, w_group_rows=n1, dtype=eng.dt)
            ms = mk.gather(1, order[:, :, None].expand(-1, -1, P))
            assert torch.equal(inter, torch.bmm(ms, ms.transpose(1, 2)))     # intersections are exact integers
            return inter

        return seg32, m16, area, ssum, gram_fn

    p = dict(score_thr=0.1, mask_thr=0.1, update_thr=0.15, nms_pre=500, top_k=100, sigma=2.0, kernel="gaussian")
    dets, seg32 = PP.select(scores.to(dev), seg_fn, strides_all.to(dev), 2, p)
    n = seg32.shape[0] // B if seg32 is not None else 0
    for b in range(B):
        rows, sc, lab = dets[b]
        if exp[b] is None:
            assert rows is None
            continue
        e_flat, e_sc, e_lab, e_seg = exp[b]
        assert rows is not None and rows.numel() == e_flat.numel(), "detection count differs"
        assert torch.equal(lab.cpu(), e_lab)
        assert torch.allclose(sc.cpu(), e_sc, rtol=1e-5, atol=1e-6)
        # same candidates in the same order: compare the selected mask rows themselves
        assert torch.equal(seg32[rows].cpu(), e_seg.reshape(-1, P))


def test_point_nms_kernel_matches_oracle(cuda_lib):
    from planerecnet_b200.engine import Engine
    eng = Engine("f16")
    B = 3
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(B, TOTAL, 16, generator=g) * 3
    logits[0, :40, 0] = 2.5            # plateaus: equal neighbours must all survive (== comparison)
    grids = torch.tensor(GRIDS, dtype=torch.int32, device="cuda")
    out = torch.empty(B, TOTAL, 2, device="cuda")
    eng._call(eng.lib.prn_point_nms_sigmoid, C.c_void_p(logits.cuda().data_ptr()), C.c_void_p(out.data_ptr()), B, TOTAL, 16, 2, 4,
              C.c_void_p(grids.data_ptr()), eng._st())
    off = 0
    for S in GRIDS:
        lg = logits[:, off:off + S * S, :2].reshape(B, S, S, 2).permute(0, 3, 1, 2)
        exp = O.point_nms(lg.sigmoid()).permute(0, 2, 3, 1).reshape(B, S * S, 2)
        got = out[:, off:off + S * S].cpu()
        assert torch.equal(got > 0, exp > 0), "kept cells differ"
        assert torch.allclose(got, exp, atol=1e-6)
        off += S * S


def test_upsample_mask_box_kernel(cuda_lib):
    from planerecnet_b200.engine import Engine
    eng = Engine("f16")
    g = torch.Generator().manual_seed(2)
    seg = torch.rand(6, 30, 40, generator=g) * 0.3
    seg[1, 5:12, 7:19] += 0.5
    seg[4, 20:25, 30:40] += 0.4
    sel = torch.tensor([4, 1, 0], dtype=torch.int32)
    Hh, Ww = 120, 160
    masks = torch.empty(3, Hh, Ww, dtype=torch.bool, device="cuda")
    boxes = torch.tensor([Ww, Hh, -1, -1], dtype=torch.int32, device="cuda").repeat(3, 1)
    eng._call(eng.lib.prn_upsample_mask_box, C.c_void_p(seg.cuda().data_ptr()), C.c_void_p(sel.cuda().data_ptr()),
              C.c_void_p(masks.data_ptr()), C.c_void_p(boxes.data_ptr()), 3, 30, 40, Hh, Ww, C.c_float(0.1), eng._st())
    up = torch.nn.functional.interpolate(seg[sel.long()].unsqueeze(0), size=(Hh, Ww), mode="bilinear", align_corners=False)[0]
    exp = up > 0.1
    got = masks.cpu()
    borderline = (up - 0.1).abs() < 1e-6           # fp contraction differences may flip exact ties only
    assert torch.equal(got | borderline, exp | borderline)
    for i in range(3):
        ys, xs = torch.where(got[i])
        assert boxes[i].tolist() == [int(xs.min()), int(ys.min()), int(xs.max()), int(ys.max())]


def test_mask_nms_kernel_matches_reference_golden(cuda_lib):
    """prn_mask_nms_greedy on the Gram matrix of the golden masks == the unmodified reference's keep vector."""
    import os
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mask_nms.pt"))
    for c in cases:
        n = c["masks"].shape[0]
        m = c["masks"].reshape(n, -1).float().cuda()
        inter = (m @ m.t()).contiguous()[None]
        keep = torch.empty(1, n, dtype=torch.uint8, device="cuda")
        valid = torch.ones(1, n, dtype=torch.uint8, device="cuda")
        sums, labels = c["sums"].cuda(), c["labels"].long().cuda()       # named: the pointers must outlive the call
        L.check(L.lib().prn_mask_nms_greedy(C.c_void_p(inter.data_ptr()), C.c_void_p(sums.data_ptr()),
                                            C.c_void_p(labels.data_ptr()), C.c_void_p(valid.data_ptr()),
                                            C.c_void_p(keep.data_ptr()), 1, n, C.c_float(c["thr"]), L.current_stream()))
        torch.cuda.synchronize()
        assert torch.equal(keep[0].bool().cpu(), c["keep"]), (n, c["thr"], keep[0].tolist(), c["keep"].tolist())


@pytest.mark.parametrize("seed,n_cand", [(0, 40), (1, 150)])
def test_selection_with_mask_nms_matches_oracle_exactly(cuda_lib, seed, n_cand):
    """nms_type == 'mask' (planerecnet.py:249-252) through the batched device selection vs the oracle, same inputs."""
    B, h, w = 2, 16, 20
    P = h * w
    scores, seg_table = _synthetic(B, h, w, seed, n_cand)
    strides_all = torch.tensor([s for gsz, s in zip(GRIDS, O.INSTANCE_STRIDES) for _ in range(gsz * gsz)], dtype=torch.float32)
    pm = dict(O.INFER, nms_type="mask")
    exp = []
    for b in range(B):
        cate = scores[b]
        inds = cate > pm["score_thr"]
        nz = inds.nonzero(as_tuple=False)
        seg = seg_table[b, nz[:, 0]].reshape(-1, h, w)
        det = O.bookkeeping(seg, cate[inds], nz[:, 1], strides_all[nz[:, 0]], p=pm)
        exp.append(None if det is None else (det[1], det[2], det[0]))
    from planerecnet_b200.engine import Engine
    eng = Engine("f16")
    tab = seg_table.cuda()

    def seg_fn(rows, valid, n):
        seg32 = tab.gather(1, rows[:, :, None].expand(-1, -1, P)).reshape(B * n, P).contiguous()
        m16 = torch.empty(B * n, P, dtype=eng.tdt, device="cuda")
        area = torch.empty(B * n, device="cuda")
        ssum = torch.empty(B * n, device="cuda")
        eng._call(eng.lib.prn_mask_stats, C.c_void_p(seg32.data_ptr()), C.c_void_p(m16.data_ptr()), C.c_void_p(area.data_ptr()),
                  C.c_void_p(ssum.data_ptr()), B * n, P, C.c_float(pm["mask_thr"]), eng.dt, eng._st())

        def gram_fn(order, n1):
            msel = m16.view(B, n, P).gather(1, order[:, :, None].expand(-1, -1, P))
            inter = torch.empty(B, n1, n1, device="cuda")
            ops.conv2d(msel.view(B, n1, 1, P), msel.view(B * n1, P), batch=B, h_in=n1, w_in=1, ksize=1, out32=inter,
                       ld_out32=n1, n_pad=n1, w_group_rows=n1, dtype=eng.dt)
            return inter

        return seg32, m16, area, ssum, gram_fn

    p = dict(score_thr=0.1, mask_thr=0.1, update_thr=0.15, nms_pre=500, top_k=100, sigma=2.0, kernel="gaussian", nms_type="mask")
    dets, seg32 = PP.select(scores.cuda(), seg_fn, strides_all.cuda(), 2, p)
    for b in range(B):
        rows, sc, lab = dets[b]
        if exp[b] is None:
            assert rows is None
            continue
        e_sc, e_lab, e_seg = exp[b]
        assert rows is not None and rows.numel() == e_sc.numel(), "detection count differs"
        assert torch.equal(lab.cpu(), e_lab)
        assert torch.allclose(sc.cpu(), e_sc, rtol=1e-5, atol=1e-6)
        assert torch.equal(seg32[rows].cpu(), e_seg.reshape(-1, P))


def test_module_level_mask_nms_reference_signature(cuda_lib):
    import os
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mask_nms.pt"))
    for c in cases:
        keep = PP.mask_nms(c["labels"].cuda(), c["masks"].cuda(), c["sums"].cuda(), c["scores"].cuda(), nms_thr=c["thr"])
        assert torch.equal(keep.cpu(), c["keep"])
    assert PP.mask_nms(torch.zeros(0), torch.zeros(0, 4, 4), torch.zeros(0), torch.zeros(0)) == []


@pytest.mark.gpu
def test_pack_mask_bits_equals_numpy_packbits(cuda_lib):
    import numpy as np
    from planerecnet_b200.postprocess import pack_mask_bits
    g = torch.Generator().manual_seed(0)
    base = (torch.rand(5, 48, 64, generator=g) < 0.4).cuda()
    views = [base[0:2], base[2:2], base[2:5]]                      # per-image views of one buffer, one of them empty
    got = pack_mask_bits(views).cpu().numpy()
    ref = np.packbits(base.cpu().numpy().reshape(-1), bitorder="little")
    assert np.array_equal(got, ref)
    sep = [base[0:2].clone(), base[2:5].clone()]                    # separate allocations: concatenated first
    assert np.array_equal(pack_mask_bits(sep).cpu().numpy(), ref)
    assert pack_mask_bits([None, base[0:0]]) is None
