"""Inference bookkeeping on top of the dense outputs (planerecnet.py:106-111, 155-289; nms.py:8-50),
batched over the images of a step with three host synchronisations per batch instead of ~8 per image.

Dense / reduction work runs in libprn_b200 kernels: sigmoid + point-NMS (`prn_point_nms_sigmoid`), the
per-candidate mask contraction and the matrix-NMS Gram matrix on the tensor-core conv kernel
(`prn_conv2d_fwd`, per-image operands), mask area / maskness sums (`prn_mask_stats`), final bilinear
upsampling + thresholding + boxes (`prn_upsample_mask_box`).  Index bookkeeping (compaction, sorting,
the n x n decay algebra on <= 500 candidates) is expressed with torch device ops and follows the
reference's order of operations, so identical dense inputs give identical indices and counts."""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib as L
from . import ops


def point_nms(heat, kernel=2):
    """nms.py:8-12 (torch formulation, kept for callers of models.functions.nms)."""
    hmax = F.max_pool2d(heat, (kernel, kernel), stride=1, padding=1)
    keep = (hmax[:, :, :-1, :-1] == heat).float()
    return heat * keep


def matrix_nms(cate_labels, seg_masks, sum_masks, cate_scores, sigma=2.0, kernel="gaussian"):
    """nms.py:15-50 for one image (seg_masks given as [n, pixels] float)."""
    n = len(cate_labels)
    if n == 0:
        return []
    inter = torch.mm(seg_masks, seg_masks.t())
    return _decay(inter[None], sum_masks[None], cate_labels[None], cate_scores[None],
                  torch.ones(1, n, dtype=torch.bool, device=cate_scores.device), sigma, kernel)[0]


def mask_nms(cate_labels, seg_masks, sum_masks, cate_scores, nms_thr=0.5):
    """nms.py:53-80 for one image, reference signature (seg_masks [n, h, w] or [n, pixels]; candidates sorted by score):
    returns the bool keep vector.  The greedy pass runs in prn_mask_nms_greedy on the masks' Gram matrix."""
    n = len(cate_scores)
    if n == 0:
        return []
    m = seg_masks.reshape(n, -1).float()
    inter = torch.mm(m, m.t()).contiguous()[None]
    area = sum_masks.float().contiguous()[None]
    labels = cate_labels.long().contiguous()[None]
    valid = torch.ones(1, n, dtype=torch.uint8, device=m.device)
    keep = torch.empty(1, n, dtype=torch.uint8, device=m.device)
    L.check(L.lib().prn_mask_nms_greedy(C.c_void_p(inter.data_ptr()), C.c_void_p(area.data_ptr()), C.c_void_p(labels.data_ptr()),
                                        C.c_void_p(valid.data_ptr()), C.c_void_p(keep.data_ptr()), 1, n, C.c_float(nms_thr),
                                        L.current_stream()), "prn_mask_nms_greedy")
    return keep[0].bool()


def _decay(inter, areas, labels, scores, valid, sigma, kernel):
    """Batched nms.py:24-48.  inter [B,n,n] mask intersections (sorted order), areas/labels/scores/valid [B,n]."""
    n = inter.shape[-1]
    sx = areas[:, None, :].expand(-1, n, n)                     # sum_masks_x[i][j] = area_j
    pair = valid[:, :, None] & valid[:, None, :]
    union = sx + sx.transpose(1, 2) - inter
    iou = torch.where(pair, inter / union, torch.zeros_like(inter)).triu(diagonal=1)
    label = ((labels[:, None, :] == labels[:, :, None]) & pair).float().triu(diagonal=1)
    decay_iou = iou * label
    comp = decay_iou.max(1).values                               # per column j: max over i
    comp_m = comp[:, :, None].expand(-1, n, n)                   # [i][j] = comp[i]
    if kernel == "linear":
        coef = ((1 - decay_iou) / (1 - comp_m)).min(1).values
    else:
        coef = (torch.exp(-1 * sigma * (decay_iou ** 2)) / torch.exp(-1 * sigma * (comp_m ** 2))).min(1).values
    return scores * coef


def _strides_table(net, device):
    t = []
    for g, s in zip(net.num_grids, net.instance_strides):
        t += [float(s)] * (g * g)
    return torch.tensor(t, device=device)


def select(scores, seg_fn, strides_all, nc, p):
    """Candidate selection + matrix-NMS for a batch (planerecnet.py:189-269).

    scores: fp32 [B, total, nc] after point-NMS.  seg_fn(rows [B,n] long, valid, n) -> (seg32 [B*n, P] fp32 sigmoid
    masks, mask16 [B*n, P] 0/1, area [B*n], ssum [B*n], gram_fn) for the candidates' kernel rows;
    gram_fn(order [B,n1], n1) -> [B, n1, n1] fp32 mask intersections of the selected candidates, in that order.
    Returns per image (sel_rows into seg32, scores, labels) of the final detections (possibly empty) + seg32."""
    B, total, _ = scores.shape
    dev = scores.device
    flat = scores.reshape(B, total * nc)
    cand = flat > p["score_thr"]
    counts = cand.sum(1)
    n_max = int(counts.max().item())                              # host sync 1
    empty = [(None, None, None)] * B
    if n_max == 0:
        return empty, None
    n = ops.round_up(n_max, 16)
    # stable compaction in reference order (row-major over [total, nc]): rank of every candidate within its image
    rank = torch.cumsum(cand, 1) - 1
    slot = torch.where(cand, rank, torch.full_like(rank, n))      # non-candidates -> overflow slot n
    src = torch.arange(total * nc, device=dev).expand(B, -1)
    idx = torch.full((B, n + 1), -1, dtype=torch.long, device=dev)
    idx.scatter_(1, slot, src)
    idx = idx[:, :n]                                              # [B, n] flat index or -1
    valid = idx >= 0
    idxc = idx.clamp(min=0)
    rows = idxc // nc
    labels = idxc % nc
    sc = torch.where(valid, flat.gather(1, idxc), torch.zeros((), device=dev))
    strides = strides_all[rows]
    seg32, mask16, area, ssum, gram_fn = seg_fn(rows, valid, n)
    area = area.view(B, n)
    ssum = ssum.view(B, n)
    keep = valid & (area > strides)                               # planerecnet.py:219-220
    sc = sc * (ssum / area)                                       # maskness (planerecnet.py:231-232)
    key = torch.where(keep, sc, torch.full_like(sc, -1.0))
    order = torch.argsort(key, dim=1, descending=True)
    n_pre = min(n, ops.round_up(p["nms_pre"], 16))                # sort, keep the best nms_pre (planerecnet.py:235-242)
    order = order[:, :n_pre]
    v1 = keep.gather(1, order)
    if n_pre > p["nms_pre"]:
        v1[:, p["nms_pre"]:] = False
    s1 = key.gather(1, order)
    a1 = area.gather(1, order)
    l1 = labels.gather(1, order)
    inter1 = gram_fn(order, n_pre)                                # intersections of the sorted survivors only
    if p.get("nms_type", "matrix") == "mask":                     # planerecnet.py:249-252: greedy, scores unchanged
        keep_u8 = torch.empty(B, n_pre, dtype=torch.uint8, device=dev)
        v1_u8 = v1.to(torch.uint8).contiguous()
        a1c, l1c = a1.contiguous(), l1.contiguous()
        L.check(L.lib().prn_mask_nms_greedy(C.c_void_p(inter1.data_ptr()), C.c_void_p(a1c.data_ptr()), C.c_void_p(l1c.data_ptr()),
                                            C.c_void_p(v1_u8.data_ptr()), C.c_void_p(keep_u8.data_ptr()), B, n_pre,
                                            C.c_float(p["mask_thr"]), L.current_stream()), "prn_mask_nms_greedy")
        s2 = s1
        keep2 = keep_u8.bool()
    else:
        s2 = _decay(inter1, a1, l1, s1, v1, p["sigma"], p["kernel"])  # nms.py
        keep2 = v1 & (s2 >= p["update_thr"])
    key2 = torch.where(keep2, s2, torch.full_like(s2, -1.0))
    order2 = torch.argsort(key2, dim=1, descending=True)[:, :p["top_k"]]
    fin_valid = keep2.gather(1, order2)
    fin_score = key2.gather(1, order2)
    fin_label = l1.gather(1, order2)
    fin_slot = order.gather(1, order2)                            # slot in the candidate list
    n_fin = fin_valid.sum(1).tolist()                             # host sync 2
    out = []
    base = torch.arange(B, device=dev)[:, None] * n
    fin_rows = (base + fin_slot)
    for b in range(B):
        k = n_fin[b]
        if k == 0:
            out.append((None, None, None))
        else:
            out.append((fin_rows[b, :k], fin_score[b, :k], fin_label[b, :k]))
    return out, seg32


def inference(eng, net, st, x):
    """Returns list[dict] with the reference's keys in the reference's order (planerecnet.py:183)."""
    B, _, H, W = x.shape
    dev = x.device
    inst = st["inst"]
    nc = net.num_classes
    total = inst["cate32"].shape[1]
    lib = eng.lib
    p = dict(score_thr=net.score_threshold, mask_thr=net.mask_threshold, update_thr=net.update_threshold,
             nms_pre=net.max_before_nms, top_k=net.max_per_img, sigma=net.nms_sigma, kernel=net.nms_kernel,
             nms_type=net.nms_type)
    if net.nms_type not in ("matrix", "mask"):
        raise NotImplementedError          # planerecnet.py:253-254

    grids = eng._zero_pool.get(("grids", tuple(net.num_grids)))
    if grids is None:
        grids = torch.tensor(list(net.num_grids), dtype=torch.int32, device=dev)
        eng._zero_pool[("grids", tuple(net.num_grids))] = grids
        eng._zero_pool[("strides", tuple(net.num_grids))] = _strides_table(net, dev)
    strides_all = eng._zero_pool[("strides", tuple(net.num_grids))]

    scores = torch.empty(B, total, nc, device=dev)
    eng._call(lib.prn_point_nms_sigmoid, C.c_void_p(inst["cate32"].data_ptr()), C.c_void_p(scores.data_ptr()), B, total,
              inst["cate32"].shape[-1], nc, len(net.num_grids), C.c_void_p(grids.data_ptr()), eng._st())

    mask16 = st["mask16"]
    _, mh, mw, mc = mask16.shape
    P = mh * mw

    def seg_fn(rows, valid, n):
        # per-image contraction sigmoid(K_sel . mask^T): rows of A = selected kernels, "weights" = mask pixels
        wsel = inst["kern16"].gather(1, rows[:, :, None].expand(-1, -1, mc)) * valid[:, :, None].to(eng.tdt)
        seg32 = torch.empty(B * n, P, device=dev)
        eng.launches += 1
        ops.conv2d(wsel, mask16.reshape(B * P, mc), batch=B, h_in=n, w_in=1, ksize=1, act=L.ACT_SIGMOID, out32=seg32,
                   ld_out32=P, n_pad=P, w_group_rows=P, dtype=eng.dt)
        m16 = torch.empty(B * n, P, dtype=eng.tdt, device=dev)
        area = torch.empty(B * n, device=dev)
        ssum = torch.empty(B * n, device=dev)
        eng._call(lib.prn_mask_stats, C.c_void_p(seg32.data_ptr()), C.c_void_p(m16.data_ptr()), C.c_void_p(area.data_ptr()),
                  C.c_void_p(ssum.data_ptr()), B * n, P, C.c_float(p["mask_thr"]), eng.dt, eng._st())
        def gram_fn(order, n1):
            # Gram matrix of the binary masks of the sorted top-nms_pre candidates (nms.py:20-22): exact in fp32
            # (integer counts < 2^24), on the tensor cores with per-image operands
            msel = m16.view(B, n, P).gather(1, order[:, :, None].expand(-1, -1, P))
            inter = torch.empty(B, n1, n1, device=dev)
            eng.launches += 1
            ops.conv2d(msel.view(B, n1, 1, P), msel.view(B * n1, P), batch=B, h_in=n1, w_in=1, ksize=1, act=L.ACT_NONE,
                       out32=inter, ld_out32=n1, n_pad=n1, w_group_rows=n1, dtype=eng.dt)
            return inter

        return seg32, m16, area, ssum, gram_fn

    dets, seg32 = select(scores, seg_fn, strides_all, nc, p)

    depth = st["depth32"][..., 0].unsqueeze(1)                         # [B,1,H/2,W/2]
    depth_up = F.interpolate(depth, size=(H, W), mode="bilinear", align_corners=False)
    sel_all = [d[0] for d in dets if d[0] is not None]
    masks_all = boxes_all = None
    if sel_all:
        sel = torch.cat(sel_all).to(torch.int32)
        nf = sel.numel()
        masks_all = torch.empty(nf, H, W, dtype=torch.bool, device=dev)
        boxes_i = torch.tensor([W, H, -1, -1], dtype=torch.int32, device=dev).repeat(nf, 1)
        eng._call(lib.prn_upsample_mask_box, C.c_void_p(seg32.data_ptr()), C.c_void_p(sel.data_ptr()),
                  C.c_void_p(masks_all.data_ptr()), C.c_void_p(boxes_i.data_ptr()), nf, mh, mw, H, W, C.c_float(p["mask_thr"]),
                  eng._st())
        boxes_all = boxes_i.float()
    results, off = [], 0
    for b in range(B):
        r = {"pred_masks": None, "pred_boxes": None, "pred_classes": None, "pred_scores": None, "pred_depth": depth_up[b:b + 1]}
        rows, sc, lab = dets[b]
        if rows is not None:
            k = rows.numel()
            r.update(pred_masks=masks_all[off:off + k], pred_boxes=boxes_all[off:off + k], pred_classes=lab, pred_scores=sc)
            off += k
        results.append(r)
    return results


def pack_mask_bits(masks):
    """pred_masks (bool [n, H, W] tensors, e.g. one per image of a batch) -> one uint8 tensor of bits, LSB first
    (numpy.unpackbits(..., bitorder="little") restores them): what a serving loop copies to the host instead of one byte
    per pixel.  The per-image masks of one `net(x)` call are views of one buffer: packed with a single launch, no copy."""
    masks = [m for m in masks if m is not None and m.numel()]
    if not masks:
        return None
    contiguous = all(m.is_contiguous() for m in masks) and all(
        masks[i + 1].data_ptr() == masks[i].data_ptr() + masks[i].numel() for i in range(len(masks) - 1))
    n = sum(m.numel() for m in masks)
    src = masks[0] if contiguous else torch.cat([m.reshape(-1) for m in masks])
    assert n % 16 == 0 and src.dtype == torch.bool
    out = torch.empty(n // 8, dtype=torch.uint8, device=src.device)
    L.check(L.lib().prn_pack_mask_bits(C.c_void_p(src.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int64(n), L.current_stream()),
            "prn_pack_mask_bits")
    return out
