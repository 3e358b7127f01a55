"""Inference bookkeeping on top of the dense outputs (planerecnet.py:106-111, 155-289; nms.py:8-50).

The per-candidate mask contraction (F.conv2d with the selected kernels, planerecnet.py:210-212) runs on
the tensor-core conv kernel; index bookkeeping (thresholding, sorting, top-k, matrix-NMS on <= 500
candidates) is expressed with torch device ops on the engine's outputs and keeps the reference's
order of operations so that, given identical dense inputs, indices and counts are identical."""
import torch
import torch.nn.functional as F

from . import _lib as L
from . import ops


def point_nms(heat, kernel=2):
    """nms.py:8-12."""
    hmax = F.max_pool2d(heat, (kernel, kernel), stride=1, padding=1)
    keep = (hmax[:, :, :-1, :-1] == heat).float()
    return heat * keep


def matrix_nms(cate_labels, seg_masks, sum_masks, cate_scores, sigma=2.0, kernel="gaussian"):
    """nms.py:15-50 (seg_masks given as [n, pixels] float)."""
    n = len(cate_labels)
    if n == 0:
        return []
    inter = torch.mm(seg_masks, seg_masks.t())
    sx = sum_masks.expand(n, n)
    iou = (inter / (sx + sx.t() - inter)).triu(diagonal=1)
    lx = cate_labels.expand(n, n)
    label = (lx == lx.t()).float().triu(diagonal=1)
    comp, _ = (iou * label).max(0)
    comp = comp.expand(n, n).t()
    decay = iou * label
    if kernel == "linear":
        coef, _ = ((1 - decay) / (1 - comp)).min(0)
    else:
        coef, _ = (torch.exp(-1 * sigma * (decay ** 2)) / torch.exp(-1 * sigma * (comp ** 2))).min(0)
    return cate_scores * coef


def _strides_table(net, device):
    t = []
    for g, s in zip(net.num_grids, net.instance_strides):
        t += [float(s)] * (g * g)
    return torch.tensor(t, device=device)


def inference(eng, net, st, x):
    """Returns list[dict] with the reference's keys in the reference's order (planerecnet.py:183)."""
    B, _, H, W = x.shape
    inst = st["inst"]
    nc = net.num_classes
    total = inst["cate32"].shape[1]
    # sigmoid + point-NMS per level on [B, nc, S, S] views of the category logits
    scores = torch.empty(B, total, nc, device=x.device)
    off = 0
    for S in net.num_grids:
        lg = inst["cate32"][:, off:off + S * S, :nc].reshape(B, S, S, nc).permute(0, 3, 1, 2)
        scores[:, off:off + S * S] = point_nms(lg.sigmoid()).permute(0, 2, 3, 1).reshape(B, S * S, nc)
        off += S * S
    depth = st["depth32"][..., 0].unsqueeze(1)                         # [B,1,H/2,W/2]
    strides_all = _strides_table(net, x.device)
    mask16 = st["mask16"]
    _, mh, mw, mc = mask16.shape
    results = []
    for b in range(B):
        result = {"pred_masks": None, "pred_boxes": None, "pred_classes": None, "pred_scores": None, "pred_depth": None}
        result["pred_depth"] = F.interpolate(depth[b:b + 1], size=(H, W), mode="bilinear", align_corners=False)
        results.append(result)
        cate = scores[b]
        inds = cate > net.score_threshold
        cate_scores = cate[inds]
        if len(cate_scores) == 0:
            continue
        inds = inds.nonzero(as_tuple=False)
        cate_labels = inds[:, 1]
        sel = inds[:, 0]
        strides = strides_all[sel]
        n = sel.numel()
        n_pad = ops.round_up(n, 16)
        wsel = torch.zeros(n_pad, mc, dtype=eng.tdt, device=x.device)
        wsel[:n] = inst["kern16"][b, sel]
        seg = torch.empty(mh * mw, n_pad, dtype=torch.float32, device=x.device)
        eng.launches += 1
        ops.conv2d(mask16[b:b + 1], wsel, batch=1, h_in=mh, w_in=mw, ksize=1, act=L.ACT_SIGMOID, out32=seg,
                   dtype=eng.dt)
        seg = seg[:, :n].t().contiguous()                              # [n, mh*mw]
        seg_masks = seg > net.mask_threshold
        sum_masks = seg_masks.sum(1).float()
        keep = sum_masks > strides
        if keep.sum() == 0:
            continue
        seg_masks, seg, sum_masks = seg_masks[keep], seg[keep], sum_masks[keep]
        cate_scores, cate_labels = cate_scores[keep], cate_labels[keep]
        seg_scores = (seg * seg_masks.float()).sum(1) / sum_masks
        cate_scores = cate_scores * seg_scores
        sort_inds = torch.argsort(cate_scores, descending=True)
        if len(sort_inds) > net.max_before_nms:
            sort_inds = sort_inds[:net.max_before_nms]
        seg_masks, seg, sum_masks = seg_masks[sort_inds], seg[sort_inds], sum_masks[sort_inds]
        cate_scores, cate_labels = cate_scores[sort_inds], cate_labels[sort_inds]
        if net.nms_type == "matrix":
            cate_scores = matrix_nms(cate_labels, seg_masks.float(), sum_masks, cate_scores, sigma=net.nms_sigma,
                                     kernel=net.nms_kernel)
            keep = cate_scores >= net.update_threshold
        else:
            raise NotImplementedError("nms_type %r (only 'matrix', the presets' default, is implemented)" % net.nms_type)
        if keep.sum() == 0:
            continue
        seg, cate_scores, cate_labels = seg[keep], cate_scores[keep], cate_labels[keep]
        sort_inds = torch.argsort(cate_scores, descending=True)
        if len(sort_inds) > net.max_per_img:
            sort_inds = sort_inds[:net.max_per_img]
        seg, cate_scores, cate_labels = seg[sort_inds], cate_scores[sort_inds], cate_labels[sort_inds]
        masks = F.interpolate(seg.reshape(1, -1, mh, mw), size=(H, W), mode="bilinear", align_corners=False).squeeze(0)
        masks = masks > net.mask_threshold
        # bbox from mask without the per-instance python loop of planerecnet.py:282-286
        ys = masks.any(2)
        xs = masks.any(1)
        ar_y = torch.arange(H, device=x.device)
        ar_x = torch.arange(W, device=x.device)
        y0 = torch.where(ys, ar_y, H).min(1).values
        y1 = torch.where(ys, ar_y, -1).max(1).values
        x0 = torch.where(xs, ar_x, W).min(1).values
        x1 = torch.where(xs, ar_x, -1).max(1).values
        boxes = torch.stack([x0, y0, x1, y1], 1).float()
        result.update(pred_scores=cate_scores, pred_classes=cate_labels, pred_masks=masks, pred_boxes=boxes)
    return results
