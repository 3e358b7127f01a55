"""CPU oracle for PlaneRecNetLoss — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (plain torch / numpy / cv2 on the CPU) of the reference's joint loss: target assignment
(models/functions/losses.py:200-275), dice instance loss (:81-118, 355-368), sigmoid focal category loss (:121-138,
331-352), RMSE-log depth loss (:141-147, 371-392), plane surface-normal loss (:150-165, models/functions/vnl.py:6-165) and
the depth-gradient ("lava") instance loss (:168-197, 277-329), with the presets' weights (data/config.py:459-468,
511-514).  Only tests/ may import this module.

Parity status: PINNED against outputs of the unmodified reference (`PlaneRecNetLoss` imported from /root/reference in the
build container with the runtime patches of SURVEY.md §8c) by tests/golden/make_loss_golden.py -> tests/golden/loss_*.pt.

Quirks of the reference that are kept on purpose: the result of `gt_depths.clamp(max=...)` is discarded
(:145); the lava valid mask stays None because the dataset name is 'ScanNetDataset', not 'ScanNet' (:172); the plane loss
hard-codes a 480x640 image and takes the principal point from the image size, not from the intrinsics (vnl.py:11-13).
"""
import cv2
import numpy as np
import torch
import torch.nn.functional as F

CFG = dict(
    num_classes=2, grids=(40, 36, 24, 16), strides=(8, 8, 16, 32),
    scale_ranges=((1, 128), (64, 256), (128, 512), (256, 2048)), sigma=0.2,             # data/config.py:367-375, 507-508
    dice_weight=3.0, focal_weight=1.0, depth_weight=5.0, lava_weight=1.0, pln_weight=1.0,   # :459-461, 513-514
    focal_gamma=2.0, focal_alpha=0.25,                                                       # :467-468
    min_depth=1 / 1000, max_depth=40, depth_resolution=1 / 1000,                             # scannet_dataset :132-134
    vnl_size=(480, 640), vnl_sample_ratio=0.3, vnl_delta_z=1e-4,                              # losses.py:50, vnl.py:7-9
)


# ------------------------------------------------------------------------------------------ elementary losses
def dice_loss(pred, target):
    """losses.py:355-368: 1 - 2<p,t> / (<p,p> + <t,t> + 0.002) per row."""
    p = pred.reshape(pred.shape[0], -1)
    t = target.reshape(target.shape[0], -1).float()
    return 1 - 2 * (p * t).sum(1) / ((p * p).sum(1) + 0.001 + (t * t).sum(1) + 0.001)


def sigmoid_focal_sum(logits, onehot, alpha, gamma):
    """losses.py:331-352 with reduction 'sum'."""
    p = torch.sigmoid(logits)
    ce = F.binary_cross_entropy_with_logits(logits, onehot, reduction="none")
    p_t = p * onehot + (1 - p) * (1 - onehot)
    loss = ce * (1 - p_t) ** gamma
    if alpha >= 0:
        loss = (alpha * onehot + (1 - alpha) * (1 - onehot)) * loss
    return loss.sum()


def rmse_log_mean(pred, target, valid, clamp_val=1e-9):
    """losses.py:371-392 with reduction 'mean': per image sqrt(sum((|log p - log t| * valid)^2) / sum(valid))."""
    n = pred.shape[0]
    d = (torch.log(pred.reshape(n, -1).clamp(min=clamp_val)) - torch.log(target.reshape(n, -1).clamp(min=clamp_val))).abs()
    d = d * valid.reshape(n, -1)
    return torch.sqrt((d ** 2).sum(1) / valid.reshape(n, -1).sum(1)).mean()


@torch.no_grad()
def gradient_map(depth, valid=None):
    """losses.py:288-329: squared Sobel/8 gradient magnitude of a reflect-padded depth map."""
    sx = torch.tensor([[1.0, 0.0, -1.0], [2.0, 0.0, -2.0], [1.0, 0.0, -1.0]]).view(1, 1, 3, 3) / 8.0
    sy = torch.tensor([[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]]).view(1, 1, 3, 3) / 8.0
    d = F.pad(depth, (1, 1, 1, 1), mode="reflect")
    g = F.conv2d(d, sx) ** 2 + F.conv2d(d, sy) ** 2
    return g if valid is None else g * valid


def lava_loss(seg, gmap):
    """losses.py:277-286: mask probabilities (bilinear to the depth resolution) weighted by the depth-gradient map."""
    up = F.interpolate(seg.unsqueeze(0), size=gmap.shape[1:], mode="bilinear").squeeze(0)
    return (up * gmap).sum() / (gmap.sum() * seg.shape[0])


# ------------------------------------------------------------------------------------------ target assignment
def center_of_mass(masks):
    """funcs.py:213-224."""
    _, h, w = masks.shape
    ys = torch.arange(h, dtype=torch.float32)
    xs = torch.arange(w, dtype=torch.float32)
    m00 = masks.sum(-1).sum(-1).clamp(min=1e-6)
    return (masks * xs).sum(-1).sum(-1) / m00, (masks * ys[:, None]).sum(-1).sum(-1) / m00


def quarter_masks(masks_u8):
    """losses.py:243-247: [n,H,W] uint8 -> cv2 bilinear rescale by 1/4 of the HWC array (funcs.py:74-193) -> [n,h,w]."""
    arr = masks_u8.permute(1, 2, 0).to(torch.uint8).numpy()
    h, w = arr.shape[:2]
    size = (int(w * 0.25 + 0.5), int(h * 0.25 + 0.5))
    out = cv2.resize(arr, size, interpolation=cv2.INTER_LINEAR)
    if out.ndim == 2:
        out = out[..., None]
    return torch.from_numpy(out).to(torch.uint8).permute(2, 0, 1)


@torch.no_grad()
def assign_targets(gt, feat_hw, cfg=CFG):
    """losses.py:200-275 for one image.  Returns per level: (ins_label uint8 [n,h,w], cate_label int64 [S,S],
    ins_ind bool [S*S], grid_order list[int])."""
    boxes, labels, masks = gt["boxes"], gt["classes"], gt["masks"]
    areas = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    fh, fw = feat_hw
    up_h, up_w = fh * 4, fw * 4
    out = []
    for (lo, hi), S in zip(cfg["scale_ranges"], cfg["grids"]):
        hit = ((areas >= lo) & (areas <= hi)).nonzero().flatten()
        cate = torch.full((S, S), cfg["num_classes"], dtype=torch.int64)
        ind = torch.zeros(S * S, dtype=torch.bool)
        ins, order = [], []
        if len(hit):
            bx, lb, mk = boxes[hit], labels[hit], masks[hit]
            half_w = 0.5 * (bx[:, 2] - bx[:, 0]) * cfg["sigma"]
            half_h = 0.5 * (bx[:, 3] - bx[:, 1]) * cfg["sigma"]
            cw, ch = center_of_mass(mk)
            nonempty = mk.sum(-1).sum(-1) > 0
            small = quarter_masks(mk)
            for k in range(len(hit)):
                if not nonempty[k]:
                    continue
                cell = lambda v, extent: int((v / extent) // (1.0 / S))     # noqa: E731
                cx, cy = cell(cw[k], up_w), cell(ch[k], up_h)
                top = max(max(0, cell(ch[k] - half_h[k], up_h)), cy - 1)
                down = min(min(S - 1, cell(ch[k] + half_h[k], up_h)), cy + 1)
                left = max(cx - 1, max(0, cell(cw[k] - half_w[k], up_w)))
                right = min(min(S - 1, cell(cw[k] + half_w[k], up_w)), cx + 1)
                cate[top:down + 1, left:right + 1] = lb[k]
                for i in range(top, down + 1):
                    for j in range(left, right + 1):
                        canvas = torch.zeros(fh, fw, dtype=torch.uint8)
                        canvas[:small.shape[1], :small.shape[2]] = small[k]
                        ins.append(canvas)
                        ind[i * S + j] = True
                        order.append(i * S + j)
        ins_t = torch.stack(ins, 0) if ins else torch.zeros(0, fh, fw, dtype=torch.uint8)
        out.append((ins_t, cate, ind, order))
    return out


# ------------------------------------------------------------------------------------------ plane surface-normal loss
class PlaneNormalLoss:
    """vnl.py:6-165 (VNL_Loss): triplets sampled with numpy's global RNG inside every plane mask (and in the non-planar
    rest against the ground-truth point cloud); 1 - |cos| between triplet normals and the plane normal, keeping the
    worst 75 %."""

    def __init__(self, size=CFG["vnl_size"], sample_ratio=CFG["vnl_sample_ratio"], delta_z=CFG["vnl_delta_z"]):
        h, w = size
        self.size, self.ratio, self.delta_z = size, sample_ratio, delta_z
        self.u = (torch.arange(w, dtype=torch.float32)[None, None, :] - float(w // 2)).expand(1, h, w)
        self.v = (torch.arange(h, dtype=torch.float32)[None, :, None] - float(h // 2)).expand(1, h, w)

    def points(self, depth, K):
        x = self.u * depth.abs() / K[0, 0]
        y = self.v * depth.abs() / K[1, 1]
        return torch.cat([x, y, depth], 0).permute(1, 2, 0)

    def sample(self, num):
        assert num <= self.size[0] * self.size[1]
        idx = []
        for _ in range(3):           # vnl.py:48-53: choice then shuffle, three times, on the global numpy RNG
            p = np.random.choice(num, int(num * self.ratio), replace=True)
            np.random.shuffle(p)
            idx.append(p)
        return idx

    @staticmethod
    def groups(idx, pts):
        return torch.stack([pts[idx[0]], pts[idx[1]], pts[idx[2]]], 2)          # [n, xyz, p123]

    def usable(self, idx, pts, delta_cos=0.985, delta_diff=0.005):
        g = self.groups(idx, pts)
        diff = torch.stack([g[:, :, 1] - g[:, :, 0], g[:, :, 2] - g[:, :, 0], g[:, :, 2] - g[:, :, 1]], 2)
        q = diff.permute(0, 2, 1)
        qn = q.norm(2, dim=2)
        cosm = torch.bmm(q, diff) / (torch.bmm(qn.unsqueeze(2), qn.unsqueeze(1)) + 1e-8)
        cosm = cosm.reshape(diff.shape[0], -1)
        colinear = ((cosm > delta_cos) | (cosm < -delta_cos)).sum(1) > 3
        in_front = (g[:, 2, :] > self.delta_z).sum(1) == 3
        near = (((diff[:, 0, :].abs() < delta_diff).sum(1) > 0) & ((diff[:, 1, :].abs() < delta_diff).sum(1) > 0) &
                ((diff[:, 2, :].abs() < delta_diff).sum(1) > 0))
        return in_front & ~(near | colinear), g

    @staticmethod
    def normals(g, keep):
        g = g[keep]
        n = torch.cross(g[:, :, 1] - g[:, :, 0], g[:, :, 2] - g[:, :, 0], dim=1)
        nrm = torch.norm(n, 2, dim=1, keepdim=True)
        return n / (nrm + (nrm == 0.0).float() * 0.01)

    @staticmethod
    def worst_three_quarters(loss):
        loss, _ = torch.sort(loss, dim=0, descending=False)
        loss = loss[int(loss.shape[0] * 0.25):]
        return torch.nansum(loss) / loss.shape[0]

    def __call__(self, pred_depth, gt_masks, gt_normals, gt_depth, K):
        pred_pts = self.points(pred_depth, K)
        n_planes = gt_normals.shape[0]
        total = 0
        rest = torch.logical_not(gt_masks.sum(0).bool())
        for i in range(n_planes):
            seg = pred_pts[gt_masks[i], :]
            idx = self.sample(seg.shape[0])
            keep, g = self.usable(idx, seg)
            cos = F.cosine_similarity(self.normals(g, keep), gt_normals[i].unsqueeze(0), dim=1).abs()
            total = total + self.worst_three_quarters(1 - cos)
        if rest.sum() > 0:
            gt_pts = self.points(gt_depth, K)
            idx = self.sample(int(rest.sum()))
            keep, g_gt = self.usable(idx, gt_pts[rest, :], delta_diff=0.1)
            if keep.sum() == 0:
                return total / n_planes
            g_pred = self.groups(idx, pred_pts[rest, :])
            g_pred[g_pred[:, 2, :] == 0] = 0.0001
            cos = F.cosine_similarity(self.normals(g_pred, keep), self.normals(g_gt, keep), dim=1).abs()
            total = total + self.worst_three_quarters(1 - cos)
            return total / (n_planes + 1)
        return total / n_planes


# ------------------------------------------------------------------------------------------ the joint loss
def loss_forward(mask_preds, cate_preds, kernel_preds, depth_preds, gt_instances, gt_depths, cfg=CFG, taps=None):
    """losses.py:53-198.  Returns {'ins','cat','dpt','pln','lav'}."""
    B = len(gt_instances)
    fh, fw = mask_preds.shape[-2:]
    targets = [assign_targets(g, (fh, fw), cfg) for g in gt_instances]            # [image][level]
    n_levels = len(cfg["grids"])
    if taps is not None:
        taps["targets"] = targets

    # instance masks: dynamic 1x1 convolution of the mask features with the kernels of the positive cells
    per_level, per_image = [], [torch.empty(0)] * B
    for lvl in range(n_levels):
        chunks = []
        for b in range(B):
            order = targets[b][lvl][3]
            k = kernel_preds[lvl][b].reshape(kernel_preds[lvl][b].shape[0], -1)[:, order]        # [128, n]
            if k.shape[-1] == 0:
                continue
            pred = F.conv2d(mask_preds[b:b + 1], k.permute(1, 0).reshape(k.shape[1], -1, 1, 1)).reshape(-1, fh, fw)
            chunks.append(pred)
            per_image[b] = torch.cat((per_image[b], pred), 0)
        per_level.append(torch.cat(chunks, 0) if chunks else None)
    ins_labels = [torch.cat([targets[b][lvl][0] for b in range(B)], 0) for lvl in range(n_levels)]
    num_ins = torch.cat([torch.cat([targets[b][lvl][2].flatten() for b in range(B)]) for lvl in range(n_levels)]).sum()
    dice = [dice_loss(torch.sigmoid(p), t) for p, t in zip(per_level, ins_labels) if p is not None]
    out = {"ins": torch.cat(dice).mean() * cfg["dice_weight"]}

    # category: focal loss over all grid cells
    labels = torch.cat([torch.cat([targets[b][lvl][1].flatten() for b in range(B)]) for lvl in range(n_levels)])
    logits = torch.cat([c.permute(0, 2, 3, 1).reshape(-1, cfg["num_classes"]) for c in cate_preds])
    pos = torch.nonzero(labels != cfg["num_classes"]).squeeze(1)
    onehot = torch.zeros_like(logits)
    onehot[pos, labels[pos]] = 1
    out["cat"] = cfg["focal_weight"] * sigmoid_focal_sum(logits, onehot, cfg["focal_alpha"], cfg["focal_gamma"]) / (num_ins + 1)

    # depth
    depth_up = F.interpolate(depth_preds, scale_factor=2, mode="bilinear", align_corners=False)
    valid = gt_depths > cfg["min_depth"]
    out["dpt"] = cfg["depth_weight"] * rmse_log_mean(depth_up, gt_depths, valid)

    # plane surface normals
    vnl = PlaneNormalLoss()
    pln = [vnl(depth_up[b], gt_instances[b]["masks"].bool(), gt_instances[b]["plane_paras"][:, :3], gt_depths[b],
               gt_instances[b]["k_matrix"]) for b in range(B)]
    out["pln"] = torch.stack(pln).mean() * cfg["pln_weight"]

    # depth-gradient constraint on the instance masks (valid mask None: see the module docstring)
    g = gradient_map(gt_depths, None) / torch.pow(gt_depths.clamp(min=cfg["depth_resolution"]), 2)
    g = g.clamp(max=1e-2)
    g[g < 1e-4] = 0
    lav = [lava_loss(per_image[b].sigmoid(), g[b]) for b in range(B) if per_image[b].shape[0] > 0 and g[b].sum() > 0]
    out["lav"] = torch.stack(lav).mean() * cfg["lava_weight"] if lav else torch.tensor([0.0])
    return out
