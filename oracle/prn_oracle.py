"""CPU oracle for PlaneRecNet's dense hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (plain torch CPU ops driven by a `state_dict`, no nn.Module graph, no CUDA)
of the reference algorithm, every function citing the reference file:line it follows (paths under
/root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (planerecnet_b200/) never does.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the unmodified reference itself, imported from /root/reference in the
build container by tests/golden/make_golden.py (committed, with its outputs under tests/golden/), and
`deform_conv2d` — third-party arithmetic (torchvision 0.11.1 pinned by the reference's
environment.yml:149; 0.26.0 in this image) — is additionally checked against torchvision's CPU
operator in tests/test_oracle_cpu.py.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# presets (data/config.py:208-250, 286-403, 485-528) — only the fields the dense path reads
# ----------------------------------------------------------------------------------------------
PRESETS = {
    "PlaneRecNet_101_config": dict(layers=[3, 4, 23, 3], dcn_layers=[0, 4, 23, 3], dcn_interval=3),
    "PlaneRecNet_50_config": dict(layers=[3, 4, 6, 3], dcn_layers=[0, 4, 6, 3], dcn_interval=1),
}
NUM_GRIDS = [40, 36, 24, 16]          # data/config.py:507
INSTANCE_STRIDES = [8, 8, 16, 32]     # data/config.py:508
NUM_CLASSES = 2                       # data/config.py:413
NUM_KERNELS = 128                     # data/config.py:349
INFER = dict(nms_pre=500, score_thr=0.1, mask_thr=0.1, update_thr=0.15, top_k=100, sigma=2.0,
             kernel="gaussian")       # data/config.py:376-390


def dcn_block_flags(layers, dcn_layers, dcn_interval):
    """Which bottlenecks use a deformable 3x3: models/backbone.py:170,184."""
    flags = []
    for blocks, dl in zip(layers, dcn_layers):
        f = [dl >= blocks]
        for i in range(1, blocks):
            f.append(((i + dl) >= blocks) and (i % dcn_interval == 0))
        flags.append(f)
    return flags


# ----------------------------------------------------------------------------------------------
# torchvision.ops.deform_conv2d restated (SURVEY.md §8 a4'; call site models/dcn.py:59-66)
# ----------------------------------------------------------------------------------------------
def _bilinear_zero(x, py, px):
    """x [B,C,H,W]; py/px [B,Ho,Wo] absolute sample coordinates.  Value is 0 when the point lies outside
    (-1,H)x(-1,W); corners outside the image contribute 0 (torchvision deform_conv2d bilinear_interpolate)."""
    B, Cc, H, W = x.shape
    inside = (py > -1) & (py < H) & (px > -1) & (px < W)
    y0 = torch.floor(py)
    x0 = torch.floor(px)
    ly, lx = py - y0, px - x0
    hy, hx = 1 - ly, 1 - lx
    y0 = y0.long()
    x0 = x0.long()
    flat = x.reshape(B, Cc, H * W)
    out = torch.zeros(B, Cc, py.shape[1] * py.shape[2], dtype=x.dtype)
    for dy, dx, wgt in ((0, 0, hy * hx), (0, 1, hy * lx), (1, 0, ly * hx), (1, 1, ly * lx)):
        yy, xx = y0 + dy, x0 + dx
        ok = inside & (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).reshape(B, 1, -1).expand(B, Cc, -1)
        w_eff = (wgt * ok.to(x.dtype)).reshape(B, 1, -1)
        out = out + torch.gather(flat, 2, idx) * w_eff
    return out.reshape(B, Cc, py.shape[1], py.shape[2])


def deform_conv2d(x, offset, weight, bias, stride, padding, mask):
    """out[n,o,ho,wo] = b[o] + sum_{c,i,j} W[o,c,i,j] * m[n,i*kw+j,ho,wo] * bil(x[n,c], y, x^),
    y = ho*s - pad + i + off[n,2(i*kw+j)], x^ = wo*s - pad + j + off[n,2(i*kw+j)+1]."""
    B, Cc, H, W = x.shape
    Co, _, kh, kw = weight.shape
    Ho, Wo = offset.shape[-2:]
    by = (torch.arange(Ho, dtype=x.dtype) * stride - padding).view(1, Ho, 1)
    bx = (torch.arange(Wo, dtype=x.dtype) * stride - padding).view(1, 1, Wo)
    out = torch.zeros(B, Co, Ho, Wo, dtype=x.dtype)
    for k in range(kh * kw):
        i, j = divmod(k, kw)
        py = by + i + offset[:, 2 * k]
        px = bx + j + offset[:, 2 * k + 1]
        val = _bilinear_zero(x, py, px) * mask[:, k:k + 1]
        out = out + torch.einsum("bchw,oc->bohw", val, weight[:, :, i, j])
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


# ----------------------------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------------------------
class Oracle:
    """Functional PlaneRecNet forward over a reference-keyed state_dict (SURVEY.md §8b key families)."""

    def __init__(self, state_dict, preset="PlaneRecNet_50_config", dtype=torch.float32, bn_train=False):
        self.sd = {k: (v.detach().to("cpu").to(dtype) if v.is_floating_point() else v.detach().cpu())
                   for k, v in state_dict.items()}
        self.cfg = PRESETS[preset]
        self.flags = dcn_block_flags(**self.cfg)
        self.dtype = dtype
        self.bn_train = bn_train
        self.taps = {}

    # ---- primitives
    def _bn(self, x, prefix, eps):
        sd = self.sd
        if self.bn_train:  # batch statistics, biased variance (nn.BatchNorm2d training mode)
            return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.0, eps)
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                            sd[prefix + ".bias"], False, 0.0, eps)

    def _conv(self, x, prefix, stride=1, padding=0):
        return F.conv2d(x, self.sd[prefix + ".weight"], self.sd.get(prefix + ".bias"), stride, padding)

    def _gn(self, x, prefix):
        return F.group_norm(x, 32, self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], 1e-5)

    # ---- models/dcn.py:52-67
    def _dcn(self, x, prefix, stride):
        h, w = x.shape[2:]
        max_offset = max(h, w) / 4.0
        off = self._conv(x, prefix + ".offset_conv", stride, 1).clamp(-max_offset, max_offset)
        mod = 2.0 * torch.sigmoid(self._conv(x, prefix + ".modulator_conv", stride, 1))
        return deform_conv2d(x, off, self.sd[prefix + ".regular_conv.weight"],
                             self.sd.get(prefix + ".regular_conv.bias"), stride, 1, mod)

    # ---- models/backbone.py:53-73
    def _bottleneck(self, x, prefix, stride, use_dcn, has_down):
        out = F.relu(self._bn(self._conv(x, prefix + ".conv1"), prefix + ".bn1", 1e-5))
        if use_dcn:
            out = self._dcn(out, prefix + ".conv2", stride)
        else:
            out = self._conv(out, prefix + ".conv2", stride, 1)
        out = F.relu(self._bn(out, prefix + ".bn2", 1e-5))
        out = self._bn(self._conv(out, prefix + ".conv3"), prefix + ".bn3", 1e-5)
        res = x
        if has_down:
            res = self._bn(self._conv(x, prefix + ".downsample.0", stride), prefix + ".downsample.1", 1e-5)
        return F.relu(out + res)

    # ---- models/backbone.py:197-209
    def backbone(self, x):
        x = F.relu(self._bn(self._conv(x, "backbone.conv1", 2, 3), "backbone.bn1", 1e-5))
        x = F.max_pool2d(x, 3, 2, 1)
        outs = []
        for s, blocks in enumerate(self.cfg["layers"]):
            for b in range(blocks):
                stride = 2 if (b == 0 and s > 0) else 1
                x = self._bottleneck(x, f"backbone.layers.{s}.{b}", stride, self.flags[s][b], b == 0)
            outs.append(x)
        return outs

    # ---- models/fpn.py:45-63 (bottom-up running sum; 'bilinear' to the coarser size)
    def fpn(self, cs):
        lats = []
        x = torch.zeros(1, dtype=self.dtype)
        for i, c in enumerate(cs):
            if i > 0:
                x = F.interpolate(x, size=c.shape[-2:], mode="bilinear", align_corners=False)
            x = self._conv(c, f"fpn.lateral_convs.{i}") + x
            lats.append(x)
        return [F.relu(self._conv(l, f"fpn.fpn_convs.{i}", 1, 1)) for i, l in enumerate(lats)]

    @staticmethod
    def _coords(feat):
        """planerecnet.py:370-376 / 483-489: x then y, linspace(-1,1)."""
        B, _, H, W = feat.shape
        xr = torch.linspace(-1, 1, W, dtype=feat.dtype)
        yr = torch.linspace(-1, 1, H, dtype=feat.dtype)
        y, x = torch.meshgrid(yr, xr, indexing="ij")
        return torch.cat([x.expand(B, 1, H, W), y.expand(B, 1, H, W)], 1)

    # ---- planerecnet.py:113-118
    @staticmethod
    def split_feats(ps):
        return [F.interpolate(ps[0], scale_factor=0.5, mode="bilinear", align_corners=False,
                              recompute_scale_factor=False), ps[1], ps[2], ps[3]]

    # ---- planerecnet.py:355-391
    def inst_head(self, feats):
        cate_pred, kernel_pred = [], []
        for idx, feat in enumerate(feats):
            kf = torch.cat([feat, self._coords(feat)], 1)
            kf = F.interpolate(kf, size=NUM_GRIDS[idx], mode="bilinear", align_corners=False)
            cf = kf[:, :-2]
            for i in (0, 3, 6):
                kf = F.relu(self._gn(self._conv(kf, f"inst_head.kernel_tower.{i}", 1, 1), f"inst_head.kernel_tower.{i + 1}"))
                cf = F.relu(self._gn(self._conv(cf, f"inst_head.cate_tower.{i}", 1, 1), f"inst_head.cate_tower.{i + 1}"))
            kernel_pred.append(self._conv(kf, "inst_head.kernel_pred", 1, 1))
            cate_pred.append(self._conv(cf, "inst_head.cate_pred", 1, 1))
        return cate_pred, kernel_pred

    # ---- planerecnet.py:467-496
    def mask_head(self, ps):
        def tower(x, lvl, j):
            p = f"mask_head.convs_all_levels.{lvl}.conv{j}"
            return F.relu(self._gn(self._conv(x, p + ".0", 1, 1), p + ".1"))

        def up(x):
            return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)

        total = tower(ps[0], 0, 0)
        for lvl in range(1, 4):
            x = ps[lvl]
            if lvl == 3:
                x = torch.cat([x, self._coords(x)], 1)
            for j in range(lvl):
                x = up(tower(x, lvl, j))
            total = total + x
        return F.relu(self._gn(self._conv(total, "mask_head.conv_pred.0"), "mask_head.conv_pred.1"))

    # ---- planerecnet.py:586-607
    def depth_decoder(self, cs, mask_pred, kernel_pred):
        B = cs[0].shape[0]
        # planerecnet.py:589,592: kernels, masks and the sigmoid attention are detached (no gradient into the heads)
        flat = torch.cat([k.permute(0, 2, 3, 1).reshape(B, -1, NUM_KERNELS) for k in kernel_pred], 1).detach()
        attn = torch.cat([F.conv2d(mask_pred[b:b + 1].detach(), flat[b].view(-1, NUM_KERNELS, 1, 1)) for b in range(B)], 0).sigmoid().detach()
        self.taps["ppa_sigmoid"] = attn
        attn = self._conv(attn, "depth_decoder.conv1x1.0")
        attn = F.interpolate(attn, scale_factor=0.25, mode="bilinear", align_corners=False, recompute_scale_factor=False)
        self.taps["ppa_attn"] = attn

        def rconv(x, name, idx_conv, upsample=False, act=True):
            if upsample:
                x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.pad(x, (1, 1, 1, 1), mode="reflect")
            x = self._conv(x, f"depth_decoder.{name}.{idx_conv}")
            if act:
                x = F.relu(self._bn(x, f"depth_decoder.{name}.{idx_conv + 1}", 1e-3))
            return x

        feats = list(reversed(cs))
        x = rconv(rconv(self._conv(feats[0], "depth_decoder.latlayer1"), "conv1", 1), "deconv1", 2, upsample=True)
        x = rconv(torch.cat([x, x * attn], 1), "refine_conv", 1)
        for k in (2, 3, 4):
            skip = rconv(self._conv(feats[k - 1], f"depth_decoder.latlayer{k}"), f"conv{k}", 1)
            x = rconv(torch.cat([skip, x], 1), f"deconv{k}", 2, upsample=True)
        x = rconv(x, "depth_pred", 1, act=False)
        return F.softplus(x)

    # ---- planerecnet.py:73-103 (training-branch return value = the dense outputs)
    def forward_dense(self, x):
        x = x.to(self.dtype)
        cs = self.backbone(x)
        ps = self.fpn(cs)
        cate, kern = self.inst_head(self.split_feats(ps))
        mask = self.mask_head(ps)
        depth = self.depth_decoder(cs, mask, kern)
        self.taps.update(cs=cs, ps=ps)
        return mask, cate, kern, depth

    # ---- planerecnet.py:106-111
    def forward_eval(self, x):
        mask, cate, kern, depth = self.forward_dense(x)
        cate = [point_nms(c.sigmoid()).permute(0, 2, 3, 1) for c in cate]
        return inference(mask, cate, kern, depth, x.shape[-2:])


# ----------------------------------------------------------------------------------------------
# inference bookkeeping (models/functions/nms.py, planerecnet.py:155-289)
# ----------------------------------------------------------------------------------------------
def point_nms(heat, kernel=2):
    """nms.py:8-12."""
    hmax = F.max_pool2d(heat, (kernel, kernel), stride=1, padding=1)
    keep = (hmax[:, :, :-1, :-1] == heat).to(heat.dtype)
    return heat * keep


def matrix_nms(cate_labels, seg_masks, sum_masks, cate_scores, sigma=2.0, kernel="gaussian"):
    """nms.py:15-50."""
    n = len(cate_labels)
    if n == 0:
        return []
    m = seg_masks.reshape(n, -1).to(cate_scores.dtype)
    inter = m @ m.t()
    sx = sum_masks.expand(n, n)
    iou = (inter / (sx + sx.t() - inter)).triu(diagonal=1)
    lx = cate_labels.expand(n, n)
    label = (lx == lx.t()).to(iou.dtype).triu(diagonal=1)
    comp, _ = (iou * label).max(0)
    comp = comp.expand(n, n).t()
    decay = iou * label
    if kernel == "linear":
        coef, _ = ((1 - decay) / (1 - comp)).min(0)
    else:
        coef, _ = (torch.exp(-sigma * decay ** 2) / torch.exp(-sigma * comp ** 2)).min(0)
    return cate_scores * coef


def mask_nms(cate_labels, seg_masks, sum_masks, cate_scores, nms_thr=0.5):
    """nms.py:53-80: greedy O(n^2) suppression over the score-sorted candidates; returns the bool keep vector."""
    n = len(cate_scores)
    if n == 0:
        return []
    keep = torch.ones(n, dtype=torch.bool)
    m = seg_masks.reshape(n, -1).float()
    for i in range(n - 1):
        if not keep[i]:
            continue
        for j in range(i + 1, n):
            if not keep[j] or cate_labels[i] != cate_labels[j]:
                continue
            inter = (m[i] * m[j]).sum()
            union = sum_masks[i] + sum_masks[j] - inter
            if union > 0:
                if inter / union > nms_thr:
                    keep[j] = False
            else:
                keep[j] = False
    return keep


def inference_single(seg_preds, cate_preds, kernel_preds, depth_pred, ori_size, p=INFER):
    """planerecnet.py:182-289, one image.  seg_preds [1,128,h,w]; cate_preds [3728,2]; kernel_preds [3728,128]."""
    result = {"pred_masks": None, "pred_boxes": None, "pred_classes": None, "pred_scores": None, "pred_depth": None}
    result["pred_depth"] = F.interpolate(depth_pred, size=ori_size, mode="bilinear", align_corners=False)
    inds = cate_preds > p["score_thr"]
    cate_scores = cate_preds[inds]
    if len(cate_scores) == 0:
        return result
    inds = inds.nonzero(as_tuple=False)
    cate_labels = inds[:, 1]
    kernel_preds = kernel_preds[inds[:, 0]]
    size_trans = torch.tensor(NUM_GRIDS).pow(2).cumsum(0)
    strides = torch.ones(int(size_trans[-1]), dtype=kernel_preds.dtype)
    strides[:size_trans[0]] *= INSTANCE_STRIDES[0]
    for i in range(1, len(NUM_GRIDS)):
        strides[size_trans[i - 1]:size_trans[i]] *= INSTANCE_STRIDES[i]
    strides = strides[inds[:, 0]]
    N, I = kernel_preds.shape
    seg = F.conv2d(seg_preds, kernel_preds.view(N, I, 1, 1)).squeeze(0).sigmoid()
    det = bookkeeping(seg, cate_scores, cate_labels, strides, p)
    if det is None:
        return result
    seg, cate_scores, cate_labels, _ = det
    masks = F.interpolate(seg.unsqueeze(0), size=ori_size, mode="bilinear", align_corners=False).squeeze(0) > p["mask_thr"]
    boxes = torch.zeros(masks.size(0), 4)
    for i in range(masks.size(0)):
        ys, xs = torch.where(masks[i])
        boxes[i] = torch.tensor([xs.min(), ys.min(), xs.max(), ys.max()]).float()
    result.update(pred_scores=cate_scores, pred_classes=cate_labels, pred_masks=masks, pred_boxes=boxes)
    return result


def bookkeeping(seg, cate_scores, cate_labels, strides, p=INFER):
    """planerecnet.py:216-269 given the candidates' sigmoid masks seg [N,h,w]: area filter, maskness rescoring,
    sort/top-500, matrix-NMS, update threshold, sort/top-100.  Returns (seg, scores, labels, candidate index) of the
    surviving detections in output order, or None."""
    cand = torch.arange(seg.shape[0])
    seg_masks = seg > p["mask_thr"]
    sum_masks = seg_masks.sum((1, 2)).to(seg.dtype)
    keep = sum_masks > strides
    if keep.sum() == 0:
        return None
    seg_masks, seg, sum_masks = seg_masks[keep], seg[keep], sum_masks[keep]
    cate_scores, cate_labels, cand = cate_scores[keep], cate_labels[keep], cand[keep]
    seg_scores = (seg * seg_masks.to(seg.dtype)).sum((1, 2)) / sum_masks
    cate_scores = cate_scores * seg_scores
    order = torch.argsort(cate_scores, descending=True)[:p["nms_pre"]]
    seg_masks, seg, sum_masks = seg_masks[order], seg[order], sum_masks[order]
    cate_scores, cate_labels, cand = cate_scores[order], cate_labels[order], cand[order]
    if p.get("nms_type", "matrix") == "mask":      # planerecnet.py:249-252
        keep = mask_nms(cate_labels, seg_masks, sum_masks, cate_scores, nms_thr=p["mask_thr"])
    else:
        cate_scores = matrix_nms(cate_labels, seg_masks, sum_masks, cate_scores, sigma=p["sigma"], kernel=p["kernel"])
        keep = cate_scores >= p["update_thr"]
    if keep.sum() == 0:
        return None
    seg, cate_scores, cate_labels, cand = seg[keep], cate_scores[keep], cate_labels[keep], cand[keep]
    order = torch.argsort(cate_scores, descending=True)[:p["top_k"]]
    return seg[order], cate_scores[order], cate_labels[order], cand[order]


def inference(mask_pred, cate_preds, kernel_preds, depth_pred, ori_size):
    """planerecnet.py:155-180.  cate_preds: list of [B,S,S,2] after point-NMS; kernel_preds: list of [B,128,S,S]."""
    results = []
    for b in range(mask_pred.shape[0]):
        cate = torch.cat([c[b].reshape(-1, NUM_CLASSES) for c in cate_preds], 0)
        kern = torch.cat([k[b].permute(1, 2, 0).reshape(-1, NUM_KERNELS) for k in kernel_preds], 0)
        results.append(inference_single(mask_pred[b:b + 1], cate, kern, depth_pred[b:b + 1], tuple(ori_size)))
    return results


def bias_init_with_prob(prior_prob):
    """models/functions/funcs.py:329-332."""
    return float(-math.log((1 - prior_prob) / prior_prob))
