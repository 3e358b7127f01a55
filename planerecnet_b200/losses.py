"""Dense parts of PlaneRecNetLoss on libprn_b200 — first two terms, each one autograd node:

  focal_cate_loss(cate32, labels, num_ins)   losses.py:121-138  sigmoid focal over all grid cells / (num_ins + 1)
  depth_rmselog_loss(depth_pred, gt_depths)  losses.py:141-147  depth_weight * RMSE-log on the x2-upsampled prediction

  ins_lava_losses(mask_pred, kernel_preds, targets, gt_depths)   losses.py:81-118, 168-197  dice + depth-gradient terms

  _PlaneNormalBatched                                            losses.py:150-165, vnl.py:6-165  plane surface-normal term of a
                                                                 whole batch in a few dozen tensor ops (numpy-RNG-exact triplets)
  PlaneRecNetLoss                                                losses.py:12-198  the joint loss behind the reference's signature

Targets come from planerecnet_b200.targets (device-resident assignment, one host round trip per batch)."""
import ctypes as C
import math

import torch

from . import _lib as L


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _FocalSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, alpha, gamma, nc):
        assert logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
        n, ld = logits.shape[0], logits.stride(0)
        labels = labels.to(torch.int64).contiguous()
        loss = torch.zeros(1, device=logits.device)
        dl = torch.empty(n, nc, device=logits.device)
        L.check(L.lib().prn_focal_loss(_p(logits), _p(labels), C.c_float(alpha), C.c_float(gamma), _p(loss), _p(dl), C.c_int64(n),
                                       ld, nc, L.current_stream()), "prn_focal_loss")
        ctx.save_for_backward(dl)
        ctx.shape = logits.shape
        ctx.nc = nc
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        out = torch.zeros(ctx.shape, device=dl.device)
        out[:, :ctx.nc] = dl * g
        return out, None, None, None, None


def focal_cate_loss(logits, labels, num_ins, num_classes=2, alpha=0.25, gamma=2.0, weight=1.0):
    """logits fp32 [n, >= num_classes] (row pitch may exceed num_classes: the instance head's cate32 buffer), labels [n]
    with num_classes = background.  == conf_loss_weight * SigmoidFocalLoss(sum)(...) / (num_ins + 1)."""
    return weight * _FocalSum.apply(logits, labels, alpha, gamma, num_classes) / (num_ins + 1)


class _DepthRMSELog(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, gt, min_depth, clamp_val, weight):
        assert depth.is_cuda and depth.dtype == torch.float32 and depth.dim() == 4 and depth.shape[1] == 1
        B, _, h, w = depth.shape
        assert tuple(gt.shape) == (B, 1, 2 * h, 2 * w)
        depth, gt = depth.contiguous(), gt.float().contiguous()
        sums = torch.zeros(B, 2, device=depth.device)
        loss = torch.empty(1, device=depth.device)
        coef = torch.empty(B, device=depth.device)
        L.check(L.lib().prn_depth_rmselog_fwd(_p(depth), _p(gt), _p(sums), _p(loss), _p(coef), B, h, w, C.c_float(min_depth),
                                              C.c_float(clamp_val), C.c_float(weight), L.current_stream()), "prn_depth_rmselog_fwd")
        ctx.save_for_backward(depth, gt, coef)
        ctx.args = (min_depth, clamp_val)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        depth, gt, coef = ctx.saved_tensors
        B, _, h, w = depth.shape
        dd = torch.zeros_like(depth)
        L.check(L.lib().prn_depth_rmselog_bwd(_p(depth), _p(gt), _p(coef), _p(dd), B, h, w, C.c_float(ctx.args[0]),
                                              C.c_float(ctx.args[1]), L.current_stream()), "prn_depth_rmselog_bwd")
        return dd * g, None, None, None, None


def depth_rmselog_loss(depth_pred, gt_depths, min_depth=1 / 1000, clamp_val=1e-9, weight=5.0):
    """depth_pred fp32 [B,1,h,w] (the model's training output), gt_depths [B,1,2h,2w]."""
    return _DepthRMSELog.apply(depth_pred, gt_depths, min_depth, clamp_val, weight)


# ------------------------------------------------------------------------------------------ dice + lava terms
class CudaBackend:
    """The five device steps of the instance-mask losses on libprn_b200 (tests substitute a torch emulation to check the
    orchestration and the gradient algebra on the CPU)."""

    def __init__(self, dtype=L.PRN_F16):
        # f16 operands: the dice gradient is cancellation-dominated (positive and negative regions of a mask nearly
        # cancel in the sums over pixels), so 10 mantissa bits matter; the gradient rows are pre-scaled into f16's range
        from . import ops
        self.ops, self.dt, self.tdt = ops, dtype, ops.torch_dtype(dtype)

    def to16(self, t):
        return t.to(self.tdt).contiguous()

    def seg_rows(self, wsel16, mask16):
        """sigmoid(K_sel . mask^T): wsel16 [B,n,C], mask16 [B,P,C] -> fp32 [B*n, P] (grouped contraction, per-image operands)."""
        B, n, Cc = wsel16.shape
        P = mask16.shape[1]
        seg = torch.empty(B * n, P, device=wsel16.device)
        self.ops.conv2d(wsel16, mask16.reshape(B * P, Cc), batch=B, h_in=n, w_in=1, ksize=1, act=L.ACT_SIGMOID, out32=seg,
                        ld_out32=P, n_pad=P, w_group_rows=P, dtype=self.dt)
        return seg

    def row_stats(self, seg, target, gw, n):
        stats = torch.empty(seg.shape[0], 4, device=seg.device)
        L.check(L.lib().prn_dice_lava_rows(_p(seg), _p(target), _p(gw), _p(stats), seg.shape[0], seg.shape[1], n, L.current_stream()),
                "prn_dice_lava_rows")
        return stats

    def row_bwd(self, seg, target, gw, coef, n):
        dx = torch.empty(seg.shape, dtype=self.tdt, device=seg.device)
        L.check(L.lib().prn_dice_lava_bwd(_p(seg), _p(target), _p(gw), _p(coef), _p(dx), seg.shape[0], seg.shape[1], n, self.dt,
                                          L.current_stream()), "prn_dice_lava_bwd")
        return dx

    def grouped_nt(self, a16, w16):
        """out[b] = a16[b] @ w16[b]^T in fp32: a16 [B,m,K], w16 [B,N,K] (K multiple of 64, N multiple of 16)."""
        B, m, K = a16.shape
        N = w16.shape[1]
        out = torch.empty(B * m, N, device=a16.device)
        self.ops.conv2d(a16.reshape(B, m, 1, K), w16.reshape(B * N, K), batch=B, h_in=m, w_in=1, ksize=1, out32=out, ld_out32=N,
                        n_pad=N, w_group_rows=N, dtype=self.dt)
        return out.view(B, m, N)

    def focal_sum(self, logits, labels, alpha, gamma, nc):
        return _FocalSum.apply(logits, labels, alpha, gamma, nc)

    def depth_rmselog(self, depth, gt, min_depth, clamp_val, weight):
        return _DepthRMSELog.apply(depth, gt, min_depth, clamp_val, weight)

    def lava_weights(self, gt, h, w, depth_res):
        B, _, H, W = gt.shape
        gw = torch.zeros(B, h * w, device=gt.device)
        gsum = torch.zeros(B, device=gt.device)
        gt32 = gt.float().contiguous()             # named: the pointer must outlive the launch
        L.check(L.lib().prn_lava_weights(_p(gt32), _p(gw), _p(gsum), B, H, W, h, w, C.c_float(depth_res), L.current_stream()),
                "prn_lava_weights")
        return gw, gsum


def _round_up(x, m):
    return (x + m - 1) // m * m


class _InsLava(torch.autograd.Function):
    """(dice instance loss, lava loss) of losses.py:81-118, 168-197 as one node: the positive cells' kernels are gathered
    into per-image row blocks, their masks come from one grouped tensor-core contraction, the per-row sums and the
    gradient rows from two passes, and the two gradient contractions (w.r.t. kernels: over pixels; w.r.t. mask features:
    over instances) from the same grouped contraction.  Per-row scalar algebra (a few hundred numbers) stays in torch."""

    @staticmethod
    def _gather_rows(targets, kernel_preds, wsel, tgt, n_b, idx_views, dev):
        """Per-(image, level) form (targets from the per-image assign_targets): fills wsel / tgt row blocks in a loop."""
        B, n, Cc = wsel.shape
        P = tgt.shape[2]
        n_levels = len(kernel_preds)
        # all positive-cell indices of the batch in ONE upload (a torch.tensor(list, device=cuda) per image and level is a
        # synchronous copy each: 32 per step)
        flat_idx = [i for b in range(B) for l in range(n_levels) for i in targets[b][l][3]]
        idx_all = torch.tensor(flat_idx, dtype=torch.int64, device="cpu").to(dev, non_blocking=True) if flat_idx else None
        pos = 0
        for b in range(B):
            off = 0
            for l in range(n_levels):
                order = targets[b][l][3]
                if order:
                    idx = idx_all[pos:pos + len(order)]
                    idx_views[(b, l)] = idx
                    pos += len(order)
                    wsel[b, off:off + len(order)] = kernel_preds[l][b].reshape(Cc, -1)[:, idx].t()
                    tgt[b, off:off + len(order)] = targets[b][l][0].reshape(len(order), P).to(dev)
                    off += len(order)
        valid_host = torch.zeros(B, n, dtype=torch.bool, device="cpu")
        for b in range(B):
            valid_host[b, :n_b[b]] = True
        return wsel, tgt, valid_host.to(dev, non_blocking=True)

    @staticmethod
    def forward(ctx, be, targets, gw, gsum, w_dice, w_lava, mask_pred, *kernel_preds):
        B, Cc, fh, fw = mask_pred.shape
        P = fh * fw
        dev = mask_pred.device
        n_levels = len(kernel_preds)
        counts = [[len(targets[b][l][3]) for l in range(n_levels)] for b in range(B)]
        n_b = [sum(c) for c in counts]
        dense = getattr(targets, "dense", None)
        if dense is not None and dense["tgt"].shape[2] == P and dense["tgt"].device == dev:
            # batch layout prepared by targets.assign_targets_batch: one gather for the positive cells' kernels of all images and
            # levels (the per-(image, level) loop below costs ~5 launches per pair), targets already in place
            n = dense["n"]
            kern_all = torch.cat([k.reshape(B, Cc, -1) for k in kernel_preds], 2)                   # [B, C, sum S^2]
            gsel = dense["gidx"][:, None, :].expand(B, Cc, n)
            wsel = kern_all.gather(2, gsel).transpose(1, 2) * dense["valid"][:, :, None]            # [B, n, C]; padding rows 0
            tgt, valid = dense["tgt"], dense["valid"]
            idx_views = None
            ctx.gsel = gsel
        else:
            n = _round_up(max(max(n_b), 1), 16)
            wsel = torch.zeros(B, n, Cc, device=dev)
            tgt = torch.zeros(B, n, P, dtype=torch.uint8, device=dev)
            ctx.gsel = None
            idx_views = {}
            wsel, tgt, valid = _InsLava._gather_rows(targets, kernel_preds, wsel, tgt, n_b, idx_views, dev)
        mask16 = be.to16(mask_pred.reshape(B, Cc, P).transpose(1, 2))        # [B, P, C]
        wsel16 = be.to16(wsel)
        seg = be.seg_rows(wsel16, mask16)                                     # [B*n, P] sigmoid probabilities
        stats = be.row_stats(seg, tgt.view(B * n, P), gw, n).view(B, n, 4)
        a, bq, c, lv = stats.unbind(-1)
        n_total = max(sum(n_b), 1)
        den = bq + c + 0.002
        dice = torch.where(valid, 1 - 2 * a / den, torch.zeros_like(a))
        loss_ins = w_dice * dice.sum() / n_total
        nb_t = torch.tensor(n_b, dtype=torch.float32, device="cpu").to(dev, non_blocking=True)
        elig = (nb_t > 0) & (gsum > 0)
        # number of eligible images stays on the device: an int() here would stall the host until the network's forward has
        # finished (every host sync between net(x) and loss.backward() is a bubble on the GPU later)
        n_elig = elig.sum().to(torch.float32).clamp(min=1.0)
        per_img = torch.where(elig, (lv * valid).sum(1) / (gsum * nb_t).clamp(min=1e-30), torch.zeros_like(gsum))
        loss_lav = w_lava * per_img.sum() / n_elig                      # no eligible image: per_img == 0 -> 0
        # d(loss)/d{a, b, lv} per row
        ca = torch.where(valid, -w_dice / n_total * 2 / den, torch.zeros_like(a))
        cb = torch.where(valid, w_dice / n_total * 2 * a / (den * den), torch.zeros_like(a))
        cl_img = torch.where(elig, w_lava / n_elig / (gsum * nb_t).clamp(min=1e-30), torch.zeros_like(gsum))
        cl = cl_img[:, None] * valid
        ctx.be, ctx.targets, ctx.counts, ctx.shapes = be, targets, counts, (B, Cc, fh, fw, n, [k.shape for k in kernel_preds])
        ctx.idx_views = idx_views
        ctx.valid = valid
        ctx.save_for_backward(seg, tgt, gw, mask16, wsel16, ca, cb, cl)
        return loss_ins, loss_lav

    @staticmethod
    def backward(ctx, g_ins, g_lav):
        be = ctx.be
        seg, tgt, gw, mask16, wsel16, ca, cb, cl = ctx.saved_tensors
        B, Cc, fh, fw, n, kshapes = ctx.shapes
        P = fh * fw
        dev = seg.device
        coef = torch.stack([ca * g_ins, cb * g_ins, cl * g_lav], -1).reshape(B * n, 3)
        # power-of-two scale that puts the largest possible |dx| near 2^12 (16-bit storage of the gradient rows)
        # (computed on the device: no host round trip in the middle of the backward)
        bound = ((coef[:, 0].abs() + 2 * coef[:, 1].abs()).max() + coef[:, 2].abs().max() * gw.max()) * 0.25
        usable = (bound > 0) & torch.isfinite(bound)
        # (torch.pow, not torch.exp2: exp2 is a jiterator op — an NVRTC compile on its first call)
        scale = torch.where(usable, torch.pow(2.0, torch.floor(torch.log2(4096.0 / bound.clamp(min=1e-38)))), torch.ones_like(bound))
        coef = (coef * scale).contiguous()
        dx16 = be.row_bwd(seg, tgt.view(B * n, P), gw, coef, n).view(B, n, P)        # scaled gradient of the pre-sigmoid rows
        # w.r.t. the selected kernels: dK[b] = dX[b] (n x P) . mask[b] (P x C)
        dK = be.grouped_nt(dx16, mask16.transpose(1, 2).contiguous()) / scale       # [B, n, C]
        # w.r.t. the mask features: dM[b] = dX[b]^T (P x n) . K_sel[b] (n x C)
        n64 = _round_up(n, 64)
        dxT = torch.zeros(B, P, n64, dtype=dx16.dtype, device=dev)
        dxT[:, :, :n] = dx16.transpose(1, 2)
        kT = torch.zeros(B, Cc, n64, dtype=wsel16.dtype, device=dev)
        kT[:, :, :n] = wsel16.transpose(1, 2)
        dM = be.grouped_nt(dxT, kT) / scale                                           # [B, P, C]
        d_mask = dM.transpose(1, 2).reshape(B, Cc, fh, fw).contiguous()
        d_kern = []
        if ctx.gsel is not None:
            # one scatter-add over the level-concatenated kernel map (duplicated cells accumulate), then per-level views
            total = sum(shp[2] * shp[3] for shp in kshapes)
            dk_rows = (dK.float() * ctx.valid[:, :, None]).transpose(1, 2)                            # [B, C, n]
            d_all = torch.zeros(B, Cc, total, device=dev).scatter_add_(2, ctx.gsel, dk_rows)
            off = 0
            for shp in kshapes:
                d_kern.append(d_all[:, :, off:off + shp[2] * shp[3]].reshape(shp))
                off += shp[2] * shp[3]
            return (None, None, None, None, None, None, d_mask, *d_kern)
        for l, shp in enumerate(kshapes):
            gk = torch.zeros(shp, device=dev)
            for b in range(B):
                order = ctx.targets[b][l][3]
                if order:
                    off = sum(ctx.counts[b][:l])
                    gk[b].view(Cc, -1).index_add_(1, ctx.idx_views[(b, l)], dK[b, off:off + len(order)].t().float())
            d_kern.append(gk)
        return (None, None, None, None, None, None, d_mask, *d_kern)


def ins_lava_losses(mask_pred, kernel_preds, targets, gt_depths, backend=None, dice_weight=3.0, lava_weight=1.0,
                    depth_resolution=1 / 1000):
    """(losses['ins'], losses['lav']) of PlaneRecNetLoss.forward for the model's training outputs mask_pred [B,128,h,w] and
    kernel_preds [[B,128,S,S]] x levels; targets[b][level] from planerecnet_b200.targets.assign_targets."""
    be = backend or CudaBackend()
    fh, fw = mask_pred.shape[-2:]
    gw, gsum = be.lava_weights(gt_depths, fh, fw, depth_resolution)
    return _InsLava.apply(be, targets, gw, gsum, dice_weight, lava_weight, mask_pred, *kernel_preds)


# ------------------------------------------------------------------------------------------ plane surface-normal term
class _PlaneNormal:
    """models/functions/vnl.py:6-165.  Host-side by design: the triplets are drawn with numpy's global
    RNG per plane, so the sampling stays in Python in the reference's call order; the few thousand 3x3 cross products run
    as torch ops on the depth map's device."""

    def __init__(self, size=(480, 640), sample_ratio=0.3, delta_z=1e-4):
        self.size, self.ratio, self.delta_z = size, sample_ratio, delta_z
        self._grid = {}

    def _uv(self, dev):
        if dev not in self._grid:
            h, w = self.size
            u = (torch.arange(w, dtype=torch.float32, device=dev)[None, None, :] - float(w // 2)).expand(1, h, w)
            v = (torch.arange(h, dtype=torch.float32, device=dev)[None, :, None] - float(h // 2)).expand(1, h, w)
            self._grid[dev] = (u, v)
        return self._grid[dev]

    def points(self, depth, K):
        u, v = self._uv(depth.device)
        return torch.cat([u * depth.abs() / K[0, 0], v * depth.abs() / K[1, 1], depth], 0).permute(1, 2, 0)

    def sample(self, num):
        import numpy as np
        assert num <= self.size[0] * self.size[1]
        idx = []
        for _ in range(3):
            p = np.random.choice(num, int(num * self.ratio), replace=True)
            np.random.shuffle(p)
            idx.append(p)
        return idx

    @staticmethod
    def groups(idx, pts):
        return torch.stack([pts[idx[0]], pts[idx[1]], pts[idx[2]]], 2)

    def usable(self, idx, pts, delta_cos=0.985, delta_diff=0.005):
        g = self.groups(idx, pts)
        diff = torch.stack([g[:, :, 1] - g[:, :, 0], g[:, :, 2] - g[:, :, 0], g[:, :, 2] - g[:, :, 1]], 2)
        q = diff.permute(0, 2, 1)
        qn = q.norm(2, dim=2)
        cosm = (torch.bmm(q, diff) / (torch.bmm(qn.unsqueeze(2), qn.unsqueeze(1)) + 1e-8)).reshape(diff.shape[0], -1)
        colinear = ((cosm > delta_cos) | (cosm < -delta_cos)).sum(1) > 3
        in_front = (g[:, 2, :] > self.delta_z).sum(1) == 3
        near = (((diff[:, 0, :].abs() < delta_diff).sum(1) > 0) & ((diff[:, 1, :].abs() < delta_diff).sum(1) > 0) &
                ((diff[:, 2, :].abs() < delta_diff).sum(1) > 0))
        return in_front & ~(near | colinear), g

    @staticmethod
    def normals(g, keep):
        g = g[keep]
        nrm_v = torch.cross(g[:, :, 1] - g[:, :, 0], g[:, :, 2] - g[:, :, 0], dim=1)
        nrm = torch.norm(nrm_v, 2, dim=1, keepdim=True)
        return nrm_v / (nrm + (nrm == 0.0).float() * 0.01)

    @staticmethod
    def tail(loss):
        loss, _ = torch.sort(loss, dim=0, descending=False)
        loss = loss[int(loss.shape[0] * 0.25):]
        return torch.nansum(loss) / loss.shape[0]

    def __call__(self, pred_depth, gt_masks, gt_normals, gt_depth, K):
        import torch.nn.functional as F
        pts = self.points(pred_depth, K)
        n_planes = gt_normals.shape[0]
        total = 0
        rest = torch.logical_not(gt_masks.sum(0).bool())
        for i in range(n_planes):
            seg = pts[gt_masks[i], :]
            idx = self.sample(seg.shape[0])
            keep, g = self.usable(idx, seg)
            cos = F.cosine_similarity(self.normals(g, keep), gt_normals[i].unsqueeze(0), dim=1).abs()
            total = total + self.tail(1 - cos)
        if rest.sum() > 0:
            gt_pts = self.points(gt_depth, K)
            idx = self.sample(int(rest.sum()))
            keep, g_gt = self.usable(idx, gt_pts[rest, :], delta_diff=0.1)
            if keep.sum() == 0:
                return total / n_planes
            g_pred = self.groups(idx, pts[rest, :])
            g_pred[g_pred[:, 2, :] == 0] = 0.0001
            cos = F.cosine_similarity(self.normals(g_pred, keep), self.normals(g_gt, keep), dim=1).abs()
            return (total + self.tail(1 - cos)) / (n_planes + 1)
        return total / n_planes


class _VnlTriplets(torch.autograd.Function):
    """Per-triplet geometry of the plane term on the device (prn_vnl_triplets_fwd / _bwd, csrc/prn_loss.cu): points, selection
    mask and 1 - |cos| for every triplet in one launch; the backward recomputes the geometry and adds the depth gradient with fp32
    reductions.  Replaces ~80 torch launches forward and ~200 autograd launches backward of the tensor formulation below (which
    stays as the CPU / reference form: tests compare the two on the GPU)."""

    @staticmethod
    def forward(ctx, depth_up, gt32, fxfy, P, region, rest_u8, tgt64, delta_z):
        B, _, H, W = depth_up.shape
        T = region.numel()
        dev = depth_up.device
        assert depth_up.dtype == torch.float32 and depth_up.is_contiguous() and gt32.is_contiguous() and P.is_contiguous()
        loss_t = torch.empty(T, dtype=torch.float64, device=dev)
        keep = torch.empty(T, dtype=torch.uint8, device=dev)
        L.check(L.lib().prn_vnl_triplets_fwd(_p(depth_up), _p(gt32), _p(fxfy), _p(P), _p(region), _p(rest_u8), _p(tgt64), _p(loss_t),
                                             _p(keep), C.c_int64(T), H, W, C.c_float(delta_z), L.current_stream()), "prn_vnl_triplets_fwd")
        ctx.save_for_backward(depth_up, gt32, fxfy, P, region, rest_u8, tgt64)
        ctx.delta_z = delta_z
        ctx.mark_non_differentiable(keep)
        return loss_t, keep

    @staticmethod
    def backward(ctx, g_loss, _g_keep):
        depth_up, gt32, fxfy, P, region, rest_u8, tgt64 = ctx.saved_tensors
        B, _, H, W = depth_up.shape
        coef = g_loss.to(torch.float64).contiguous()
        d = torch.zeros_like(depth_up)
        L.check(L.lib().prn_vnl_triplets_bwd(_p(depth_up), _p(gt32), _p(fxfy), _p(P), _p(region), _p(rest_u8), _p(tgt64), _p(coef), _p(d),
                                             C.c_int64(region.numel()), H, W, C.c_float(ctx.delta_z), L.current_stream()),
                "prn_vnl_triplets_bwd")
        return d, None, None, None, None, None, None, None


class _PlaneNormalBatched:
    """The same plane surface-normal term (models/functions/vnl.py:6-165) for a whole batch in a few dozen tensor ops instead of a
    Python loop over every plane of every image (48 planes + 8 "rest" regions per batch of 8: ~1000 tiny launches and a host
    round trip per plane before).  One host sync fetches the pixel count of every region; the triplets are then drawn

      sampling="numpy"   with numpy's GLOBAL RNG in exactly the reference's call order (image by image, plane by plane, then the
                         non-planar rest: `choice` + `shuffle`, three times, vnl.py:48-53) — bit-identical triplets, hence loss
                         values and gradients equal to the reference's for a given `np.random.seed`.  The draws are made by a C
                         restatement of numpy's legacy choice + shuffle on numpy's own MT19937 state (prn_numpy_choice_shuffle;
                         "numpy_py" makes the literal numpy calls instead: 3x slower, the checker of the C path);
      sampling="device"  with torch's device RNG (i.i.d. uniform indices: the reference's shuffle of i.i.d. draws is a
                         statistical no-op) — same distribution, no host work;

    and everything after the sampling (gather of the triplets' 3-D points, the colinear / in-front / too-close tests, normals,
    |cos| against the plane normal resp. the ground-truth normals, worst-75 % tail per region) is evaluated for all triplets of
    all regions at once; the per-region sort of the tail is one global sort on the composite key (region, loss)."""

    def __init__(self, size=(480, 640), sample_ratio=0.3, delta_z=1e-4, sampling="numpy", threaded=True, kernels=True):
        assert sampling in ("numpy", "numpy_py", "device")
        self.size, self.ratio, self.delta_z, self.sampling, self.threaded = size, sample_ratio, delta_z, sampling, threaded
        self.kernels = kernels        # CUDA tensors: per-triplet geometry by prn_vnl_triplets_* instead of the tensor formulation
        self._grid = {}
        self._pinned, self._pinned_ev = None, None

    def _uv(self, dev):
        if dev not in self._grid:
            h, w = self.size
            u = (torch.arange(w, dtype=torch.float32, device=dev)[None, None, :] - float(w // 2)).expand(1, h, w)
            v = (torch.arange(h, dtype=torch.float32, device=dev)[None, :, None] - float(h // 2)).expand(1, h, w)
            self._grid[dev] = (u, v)
        return self._grid[dev]

    def _points(self, depth, fx, fy):
        """depth [B,1,H,W] -> [B*H*W, 3] (vnl.py:20-38: u0 / v0 are the image centre, not the intrinsics' principal point)."""
        u, v = self._uv(depth.device)
        d = depth[:, 0]
        x = u * d.abs() / fx
        y = v * d.abs() / fy
        return torch.stack([x, y, d], -1).reshape(-1, 3)

    def _region_ks(self, counts, is_rest):
        """k = int(n * ratio) per region; 0 for an empty rest region (vnl.py:137: the rest term only exists when there are
        non-planar pixels)."""
        ks = []
        for c, rest in zip(counts, is_rest):
            assert c <= self.size[0] * self.size[1]
            ks.append(0 if (rest and c == 0) else int(c * self.ratio))
        return ks

    def _sample_host_numpy(self, counts, is_rest, out):
        """The reference's own calls (np.random.choice + np.random.shuffle, three times per region): ~30 ns per index.  Kept as
        the checker of the C restatement below (tests/test_plane_normal_cpu.py) and selectable with sampling="numpy_py"."""
        import numpy as np
        ks = self._region_ks(counts, is_rest)
        off = 0
        for c, k in zip(counts, ks):
            if k == 0 and c == 0:
                continue                                 # the reference never reaches choice() for such a region
            for j in range(3):
                p = np.random.choice(c, k, replace=True)
                np.random.shuffle(p)
                out[j, off:off + k] = p
            off += k
        return ks

    def _sample_host(self, counts, is_rest, out):
        """Triplet indices of every region, drawn from numpy's GLOBAL RNG stream in the reference's call order, by the C
        restatement of legacy choice + shuffle (prn_numpy_choice_shuffle, csrc/prn_hostrng.cu): same indices, same stream
        position afterwards, ~3x less host time.  `out`: int32 numpy array [3, >= sum k].  Returns the per-region k."""
        import ctypes as C
        import numpy as np
        from . import _lib as L
        ks = self._region_ks(counts, is_rest)
        st = np.random.get_state()
        assert st[0] == "MT19937"
        key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
        pos = C.c_int32(int(st[2]))
        # regions with n == 0 draw nothing in the reference either (k == 0)
        n_arr = np.asarray(counts, dtype=np.int64)
        k_arr = np.asarray(ks, dtype=np.int64)
        assert out.dtype == np.int32 and out.ndim == 2 and out.strides[1] == 4
        L.check(L.lib().prn_numpy_choice_shuffle(key.ctypes.data_as(C.c_void_p), C.byref(pos), n_arr.ctypes.data_as(C.c_void_p),
                                                 k_arr.ctypes.data_as(C.c_void_p), len(ks), 3, out.ctypes.data_as(C.c_void_p),
                                                 out.strides[0] // 4), "prn_numpy_choice_shuffle")
        np.random.set_state((st[0], key, pos.value, st[3], st[4]))
        return ks

    def prepare(self, gt_instances):
        """Everything of the term that depends on the ground truth only: the region list, the pixel count of every region (the
        one host sync of this term) and — for the numpy-exact sampling — the triplet draws, started on a worker thread (the C
        sampler releases the GIL; ~25 ms per batch of 8 that now overlap the network's forward and the other loss terms).
        No np.random call may be made by the caller between prepare() and the __call__ that consumes it (the worker owns the
        global stream in between).  Returns an opaque dict for __call__(..., prep=...)."""
        H, W = self.size
        HW = H * W
        dev = gt_instances[0]["masks"].device
        # ---- regions in the reference's order: per image its planes, then the non-planar rest
        stacks, reg_img, reg_rest, n_planes = [], [], [], []
        for b, g in enumerate(gt_instances):
            assert tuple(g["masks"].shape[-2:]) == (H, W), "the plane term assumes the configured image size (losses.py:50)"
            m = g["masks"].bool().reshape(-1, HW)
            stacks += [m, torch.logical_not(m.any(0, keepdim=True))]
            n = m.shape[0]
            n_planes.append(n)
            reg_img += [b] * (n + 1)
            reg_rest += [False] * n + [True]
        member = torch.cat(stacks, 0)                                   # [R, HW] bool
        counts_dev = member.sum(1)
        counts = counts_dev.tolist()                                    # the one host sync of this term
        prep = dict(key=tuple(id(g["masks"]) for g in gt_instances), keep=[g["masks"] for g in gt_instances], R=member.shape[0],
                    counts_dev=counts_dev, counts=counts, reg_img=reg_img, reg_rest=reg_rest, n_planes=n_planes, dev=dev,
                    thread=None, ks=None, error=None)
        if self.sampling in ("numpy", "numpy_py"):
            T_max = sum(int(c * self.ratio) for c in counts)
            cuda = dev.type == "cuda"
            if self._pinned is None or self._pinned.shape[1] < T_max:
                self._pinned = torch.empty(3, max(T_max, 1 << 20), dtype=torch.int32, device="cpu")
                if cuda:
                    self._pinned = self._pinned.pin_memory()
            elif cuda and self._pinned_ev is not None:
                # the previous step's upload reads this buffer asynchronously: it has long finished (a whole backward lies in
                # between), but make that explicit before overwriting it
                self._pinned_ev.synchronize()
            sampler = self._sample_host if self.sampling == "numpy" else self._sample_host_numpy
            out = self._pinned.numpy()                                   # written straight into the (pinned) staging buffer

            def work():
                try:
                    prep["ks"] = sampler(counts, reg_rest, out)
                except BaseException as exc:                             # re-raised by the consumer
                    prep["error"] = exc

            if self.sampling == "numpy" and self.threaded:
                import threading
                prep["thread"] = threading.Thread(target=work, name="prn-vnl-sampler", daemon=True)
                prep["thread"].start()
            else:
                work()
        else:
            prep["ks"] = self._region_ks(counts, reg_rest)
        # pixel lists of the regions (row-major inside a region = the order of `pts[mask]`); nonzero() is a host sync, which is
        # why it lives here (ground truth only) and not after the forward has been enqueued
        img_of = torch.tensor(reg_img, dtype=torch.int64, device="cpu").to(dev, non_blocking=True)
        nz = member.nonzero()
        prep["gpix"] = img_of[nz[:, 0]] * HW + nz[:, 1]
        prep["base"] = torch.cumsum(counts_dev, 0) - counts_dev          # first entry of each region in gpix
        prep["rest_rows"] = torch.tensor([i for i, r in enumerate(reg_rest) if r], dtype=torch.int64,
                                         device="cpu").to(dev, non_blocking=True)   # one rest region per image, in image order
        return prep

    def __call__(self, depth_up, gt_instances, gt_depths, prep=None):
        """depth_up [B,1,H,W] (x2 bilinear of the prediction), gt_depths [B,1,H,W].  Returns the per-image losses [B] (float64; NaN
        where the reference yields NaN).  prep: the result of prepare(gt_instances) when the caller made it ahead of time."""
        import numpy as np
        import torch.nn.functional as F
        B, _, H, W = depth_up.shape
        assert (H, W) == tuple(self.size), "the plane term assumes the configured image size (losses.py:50)"
        HW = H * W
        dev = depth_up.device
        if prep is None or prep["key"] != tuple(id(g["masks"]) for g in gt_instances):
            prep = self.prepare(gt_instances)
        if prep["thread"] is not None:
            prep["thread"].join()
            prep["thread"] = None
        if prep["error"] is not None:
            raise prep["error"]
        R, counts_dev, reg_img, reg_rest, n_planes, ks = (prep[k] for k in ("R", "counts_dev", "reg_img", "reg_rest", "n_planes", "ks"))
        T = sum(ks)
        # ---- triplet indices
        if self.sampling in ("numpy", "numpy_py"):
            idx = self._pinned[:, :T].to(dev, non_blocking=True).long()
            if dev.type == "cuda":
                self._pinned_ev = torch.cuda.Event()
                self._pinned_ev.record()
        # the per-region host tables in ONE upload: k, image of the region, rest flag
        tab = torch.tensor([ks, reg_img, [int(r) for r in reg_rest]], dtype=torch.int64, device="cpu").to(dev, non_blocking=True)
        ks_t, img_of, is_rest_r = tab[0], tab[1], tab[2].bool()
        region = torch.repeat_interleave(torch.arange(R, device=dev), ks_t, output_size=T)            # [T]
        counts_t = counts_dev
        if self.sampling == "device":
            idx = (torch.rand(3, T, device=dev, dtype=torch.float64) * counts_t[region].double()).long()
            idx = torch.minimum(idx, (counts_t[region] - 1).clamp(min=0))
        gpix, base = prep["gpix"], prep["base"]
        P = gpix[(base[region][None, :] + idx).reshape(-1)].reshape(3, T)                              # global pixel of every point
        # ---- point clouds (vnl.py:20-38)
        fx = torch.stack([g["k_matrix"][0, 0] for g in gt_instances]).to(device=dev, dtype=torch.float32)[:, None, None]
        fy = torch.stack([g["k_matrix"][1, 1] for g in gt_instances]).to(device=dev, dtype=torch.float32)[:, None, None]
        is_rest_t = is_rest_r[region]
        # target normal of every triplet: the plane's ground-truth normal (float64 like the reference's plane_paras)
        tgt = torch.zeros(R, 3, dtype=torch.float64, device=dev)
        r0 = 0
        for b, g in enumerate(gt_instances):
            tgt[r0:r0 + n_planes[b]] = g["plane_paras"][:, :3].to(device=dev, dtype=torch.float64)
            r0 += n_planes[b] + 1
        if self.kernels and dev.type == "cuda" and depth_up.dtype == torch.float32:
            fxfy = torch.stack([fx.flatten(), fy.flatten()], 1).contiguous()
            loss_t, keep_u8 = _VnlTriplets.apply(depth_up.contiguous(), gt_depths.to(torch.float32).contiguous(), fxfy, P.contiguous(),
                                                 region, is_rest_r.to(torch.uint8), tgt, float(self.delta_z))
            return self._tail(loss_t, keep_u8.bool(), region, is_rest_t, is_rest_r, img_of, counts_t, prep, n_planes, R, T, B, dev)
        pred_pts = self._points(depth_up, fx, fy)
        gt_pts = self._points(gt_depths.to(depth_up.dtype), fx, fy)
        g_pred = pred_pts[P].permute(1, 2, 0)                            # [T, xyz, p123]
        g_gt = gt_pts[P].permute(1, 2, 0)
        g_test = torch.where(is_rest_t[:, None, None], g_gt, g_pred)     # planes are tested on the prediction, the rest on the GT
        # ---- vnl.py:56-98: usable triplets
        # (a selection mask: no gradient flows through it.  The 3x3 products are written out element-wise — torch.bmm with
        # 7e5 batches of 3x3 matrices runs as 36 SIMT sgemm launches of 32x32 tiles, 5 ms per step)
        with torch.no_grad():
            gt_ = g_test.detach()
            diff = torch.stack([gt_[:, :, 1] - gt_[:, :, 0], gt_[:, :, 2] - gt_[:, :, 0], gt_[:, :, 2] - gt_[:, :, 1]], 2)   # [T, xyz, 3]
            q = diff.permute(0, 2, 1)                                    # [T, 3, xyz]
            qn = q.norm(2, dim=2)                                        # [T, 3]
            dots = (q[:, :, None, :] * q[:, None, :, :]).sum(3)          # q . diff: [T, 3, 3]
            cosm = (dots / (qn[:, :, None] * qn[:, None, :] + 1e-8)).reshape(T, -1)
            colinear = ((cosm > 0.985) | (cosm < -0.985)).sum(1) > 3
            in_front = (gt_[:, 2, :] > self.delta_z).sum(1) == 3
            dd = torch.where(is_rest_t, 0.1, 0.005).to(diff.dtype)[:, None]
            near = (((diff[:, 0, :].abs() < dd).sum(1) > 0) & ((diff[:, 1, :].abs() < dd).sum(1) > 0) &
                    ((diff[:, 2, :].abs() < dd).sum(1) > 0))
            keep = in_front & ~(near | colinear)

        def normals(g):
            nv = torch.cross(g[:, :, 1] - g[:, :, 0], g[:, :, 2] - g[:, :, 0], dim=1)
            nrm = torch.norm(nv, 2, dim=1, keepdim=True)
            return nv / (nrm + (nrm == 0.0).float() * 0.01)

        # vnl.py:146: a predicted point with z == 0 (never, behind a softplus) is nudged; indexing quirk of the reference kept
        g_pred_rest = torch.where((g_pred[:, 2, :] == 0)[:, :, None], torch.full_like(g_pred, 0.0001), g_pred)
        n_plane = normals(g_pred)
        n_rest = normals(g_pred_rest)
        n_gt = normals(g_gt)
        cos_plane = F.cosine_similarity(n_plane, tgt[region], dim=1).abs()            # float64
        cos_rest = F.cosine_similarity(n_rest, n_gt, dim=1).abs()                     # float32
        loss_t = torch.where(is_rest_t, (1 - cos_rest).double(), 1 - cos_plane)        # [T] float64 (rest values are exact fp32)
        return self._tail(loss_t, keep, region, is_rest_t, is_rest_r, img_of, counts_t, prep, n_planes, R, T, B, dev)

    def _tail(self, loss_t, keep, region, is_rest_t, is_rest_r, img_of, counts_t, prep, n_planes, R, T, B, dev):
        """Per-region worst-75 % tails and per-image means (vnl.py:100-165) from the per-triplet losses and selection mask."""
        # ---- vnl.py:100-117: worst 75 % per region: one global sort on (region, loss); NaNs sort last inside their region
        key = torch.where(keep, region.double() * 2 + torch.where(torch.isnan(loss_t), torch.full_like(loss_t, 1.5), loss_t),
                          torch.full_like(loss_t, float("inf")))
        order = torch.argsort(key)
        pos = torch.empty_like(order)
        pos[order] = torch.arange(T, device=dev)
        # per-region sums: T ~ 7e5 values into R ~ 60 rows.  index_add_ straight into [R] serialises on 60 addresses (3 ms per
        # step); spread every region over 128 slots first and fold the slots afterwards
        slots = 128
        slot_of = region * slots + (torch.arange(T, device=dev) & (slots - 1))

        def region_sum(values, dtype):
            return torch.zeros(R * slots, dtype=dtype, device=dev).index_add_(0, slot_of, values).view(R, slots).sum(1)

        n_keep = region_sum(keep.long(), torch.long)
        start = torch.cumsum(n_keep, 0) - n_keep
        drop = n_keep // 4                                                # int(n * 0.25)
        incl = keep & ((pos - start[region]) >= drop[region])
        contrib = torch.where(incl & ~torch.isnan(loss_t), loss_t, torch.zeros_like(loss_t))
        # the rest regions' tail is a float32 sum in the reference: accumulate it in float32, the planes' in float64
        sum_plane = region_sum(torch.where(is_rest_t, 0.0, contrib), torch.float64)
        sum_rest = region_sum(torch.where(is_rest_t, contrib, 0.0).float(), torch.float32)
        den = (n_keep - drop).double()
        loss_r = torch.where(is_rest_r, sum_rest.double() / den.float().double(), sum_plane / den)       # 0 / 0 -> NaN like the reference
        # ---- per image (vnl.py:119-165)
        total = torch.zeros(B, dtype=torch.float64, device=dev).index_add_(0, img_of, torch.where(is_rest_r, 0.0, loss_r))
        npl = torch.tensor(n_planes, dtype=torch.float64, device="cpu").to(dev, non_blocking=True)
        rest_rows = prep["rest_rows"]
        rest_used = (counts_t[rest_rows] > 0) & (n_keep[rest_rows] > 0)
        with_rest = (total + torch.where(rest_used, loss_r[rest_rows], torch.zeros_like(total))) / (npl + 1)
        return torch.where(rest_used, with_rest, total / npl)


# ------------------------------------------------------------------------------------------ the joint loss
class PlaneRecNetLoss(torch.nn.Module):
    """Drop-in for models/functions/losses.py:PlaneRecNetLoss (same constructor-time cfg fields, same forward signature and
    result dict {'ins','cat','dpt','pln','lav'}): target assignment on the device (targets.py), dice + lava and focal and
    RMSE-log terms on libprn_b200 kernels, plane-normal term host-side.  Reference quirks kept: the lava valid mask stays
    None for the presets' dataset name (losses.py:172), the plane term assumes 480x640 (losses.py:50)."""

    def __init__(self, cfg=None, backend=None, vnl_sampling="numpy"):
        """vnl_sampling: 'numpy' (default) draws the plane term's triplets with numpy's global RNG in the reference's call order
        (bit-identical samples for a given np.random.seed; C restatement of numpy's legacy choice + shuffle, ~25 ms per batch of 8
        on a worker thread), 'device' draws the same distribution with torch's device RNG (no host work)."""
        super().__init__()
        if cfg is None:
            from .config import cfg as _cfg
            cfg = _cfg
        self.num_classes = cfg.num_classes
        self.num_grids = list(cfg.solov2.num_grids)
        self.scale_ranges = cfg.solov2.fpn_scale_ranges
        self.sigma = cfg.solov2.sigma
        self.focal_alpha, self.focal_gamma = cfg.focal_alpha, cfg.focal_gamma
        self.w_ins, self.w_cat, self.w_dpt = cfg.dice_weight, cfg.focal_weight, cfg.depth_weight
        self.w_lav, self.w_pln = cfg.lava_weight, cfg.pln_weight
        self.use_lava, self.use_plane = cfg.use_lava_loss, cfg.use_plane_loss
        self.min_depth, self.depth_resolution = cfg.dataset.min_depth, cfg.dataset.depth_resolution
        self.backend = backend
        self.vnl = _PlaneNormal((480, 640))                                  # per-plane formulation (kept as the readable mirror of vnl.py)
        self.vnl_batched = _PlaneNormalBatched((480, 640), sampling=vnl_sampling)
        self._prep, self._side = None, None

    def prepare(self, gt_instances, feat_hw=None, background=None):
        """The ground-truth-only part of a step — SOLOv2 target assignment and the plane term's region counts, pixel lists and
        triplet sampling — started BEFORE the network's forward is enqueued (or, prefetching, for the NEXT batch before this
        step's backward):

            crit.prepare(gts); outs = net(x); losses = crit(net, *outs, gts, gt_depths)

        On a CUDA device the work runs on a worker thread and a side stream (`background`, default there): its host round trips
        wait for the ground truth's upload only, not for the forward, the C sampler (~25 ms per batch of 8, GIL released)
        overlaps the forward / the previous backward, and prepare() itself returns at once.  forward() joins it.  Optional:
        forward() does the same work inline when no matching preparation exists (the reference's call pattern, train.py:339-346).
        feat_hw: size of the mask feature map (default: a quarter of the ground-truth masks — the mask head runs at stride 4).
        No np.random call may be made between prepare() and the forward() that consumes it (the sampler owns numpy's global
        stream in between)."""
        from .targets import assign_targets_batch
        if feat_hw is None:
            h, w = gt_instances[0]["masks"].shape[-2:]
            feat_hw = (h // 4, w // 4)
        feat_hw = tuple(feat_hw)
        dev = gt_instances[0]["masks"].device
        if background is None:
            background = dev.type == "cuda"
        if self._prep is not None:
            # an earlier preparation that was never consumed: let its threads end before a new sampler touches numpy's stream
            self._drain(self._prep)
            self._prep = None
        prep = dict(key=tuple(id(g["masks"]) for g in gt_instances), keep=[g["masks"] for g in gt_instances], feat_hw=feat_hw,
                    targets=None, vnl=None, thread=None, error=None, done=None)

        def work():
            prep["vnl"] = self.vnl_batched.prepare(gt_instances) if self.use_plane else None   # first: its sampler starts earliest
            prep["targets"] = assign_targets_batch(gt_instances, feat_hw, self.num_grids, self.scale_ranges, self.num_classes, self.sigma)

        if background and dev.type == "cuda":
            import threading
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(dev)
            side = self._side
            uploaded = torch.cuda.Event()
            uploaded.record(torch.cuda.current_stream(dev))              # the ground truth's H2D copies precede this point

            def run():
                try:
                    torch.cuda.set_device(dev)
                    with torch.cuda.stream(side):
                        side.wait_event(uploaded)
                        work()
                        prep["done"] = torch.cuda.Event()
                        prep["done"].record(side)
                except BaseException as exc:                              # re-raised by the consumer
                    prep["error"] = exc

            prep["thread"] = threading.Thread(target=run, name="prn-loss-prepare", daemon=True)
            prep["thread"].start()
        else:
            work()
        self._prep = prep
        return prep

    @staticmethod
    def _drain(prep):
        """Wait for every thread of a preparation (its worker and the plane term's sampler) without consuming it."""
        if prep.get("thread") is not None:
            prep["thread"].join()
            prep["thread"] = None
        vnl = prep.get("vnl")
        if vnl is not None and vnl.get("thread") is not None:
            vnl["thread"].join()
            vnl["thread"] = None

    @staticmethod
    def _finish(prep, dev):
        """Join a background preparation and order the consumer's stream after it."""
        if prep["thread"] is not None:
            prep["thread"].join()
            prep["thread"] = None
        if prep["error"] is not None:
            raise prep["error"]
        if prep["done"] is not None:
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(prep["done"])

            def rec(o):                    # tensors allocated under the side stream, consumed on this one
                if torch.is_tensor(o):
                    if o.is_cuda:
                        o.record_stream(cur)
                elif isinstance(o, dict):
                    for k, v in o.items():
                        if k != "keep":
                            rec(v)
                elif isinstance(o, (list, tuple)):
                    for v in o:
                        rec(v)
                    rec(getattr(o, "dense", None))          # targets.TargetsBatch: the batch-layout tensors

            rec(prep["targets"])
            rec(prep["vnl"])
            prep["done"] = None
        return prep

    def forward(self, net, mask_preds, cate_preds, kernel_preds, depth_preds, gt_instances, gt_depths):
        import torch.nn.functional as F
        be = self.backend or CudaBackend()
        B = len(gt_instances)
        fh, fw = mask_preds.shape[-2:]
        prep, self._prep = self._prep, None
        if prep is not None and (prep["key"] != tuple(id(g["masks"]) for g in gt_instances) or prep["feat_hw"] != (fh, fw)):
            self._drain(prep)                                            # a stale preparation: let its threads end, then ignore it
            prep = None
        if prep is None:
            prep = self.prepare(gt_instances, (fh, fw), background=False)
            self._prep = None
        self._finish(prep, mask_preds.device)
        targets = prep["targets"]
        losses = {}
        gw, gsum = be.lava_weights(gt_depths, fh, fw, self.depth_resolution)
        if not self.use_lava:
            gsum = torch.zeros_like(gsum)
        ins, lav = _InsLava.apply(be, targets, gw, gsum, self.w_ins, self.w_lav, mask_preds, *kernel_preds)
        losses["ins"] = ins
        # category: rows ordered (level, image, cell) like losses.py:121-133
        n_levels = len(self.num_grids)
        dense = getattr(targets, "dense", None)
        if dense is not None:                     # [B, all cells] label map of the batch: one slice per level instead of 32 tensors
            offs = dense["level_off"]
            labels = torch.cat([dense["cate"][:, o:o + S * S].reshape(-1) for o, S in zip(offs, self.num_grids)])
        else:
            labels = torch.cat([torch.cat([targets[b][l][1].flatten() for b in range(B)]) for l in range(n_levels)])
        logits = torch.cat([c.permute(0, 2, 3, 1).reshape(-1, self.num_classes) for c in cate_preds]).contiguous()
        # number of distinct positive cells (losses.py:114: sum of the boolean cell map, not of the instance rows)
        # (counted from the host-side cell lists: the device maps would cost 32 host syncs)
        num_ins = sum(len(set(targets[b][l][3])) for b in range(B) for l in range(n_levels))
        losses["cat"] = self.w_cat * be.focal_sum(logits, labels, self.focal_alpha, self.focal_gamma, self.num_classes) / (num_ins + 1)
        losses["dpt"] = be.depth_rmselog(depth_preds, gt_depths, self.min_depth, 1e-9, self.w_dpt)
        if self.use_plane:
            up = F.interpolate(depth_preds, scale_factor=2, mode="bilinear", align_corners=False)
            losses["pln"] = self.vnl_batched(up, gt_instances, gt_depths, prep=prep["vnl"]).mean() * self.w_pln
        if self.use_lava:
            losses["lav"] = lav
        return losses
