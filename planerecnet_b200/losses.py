"""Dense parts of PlaneRecNetLoss on libprn_b200 — first two terms, each one autograd node:

  focal_cate_loss(cate32, labels, num_ins)   losses.py:121-138  sigmoid focal over all grid cells / (num_ins + 1)
  depth_rmselog_loss(depth_pred, gt_depths)  losses.py:141-147  depth_weight * RMSE-log on the x2-upsampled prediction

The dice / lava terms (dynamic 1x1 convolutions of the mask features with the positive cells' kernels) and the plane
normal loss are not built yet; the reference's own implementations run on the training outputs unchanged meanwhile.
Targets come from planerecnet_b200.targets (device-resident assignment)."""
import ctypes as C

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
