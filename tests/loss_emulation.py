"""Torch (fp32, any device) emulation of planerecnet_b200.losses.CudaBackend — TEST INFRASTRUCTURE: lets the CPU suite run
the orchestration and gradient algebra of the dice / lava losses against the pinned oracle without a GPU."""
import torch
import torch.nn.functional as F


class EmuBackend:
    def to16(self, t):
        return t.float().contiguous()

    def seg_rows(self, wsel, mask):
        return torch.sigmoid(torch.bmm(wsel, mask.transpose(1, 2))).reshape(-1, mask.shape[1])

    def row_stats(self, seg, target, gw, n):
        t = target.float()
        g = gw.repeat_interleave(n, 0)
        return torch.stack([(seg * t).sum(1), (seg * seg).sum(1), (t * t).sum(1), (seg * g).sum(1)], 1)

    def row_bwd(self, seg, target, gw, coef, n):
        g = gw.repeat_interleave(n, 0)
        return (coef[:, 0:1] * target.float() + 2 * coef[:, 1:2] * seg + coef[:, 2:3] * g) * seg * (1 - seg)

    def grouped_nt(self, a, w):
        return torch.bmm(a, w.transpose(1, 2))

    def lava_weights(self, gt, h, w, depth_res):
        from oracle import prn_loss_oracle as LO
        g = LO.gradient_map(gt, None) / torch.pow(gt.clamp(min=depth_res), 2)
        g = g.clamp(max=1e-2)
        g[g < 1e-4] = 0
        probe = torch.zeros(gt.shape[0], 1, h, w, requires_grad=True)
        up = F.interpolate(probe, size=gt.shape[-2:], mode="bilinear")
        (gw,) = torch.autograd.grad((up * g).sum(), probe)
        return gw.reshape(gt.shape[0], -1), g.reshape(gt.shape[0], -1).sum(1)

    def focal_sum(self, logits, labels, alpha, gamma, nc):
        from oracle import prn_loss_oracle as LO
        oh = torch.zeros(logits.shape[0], nc)
        pos = torch.nonzero(labels != nc).squeeze(1)
        oh[pos, labels[pos]] = 1
        return LO.sigmoid_focal_sum(logits[:, :nc], oh, alpha, gamma)

    def depth_rmselog(self, depth, gt, min_depth, clamp_val, weight):
        from oracle import prn_loss_oracle as LO
        up = F.interpolate(depth, scale_factor=2, mode="bilinear", align_corners=False)
        return weight * LO.rmse_log_mean(up, gt, gt > min_depth, clamp_val)
