"""PlaneRecNet model shell — the reference's public module API (planerecnet.py:20-607): same class
names, constructor arguments, attribute names, state_dict keys and output contracts, so that the
reference's train.py / eval.py / simple_inference.py can use it unchanged.

The nn.Modules below only own parameters and buffers.  All arithmetic of `forward` runs in
libprn_b200 (hand-written sm_100a kernels behind the C ABI of include/prn_b200.h) through
`planerecnet_b200.engine.Engine`; there is no torch-operator or CPU fallback."""
import torch
import torch.nn as nn

from .models.backbone import construct_backbone
from .models.fpn import FPN
from .models.functions.funcs import bias_init_with_prob


def _reflect_conv_bn_relu(cin, cout, upsample=False):
    """[Upsample(nearest x2)] -> ReflectionPad2d(1) -> Conv2d 3x3 -> BatchNorm2d(eps 1e-3, momentum 0.01) -> ReLU
    (planerecnet.py:515-568).  Index positions inside the Sequential define the state_dict keys."""
    mods = []
    if upsample:
        mods.append(nn.Upsample(scale_factor=2, mode="nearest", align_corners=None))
    mods += [nn.ReflectionPad2d(1), nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=0),
             nn.BatchNorm2d(cout, eps=0.001, momentum=0.01), nn.ReLU(inplace=True)]
    return nn.Sequential(*mods)


class PlaneRecNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.device = torch.device(cfg.device)
        self.depth_decoder_indices = cfg.depth.selected_layers
        self.fpn_indices = cfg.fpn.selected_layers

        s = cfg.solov2
        self.num_classes = cfg.num_classes
        self.num_kernels = s.num_kernels
        self.num_grids = s.num_grids
        self.instance_in_features = s.instance_in_features
        self.instance_strides = s.fpn_instance_strides
        self.instance_in_channels = cfg.fpn.num_features
        self.instance_channels = s.instance_channels
        self.mask_in_features = s.masks_in_features
        self.mask_in_channels = cfg.fpn.num_features
        self.mask_channels = s.masks_channels
        self.num_masks = s.num_masks
        self.max_before_nms = s.nms_pre
        self.score_threshold = s.score_thr
        self.update_threshold = s.update_thr
        self.mask_threshold = s.mask_thr
        self.max_per_img = s.top_k
        self.nms_kernel = s.nms_kernel
        self.nms_sigma = s.nms_sigma
        self.nms_type = s.nms_type

        # construction order == reference (planerecnet.py:55-71): a seeded default init gives identical weights
        self.backbone = construct_backbone(cfg.backbone)
        if cfg.freeze_bn:
            self.freeze_bn()
        src_channels = self.backbone.channels
        self.fpn = FPN([src_channels[i] for i in self.fpn_indices], start_level=cfg.fpn.start_level, cfg=cfg)
        self.depth_decoder = DepthDecoder_FPN(cfg)
        self.inst_head = SOLOv2InsHead(cfg, [cfg.fpn.num_features] * len(s.instance_in_features))
        self.mask_head = SOLOv2MaskHead(cfg, [cfg.fpn.num_features] * len(s.masks_in_features))
        self._engine = None
        self.use_cuda_graph = True

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine()
        return self._engine

    @property
    def train_engine(self):
        """Executor of the training branch (bf16 activations/gradients, fp32 weight gradients)."""
        if getattr(self, "_train_engine", None) is None:
            from .train_engine import TrainEngine
            self._train_engine = TrainEngine("bf16")
        return self._train_engine

    def set_train_precision(self, name, grad_scale=1.0):
        """'bf16' (default) or 'f16' activations / gradients for the training step.  f16 carries 3 more mantissa bits
        (gradient cosine vs fp32 autograd 0.9998 instead of 0.9987 on the parity model) but needs `grad_scale` (loss
        scaling) once gradients fall below ~6e-8."""
        from .train_engine import TrainEngine
        self._train_engine = TrainEngine(name)
        self._train_engine.grad_scale = float(grad_scale)
        return self

    def set_precision(self, name):
        """'f16' (default: 10-bit mantissa, ~1e-3 end-to-end) or 'bf16' (~1e-2) storage/operand type of the
        tensor-core path; accumulation is fp32 in TMEM either way."""
        from .engine import Engine
        self._engine = Engine(dtype=name)
        return self

    # ------------------------------------------------------------------ forward (planerecnet.py:73-111)
    def forward_dense(self, x):
        """Dense outputs (mask_pred, [cate_pred]x4, [kernel_pred]x4, depth_pred), NCHW fp32 — what the
        reference's training branch returns (planerecnet.py:101-103)."""
        return self.engine.forward_dense(self, x)["outputs"]

    def forward(self, x):
        from .utils import timer
        if self.training:
            # planerecnet.py:101-103: (mask_pred, cate_pred, kernel_pred, depth_pred), autograd-connected through one
            # node whose backward replays the sm_100a tape (train_engine.py)
            from .train_engine import forward_train_autograd
            if torch.is_grad_enabled():
                return forward_train_autograd(self, x)
            outs = self.train_engine.forward_train(self, x)
            self.train_engine.reset()
            return outs
        with timer.env("dense forward"):
            st = (self.engine.forward_dense_graph if self.use_cuda_graph else self.engine.forward_dense)(self, x, False)
        with timer.env("Inferencing"):
            return self.engine.inference(self, st, x)

    def forward_frames(self, frames):
        """Eval-mode `net(FastBaseTransform()(frames))` (simple_inference.py:226-231, eval.py:85-88) for camera frames
        [B, H, W, 3] BGR, uint8 or fp32 0..255 on CUDA, with the transform (mean/std per BGR channel, BGR -> RGB) and
        `pad_even_divided` (zero raw pixels up to the next multiples of 32) folded into the stem's im2col: no NHWC -> NCHW ->
        NHWC bounce and no normalised fp32 copy of the batch.  Detections refer to the padded size, like the reference's."""
        assert not self.training, "forward_frames is an eval-mode entry point"
        B, Hi, Wi, _ = frames.shape
        Hp, Wp = (Hi + 31) // 32 * 32, (Wi + 31) // 32 * 32
        st = (self.engine.forward_dense_graph if self.use_cuda_graph else self.engine.forward_dense)(self, frames, False, frames=True)
        shape_only = torch.empty(1, device=frames.device).expand(B, 3, Hp, Wp)      # the bookkeeping only reads the input size
        return self.engine.inference(self, st, shape_only)

    def infer_pipelined(self, batches, depth=3):
        """Serving loop over an iterable of equally shaped input batches (pinned host or device tensors): yields, in
        order, exactly what `net(x)` returns for each batch.  Up to `depth` batches are in flight: while batch k's
        inference bookkeeping — which has to wait for the host at its data-dependent steps (planerecnet.py:189-269) —
        runs on a high-priority stream, the dense forwards of batches k+1 .. k+depth-1 are already queued on the forward
        stream (one CUDA-graph slot each) and the next input is being copied on the copy stream."""
        assert not self.training, "infer_pipelined is an eval-mode loop"
        assert depth >= 2
        from collections import deque
        eng = self.engine
        if getattr(self, "_pipe_streams", None) is None:
            # bookkeeping = many tiny kernels with host round trips in between: a high-priority stream lets their CTAs
            # in ahead of the forward's next persistent kernel instead of queueing behind it
            self._pipe_streams = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream(priority=-1))
        s_copy, s_fwd, s_book = self._pipe_streams
        start = torch.cuda.Event()   # work already queued on the caller's stream may still read the slots' static buffers
        start.record(torch.cuda.current_stream())
        done = [start] * depth       # per graph slot: the bookkeeping that read its static buffers has been enqueued
        inflight = deque()

        def finish(p):
            st, xb, ev_f, slot = p
            main = torch.cuda.current_stream()
            s_book.wait_event(ev_f)
            with torch.cuda.stream(s_book):
                res = eng.inference(self, st, xb)
                e = torch.cuda.Event()
                e.record(s_book)
            done[slot] = e
            main.wait_event(e)                     # the caller consumes the results on its own stream
            for r in res:
                for v in r.values():
                    if v is not None:
                        v.record_stream(main)
            return res

        def issue_copy(xh):
            with torch.cuda.stream(s_copy):
                xb = xh if xh.is_cuda else xh.cuda(non_blocking=True)
                ev_c = torch.cuda.Event()
                ev_c.record(s_copy)
            return xb, ev_c

        with torch.no_grad():
            it = iter(batches)
            first = next(it, None)
            pending = issue_copy(first) if first is not None else None
            k = 0
            while pending is not None:
                if len(inflight) == depth:         # every slot holds a batch: deliver the oldest first
                    yield finish(inflight.popleft())
                xb, ev_c = pending
                nxt = next(it, None)               # the copy of the following batch is issued before this forward is queued
                pending = issue_copy(nxt) if nxt is not None else None
                slot = k % depth
                s_fwd.wait_event(ev_c)
                s_fwd.wait_event(done[slot])
                with torch.cuda.stream(s_fwd):
                    st = eng.forward_dense_graph(self, xb, False, slot=slot)
                    ev_f = torch.cuda.Event()
                    ev_f.record(s_fwd)
                xb.record_stream(s_fwd)
                inflight.append((st, xb, ev_f, slot))
                k += 1
            while inflight:
                yield finish(inflight.popleft())

    @staticmethod
    def split_feats(feats):
        """(x0.5 bilinear of P2, P3, P4, P5): planerecnet.py:113-118 — NCHW fp32 in/out, computed on the engine."""
        from .engine import default_engine
        eng = default_engine()
        p2 = eng.to_nhwc(feats[0])
        return (eng.to_nchw(eng.avgpool2(p2), feats[0].shape[1]), feats[1], feats[2], feats[3])

    def _apply(self, fn, *args, **kwargs):
        """Device / dtype moves may replace parameter tensors: drop the cached parameter and BatchNorm lists of the training
        boundary (train_engine.forward_train_autograd)."""
        self._params_cache = None
        self._bn_modules_cache = None
        return super()._apply(fn, *args, **kwargs)

    # ------------------------------------------------------------------ weights (planerecnet.py:121-153)
    def save_weights(self, path):
        torch.save(self.state_dict(), path)

    def load_weights(self, path):
        self.load_state_dict(torch.load(path))

    def init_weights(self, backbone_path):
        self.backbone.init_backbone(backbone_path)
        for name, module in self.named_modules():
            if isinstance(module, nn.Conv2d) and module not in self.backbone.backbone_modules:
                nn.init.xavier_uniform_(module.weight.data)
                if module.bias is not None:
                    if "inst_head" in name and "cate_pred" in name:
                        module.bias.data.fill_(bias_init_with_prob(self.cfg.solov2.focal_loss_init_pi))
                    else:
                        module.bias.data.fill_(0)

    def freeze_bn(self, enable=False):
        for module in self.modules():
            if isinstance(module, nn.BatchNorm2d):
                module.train() if enable else module.eval()
                module.weight.requires_grad = enable
                module.bias.requires_grad = enable


class SOLOv2InsHead(nn.Module):
    """planerecnet.py:292-391: shared-weight category / kernel towers over the S x S grids."""

    def __init__(self, cfg, in_channels):
        super().__init__()
        s = cfg.solov2
        self.num_classes = cfg.num_classes
        self.num_kernels = s.num_kernels
        self.num_grids = s.num_grids
        self.instance_in_features = s.instance_in_features
        self.instance_strides = s.fpn_instance_strides
        self.instance_in_channels = cfg.fpn.num_features
        self.instance_channels = s.instance_channels
        self.num_levels = len(self.instance_in_features)
        assert self.num_levels == len(self.instance_strides), "Strides should match the features."
        assert len(set(in_channels)) == 1, "Each level must have the same channel!"
        if s.norm != "GN" or s.use_dcn_in_instance or not s.use_coord_conv:
            raise NotImplementedError("only the presets' GN / coord-conv / no-DCN instance head is implemented")
        for head, use_coord in (("cate", False), ("kernel", True)):
            tower = []
            for i in range(s.num_instance_convs):
                chn = (self.instance_in_channels + (2 if use_coord else 0)) if i == 0 else self.instance_channels
                tower += [nn.Conv2d(chn, self.instance_channels, kernel_size=3, stride=1, padding=1, bias=False),
                          nn.GroupNorm(32, self.instance_channels), nn.ReLU(inplace=True)]
            self.add_module(f"{head}_tower", nn.Sequential(*tower))
        self.cate_pred = nn.Conv2d(self.instance_channels, self.num_classes, kernel_size=3, stride=1, padding=1)
        self.kernel_pred = nn.Conv2d(self.instance_channels, self.num_kernels, kernel_size=3, stride=1, padding=1)

    def forward(self, features):
        """features: (x0.5 P2, P3, P4, P5) NCHW fp32 -> ([cate_pred]x4, [kernel_pred]x4) NCHW fp32."""
        from .engine import engine_for
        eng = engine_for(self)
        st = eng.inst_head([eng.to_nhwc(f) for f in features], self)
        return eng.inst_outputs_nchw(st, self)


class SOLOv2MaskHead(nn.Module):
    """planerecnet.py:394-496: per-level conv/GN/ReLU(+x2 bilinear) towers summed, then 1x1 conv/GN/ReLU."""

    def __init__(self, cfg, input_shape):
        super().__init__()
        s = cfg.solov2
        self.num_masks = s.num_masks
        self.mask_in_features = s.masks_in_features
        self.mask_in_channels = cfg.fpn.num_features
        self.mask_channels = s.masks_channels
        self.num_levels = len(input_shape)
        assert self.num_levels == len(self.mask_in_features), "Input shape should match the features."
        if s.norm != "GN":
            raise NotImplementedError("only the presets' GN mask head is implemented")

        def tower(cin):
            return nn.Sequential(nn.Conv2d(cin, self.mask_channels, kernel_size=3, stride=1, padding=1, bias=False),
                                 nn.GroupNorm(32, self.mask_channels), nn.ReLU(inplace=False))

        self.convs_all_levels = nn.ModuleList()
        for i in range(self.num_levels):
            level = nn.Sequential()
            if i == 0:
                level.add_module("conv0", tower(self.mask_in_channels))
            for j in range(i):
                cin = self.mask_channels if j > 0 else (self.mask_in_channels + 2 if i == 3 else self.mask_in_channels)
                level.add_module(f"conv{j}", tower(cin))
                level.add_module(f"upsample{j}", nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False))
            self.convs_all_levels.append(level)
        self.conv_pred = nn.Sequential(
            nn.Conv2d(self.mask_channels, self.num_masks, kernel_size=1, stride=1, padding=0, bias=False),
            nn.GroupNorm(32, self.num_masks), nn.ReLU(inplace=True))

    def forward(self, features):
        assert len(features) == self.num_levels, "The number of input features should be equal to the supposed level."
        from .engine import engine_for
        eng = engine_for(self)
        m = eng.mask_head([eng.to_nhwc(f) for f in features], self)
        return eng.to_nchw(m, self.num_masks)


class DepthDecoder_FPN(nn.Module):
    """planerecnet.py:499-607: plane-prior attention + reflect-padded top-down depth decoder."""

    def __init__(self, cfg=None):
        super().__init__()
        if cfg is None:
            from .config import cfg as _cfg
            cfg = _cfg
        self.num_output_channels = 1
        self.num_kernels = cfg.solov2.num_kernels
        self.num_grids = list(cfg.solov2.num_grids)
        self.channels_kernels_flatten = sum(g * g for g in cfg.solov2.num_grids)

        self.latlayer1 = nn.Conv2d(2048, 256, kernel_size=1, stride=1, padding=0)
        self.latlayer2 = nn.Conv2d(1024, 256, kernel_size=1, stride=1, padding=0)
        self.latlayer3 = nn.Conv2d(512, 256, kernel_size=1, stride=1, padding=0)
        self.latlayer4 = nn.Conv2d(256, 256, kernel_size=1, stride=1, padding=0)
        self.conv1 = _reflect_conv_bn_relu(256, 256)
        self.conv2 = _reflect_conv_bn_relu(256, 128)
        self.conv3 = _reflect_conv_bn_relu(256, 128)
        self.conv4 = _reflect_conv_bn_relu(256, 128)
        self.deconv1 = _reflect_conv_bn_relu(256, 256, upsample=True)
        self.deconv2 = _reflect_conv_bn_relu(256, 128, upsample=True)
        self.deconv3 = _reflect_conv_bn_relu(256, 128, upsample=True)
        self.deconv4 = _reflect_conv_bn_relu(256, 64, upsample=True)
        self.depth_pred = nn.Sequential(nn.ReflectionPad2d(1),
                                        nn.Conv2d(64, self.num_output_channels, kernel_size=3, stride=1, padding=0),
                                        nn.Softplus())
        self.conv1x1 = nn.Sequential(nn.Conv2d(self.channels_kernels_flatten, 256, kernel_size=1, stride=1, padding=0))
        self.refine_conv = _reflect_conv_bn_relu(512, 128)

    def forward(self, feature_maps, seg_preds, kernel_preds):
        """feature_maps (C2..C5), seg_preds [B,128,H/4,W/4], kernel_preds [[B,128,S,S]]x4, all NCHW fp32."""
        from .engine import engine_for
        eng = engine_for(self)
        cs = [eng.to_nhwc(f) for f in feature_maps]
        mask = eng.to_nhwc(seg_preds)
        B = seg_preds.shape[0]
        kern = eng.pack_kernel_preds(kernel_preds, B)
        d = eng.depth_decoder(cs, mask, kern, self)
        return eng.depth_output_nchw(d, B, feature_maps[0].shape[2] * 2, feature_maps[0].shape[3] * 2)
