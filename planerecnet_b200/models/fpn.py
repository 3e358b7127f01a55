"""Bottom-up FPN variant of the reference (models/fpn.py:7-73) — parameter container; the lateral
1x1 + 2x2-average running sum + 3x3/ReLU run in libprn_b200."""
from torch import nn


class FPN(nn.Module):
    def __init__(self, in_channels, start_level=0, cfg=None):
        super().__init__()
        assert isinstance(in_channels, list)
        if cfg is None:
            from ..config import cfg as _cfg
            cfg = _cfg
        self.in_channels = in_channels
        self.out_channels = cfg.fpn.num_features
        self.num_ins = len(in_channels)
        self.backbone_end_level = self.num_ins
        self.start_level = start_level
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(nn.Conv2d(in_channels[i], self.out_channels, kernel_size=1))
        for _ in range(self.start_level, self.backbone_end_level):
            self.fpn_convs.append(nn.Conv2d(self.out_channels, self.out_channels, kernel_size=3, padding=1))
        if cfg.fpn.high_level_mode is not None:
            raise NotImplementedError("fpn.high_level_mode is None in the PlaneRecNet_50/101 presets (data/config.py:500)")
        self.interpolation_mode = cfg.fpn.interpolation_mode
        self.relu_pred_layers = cfg.fpn.relu_pred_layers
        self.high_level_mode = cfg.fpn.high_level_mode
        if self.interpolation_mode != "bilinear" or not self.relu_pred_layers:
            raise NotImplementedError("only the presets' bilinear / relu_pred_layers FPN is implemented")

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        from ..engine import engine_for
        eng = engine_for(self)
        outs = eng.fpn([eng.to_nhwc(x) for x in inputs], self)
        return [eng.to_nchw(o, self.out_channels) for o in outs]
