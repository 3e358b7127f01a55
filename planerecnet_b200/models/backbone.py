"""ResNet-50/101 backbone with deformable 3x3s — parameter containers carrying the reference's
attribute names, constructor signatures and state_dict keys (models/backbone.py:5-243).  `forward`
hands the whole stage to the sm_100a engine (no torch operator is used for the arithmetic)."""
import torch
from torch import nn

from .dcn import DeformableConv2d


class Bottleneck(nn.Module):
    """1x1 -> 3x3 (stride here; optionally deformable) -> 1x1 (+ projection skip); models/backbone.py:5-73."""

    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm_layer=nn.BatchNorm2d, dilation=1,
                 use_dcn=False):
        super().__init__()
        if dilation != 1:
            raise NotImplementedError("atrous bottlenecks are not used by the PlaneRecNet presets")
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = norm_layer(planes)
        if use_dcn:
            self.conv2 = DeformableConv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=True)
        else:
            self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = norm_layer(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = norm_layer(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        from ..engine import engine_for
        eng = engine_for(self)
        y = eng.bottleneck(eng.to_nhwc(x), self)
        return eng.to_nchw(y, self.conv3.out_channels)


class ResNetBackbone(nn.Module):
    """models/backbone.py:76-230.  Same signature: (layers, dcn_layers, dcn_interval, atrous_layers, block, norm_layer)."""

    def __init__(self, layers, dcn_layers=[0, 0, 0, 0], dcn_interval=1, atrous_layers=[], block=Bottleneck,
                 norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.num_base_layers = len(layers)
        self.layers = nn.ModuleList()     # registered before conv1: state_dict order of the reference
        self.channels = []
        self.norm_layer = norm_layer
        self.dilation = 1
        self.atrous_layers = atrous_layers
        self.inplanes = 64

        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = norm_layer(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

        for i, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2))):
            self._make_layer(block, planes, layers[i], stride=stride, dcn_layers=dcn_layers[i],
                             dcn_interval=dcn_interval)

        # convs that a pretrained checkpoint initialises (PlaneRecNet.init_weights leaves them alone)
        self.backbone_modules = [m for m in self.modules() if isinstance(m, nn.Conv2d)]

    def _make_layer(self, block, planes, blocks, stride=1, dcn_layers=0, dcn_interval=1):
        """One stage = `blocks` bottlenecks; DCN placement rule of models/backbone.py:170,184."""
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            if len(self.layers) in self.atrous_layers:
                raise NotImplementedError("atrous stages are not used by the PlaneRecNet presets")
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                self.norm_layer(planes * block.expansion),
            )
        stage = [block(self.inplanes, planes, stride, downsample, self.norm_layer, self.dilation,
                       use_dcn=dcn_layers >= blocks)]
        self.inplanes = planes * block.expansion
        for i in range(1, blocks):
            use_dcn = ((i + dcn_layers) >= blocks) and (i % dcn_interval == 0)
            stage.append(block(self.inplanes, planes, norm_layer=self.norm_layer, use_dcn=use_dcn))
        layer = nn.Sequential(*stage)
        self.channels.append(planes * block.expansion)
        self.layers.append(layer)
        return layer

    def forward(self, x):
        """Returns the tuple (C2..C5) as NCHW fp32, like models/backbone.py:197-209."""
        from ..engine import engine_for
        eng = engine_for(self)
        outs = eng.backbone(x, self)
        return tuple(eng.to_nchw(o, c) for o, c in zip(outs, self.channels))

    def init_backbone(self, path):
        """Load a torchvision-style ResNet checkpoint, renaming layerN.* -> layers.(N-1).* (models/backbone.py:211-224)."""
        state_dict = torch.load(path)
        for key in list(state_dict):
            if key.startswith("layer"):
                idx = int(key[5])
                state_dict["layers." + str(idx - 1) + key[6:]] = state_dict.pop(key)
        self.load_state_dict(state_dict, strict=False)

    def add_layer(self, conv_channels=1024, downsample=2, depth=1, block=Bottleneck):
        self._make_layer(block, conv_channels // block.expansion, blocks=depth, stride=downsample)


def construct_backbone(cfg):
    """models/backbone.py:233-243."""
    backbone = cfg.type(*cfg.args)
    num_layers = max(cfg.selected_layers) + 1
    while len(backbone.layers) < num_layers:
        backbone.add_layer()
    return backbone
