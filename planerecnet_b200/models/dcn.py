"""Modulated deformable convolution — parameter container with the reference's names
(models/dcn.py:12-50: offset_conv, modulator_conv, regular_conv).  The arithmetic
(models/dcn.py:52-67 + torchvision.ops.deform_conv2d) runs in libprn_b200: one fused 27-channel
offset/modulator conv (clamp, 2*sigmoid in the epilogue) and a bilinear-gather + tcgen05 GEMM."""
from torch import nn


class DeformableConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False):
        super().__init__()
        self.padding = padding
        self.stride = stride
        taps = kernel_size * kernel_size
        # same construction order as the reference so that a seeded default init reproduces its weights
        self.offset_conv = nn.Conv2d(in_channels, 2 * taps, kernel_size, stride, padding, bias=True)
        nn.init.zeros_(self.offset_conv.weight)
        nn.init.zeros_(self.offset_conv.bias)
        self.modulator_conv = nn.Conv2d(in_channels, taps, kernel_size, stride, padding, bias=True)
        nn.init.zeros_(self.modulator_conv.weight)
        nn.init.zeros_(self.modulator_conv.bias)
        self.regular_conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)

    def forward(self, x):
        from ..engine import engine_for
        eng = engine_for(self)
        y = eng.dcn(eng.to_nhwc(x), self, bn=None, relu=False)
        return eng.to_nchw(y, self.regular_conv.out_channels)
