"""One eager dense forward bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402
from planerecnet_b200.planerecnet import PlaneRecNet  # noqa: E402
from planerecnet_b200.utils.synth import make_input, perturb_  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else "PlaneRecNet_101_config"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prec = sys.argv[3] if len(sys.argv) > 3 else "f16"
set_cfg(preset)
torch.manual_seed(0)
net = perturb_(PlaneRecNet(cfg)).eval().cuda().set_precision(prec)
x = make_input(B, 480, 640, 0).cuda()
with torch.no_grad():
    for _ in range(2):
        net.engine.forward_dense(net, x, False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net.engine.forward_dense(net, x, False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("profiled one step; launches:", net.engine.launches // 3)
