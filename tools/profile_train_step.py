"""One replayed training step (GraphedStep forward + backward, R101 bs 8 480x640, bf16) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off --graph-profiling node -k regex:<kernel> ...` captures of the training-side kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402
from planerecnet_b200.planerecnet import PlaneRecNet  # noqa: E402
from planerecnet_b200.train_engine import GraphedStep  # noqa: E402
from planerecnet_b200.utils.synth import make_cotangents, make_input, perturb_  # noqa: E402

set_cfg("PlaneRecNet_101_config")
torch.manual_seed(0)
net = perturb_(PlaneRecNet(cfg)).train().cuda()
x = make_input(8, 480, 640, 0).cuda()
step = GraphedStep(net.train_engine, net, x, flat_grads=True)
cots = make_cotangents(step.outs, seed=1, device="cuda")
for _ in range(2):
    step.forward(x)
    step.backward(*cots)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step.forward(x)
step.backward(*cots)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one replayed training step;", step.fwd_launches + step.bwd_launches, "launches")
