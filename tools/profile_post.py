import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from planerecnet_b200.config import cfg, set_cfg
from planerecnet_b200.planerecnet import PlaneRecNet
from planerecnet_b200.utils.synth import make_input, perturb_
set_cfg("PlaneRecNet_101_config"); torch.manual_seed(0)
net = perturb_(PlaneRecNet(cfg)).eval().cuda(); eng = net.engine
xd = make_input(8, 480, 640, 0).cuda()
with torch.no_grad():
    st = eng.forward_dense_graph(net, xd, False)
    for _ in range(2): eng.inference(net, st, xd)
    torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    eng.inference(net, st, xd)
    torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
