"""One launch each of the training step's HBM-bound BatchNorm passes on the largest layer shape (153600 x 256, bf16) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200 import _lib as L, ops  # noqa: E402

rows, Cc = 153600, 256
dt, tdt = L.PRN_BF16, torch.bfloat16
mk = lambda: torch.randn(rows, Cc, device="cuda").to(tdt)  # noqa: E731
x, res, dz = mk(), mk(), mk()
out, dx, g = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
mi = torch.stack([torch.zeros(Cc), torch.ones(Cc)], 1).contiguous().cuda()
gamma, beta = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
sums = torch.zeros(Cc, 2, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    flush.zero_()
    ops.bn_apply(x, out, mi, gamma, beta, res, True, dt)
    flush.zero_()
    ops.chan_reduce(dz, out, x, mi, sums, dt)
    flush.zero_()
    ops.bn_bwd_apply(dz, out, x, mi, gamma, sums, dx, g, dt)


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("algorithmic bytes per launch: bn_apply 3 x 78.6 MB, chan_reduce 3 x 78.6 MB, bn_bwd_apply 5 x 78.6 MB")
