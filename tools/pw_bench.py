"""Achieved HBM bandwidth of the training-step pointwise / reduction passes on the largest layer shapes
(algorithmic bytes = operands read + written, each once)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200 import _lib as L, ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def main():
    dt, tdt = L.PRN_BF16, torch.bfloat16
    res = {}
    for rows, Cc in ((153600, 256), (153600, 64), (38400, 512), (9600, 1024), (614400, 64)):
        # several independent operand sets so that consecutive launches do not hit L2 (126 MB)
        nset = max(2, int(400e6 // (rows * Cc * 2 * 5)) + 1)
        sets = []
        for _ in range(nset):
            x = torch.randn(rows, Cc, device="cuda").to(tdt)
            sets.append(dict(x=x, out=torch.empty_like(x), res=torch.randn(rows, Cc, device="cuda").to(tdt), dz=torch.randn(rows, Cc, device="cuda").to(tdt),
                             dx=torch.empty_like(x), g=torch.empty_like(x)))
        mi = torch.stack([torch.zeros(Cc), torch.ones(Cc)], 1).contiguous().cuda()
        gamma, beta = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
        sums = torch.zeros(Cc, 2, device="cuda")
        k = [0]

        def nxt():
            k[0] = (k[0] + 1) % nset
            return sets[k[0]]

        bytes1 = rows * Cc * 2
        t = timeit(lambda: (lambda s: ops.bn_apply(s["x"], s["out"], mi, gamma, beta, s["res"], True, dt))(nxt()))
        res[f"bn_apply_res {rows}x{Cc}"] = round(3 * bytes1 / t / 1e9)
        t = timeit(lambda: (lambda s: ops.chan_reduce(s["dz"], s["out"], s["x"], mi, sums, dt))(nxt()))
        res[f"chan_reduce {rows}x{Cc}"] = round(3 * bytes1 / t / 1e9)
        t = timeit(lambda: (lambda s: ops.bn_bwd_apply(s["dz"], s["out"], s["x"], mi, gamma, sums, s["dx"], s["g"], dt))(nxt()))
        res[f"bn_bwd_apply_g {rows}x{Cc}"] = round(5 * bytes1 / t / 1e9)
        t = timeit(lambda: (lambda s: ops.bn_bwd_apply(s["dz"], s["out"], s["x"], mi, gamma, sums, s["dx"], None, dt))(nxt()))
        res[f"bn_bwd_apply {rows}x{Cc}"] = round(4 * bytes1 / t / 1e9)
        del sets
    print("PW_BENCH GB/s " + json.dumps(res))


if __name__ == "__main__":
    main()
