"""Time the weight-gradient kernel (and the input-gradient contraction) on real PlaneRecNet layer shapes (bs 8, 480x640)
with CUDA events; run under `ncu -k regex:wgrad_umma --set full` for the full capture.
Usage: python tools/wgrad_probe.py [name ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200 import _lib as L  # noqa: E402
from planerecnet_b200 import ops  # noqa: E402
import ctypes as C  # noqa: E402

# name: (B, H, W, Cin, Cout, k, extra)
SHAPES = {
    "fpn0_3x3_256": (8, 120, 160, 256, 256, 3, {}),
    "l0_1x1_64_256": (8, 120, 160, 64, 256, 1, {}),
    "l0_3x3_64": (8, 120, 160, 64, 64, 3, {}),
    "l1_3x3_128": (8, 60, 80, 128, 128, 3, {}),
    "l2_3x3_256": (8, 30, 40, 256, 256, 3, {}),
    "l2_1x1_256_1024": (8, 30, 40, 256, 1024, 1, {}),
    "l2_1x1_1024_256": (8, 30, 40, 1024, 256, 1, {}),
    "l3_1x1_2048_512": (8, 15, 20, 2048, 512, 1, {}),
    "deconv4_up_256_64": (8, 120, 160, 256, 64, 3, {"up": 2, "reflect": True}),
    "mask0_3x3_256_128": (8, 120, 160, 256, 128, 3, {}),
}


def probe(name, dt=L.PRN_BF16, reps=10):
    B, H, W, Cc, N, k, ex = SHAPES[name]
    tdt = ops.torch_dtype(dt)
    up, pad = ex.get("up", 1), k // 2
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, W, Cc, device="cuda", generator=g).to(tdt)
    Ho, Wo = H * up, W * up
    M = B * Ho * Wo
    dy = torch.randn(M, N, device="cuda", generator=g).to(tdt)
    dw = torch.zeros(N, k * k * Cc, device="cuda")
    kw = dict(batch=B, h_in=H, w_in=W, n=N, ksize=k, pad=pad, pad_mode=L.PAD_REFLECT if ex.get("reflect") else L.PAD_ZERO,
              upsample=up, dtype=dt)
    d = ops.conv2d_wgrad(x, dy, dw, **kw)
    out8 = (C.c_int32 * 8)()
    L.lib().prn_conv2d_wgrad_plan(C.byref(d), out8)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()                       # > L2: operands come from HBM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d_wgrad(x, dy, dw, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    flops = 2.0 * M * N * Cc * k * k
    return {"name": name, "M": M, "N": N, "K": Cc * k * k, "us": round(ms * 1e3, 1), "tflops": round(flops / ms / 1e9, 1),
            "plan(m_tiles,n_tiles,splits,kb/split,m_sub,stages,grid,atoms)": list(out8)}


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(SHAPES)):
        print("WGRAD_PROBE " + json.dumps(probe(n)), flush=True)
