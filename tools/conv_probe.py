"""Time the conv kernel on real PlaneRecNet layer shapes (bs=8, 480x640) with CUDA events and print the
role-level stall counters of CTA 0.  Usage: python tools/conv_probe.py [name ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200 import _lib as L  # noqa: E402
from planerecnet_b200 import ops  # noqa: E402

# name: (B, H, W, C, N, k, stride, extra)
SHAPES = {
    "l0_3x3_64": (8, 120, 160, 64, 64, 3, 1, {}),
    "l0_1x1_64_256": (8, 120, 160, 64, 256, 1, 1, {}),
    "l0_1x1_256_64": (8, 120, 160, 256, 64, 1, 1, {}),
    "l1_3x3_128": (8, 60, 80, 128, 128, 3, 1, {}),
    "l1_1x1_128_512": (8, 60, 80, 128, 512, 1, 1, {}),
    "l2_3x3_256": (8, 30, 40, 256, 256, 3, 1, {}),
    "l2_1x1_256_1024": (8, 30, 40, 256, 1024, 1, 1, {}),
    "l2_1x1_1024_256": (8, 30, 40, 1024, 256, 1, 1, {}),
    "l3_3x3_512": (8, 15, 20, 512, 512, 3, 1, {}),
    "l3_1x1_2048_512": (8, 15, 20, 2048, 512, 1, 1, {}),
    "fpn0_3x3_256": (8, 120, 160, 256, 256, 3, 1, {}),
    "mask0_3x3_256_128": (8, 120, 160, 256, 128, 3, 1, {}),
    "deconv4_up_256_64": (8, 120, 160, 256, 64, 3, 1, {"up": 2, "reflect": True}),
    "depth_pred_64_1": (8, 240, 320, 64, 1, 3, 1, {"reflect": True, "out32": True}),
    "stem_k192_64": (8, 240, 320, 192, 64, 1, 1, {}),
    "inst_3x3_256_s40": (8, 40, 40, 256, 256, 3, 1, {}),
    "ppa_dyn": (8, 1200, 4, 128, 3728, 1, 1, {"grouped": True, "act": L.ACT_SIGMOID_AVG4}),
    "ppa_1x1_3776_256": (8, 30, 40, 3776, 256, 1, 1, {}),
    "mask2_3x3_256_128": (8, 30, 40, 256, 128, 3, 1, {}),
    "deconv4_subpixel_256_4x64": (8, 120, 160, 256, 256, 3, 1, {"clamp": True, "shuffle": 64}),
    "deconv3_subpixel_256_4x128": (8, 60, 80, 256, 512, 3, 1, {"clamp": True, "shuffle": 128}),
    "conv4_reflect_256_128": (8, 120, 160, 256, 128, 3, 1, {"reflect": True}),
    "fpn1_3x3_256": (8, 60, 80, 256, 256, 3, 1, {}),
    "inst_3x3_256_s16": (8, 16, 16, 256, 256, 3, 1, {}),
    "l1_1x1_512_128": (8, 60, 80, 512, 128, 1, 1, {}),
    "l3_1x1_512_2048": (8, 15, 20, 512, 2048, 1, 1, {}),
    "dcn_l2_256": (8, 30, 40, 256, 256, 3, 1, {"dcn": True}),
    "dcn_l1_128": (8, 60, 80, 128, 128, 3, 1, {"dcn": True}),
    "dcn_l3_512": (8, 15, 20, 512, 512, 3, 1, {"dcn": True}),
}


def probe(name, dt=L.PRN_F16, reps=20):
    B, H, W, Cc, N, k, stride, ex = SHAPES[name]
    tdt = ops.torch_dtype(dt)
    up = ex.get("up", 1)
    pad = k // 2
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, W, Cc, device="cuda", generator=g).to(tdt)
    Ho = (H * up + 2 * pad - k) // stride + 1
    Wo = (W * up + 2 * pad - k) // stride + 1
    M = B * Ho * Wo
    n_pad = ops.round_up(N, 16)
    grouped = ex.get("grouped", False)
    if grouped:
        w = (torch.randn(B * N, Cc, device="cuda", generator=g) * 0.1).to(tdt)
    else:
        w = (torch.randn(n_pad, k * k * Cc, device="cuda", generator=g) * 0.05).to(tdt)
    act = ex.get("act", L.ACT_RELU)
    m_out = M // 4 if act == L.ACT_SIGMOID_AVG4 else M
    ld = ops.round_up(n_pad, 64) if grouped else n_pad
    shuffle = ex.get("shuffle", 0)
    if shuffle:
        m_out, ld = 4 * M, shuffle
    out16 = None if ex.get("out32") else torch.empty(m_out, ld, device="cuda", dtype=tdt)
    out32 = torch.empty(m_out, n_pad, device="cuda") if ex.get("out32") else None
    bias = torch.zeros(n_pad, device="cuda")
    om = None
    if ex.get("dcn"):
        om = torch.zeros(M, 32, device="cuda")
        om[:, :18] = torch.randn(M, 18, device="cuda", generator=g) * 2
        om[:, 18:27] = torch.rand(M, 9, device="cuda", generator=g) * 2
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda")

    def run(counters=None):
        ops.conv2d(x, w, batch=B, h_in=H, w_in=W, ksize=k, stride=stride, pad=pad,
                   pad_mode=L.PAD_REFLECT if ex.get("reflect") else (L.PAD_CLAMP if ex.get("clamp") else L.PAD_ZERO),
                   upsample=up, bias=bias, act=act, shuffle_n=shuffle,
                   out16=out16, out32=out32, ld_out16=ld if out16 is not None else None, dcn_offmask=om,
                   n_pad=N if grouped else n_pad, w_group_rows=N if grouped else 0, dtype=dt, counters=counters,
                   out_img_rows=(Ho * Wo // 4 if act == L.ACT_SIGMOID_AVG4 else 0))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    run(cnt)
    torch.cuda.synchronize()
    c = cnt.tolist()
    fl = 2.0 * M * N * k * k * Cc
    nt = C_int(); st = C_int(); gr = C_int()
    d = L.PrnConv()
    tfl = fl / us / 1e6
    kb = max(1, c[3])
    print(f"{name:20s} M={M:7d} N={N:5d} K={k * k * Cc:5d} {us:8.1f} us {tfl:7.1f} TFLOP/s | "
          f"prod tot={c[0]:9d} wEmpty={c[1]:9d} wCp={c[2]:9d} kb={c[3]:5d} ({c[0] // kb} cyc/kb) | "
          f"mma tot={c[4]:9d} wFull={c[5]:9d} wTmemE={c[6]:8d} wB={c[9]:8d} | epi tot={c[7]:9d} wTfull={c[8]:9d} wLd={c[10]:7d} wStore={c[11]:7d} ldIssue={c[12]:7d} chunk={c[13]:7d} stIssue={c[14]:7d}", flush=True)


def C_int():
    import ctypes
    return ctypes.c_int32(0)


if __name__ == "__main__":
    names = sys.argv[1:] or list(SHAPES)
    for n in names:
        try:
            probe(n)
        except Exception as e:  # keep going: one bad launch configuration should not hide the others
            print(f"{n:20s} FAILED: {str(e)[-160:]}", flush=True)
