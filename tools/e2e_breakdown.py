"""Where the end-to-end step (net(x) from pinned host memory) spends its time."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402
from planerecnet_b200.planerecnet import PlaneRecNet  # noqa: E402
from planerecnet_b200.utils.synth import make_input, perturb_  # noqa: E402

set_cfg("PlaneRecNet_101_config")
torch.manual_seed(0)
net = perturb_(PlaneRecNet(cfg)).eval().cuda()
eng = net.engine
x_host = make_input(8, 480, 640, 0).pin_memory()


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


with torch.no_grad():
    xd = x_host.cuda()
    t_h2d, _ = timed(lambda: x_host.cuda(non_blocking=True))
    t_fwd, st = timed(lambda: eng.forward_dense_graph(net, xd, False))
    t_post, res = timed(lambda: eng.inference(net, st, xd))

    def d2h():
        for r in res:
            for k in ("pred_scores", "pred_classes", "pred_boxes", "pred_depth"):
                if r[k] is not None:
                    r[k].cpu()
    t_d2h, _ = timed(d2h)
    from planerecnet_b200.postprocess import pack_mask_bits
    t_pack, packed = timed(lambda: pack_mask_bits([r["pred_masks"] for r in res]))
    t_pack_d2h, _ = timed(lambda: packed.cpu())
    print(f"pack masks {t_pack:.3f} ms ({packed.numel()} bytes) | packed d2h (pageable) {t_pack_d2h:.3f} ms")
    it = net.infer_pipelined(x_host for _ in range(12))
    t0 = None
    for i, r_ in enumerate(it):
        torch.cuda.synchronize()
        if i == 3:
            t0 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"infer_pipelined, results consumed with a full sync each: {(time.perf_counter() - t0) / 8 * 1e3:.2f} ms per batch")
    t_all, _ = timed(lambda: net(x_host.cuda(non_blocking=True)))
print(f"h2d {t_h2d:.2f} ms | dense graph {t_fwd:.2f} ms | bookkeeping {t_post:.2f} ms | d2h {t_d2h:.2f} ms | net(x) {t_all:.2f} ms")
print("detections per image:", [0 if r["pred_scores"] is None else len(r["pred_scores"]) for r in res])
import cProfile, pstats
pr = cProfile.Profile()
with torch.no_grad():
    pr.enable()
    for _ in range(5):
        eng.inference(net, st, xd)
    torch.cuda.synchronize()
    pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
