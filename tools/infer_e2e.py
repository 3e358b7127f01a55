"""The serving loop of bench.py's `inference.e2e` in isolation (net.infer_pipelined from pinned host batches, D2H of the results),
with the device-to-host part switchable, to see where its time goes.  Usage: python tools/infer_e2e.py [none|fields|all] [depth]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402
from planerecnet_b200.planerecnet import PlaneRecNet  # noqa: E402
from planerecnet_b200.postprocess import pack_mask_bits  # noqa: E402
from planerecnet_b200.utils.synth import make_input, perturb_  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
set_cfg("PlaneRecNet_101_config")
torch.manual_seed(0)
net = perturb_(PlaneRecNet(cfg)).eval().cuda()
x_host = make_input(8, 480, 640, 0).pin_memory()
pinned = {}


def d2h(res):
    if mode == "none":
        torch.cuda.current_stream().synchronize()
        return
    fields = {k: [r[k] for r in res if r[k] is not None] for k in ("pred_scores", "pred_classes", "pred_boxes", "pred_depth")}
    if mode == "all":
        packed = pack_mask_bits([r["pred_masks"] for r in res])
        if packed is not None:
            fields["pred_masks_packed"] = [packed]
    for k, parts in fields.items():
        if parts:
            t = torch.cat(parts) if len(parts) > 1 else parts[0]
            buf = pinned.get(k)
            if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                buf = pinned[k] = torch.empty(max(int(t.numel() * 1.5), 4096), dtype=t.dtype).pin_memory()
            buf[:t.numel()].copy_(t.reshape(-1), non_blocking=True)
    torch.cuda.current_stream().synchronize()


warm, steps = 5, 10
t0 = t1 = None
for i, res in enumerate(net.infer_pipelined((x_host for _ in range(warm + steps + 3)), depth=depth)):
    d2h(res)
    if i == warm - 1:
        t0 = time.perf_counter()
    if i == warm + steps - 1:
        t1 = time.perf_counter()
print(f"INFER_E2E mode={mode} depth={depth} PRN_CONV_TMA={os.environ.get('PRN_CONV_TMA', '1')} PRN_PDL={os.environ.get('PRN_PDL', '1')}: "
      f"{(t1 - t0) / steps * 1e3:.2f} ms per batch = {8 * steps / (t1 - t0):.1f} img/s")
