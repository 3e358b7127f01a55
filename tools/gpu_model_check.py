"""Whole-model parity report on the GPU box: our sm_100a forward vs the CPU oracle, stage by stage
(rel-L2 and max-abs/max-ref), plus detections.  Logs to gpurun_out/model_check.log."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from oracle import prn_oracle as O  # noqa: E402

LOG = os.path.join(ROOT, "gpurun_out", "model_check.log")


def log(msg):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    with open(LOG, "a") as fh:
        fh.write(msg + "\n")
    print(msg, flush=True)


def nhwc_to_nchw(t, c):
    return t[..., :c].float().permute(0, 3, 1, 2).cpu()


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "PlaneRecNet_50_config"
    B, Hh, Ww = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (2, 192, 256)))
    prec = sys.argv[5] if len(sys.argv) > 5 else "bf16"
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)
    net = H.perturb_(H.build_ours(preset, 0)).eval()
    x = H.make_input(B, Hh, Ww, 0)
    t0 = time.time()
    orc = O.Oracle(net.state_dict(), preset)
    with torch.no_grad():
        omask, ocate, okern, odepth = orc.forward_dense(x)
        ores = orc.forward_eval(x)
    log(f"== {preset} B={B} {Hh}x{Ww} {prec}: oracle done in {time.time() - t0:.1f}s")
    net = net.cuda()
    net.set_precision(prec)
    eng = net.engine
    xc = x.cuda()
    with torch.no_grad():
        st = eng.forward_dense(net, xc)
    torch.cuda.synchronize()
    log(f"   device forward done; launches={eng.launches}")
    mask, cates, kerns, depth = st["outputs"]
    rows = []
    for i in range(4):
        rows.append((f"C{i + 2}", nhwc_to_nchw(st["cs"][i], orc.taps["cs"][i].shape[1]), orc.taps["cs"][i]))
    for i in range(4):
        rows.append((f"P{i + 2}", nhwc_to_nchw(st["ps"][i], 256), orc.taps["ps"][i]))
    for i in range(4):
        rows.append((f"cate{i}", cates[i].cpu(), ocate[i]))
        rows.append((f"kern{i}", kerns[i].cpu(), okern[i]))
    rows.append(("mask", mask.cpu(), omask))
    rows.append(("attn", nhwc_to_nchw(st["attn"], 256), orc.taps["ppa_attn"]))
    rows.append(("depth", depth.cpu(), odepth))
    worst = 0.0
    for name, got, ref in rows:
        ok = torch.isfinite(got).all().item()
        e, m = H.rel_l2(got, ref), H.max_rel(got, ref)
        worst = max(worst, e)
        log(f"   {name:7s} shape={list(ref.shape)} rel_l2={e:.3e} max_rel={m:.3e} finite={ok}")
    log(f"   worst rel_l2 = {worst:.3e}")
    with torch.no_grad():
        res = eng.inference(net, st, xc)
    torch.cuda.synchronize()
    for b, (r, o) in enumerate(zip(res, ores)):
        n_r = 0 if r["pred_scores"] is None else len(r["pred_scores"])
        n_o = 0 if o["pred_scores"] is None else len(o["pred_scores"])
        msg = f"   img{b}: detections ours={n_r} oracle={n_o}"
        if n_r and n_o:
            k = min(n_r, n_o)
            msg += f" top-score ours={float(r['pred_scores'][0]):.4f} oracle={float(o['pred_scores'][0]):.4f}"
            msg += f" |dscore|max(first {k})={float((r['pred_scores'][:k].cpu() - o['pred_scores'][:k]).abs().max()):.4f}"
            iou = (r["pred_masks"][0].cpu() & o["pred_masks"][0]).sum().item() / max(1, (r["pred_masks"][0].cpu() | o["pred_masks"][0]).sum().item())
            msg += f" top-mask IoU={iou:.3f}"
        msg += f" depth rel_l2={H.rel_l2(r['pred_depth'].cpu(), o['pred_depth']):.3e}"
        log(msg)
    # timing (no graph): 5 runs
    for _ in range(2):
        eng.forward_dense(net, xc, want_nchw=False)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        eng.forward_dense(net, xc, want_nchw=False)
    torch.cuda.synchronize()
    log(f"   eager dense forward: {(time.time() - t0) / 5 * 1e3:.2f} ms / batch of {B}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
