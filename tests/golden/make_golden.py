"""Generate the golden fixtures from the UNMODIFIED reference (imported from /root/reference, CPU fp32)
and pin the oracle to it.  Run in the build container only (the reference does not travel to the GPU
box):  python tests/golden/make_golden.py

For every case: seed -> reference PlaneRecNet(cfg) -> tests.helpers.perturb_ -> eval forward on a seeded
input.  Writes tests/golden/<case>.pt holding (a) fixed pseudo-random samples + moments of every stage
tensor, (b) the full small outputs (category logits, detections), and asserts that
oracle/prn_oracle.py reproduces every stage of the reference to <= 2e-4 rel-L2 before saving."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

CASES = {
    # name: (preset, B, H, W)
    "r50_b2_192x256": ("PlaneRecNet_50_config", 2, 192, 256),
    "r101_b1_192x256": ("PlaneRecNet_101_config", 1, 192, 256),
    "r50_b1_480x640": ("PlaneRecNet_50_config", 1, 480, 640),
    "r101_b1_480x640": ("PlaneRecNet_101_config", 1, 480, 640),      # the benchmarked preset at the benchmarked resolution
}
NSAMPLE = 4096


def sample_idx(numel, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(NSAMPLE, numel),), generator=g)


def summarize(t):
    t = t.detach().float().contiguous()
    flat = t.flatten()
    idx = sample_idx(flat.numel())
    return {"shape": list(t.shape), "mean": float(flat.mean()), "std": float(flat.std()), "absmax": float(flat.abs().max()),
            "l2": float(flat.double().norm()), "samples": flat[idx].clone()}


def main():
    sys.path.insert(0, REF)
    sys.path.insert(1, os.path.join(ROOT, "tests"))
    torch.cuda.current_device = lambda: 0            # planerecnet.py:18 touches CUDA at import
    import planerecnet as ref_mod                     # the reference module
    from data.config import cfg, set_cfg
    from utils import timer
    timer.disable_all()
    sys.path.append(ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("prn_oracle", os.path.join(ROOT, "oracle", "prn_oracle.py"))
    O = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(O)
    spec = importlib.util.spec_from_file_location("t_helpers", os.path.join(ROOT, "tests", "helpers.py"))
    Hh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(Hh)

    only = sys.argv[1:]
    for name, (preset, B, H, W) in CASES.items():
        if only and name not in only:
            continue
        set_cfg(preset)
        torch.manual_seed(0)
        net = ref_mod.PlaneRecNet(cfg)
        Hh.perturb_(net)
        net.eval()
        x = Hh.make_input(B, H, W, seed=0)
        taps = {}
        hooks = [net.depth_decoder.conv1x1.register_forward_hook(lambda m, i, o: taps.__setitem__("ppa_conv1x1", o))]
        with torch.no_grad():
            cs = net.backbone(x)
            ps = net.fpn([cs[i] for i in net.fpn_indices])
            cate, kern = net.inst_head(net.split_feats(ps))
            mask = net.mask_head(ps)
            depth = net.depth_decoder([cs[i] for i in net.depth_decoder_indices], mask, kern)
            attn = torch.nn.functional.interpolate(taps["ppa_conv1x1"], scale_factor=0.25, mode="bilinear",
                                                   align_corners=False, recompute_scale_factor=False)
            results = net(x)
        for h in hooks:
            h.remove()
        # ---- pin the oracle against the reference, stage by stage
        orc = O.Oracle(net.state_dict(), preset)
        with torch.no_grad():
            omask, ocate, okern, odepth = orc.forward_dense(x)
            ores = orc.forward_eval(x)
        checks = {"C%d" % (i + 2): (orc.taps["cs"][i], cs[i]) for i in range(4)}
        checks.update({"P%d" % (i + 2): (orc.taps["ps"][i], ps[i]) for i in range(4)})
        checks.update({"cate%d" % i: (ocate[i], cate[i]) for i in range(4)})
        checks.update({"kern%d" % i: (okern[i], kern[i]) for i in range(4)})
        checks.update(mask=(omask, mask), attn=(orc.taps["ppa_attn"], attn), depth=(odepth, depth))
        worst = 0.0
        for k, (a, b) in checks.items():
            e = Hh.rel_l2(a, b)
            worst = max(worst, e)
            assert e <= 2e-4, f"{name}: oracle != reference at {k}: rel-L2 {e:.3g}"
        for b in range(B):
            r, o = results[b], ores[b]
            assert list(r.keys()) == list(o.keys())
            if r["pred_scores"] is None:
                assert o["pred_scores"] is None
                continue
            assert r["pred_scores"].shape == o["pred_scores"].shape, f"{name}: detection count differs"
            assert torch.equal(r["pred_classes"], o["pred_classes"])
            d = (r["pred_scores"] - o["pred_scores"]).abs().max().item()
            print("   img", b, "n", len(r["pred_scores"]), "max score diff", d, "score range", r["pred_scores"].min().item(), r["pred_scores"].max().item())
            assert d <= 2e-3, f"{name}: detection scores differ by {d}"
        print(f"{name}: oracle pinned to reference, worst stage rel-L2 {worst:.3g}; detections per image:",
              [0 if r["pred_scores"] is None else len(r["pred_scores"]) for r in results])
        gold = {"preset": preset, "B": B, "H": H, "W": W, "seed": 0, "n_state": len(net.state_dict()),
                "state_l2": float(sum(v.double().norm() ** 2 for v in net.state_dict().values() if v.is_floating_point()) ** 0.5),
                "stages": {k: summarize(b) for k, (a, b) in checks.items()},
                "cate_full": [c.clone() for c in cate],
                "depth_ds4": depth[:, :, ::4, ::4].clone(),
                "results": [{k: (None if v is None else (v.clone() if k != "pred_masks" else v.flatten(1).sum(1)))
                             for k, v in r.items() if k != "pred_depth"} for r in results]}
        torch.save(gold, os.path.join(HERE, name + ".pt"))
        print("  wrote", name + ".pt", os.path.getsize(os.path.join(HERE, name + ".pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
