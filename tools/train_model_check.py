"""Whole-model training step on the GPU against autograd through the CPU oracle: prints per-parameter-family
gradient agreement.  Usage: python tools/train_model_check.py [preset] [B H W] [bf16|f16] [train|frozen] [cond|plain]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import train_cases as TC  # noqa: E402


def main():
    a = sys.argv[1:]
    preset = a[0] if len(a) > 0 else "PlaneRecNet_50_config"
    B, H, W = (int(v) for v in a[1:4]) if len(a) > 3 else (2, 128, 160)
    prec = a[4] if len(a) > 4 else "bf16"
    bn_mode = a[5] if len(a) > 5 else "train"
    cond = (a[6] if len(a) > 6 else "cond") == "cond"
    print(f"== {preset} {B}x{H}x{W} precision={prec} bn={bn_mode} cond={cond}")
    r = TC.run_model_check(preset, B, H, W, prec, bn_mode, cond)
    print("  outputs rel-L2:", " ".join(f"{v:.2e}" for v in r["outs"]))
    for f, (mn, mean) in sorted(r["fam"].items()):
        print(f"  {f:28s} cos min {mn:.4f} mean {mean:.4f}")
    print(f"  ALL: cos {r['all_cos']:.5f} rel_l2 {r['all_rel']:.4f}  launches {r['launches']}  missing {r['missing']}")


if __name__ == "__main__":
    main()
