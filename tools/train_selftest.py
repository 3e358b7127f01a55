"""Run the training-kernel parity cases group by group, each group in its own subprocess with a hard timeout
(a trapped kernel poisons its CUDA context), logging to gpurun_out/train_selftest.log."""
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
LOG = os.path.join(ROOT, "gpurun_out", "train_selftest.log")

GROUPS = ["wgrad0", "dgrad", "bn", "misc", "dcn", "heads"]   # "wgrad1" = swapped LBO/SBO probe (expected to fail)


def log(msg):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    with open(LOG, "a") as fh:
        fh.write(msg + "\n")
    print(msg, flush=True)


def run_group(group):
    import train_cases as TC
    jobs = []
    if group == "wgrad0":
        jobs = [(n, lambda n=n: TC.run_wgrad_case(n)) for n in TC.WGRAD_CASES]
    elif group == "wgrad1":
        jobs = [(n + "/swapped", lambda n=n: TC.run_wgrad_case(n, flags=1)) for n in list(TC.WGRAD_CASES)[:3]]
    elif group == "dgrad":
        jobs = [(n, lambda n=n: TC.run_dgrad_case(n)) for n in TC.DGRAD_CASES]
    elif group == "bn":
        for kw in [dict(), dict(residual=False), dict(relu=False, residual=False, Cc=64, rows=777), dict(Cc=2048, rows=600),
                   dict(dtype="f16")]:
            jobs.append(("bn" + json.dumps(kw), lambda kw=kw: TC.run_bn_case(**kw)))
    elif group == "misc":
        jobs = [("maxpool_bwd", TC.run_maxpool_bwd_case), ("adds", TC.run_add_cases)]
    elif group == "dcn":
        jobs = [(n, lambda n=n: TC.run_dcn_case(n)) for n in TC.DCN_CASES]
    elif group == "heads":
        jobs = [("gn", TC.run_gn_case), ("gn128", lambda: TC.run_gn_case(Cc=128, HW=777, B=2)),
                ("resample", TC.run_resample_bwd_cases), ("reflect_up1", lambda: TC.run_reflect_dgrad_case(up=1)),
                ("reflect_up2", lambda: TC.run_reflect_dgrad_case(up=2)),
                ("reflect_up2_acc", lambda: TC.run_reflect_dgrad_case(up=2, accumulate=True)),
                ("reflect_3x3", lambda: TC.run_reflect_dgrad_case(up=1, H=3, W=3)), ("softplus", TC.run_softplus_case)]
    ok = 0
    for name, fn in jobs:
        try:
            r = fn()
            print("PASS " + name + " " + json.dumps(r, default=float), flush=True)
            ok += 1
        except Exception as e:  # noqa: BLE001
            print("FAIL " + name + " :: " + str(e).splitlines()[0][:300], flush=True)
            if "CUDA" in str(e) or "status -2" in str(e):
                traceback.print_exc()
                break
    print(f"GROUP {group}: {ok}/{len(jobs)}", flush=True)
    return 0 if ok == len(jobs) else 1


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--group":
        return run_group(sys.argv[2])
    groups = sys.argv[1:] or GROUPS
    bad = 0
    for gname in groups:
        t0 = time.time()
        p = subprocess.Popen([sys.executable, "-u", os.path.abspath(__file__), "--group", gname], stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        try:
            out, _ = p.communicate(timeout=int(os.environ.get("PRN_GROUP_TIMEOUT", "240")))
        except subprocess.TimeoutExpired:
            p.kill()
            out = "TIMEOUT\n" + (p.communicate()[0] or "")
        lines = [l for l in (out or "").splitlines() if l.startswith(("PASS", "FAIL", "GROUP", "TIMEOUT")) or "Error" in l]
        log(f"---- {gname} ({time.time() - t0:.1f}s, rc={p.returncode})")
        for l in lines:
            log("  " + l)
        bad += p.returncode != 0
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
