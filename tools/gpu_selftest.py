"""Run conv parity cases each in its own subprocess with a hard timeout, logging incrementally to
gpurun_out/selftest.log so that partial progress survives a killed command."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
LOG = os.path.join(ROOT, "gpurun_out", "selftest.log")


def log(msg):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    with open(LOG, "a") as fh:
        fh.write(msg + "\n")
    print(msg, flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        t0 = time.time()
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("PRN_FAULT_AFTER", "20")), exit=False)
        import torch
        sys.stderr.write(f"[{time.time() - t0:.1f}s] torch imported\n")
        torch.zeros(1, device="cuda")
        torch.cuda.synchronize()
        sys.stderr.write(f"[{time.time() - t0:.1f}s] cuda context up\n")
        import conv_cases
        d = conv_cases.run_case(sys.argv[2])
        sys.stderr.write(f"[{time.time() - t0:.1f}s] case done\n")
        print("RESULT " + json.dumps(d))
        return 0
    import conv_cases
    per_case = int(os.environ.get("PRN_CASE_TIMEOUT", "60"))
    names = sys.argv[1:] or list(conv_cases.CASES)
    n_ok = 0
    for n in names:
        t0 = time.time()
        p = subprocess.Popen([sys.executable, "-u", os.path.abspath(__file__), "--one", n], stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        try:
            out, _ = p.communicate(timeout=per_case)
            ok = p.returncode == 0
        except subprocess.TimeoutExpired:
            p.kill()
            try:
                out, _ = p.communicate(timeout=10)
            except subprocess.TimeoutExpired:
                out = "(child did not die after SIGKILL)"
            ok, out = False, "TIMEOUT\n" + (out or "")
        tail = [l for l in (out or "").strip().splitlines() if "Warning" not in l][-8:]
        n_ok += ok
        log(("PASS " if ok else "FAIL ") + f"{n} ({time.time() - t0:.1f}s) | " + (tail[-1] if tail else ""))
        if not ok:
            for l in tail:
                log("      " + l)
    log(f"== {n_ok}/{len(names)} cases passed")
    return 0 if n_ok == len(names) else 1


if __name__ == "__main__":
    sys.exit(main())
