"""Run every conv parity case in its own subprocess with a timeout, so that a trap or a hang in one
case cannot take the rest (or the GPU box) down.  Writes gpurun_out/selftest.json."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        import conv_cases
        d = conv_cases.run_case(sys.argv[2])
        print("RESULT " + json.dumps(d))
        return 0
    import conv_cases
    names = sys.argv[1:] or list(conv_cases.CASES)
    out = {}
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True,
                               text=True, timeout=120)
            ok = p.returncode == 0
            tail = (p.stdout + p.stderr).strip().splitlines()[-6:]
        except subprocess.TimeoutExpired:
            ok, tail = False, ["TIMEOUT"]
        out[n] = {"ok": ok, "sec": round(time.time() - t0, 1), "tail": tail}
        print(("PASS " if ok else "FAIL ") + n, "|", tail[-1] if tail else "", flush=True)
        if not ok:
            for l in tail:
                print("     ", l)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "selftest.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    return 0 if all(v["ok"] for v in out.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
