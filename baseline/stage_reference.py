"""Stage the UNMODIFIED reference (EryiXie/PlaneRecNet, /root/reference) under baseline/_ref/ so that it travels to the
GPU box with the repo snapshot (baseline/_ref/ is git-ignored, not gpurun-ignored; nothing of it enters the history).

The reference is not an installable package (no setup.py / pyproject: `pip install /root/reference` fails with
"neither 'setup.py' nor 'pyproject.toml' found"), so "install" is a byte-for-byte copy of its source tree; a manifest
with the sha256 of every copied file is written next to it (baseline/_ref/MANIFEST.sha256) and checked by ref_runner.py.

Usage: python baseline/stage_reference.py [--src /root/reference]"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SKIP_EXT = (".png",)          # 4.3 MB of README figures: not code


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def stage(src="/root/reference", quiet=False):
    if not os.path.isdir(src):
        return None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    lines = []
    for root, dirs, files in os.walk(src):
        dirs[:] = sorted(d for d in dirs if d not in (".git", "__pycache__"))
        for f in sorted(files):
            if f.endswith(SKIP_EXT) or f.endswith(".pyc"):
                continue
            s = os.path.join(root, f)
            rel = os.path.relpath(s, src)
            d = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            lines.append(f"{sha(d)}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if not quiet:
        print(f"staged {len(lines)} files of {src} under {DST}")
    return DST


def verify():
    """True when every staged file still has the digest recorded at staging time."""
    man = os.path.join(DST, "MANIFEST.sha256")
    if not os.path.exists(man):
        return False
    with open(man) as fh:
        for line in fh:
            digest, rel = line.rstrip("\n").split("  ", 1)
            p = os.path.join(DST, rel)
            if not os.path.exists(p) or sha(p) != digest:
                return False
    return True


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "/root/reference"
    if stage(src) is None:
        print(f"{src} not found: nothing staged")
        sys.exit(1)
