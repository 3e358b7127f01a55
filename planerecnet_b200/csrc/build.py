"""Build libprn_b200.so (sm_100a only) in-tree with nvcc.  No torch involved: the library is a plain
C-ABI shared object (see include/prn_b200.h)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libprn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "-I", os.path.join(ROOT, "include"),
]


def sources():
    return sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cu"))


def _fingerprint():
    h = hashlib.sha256()
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".cu", ".cuh", ".h", ".py")):
            with open(os.path.join(HERE, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(ROOT, "include", "prn_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def fingerprint():
    """sha256 over the kernel sources, this script and the public header: compiled into the library
    (prn_build_fingerprint()) so that a stale .so left over from another checkout is detected at load time."""
    return _fingerprint()


def embedded_fingerprint(path=LIB):
    """The fingerprint the shared library was built from ('' if none).  Read from the file's bytes (marker "PRN_FP:"), not
    through dlopen: a handle opened here would be served again, stale, after a rebuild in the same process."""
    try:
        with open(path, "rb") as fh:
            blob = fh.read()
    except OSError:
        return ""
    i = blob.find(b"PRN_FP:")
    return blob[i + 7:i + 7 + 64].decode("ascii", "replace") if i >= 0 else ""


def build(force=False, verbose=False):
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and embedded_fingerprint() == fp:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f'-DPRN_BUILD_FP="{fp}"', "-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-shared", "-arch=sm_100a", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
