"""Builds libu2pl_b200.so (sm_100a only) in-tree with nvcc.  No torch involved: the
library's boundary is the plain C ABI of include/u2pl_b200.h.

Every csrc/*.cu is compiled to its own object (in parallel, re-compiled only when its digest --
the source, every header, the flags -- changes) and the objects are linked into a temporary
file that is renamed over the library atomically.  The whole build runs under an exclusive file
lock, so N ranks of a torchrun job that all find a stale library build it once: the first
holder compiles, the others block on the lock, re-check the stamp and return."""
import concurrent.futures
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libu2pl_b200.so")
STAMP = os.path.join(HERE, ".libu2pl_b200.stamp")
LOCK = os.path.join(HERE, ".libu2pl_b200.lock")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _header_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _source_digest(src, hdr):
    h = hashlib.sha256(hdr.encode())
    with open(src, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _digest():
    hdr = _header_digest()
    h = hashlib.sha256()
    for s in _sources():
        h.update(os.path.basename(s).encode())
        h.update(_source_digest(s, hdr).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def _compile_one(nvcc, src, hdr):
    name = os.path.splitext(os.path.basename(src))[0]
    obj, stamp = os.path.join(OBJ, name + ".o"), os.path.join(OBJ, name + ".digest")
    log = os.path.join(OBJ, name + ".ptxas")
    dig = _source_digest(src, hdr)
    if os.path.exists(obj) and os.path.exists(stamp) and os.path.exists(log) and open(stamp).read().strip() == dig:
        return obj, open(log).read(), 0, ""
    tmp = obj + f".tmp{os.getpid()}"
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", "-o", tmp, src], capture_output=True, text=True)
    if res.returncode != 0:
        return obj, "", res.returncode, res.stdout + res.stderr
    os.replace(tmp, obj)
    with open(log, "w") as fh:
        fh.write(res.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, res.stderr, 0, res.stdout + res.stderr


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Returns the library path."""
    if not force and is_current():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libu2pl_b200.so cannot be built (and there is no fallback path)")
    os.makedirs(OBJ, exist_ok=True)
    with open(LOCK, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)                       # one builder per checkout; the others wait here
        try:
            if not force and is_current():                     # another process built it while we waited
                return LIB
            if force:
                shutil.rmtree(OBJ, ignore_errors=True)
                os.makedirs(OBJ, exist_ok=True)
            hdr = _header_digest()
            srcs = _sources()
            with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
                results = list(ex.map(lambda s: _compile_one(nvcc, s, hdr), srcs))
            failed = [r for r in results if r[2] != 0]
            if verbose or failed:
                sys.stderr.write("".join(r[3] for r in results))
            if failed:
                raise RuntimeError("nvcc failed building libu2pl_b200.so")
            tmp = LIB + f".tmp{os.getpid()}"
            res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + [r[0] for r in results],
                                 capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed linking libu2pl_b200.so")
            os.replace(tmp, LIB)                               # atomic: a concurrent dlopen sees the old or the new file
            with open(os.path.join(HERE, "ptxas_report.txt"), "w") as fh:      # keep the report diff-stable
                fh.write("".join(l for r in results for l in r[1].splitlines(True) if "Compile time" not in l))
            with open(STAMP, "w") as fh:
                fh.write(_digest())
            return LIB
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
