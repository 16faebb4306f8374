"""Build libmvk.so (hand-written sm_100a CUDA behind the C ABI in include/mvk.h) in-tree.

    python -m mvkpconv_b200.build        (or __graft_entry__.build())

nvcc cross-compiles without a GPU; the .so stays next to this file (git-ignored, but it travels
to the GPU box with the repo snapshot).
"""
import glob
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmvk.so")
STAMP = os.path.join(HERE, "libmvk.so.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
        os.path.join(os.path.dirname(HERE), "include", "mvk.h")]
    for f in files:
        h.update(os.path.basename(f).encode())  # NOT the absolute path: the repo is copied to other roots (GPU box)
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    return shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)


def is_current():
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into libmvk.so for sm_100a.  Returns the library path."""
    if not force and is_current():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("libmvk.so is missing/stale and nvcc was not found: cannot build the CUDA path")
    # several ranks may import the package at once: serialise the build and publish the library atomically
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return LIB  # another process built it while we waited
            tmp = LIB + ".tmp.%d" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + sources()
            if verbose:
                cmd += ["-Xptxas", "-v"]
                print(" ".join(cmd))
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
            os.replace(tmp, LIB)
            open(STAMP, "w").write(_digest())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
