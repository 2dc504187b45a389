"""In-tree build of libhtf_b200.so: nvcc, sm_100a only, no torch dependency.

The library lands in ``hoomd-tf_b200/lib/`` (git-ignored, but it travels to the GPU box
with the gpurun snapshot).  ``python hoomd-tf_b200/build.py [--force] [--verbose]``.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhtf_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build(force=False, verbose=False, defines=(), out=None):
    """Compile every .cu under csrc/ into lib/libhtf_b200.so (returns the path).
    ``defines`` / ``out`` build an experimental variant next to it (kernel A/B timing)."""
    if out is not None:
        nvcc = find_nvcc()
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out] + sources()
        subprocess.check_call(cmd)
        return out
    if not force and not _stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found; cannot build libhtf_b200.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    # a host compiler that nvcc accepts; the image's CC wrapper is fine for C++ too, but
    # /usr/bin/g++ is what nvcc picks by default
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stdout.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (%d): %s" % (res.returncode, " ".join(cmd)))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
