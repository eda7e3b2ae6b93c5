"""Build the CUDA shared library (libqlcuda.so) in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("QLC_LIB") or os.path.join(_HERE, "libqlcuda.so")   # QLC_LIB: A/B-test another build of the library
SOURCES = [os.path.join(_HERE, "csrc", "qlc_api.cu")]
HEADERS = [
    os.path.join(_HERE, "csrc", "kernels.cuh"),
    os.path.join(_HERE, "csrc", "physics.cuh"),
    os.path.join(_HERE, "csrc", "qnet.cuh"),
    os.path.join(_HERE, "csrc", "qnet_conv.cuh"),
    os.path.join(os.path.dirname(_HERE), "include", "ql_cuda.h"),
]
# -fmad=false: the reference's Rust never contracts a*b+c; results must be bit-identical to that arithmetic.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def needs_build():
    if os.environ.get("QLC_LIB"):
        return False
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile libqlcuda.so if missing or stale. Returns the path."""
    if not force and not needs_build():
        return SO_PATH
    try:
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH + ".tmp"] + SOURCES
        env = dict(os.environ)
        env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc; let nvcc use the system g++
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout)
    except (RuntimeError, OSError):
        # a prebuilt library that travelled with the tree (file times are not preserved by every copy) is still the
        # CUDA product; only a missing library is fatal
        if not force and os.path.exists(SO_PATH):
            return SO_PATH
        raise
    os.replace(SO_PATH + ".tmp", SO_PATH)
    if verbose:
        print(res.stdout)
    return SO_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
