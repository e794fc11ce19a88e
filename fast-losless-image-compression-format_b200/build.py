"""In-tree nvcc build of libflicb200.so for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SOURCES = ["csrc/encode.cu", "csrc/decode.cu", "csrc/decode_one.cu", "csrc/splice.cu", "csrc/api.cu"]
_DEPS = _SOURCES + ["csrc/common.cuh", "csrc/decode_common.cuh", "../include/flic_b200.h"]


def library_path() -> str:
    # FLIC_LIB: A/B experiments load another build of the same sources (tools/ab.sh); never set in tests or the bench
    return os.environ.get("FLIC_LIB") or os.path.join(_HERE, "libflicb200.so")


def _stale(out: str) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(_HERE, d)) > t for d in _DEPS)


def build_library(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """Compile the CUDA engine + C ABI. nvcc cross-compiles without a GPU.
    out / defines: an A/B build of the same sources under another name (tools/ab.py), e.g. defines=("FLIC_SUBHIST=8",)."""
    if out is None:
        out = library_path()
        if os.environ.get("FLIC_LIB"):
            return out
        if not force and not _stale(out):
            return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o", out,
    ] + ["-D" + d for d in defines] + _SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, cwd=_HERE, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out
