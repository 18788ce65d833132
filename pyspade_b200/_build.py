"""Build ``libspade_b200.so`` in-tree with nvcc for sm_100a.

The shared library is the product's only native artefact; it is git-ignored
(``*.so``) but travels to the GPU box with the ``gpurun`` snapshot.
"""
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "spade_kernels.cu")
HDR = os.path.join(ROOT, "include", "spade.h")
OUT = os.path.join(PKG, "libspade_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-shared", "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
]


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    deps = [HDR] + [os.path.join(os.path.dirname(SRC), f) for f in os.listdir(os.path.dirname(SRC))]
    newest = max(os.path.getmtime(p) for p in deps)
    return os.path.getmtime(OUT) < newest


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """``out`` / ``defines``: tuning variants (e.g. ``defines=["SPADE_MIN_CTAS=5"]``) built next
    to the product library and selected with the ``SPADE_LIB`` environment variable."""
    if out == OUT and not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = ([nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
           + ["-o", out, SRC])
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
