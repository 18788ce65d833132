"""Compile the plain-C oracle (``hypergeom_oracle.c``) with gcc.

TEST INFRASTRUCTURE ONLY.  Output goes to ``oracle/_build/`` (git-ignored via
``*.so``; it travels to the GPU box with the snapshot).  There is no
``oracle/_ref``: the reference is pure Python (no compiled sources to build),
see DESIGN.md.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hypergeom_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libspade_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.isfile(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
