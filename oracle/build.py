"""Compile the plain-C oracle (``hypergeom_oracle.c``) with gcc.

TEST INFRASTRUCTURE ONLY.  Output goes to ``oracle/_build/`` (git-ignored via
``*.so``; it travels to the GPU box with the snapshot).  The reference is pure
Python - nothing to compile; ``stage_reference`` places its hot-path modules,
unmodified, under ``oracle/_ref/`` (git-ignored build output, never committed) so
that ``bench.py --impl reference`` / ``cpu_baseline`` can time the reference
itself where ``/root/reference`` does not exist (the GPU box).
"""
import os
import shutil
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


REF_SRC = "/root/reference/pySpade"
REF_DST = os.path.join(HERE, "_ref", "pySpade")
REF_MODULES = ("__init__.py", "utils.py", "DEobs.py", "DErand.py")      # what oracle/ref_loader imports


def stage_reference() -> bool:
    """Copy the reference's hot-path modules byte for byte into ``oracle/_ref/pySpade``.
    Returns False (and changes nothing) where the reference tree is absent."""
    if not os.path.isfile(os.path.join(REF_SRC, "utils.py")):
        return os.path.isfile(os.path.join(REF_DST, "utils.py"))
    os.makedirs(REF_DST, exist_ok=True)
    for name in REF_MODULES:
        shutil.copyfile(os.path.join(REF_SRC, name), os.path.join(REF_DST, name))
    return True


if __name__ == "__main__":
    print(build(force=True))
