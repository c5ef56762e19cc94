"""Build the C half of the oracle (TEST INFRASTRUCTURE, never on the product path).

    python oracle/build_oracle.py

compiles oracle/sift_oracle.c into oracle/_build/libsift_oracle.so with gcc.  The reference is a
pure-Python package (no C/C++ sources to compile), so there is no `oracle/_ref`; the reference is
instead imported in the build container by `oracle/make_golden.py` to freeze golden vectors.
"""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sift_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libsift_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = [
        "gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
        "-o", OUT, SRC, "-lm",
    ]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
