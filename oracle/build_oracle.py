"""Build the CPU oracle shared library (test infrastructure only).

    python oracle/build_oracle.py

Output: oracle/libfwi_oracle.so (git-ignored; travels to the GPU box with the snapshot).
-ffp-contract=off keeps one rounding per arithmetic op, which is what makes the fp32 forward
bit-identical to the reference's eager tensor expression (solvers/pde.py:79).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libfwi_oracle.so")
SRCS = [os.path.join(HERE, "fwi_oracle.c")]
DEPS = SRCS + [os.path.join(HERE, "fwi_oracle_impl.h"), os.path.join(HERE, "fwi_oracle.h")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-Wall", "-Wextra", "-Wno-unused-function", "-o", LIB] + SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
