"""Compiles the C part of the oracle (oracle/c/*.c) into oracle/_build/liboracle.so with gcc.
Test infrastructure only; the reference itself ships no compilable native code, so there
is no oracle/_ref."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c")
OUT = os.path.join(HERE, "_build", "liboracle.so")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(SRC, f)) > t for f in os.listdir(SRC))


def build(force=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith(".c"))
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-pthread", "-o", OUT,
                           *srcs, "-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
