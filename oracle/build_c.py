"""TEST INFRASTRUCTURE ONLY -- builds oracle/_build/libosq_oracle.so from oracle/osq_oracle.c with gcc.
(The reference is pure Python, so there is no oracle/_ref to compile; see DESIGN.md.)"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libosq_oracle.so")
SRC = os.path.join(HERE, "osq_oracle.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(True))
